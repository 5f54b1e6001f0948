"""Rebind the reference's hot-path entry points to the device implementations.

`install()` patches, in an imported `EasyFEA` package:
  * level 1 — `EasyFEA.FEM.Operators.Bilinear.{LinearizedElasticity, UV, GradUGradV, GradU_A_GradV}` and
    `Operators.Linear.{V, InternalForce}` (looked up at call time by the simulations: `_elastic.py:132,135`,
    `Simulations/_phasefield.py:472,554,557,560`, `_thermal.py:122,126`);
  * level 2 — `_Simu._Simu__Get_csr_map` / `_Simu._Simu__Assemble_csr` (`_simu.py:989-1102`);
  * level 3 — `Models.PhaseField.{Calc_C, Calc_psi_e_pg, Get_g_e_pg}` for homogeneous isotropic materials and the splits
    on the path (other models fall through to the reference's own code).
`uninstall()` restores the originals.  Nothing here computes on the CPU: the replacements raise `EfbError` without a GPU.
"""
from __future__ import annotations

import numpy as np

from . import assembly, operators, phasefield

_saved = {}
_LEVEL1 = {"Bilinear": ("LinearizedElasticity", "UV", "GradUGradV", "GradU_A_GradV"), "Linear": ("V", "InternalForce")}


def _wrap_fearray(EasyFEA, arr):
    """operator results are plain ndarrays in the reference; law results are FeArrays (Models/_phasefield.py:396-431)"""
    return EasyFEA.FEM.FeArray.asfearray(arr)


def install(EasyFEA=None, levels=(1, 2, 3)):
    """Patch the imported reference package (default: `import EasyFEA`).  Returns the list of patched attribute names."""
    if EasyFEA is None:
        import EasyFEA  # noqa: F811
    if _saved:
        raise RuntimeError("easyfea_b200.dropin is already installed")
    patched = []
    Operators = EasyFEA.FEM.Operators
    if 1 in levels:
        for mod, names in _LEVEL1.items():
            m = getattr(Operators, mod)
            for name in names:
                _saved[(m, name)] = getattr(m, name)
                setattr(m, name, getattr(operators, name))
                patched.append(f"Operators.{mod}.{name}")
    if 2 in levels:
        from EasyFEA.Simulations._simu import _Simu

        asm = assembly.Assembler()

        def Get_csr_map(self, dof_n, isMatrix, Ndof, groups):
            return asm.Get_csr_map(dof_n, isMatrix, Ndof, tuple(groups))

        def Assemble_csr(self, dict_group_data, dof_n, Ndof, isMatrix):
            return asm.Assemble_csr(dict_group_data, dof_n, Ndof, isMatrix)

        for name, fn in (("_Simu__Get_csr_map", Get_csr_map), ("_Simu__Assemble_csr", Assemble_csr)):
            _saved[(_Simu, name)] = _Simu.__dict__[name]
            setattr(_Simu, name, fn)
            patched.append(f"_Simu.{name}")
        _saved[("assembler",)] = asm
    if 3 in levels:
        PF = EasyFEA.Models.PhaseField
        orig_C, orig_psi, orig_g = PF.Calc_C, PF.Calc_psi_e_pg, PF.Get_g_e_pg

        def _device_model(self):
            key = "_efb_model"
            if getattr(self, key, None) is None:
                try:
                    object.__setattr__(self, key, phasefield.PhaseFieldModel.from_reference(self))
                except NotImplementedError:
                    object.__setattr__(self, key, False)
            return getattr(self, key)

        def Calc_C(self, Epsilon_e_pg, verif=False):
            m = _device_model(self)
            if not m:
                return orig_C(self, Epsilon_e_pg, verif)
            cP, cM = m.Calc_C(np.asarray(Epsilon_e_pg), verif)
            return _wrap_fearray(EasyFEA, cP), _wrap_fearray(EasyFEA, cM)

        def Calc_psi_e_pg(self, Epsilon_e_pg):
            m = _device_model(self)
            if not m:
                return orig_psi(self, Epsilon_e_pg)
            pP, pM = m.Calc_psi_e_pg(np.asarray(Epsilon_e_pg))
            return _wrap_fearray(EasyFEA, pP), _wrap_fearray(EasyFEA, pM)

        def Get_g_e_pg(self, d_n, groupElem, matrixType, k_res=1e-12):
            m = _device_model(self)
            if not m:
                return orig_g(self, d_n, groupElem, matrixType, k_res)
            return _wrap_fearray(EasyFEA, m.Get_g_e_pg(d_n, groupElem, matrixType, k_res))

        for name, fn in (("Calc_C", Calc_C), ("Calc_psi_e_pg", Calc_psi_e_pg), ("Get_g_e_pg", Get_g_e_pg)):
            _saved[(PF, name)] = PF.__dict__[name]
            setattr(PF, name, fn)
            patched.append(f"Models.PhaseField.{name}")
    return patched


def uninstall():
    """Restore every attribute `install()` replaced."""
    for key, val in list(_saved.items()):
        if len(key) == 2:
            setattr(key[0], key[1], val)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)
