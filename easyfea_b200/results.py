"""Post-processing fields on the device: what `Simulations.Elastic.Result` computes for the
strain / stress family and the deformation energy, and `Mesh.Get_Node_Values`.

  * strain at the Gauss points: `efb_strain` (`_Elastic.Calc_Epsilon_e_pg`, EasyFEA/Models/Elastic/_laws.py:127-157)
  * stress: `efb_hooke` (`Calc_Sigma_e_pg`, _laws.py:159-185)
  * per-element results `Sxx ... Exy, Svm, Evm, Stress, Strain`: `efb_field_result`
    (`Result_strain_or_stress_field_e`, EasyFEA/Models/_utils.py:302-430; dispatch of `Elastic.Result`, Simulations/_elastic.py:292-313)
  * `Wdef`, `Wdef_e`: `efb_energy_e` (`_Calc_Psi_Elas`, Simulations/_elastic.py:323-396, raw element stresses)
  * node values: `efb_node_values` (`Mesh.Get_Node_Values`, EasyFEA/FEM/_mesh.py:822-873) on the assembly pattern's node ->
    element lists, summed in ascending element order like scipy's `connect_n_e @ values_e`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import device as dv
from . import operators as op
from .assembly import NodeGraph
from .mesh import device_group
from .operators import tensor_mode

_COMPONENTS = {2: ["xx", "yy", "xy"], 3: ["xx", "yy", "zz", "yz", "xz", "xy"]}


def stress_dev(groupElem, eps: torch.Tensor, C) -> torch.Tensor:
    """sigma (Ne,nPg,ns) = C eps on the device; C (ns,ns), (Ne,ns,ns) or (Ne,nPg,ns,ns)"""
    Ne, nPg, ns = eps.shape
    Carr, mode = tensor_mode(C, Ne, nPg, ns)
    out = torch.empty_like(eps)
    _lib.call("efb_hooke", Ne, nPg, ns, dv.ptr(eps), dv.ptr(dv.to_device(Carr)), mode, dv.ptr(out), dv.stream_ptr())
    return out


def field_result_dev(field: torch.Tensor, result: str, coef: float = np.sqrt(2)) -> torch.Tensor:
    """per-element value(s) of a Kelvin-Mandel strain/stress field (Ne,nPg,ns): (Ne,) or (Ne,ns) for "Strain"/"Stress"."""
    Ne, nPg, ns = field.shape
    dim = 2 if ns == 3 else 3
    if result in ("Strain", "Stress"):
        what = -2
    elif "vm" in result:
        what = -1
    else:
        what = next((i for i, c in enumerate(_COMPONENTS[dim]) if c in result), None)
        if what is None:  # the reference's message, _utils.py:352-355 / 389-392
            raise Exception(f"result must be in [{', '.join(_COMPONENTS[dim])}, vm, Strain, Stress, Green-Lagrange, Piola-Kirchhoff]")
    out = dv.empty((Ne, ns) if what == -2 else (Ne,))
    _lib.call("efb_field_result", Ne, nPg, dim, dv.ptr(field.contiguous()), what, float(coef), dv.ptr(out), dv.stream_ptr())
    return out


def node_values_dev(groupElem, result_e: torch.Tensor, Nn: int = None) -> torch.Tensor:
    """element values (Ne,) / (Ne,i) -> node values (Nn,) / (Nn,i), one element group"""
    dg = device_group(groupElem)
    Nn = int(dg.Ncoords if Nn is None else Nn)
    cache = groupElem.__dict__.setdefault("_efb_node_lists", {})
    if Nn not in cache:
        gph = NodeGraph((groupElem,), Nn)
        cache[Nn] = (gph.rowptr, gph.qlist)
    rowptr, qlist = cache[Nn]
    r = result_e.contiguous()
    assert r.shape[0] == dg.Ne, "Must be of size (Ne,i)"  # _mesh.py:839
    ncols = 1 if r.ndim == 1 else int(np.prod(r.shape[1:]))
    out = dv.empty((Nn,) if r.ndim == 1 else (Nn, ncols))
    _lib.call("efb_node_values", Nn, dv.ptr(rowptr), dv.ptr(qlist), dg.nPe, dv.ptr(r), ncols, dv.ptr(out), dv.stream_ptr())
    return out


def Get_Node_Values(groupElem, result_e) -> np.ndarray:
    """`Mesh.Get_Node_Values(result_e)` for a single-group mesh: NumPy in, NumPy out"""
    return dv.to_host(node_values_dev(groupElem, dv.to_device(np.asarray(result_e, dtype=np.float64))))


class ElasticResults:
    """The strain / stress / energy results of `Simulations.Elastic.Result` for one element group and a material matrix C."""

    def __init__(self, groupElem, C, thickness: float = 1.0, coef: float = np.sqrt(2)):
        self.g, self.C, self.coef = groupElem, np.asarray(C, dtype=np.float64), float(coef)
        self.dim = int(device_group(groupElem).dim)
        self.thickness = float(thickness) if self.dim == 2 else 1.0

    def Results_Available(self):
        comps = _COMPONENTS[self.dim]
        return ["S" + c for c in comps] + ["E" + c for c in comps] + ["Svm", "Evm", "Stress", "Strain", "Wdef", "Wdef_e"]

    def result_dev(self, u, result: str, nodeValues: bool = True):
        if result in ("Wdef", "Wdef_e"):
            Wdef_e = self.Calc_Psi_Elas_dev(u)
            if result == "Wdef":
                return Wdef_e.sum()
            return node_values_dev(self.g, Wdef_e) if nodeValues else Wdef_e
        if not (("S" in result or "E" in result) and "_norm" not in result):
            raise NotImplementedError(f"The result '{result}' is not implemented yet.")
        isStrain = "E" in result or result == "Strain"  # Simulations/_elastic.py:296-300
        eps = op.strain_dev(self.g, u, op.RIGI)
        field = eps if isStrain else stress_dev(self.g, eps, self.C)
        res = result if result in ("Strain", "Stress") else result[-2:]
        val_e = field_result_dev(field, res, self.coef)
        return node_values_dev(self.g, val_e) if nodeValues else val_e

    def Result(self, u, result: str, nodeValues: bool = True):
        """NumPy result like `simu.Result(result, nodeValues)` for the displacement vector u"""
        out = self.result_dev(u, result, nodeValues)
        return float(out.item()) if result == "Wdef" else dv.to_host(out)

    def Calc_Psi_Elas_dev(self, u) -> torch.Tensor:
        """Wdef_e (Ne,) = thickness * sum_p wJ 1/2 sigma:eps"""
        eps = op.strain_dev(self.g, u, op.RIGI)
        sig = stress_dev(self.g, eps, self.C)
        wJ = op.geometry_dev(self.g, op.RIGI, ("wJ",))["wJ"]
        Ne, nPg, ns = eps.shape
        out = dv.empty((Ne,))
        _lib.call("efb_energy_e", Ne, nPg, ns, dv.ptr(eps), dv.ptr(sig), dv.ptr(wJ), self.thickness, dv.ptr(out), dv.stream_ptr())
        return out
