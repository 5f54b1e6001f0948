"""Rebind the reference's hot-path entry points to the device implementations.

`install()` patches, in an imported `EasyFEA` package:
  * level 0 — the cached Gauss-point getters of `FEM._GroupElem`: `Get_F_e_pg`, `Get_jacobian_e_pg`, `Get_invF_e_pg`,
    `Get_dN_e_pg`, `Get_B_e_pg`, `Get_leftDispPart_e_pg`, `Get_ReactionPart_e_pg`, `Get_DiffusePart_e_pg`,
    `Get_SourcePart_e_pg` (`FEM/_group_elem.py:832-1407`), kept behind the reference's own `@cache_computed_values`;
  * level 1 — `EasyFEA.FEM.Operators.Bilinear.{LinearizedElasticity, UV, GradUGradV, GradU_A_GradV}` and
    `Operators.Linear.{V, InternalForce}` (looked up at call time by the simulations: `_elastic.py:132,135`,
    `Simulations/_phasefield.py:472,554,557,560`, `_thermal.py:122,126`) and
    `Operators.NonLinear.SecondPiolaKirchhoffStressTensor` (`_hyperelastic.py:317`);
  * level 2 — `_Simu._Simu__Get_csr_map` / `_Simu._Simu__Assemble_csr` (`_simu.py:989-1102`);
  * level 3 — `Models.PhaseField.{Calc_C, Calc_psi_e_pg, Calc_Sigma_e_pg, Get_g_e_pg}` for homogeneous isotropic materials
    and the splits on the path (other models fall through to the reference's own code);
  * level 4 — the linear solve: `Solvers.Solve_simu` (`Simulations/Solvers.py:397`, also bound as `_simu.Solve_simu`,
    `_simu.py:45,1544,1710`) and `Solvers._Solve_Axb` (`Solvers.py:225`).  Systems with at least `min_dofs` unknowns go to
    the device Jacobi-PCG (Dirichlet dofs masked, no host-side slicing of `A`); smaller ones, Lagrange-multiplier systems,
    bound-constrained solves and MPI runs fall through to the reference's own solver (north star: "small meshes hand a
    scipy-compatible CSR back to the reference's own solver").
Groups the device path does not cover (1D elements, elements embedded in a higher dimension) fall through to the
reference's own operator, like unsupported models at level 3.  `uninstall()` restores the originals.  Nothing here
computes on the CPU: the replacements raise `EfbError` without a GPU.
"""
from __future__ import annotations

import numpy as np

from . import assembly, operators, phasefield

_saved = {}
_LEVEL1 = {"Bilinear": ("LinearizedElasticity", "UV", "GradUGradV", "GradU_A_GradV"), "Linear": ("V", "InternalForce"),
           "NonLinear": ("SecondPiolaKirchhoffStressTensor",)}
# level 0: reference getter name -> (device function, takes dof_n)
_LEVEL0 = ("Get_F_e_pg", "Get_jacobian_e_pg", "Get_invF_e_pg", "Get_dN_e_pg", "Get_B_e_pg", "Get_leftDispPart_e_pg",
           "Get_ReactionPart_e_pg", "Get_DiffusePart_e_pg", "Get_SourcePart_e_pg")

config = {"min_dofs": 100_000, "pcg_tol": 1e-10, "pcg_maxiter": None}
stats = {"device_solves": 0, "host_solves": 0, "last_info": None}


def _wrap_fearray(EasyFEA, arr):
    """operator results are plain ndarrays in the reference; law results are FeArrays (Models/_phasefield.py:396-431)"""
    return EasyFEA.FEM.FeArray.asfearray(arr)


def _with_fallback(device_fn, original):
    """device operator that hands groups outside the path (NotImplementedError from the device mirror) to the reference"""

    def op(*args, **kwargs):
        try:
            return device_fn(*args, **kwargs)
        except NotImplementedError:
            return original(*args, **kwargs)

    op.__name__ = getattr(original, "__name__", device_fn.__name__)
    op.__doc__ = device_fn.__doc__
    return op


# ---------------------------------------------------------------------------------------------------------
# level 4: the linear solve on the device
# ---------------------------------------------------------------------------------------------------------
def device_solve(A, b, x_start, free_mask=None, tol=None, maxiter=None):
    """Jacobi-PCG of `A x = b` on the device for a scipy CSR matrix (rows with `free_mask == 0` keep `x_start`).
    Returns (x, info).  Raises EfbError if A is not symmetric or the solve does not converge."""
    import torch
    from scipy import sparse

    from . import _lib
    from . import device as dv
    from .assembly import DeviceCsr
    from .solver import pcg, spmv

    if not sparse.isspmatrix_csr(A):
        A = sparse.csr_matrix(A)
    if not A.has_canonical_format:
        A.sum_duplicates()
    n = A.shape[0]
    Ad = DeviceCsr(dv.to_device(A.indptr), dv.to_device(A.indices), dv.to_device(A.data), A.shape)
    # the conjugate gradient needs a symmetric operator: one randomised check (two products) before trusting it
    g = torch.Generator(device="cpu").manual_seed(0)
    u = dv.to_device(torch.rand(n, generator=g, dtype=torch.float64))
    v = dv.to_device(torch.rand(n, generator=g, dtype=torch.float64))
    Au, Av = spmv(Ad, u), spmv(Ad, v)
    s1, s2 = float(torch.dot(v, Au)), float(torch.dot(u, Av))
    if abs(s1 - s2) > 1e-9 * max(abs(s1), abs(s2), 1e-300):
        raise _lib.EfbError("device_solve: the system matrix is not symmetric; the Jacobi-PCG does not apply")
    x, info = pcg(Ad, np.asarray(b, dtype=np.float64).ravel(), x0=np.asarray(x_start, dtype=np.float64).ravel(),
                  free_mask=free_mask, tol=config["pcg_tol"] if tol is None else tol,
                  maxiter=config["pcg_maxiter"] if maxiter is None else maxiter)
    if not info["converged"]:
        raise _lib.EfbError(f"device_solve: Jacobi-PCG did not converge ({info['iterations']} iterations, relative residual "
                            f"{info['rel_residual']:.3e})")
    stats["device_solves"] += 1
    stats["last_info"] = info
    return dv.to_host(x), info


def _install_level4(EasyFEA, patched):
    from EasyFEA.Simulations import Solvers, _simu

    orig_axb, orig_simu = Solvers._Solve_Axb, Solvers.Solve_simu
    ResolType = Solvers.ResolType

    def _host_only(simu, problemType):
        """cases the device solve does not cover: Lagrange multipliers (saddle point), bounds, MPI runs"""
        if len(simu.Bc_Lagrange) > 0 or Solvers.MPI_SIZE > 1:
            return True
        lb, ub = simu.Get_lb_ub(problemType)
        return len(lb) > 0 or len(ub) > 0

    def _Solve_Axb(simu, problemType, A, b, x0, lb, ub, resol=ResolType.r1, ownedDofs=None, mapping=None):
        """`Solvers._Solve_Axb` (Solvers.py:225): the reduced system of `__Solver_1` on the device when it is large"""
        if (resol != ResolType.r1 or A.shape[0] < config["min_dofs"] or ownedDofs is not None or len(lb) > 0 or len(ub) > 0
                or len(simu.Bc_Lagrange) > 0):
            stats["host_solves"] += 1
            return orig_axb(simu, problemType, A, b, x0, lb, ub, resol, ownedDofs, mapping)
        rhs = b.toarray().ravel() if hasattr(b, "toarray") else np.asarray(b).ravel()
        x, _ = device_solve(A, rhs, x0)
        return np.array(x)

    def Solve_simu(simu, problemType):
        """`Solvers.Solve_simu` (Solvers.py:397) for resolution r1: `A x = b` with the known dofs masked instead of the
        host-side `A[dofsUnknown, :].tocsc()[:, dofsUnknown]` slicing of `__Solver_1` (Solvers.py:502-553)"""
        Ndof = simu.mesh.Nn * simu.Get_dof_n(problemType)
        if Ndof < config["min_dofs"] or _host_only(simu, problemType):
            return orig_simu(simu, problemType)
        b = simu._Solver_Apply_Neumann(problemType)
        A, x = simu._Solver_Apply_Dirichlet(problemType, b, ResolType.r1)
        dofsKnown, dofsUnknown = simu.Bc_dofs_known_unknown(problemType)
        x_start = np.array(simu.Get_x0(problemType), dtype=np.float64).ravel()
        xk = np.asarray(x.toarray()).ravel()
        x_start[dofsKnown] = xk[dofsKnown]
        free = np.zeros(Ndof, dtype=np.uint8)
        free[dofsUnknown] = 1
        rhs = np.asarray(b.toarray()).ravel()
        sol, info = device_solve(A, rhs, x_start, free_mask=free)
        sol[dofsKnown] = xk[dofsKnown]
        if simu.isNonLinear:  # ||b_i - A_ic x_c||, what __Solver_1 returns as the residual norm of a Newton step
            return sol, float(info["rhs_norm"])
        return sol, None

    for mod, name, fn in ((Solvers, "_Solve_Axb", _Solve_Axb), (Solvers, "Solve_simu", Solve_simu), (_simu, "Solve_simu", Solve_simu)):
        _saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, fn)
        patched.append(f"{mod.__name__.split('.')[-1]}.{name}")


def _install_level0(EasyFEA, patched):
    from EasyFEA.FEM._group_elem import _GroupElem
    from EasyFEA.Utilities._cache import cache_computed_values

    def make(name):
        original = _GroupElem.__dict__[name]
        device_fn = getattr(operators, name)

        def getter(self, *args, **kwargs):
            if self.dim == 0:
                return None
            try:
                return _wrap_fearray(EasyFEA, device_fn(self, *args, **kwargs))
            except NotImplementedError:
                return getattr(original, "__wrapped__", original)(self, *args, **kwargs)

        getter.__name__ = name  # the cache key of @cache_computed_values
        getter.__doc__ = original.__doc__
        return cache_computed_values(getter), original

    for name in _LEVEL0:
        fn, original = make(name)
        _saved[(_GroupElem, name)] = original
        setattr(_GroupElem, name, fn)
        patched.append(f"_GroupElem.{name}")


def install(EasyFEA=None, levels=(0, 1, 2, 3, 4), min_dofs=None, pcg_tol=None):
    """Patch the imported reference package (default: `import EasyFEA`).  Returns the list of patched attribute names.
    `min_dofs`: systems with at least this many dofs are solved on the device (level 4; default 100 000);
    `pcg_tol`: relative residual of the device Jacobi-PCG (default 1e-10)."""
    if EasyFEA is None:
        import EasyFEA  # noqa: F811
    if _saved:
        raise RuntimeError("easyfea_b200.dropin is already installed")
    if min_dofs is not None:
        config["min_dofs"] = int(min_dofs)
    if pcg_tol is not None:
        config["pcg_tol"] = float(pcg_tol)
    patched = []
    Operators = EasyFEA.FEM.Operators
    if 0 in levels:
        _install_level0(EasyFEA, patched)
    if 1 in levels:
        for mod, names in _LEVEL1.items():
            m = getattr(Operators, mod)
            for name in names:
                _saved[(m, name)] = getattr(m, name)
                setattr(m, name, _with_fallback(getattr(operators, name), _saved[(m, name)]))
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
        orig_C, orig_psi, orig_g, orig_sig = PF.Calc_C, PF.Calc_psi_e_pg, PF.Get_g_e_pg, PF.Calc_Sigma_e_pg

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

        def Calc_Sigma_e_pg(self, Epsilon_e_pg):
            m = _device_model(self)
            if not m:
                return orig_sig(self, Epsilon_e_pg)
            sP, sM = m.Calc_Sigma_e_pg(np.asarray(Epsilon_e_pg))
            return _wrap_fearray(EasyFEA, sP), _wrap_fearray(EasyFEA, sM)

        def Get_g_e_pg(self, d_n, groupElem, matrixType, k_res=1e-12):
            m = _device_model(self)
            if not m:
                return orig_g(self, d_n, groupElem, matrixType, k_res)
            try:
                return _wrap_fearray(EasyFEA, m.Get_g_e_pg(d_n, groupElem, matrixType, k_res))
            except NotImplementedError:
                return orig_g(self, d_n, groupElem, matrixType, k_res)

        for name, fn in (("Calc_C", Calc_C), ("Calc_psi_e_pg", Calc_psi_e_pg), ("Calc_Sigma_e_pg", Calc_Sigma_e_pg),
                         ("Get_g_e_pg", Get_g_e_pg)):
            _saved[(PF, name)] = PF.__dict__[name]
            setattr(PF, name, fn)
            patched.append(f"Models.PhaseField.{name}")
    if 4 in levels:
        _install_level4(EasyFEA, patched)
    return patched


def uninstall():
    """Restore every attribute `install()` replaced."""
    for key, val in list(_saved.items()):
        if len(key) == 2:
            setattr(key[0], key[1], val)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)
