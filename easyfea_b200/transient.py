"""Device-resident time stepping of the linear problems of BASELINE config 5 (transient thermal, elastodynamics):
the reference's time-scheme system build — SURVEY.md section 8f rank 1 — on top of the assembled K, C, M.

What the reference does on the host with scipy every step and what runs here instead:

  * `_Simu._Solver_Apply_Neumann` (EasyFEA/Simulations/_simu.py:1758-1853): `b = F + sum (coef * Matrix) @ history vector`.
    Here the history vectors are combined FIRST (`efb_lincomb`), so a step costs at most one SpMV per matrix:
    `b = F + K @ wK + C @ wC + M @ wM` with `w_X = x_u u_n + x_v v_n + x_a a_n` (table in `_history_weights`).
  * `_Solver_Apply_Dirichlet` (:1855-1894): `A = coefK K + coefC C + coefM M` — K, C, M share one CSR pattern (same
    mesh, same dof_n), so A is a linear combination of the three value arrays; `__Solver_1` (Solvers.py:502-553)
    becomes the masked Jacobi-PCG of `easyfea_b200.solver` (prescribed values ride in the start vector).
  * `_Solver_Update_solutions` (:1552-1657): the correctors for v and a.

Schemes: parabolic (theta method), newmark, hht, midpoint, hht_newmark, euler_implicit, euler_explicit — the linear ones of
`AlgoType`.  Single GPU or row-sharded (`LocalSystem` with a partition): vectors are `[owned | halo]`, u/v/a keep their
halo entries fresh after every step because the history SpMVs read them.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import device as dv
from . import operators as op
from .solver import pcg, spmv
from .staggered import Dirichlet, LocalSystem, _apply

HYPERBOLIC = ("newmark", "hht", "midpoint", "hht_newmark", "euler_implicit", "euler_explicit")


def lincomb(terms, out=None) -> torch.Tensor:
    """out = sum c_k v_k on the device (`efb_lincomb`); terms = [(c, tensor), ...] (1..4), zero coefficients are dropped"""
    terms = [(float(c), v) for c, v in terms if c != 0.0 and v is not None]
    ref = terms[0][1] if terms else out
    if out is None:
        out = torch.empty_like(ref)
    if not terms:
        return out.zero_()
    assert len(terms) <= 4 and all(v.numel() == out.numel() and v.is_contiguous() for _, v in terms)
    coefs = (ctypes.c_double * len(terms))(*[c for c, _ in terms])
    _lib.call("efb_lincomb", out.numel(), len(terms), coefs, dv.ptr_array([v for _, v in terms]), dv.ptr(out), dv.stream_ptr())
    return out


def time_scheme_coefs(algo, dt, beta=0.25, gamma=0.5, alpha=0.5):
    """(coefK, coefC, coefM) of `_Solver_Get_K_C_M_coefs_for_time_scheme`, _simu.py:1399-1455"""
    if algo == "newmark":
        return 1.0, gamma / (beta * dt), 1 / (beta * dt**2)
    if algo == "hht":
        return 1 - alpha, (1 - alpha) * gamma / (beta * dt), (1 - alpha) / (beta * dt**2)
    if algo == "midpoint":
        return 0.5, 1 / dt, 2 / dt**2
    if algo == "hht_newmark":
        return 1 - alpha, gamma / (beta * dt), 1 / (beta * dt**2)
    if algo == "parabolic":
        return 1.0, 1 / (alpha * dt), 0.0
    if algo == "euler_implicit":
        return 1.0, 1 / dt, 1 / dt**2
    if algo == "euler_explicit":
        return 0.0, 0.0, 1.0
    raise NotImplementedError(f"Algo {algo} is not implemented here.")


def _history_weights(algo, dt, beta, gamma, alpha):
    """{matrix: (x_u, x_v, x_a)} with b = F + sum_X X @ (x_u u_n + x_v v_n + x_a a_n) — `_Solver_Apply_Neumann`, :1777-1853."""
    if algo == "parabolic":  # b += 1/(alpha dt) C (u + (1-alpha) dt v)
        return {"C": (1 / (alpha * dt), (1 - alpha) / alpha, 0.0)}
    if algo in ("newmark", "hht_newmark"):
        cC, cM = gamma / (beta * dt), 1 / (beta * dt**2)
        ut = (1.0, dt, dt**2 / 2 * (1 - 2 * beta))   # u~ = u + dt v + dt^2/2 (1-2 beta) a
        vt = (0.0, 1.0, dt * (1 - gamma))            # v~ = v + dt (1-gamma) a
        w = {"C": tuple(cC * a - b for a, b in zip(ut, vt)), "M": tuple(cM * a for a in ut)}
        if algo == "hht_newmark":
            w["K"] = (-alpha, 0.0, 0.0)
        return w
    if algo == "midpoint":
        return {"K": (-0.5, 0.0, 0.0), "C": (1 / dt, 0.0, 0.0), "M": (2 / dt**2, 2 / dt, 0.0)}
    if algo == "hht":
        cM, cC = 1 / (beta * dt**2), gamma / (beta * dt)
        return {"K": (-alpha, 0.0, 0.0),
                "C": (-(alpha - 1) * cC, -((alpha - 1) * (gamma / beta) + 1), -dt * (alpha - 1) * (gamma / (2 * beta) - 1)),
                "M": (-(alpha - 1) * cM, -(alpha - 1) / (beta * dt), -((alpha - 1) / (2 * beta) + 1))}
    if algo == "euler_implicit":
        return {"C": (1 / dt, 0.0, 0.0), "M": (1 / dt**2, 1 / dt, 0.0)}
    if algo == "euler_explicit":
        return {"K": (-1.0, 0.0, 0.0), "C": (0.0, -1.0, 0.0)}
    raise NotImplementedError(f"Algo {algo} is not implemented here.")


class TransientSolve:
    """`K u + C v + M a = F` of one element group, stepped on the device.  Element matrices come from the operator kernels
    (`set_element_matrices`) — see `thermal` / `elastodynamic` for the two systems of config 5."""

    def __init__(self, system: LocalSystem, dof_n: int):
        self.sys, self.d = system, int(dof_n)
        n = system.n_local * self.d
        dev = dv.device()
        self.u = torch.zeros(n, dtype=torch.float64, device=dev)
        self.v = torch.zeros(n, dtype=torch.float64, device=dev)
        self.a = torch.zeros(n, dtype=torch.float64, device=dev)
        self.bc = Dirichlet(n)
        self.K = self.C = self.M = None
        self.algo = "elliptic"
        self.pcg_tol, self.pcg_maxiter, self.pcg_fused, self.pcg_persistent = 1e-10, None, "auto", False
        self.pcg_single_reduction = "auto"
        self.pcg_precond_degree = "auto"
        self.info = {}

    # -- systems ---------------------------------------------------------------------------------------------------
    def set_element_matrices(self, Ke, Ce=None, Me=None):
        """assemble the owned rows of K, C, M (device element matrices `(Ne, ndof, ndof)`; None = absent)"""
        self.K = self.sys.matrix(Ke, self.d)  # every replay fills a fresh value array on the shared pattern
        self.C = None if Ce is None else self.sys.matrix(Ce, self.d)
        self.M = None if Me is None else self.sys.matrix(Me, self.d)
        self._A = None

    @staticmethod
    def _clone(A):
        from .assembly import DeviceCsr

        return DeviceCsr(A.indptr, A.indices, A.data.clone(), A.shape, A.node_graph)

    @classmethod
    def thermal(cls, system: LocalSystem, k: float, rho_c: float, thickness: float = 1.0):
        """`Thermal.Construct_local_matrix_system` (Simulations/_thermal.py:114-137): K_e = GradUGradV(k), C_e = UV(rho c)"""
        g = system.group
        scale = thickness if g.dim == 2 else 1.0
        self = cls(system, 1)
        Ke = op.diffusion_Ke_dev(g, None, k, op.RIGI, scale)
        Ce = op.mass_Me_dev(g, rho_c, 1, op.MASS, scale)
        self.set_element_matrices(Ke, Ce, None)
        return self

    @classmethod
    def elastodynamic(cls, system: LocalSystem, C_mat, rho: float, coefM: float = 0.0, coefK: float = 0.0, thickness: float = 1.0):
        """`Elastic.Construct_local_matrix_system` (Simulations/_elastic.py:123-152): K_e, M_e = UV(rho, dim),
        Rayleigh damping C_e = coefK K_e + coefM M_e (:148)"""
        g = system.group
        scale = thickness if g.dim == 2 else 1.0
        self = cls(system, g.dim)
        Ke = op.elastic_Ke_dev(g, np.asarray(C_mat, dtype=np.float64), op.RIGI, scale)
        Me = op.mass_Me_dev(g, rho, g.dim, op.MASS, scale)
        self.K = self.sys.matrix(Ke, self.d)
        del Ke
        self.M = self.sys.matrix(Me, self.d)
        del Me
        self.C = None
        if coefM != 0.0 or coefK != 0.0:  # same pattern: the damping matrix is a combination of the value arrays
            self.C = self._clone(self.K)
            lincomb([(coefK, self.K.data), (coefM, self.M.data)], out=self.C.data)
        self._A = None
        return self

    # -- schemes (names of _simu.py:1161-1290) -----------------------------------------------------------------------
    def Solver_Set_Elliptic_Algorithm(self):
        self.algo, self._A = "elliptic", None

    def Solver_Set_Parabolic_Algorithm(self, dt: float, alpha=1 / 2):
        assert dt > 0, "Time increment must be > 0"
        self.algo, self.dt, self.beta, self.gamma, self.alpha, self._A = "parabolic", float(dt), 0.25, 0.5, float(alpha), None

    def Solver_Set_Hyperbolic_Algorithm(self, dt: float, algo="newmark", beta=0.25, gamma=0.5, alpha=0.5):
        algo = getattr(algo, "name", algo)
        assert algo in HYPERBOLIC, f"algo must be in {HYPERBOLIC}"
        assert dt > 0, "Time increment must be > 0"
        if algo == "hht_newmark":
            assert 0 <= alpha <= 1 / 3, "hht_newmark requires alpha in [0, 1/3]"
            beta, gamma = 1 / 4 * (1 + alpha) ** 2, 1 / 2 + alpha
        else:
            assert 0 <= alpha < 1
        self.algo, self.dt, self.beta, self.gamma, self.alpha, self._A = algo, float(dt), float(beta), float(gamma), float(alpha), None

    # -- one step ----------------------------------------------------------------------------------------------------
    def _matrix(self):
        """A = coefK K + coefC C + coefM M on the shared pattern (cached until the scheme or the matrices change)"""
        if self.algo == "elliptic":
            return self.K
        if self._A is None:
            cK, cC, cM = time_scheme_coefs(self.algo, self.dt, self.beta, self.gamma, self.alpha)
            A = self._clone(self.K)
            lincomb([(cK, self.K.data), (cC, None if self.C is None else self.C.data), (cM, None if self.M is None else self.M.data)],
                    out=A.data)
            self._A = A
        return self._A

    def _rhs(self, F):
        n_own = self.sys.n_owned * self.d
        b = torch.zeros(n_own, dtype=torch.float64, device=self.u.device) if F is None else dv.to_device(F).reshape(-1)[:n_own].clone()
        if self.algo == "elliptic":
            return b
        w = _history_weights(self.algo, self.dt, self.beta, self.gamma, self.alpha)
        y = dv.empty((n_own,))
        for name, (xu, xv, xa) in w.items():
            Mx = getattr(self, name)
            if Mx is None or (xu == 0.0 and xv == 0.0 and xa == 0.0):
                continue
            hist = lincomb([(xu, self.u), (xv, self.v), (xa, self.a)])  # over [owned | halo]: the SpMV reads halo columns
            spmv(Mx, hist, y)
            lincomb([(1.0, b), (1.0, y)], out=b)
        return b

    def Solve(self, F=None):
        """one time step (`_Simu.Solve`): returns u_{n+1} over `[owned | halo]`; v, a are updated in place.  `F` = nodal loads
        (owned dofs), what `_Solver_Apply_Neumann` starts from."""
        d, s = self.d, self.sys
        n_own = s.n_owned * d
        A = self._matrix()
        b = self._rhs(F)
        mask, dofs, vals = self.bc.device_arrays(n_own)
        explicit = self.algo == "euler_explicit"
        x0 = self.a.clone() if explicit else self.u.clone()  # the unknown of euler_explicit is a^n, zero on constrained dofs
        _apply(x0, dofs, torch.zeros_like(vals) if explicit else vals)
        s.refresh_halo(x0, d)
        x, info = pcg(A, b, x0=x0, free_mask=mask, tol=self.pcg_tol, maxiter=self.pcg_maxiter, comm=s.comm(d), fused=self.pcg_fused,
                      persistent=self.pcg_persistent, single_reduction=self.pcg_single_reduction, precond_degree=self.pcg_precond_degree,
                      reuse_setup=True)  # the system matrix of a scheme is built once (`_matrix`) and the mask is the scheme's
        self.info = info
        x0[:n_own] = x
        s.refresh_halo(x0, d)
        self._update(x0)
        return self.u

    def _update(self, x):
        """`_Solver_Update_solutions`, _simu.py:1552-1657 (x over [owned | halo], like u, v, a)"""
        u, v, a, algo = self.u, self.v, self.a, self.algo
        if algo == "elliptic":
            self.u = x
            return
        dt = self.dt
        if algo == "parabolic":
            al = self.alpha
            self.v = lincomb([(1 / (al * dt), x), (-1 / (al * dt), u), (-(1 - al) / al, v)])
            self.u = x
            return
        be, ga = self.beta, self.gamma
        if algo in ("newmark", "hht_newmark"):
            c = 1 / (be * dt**2)
            a1 = lincomb([(c, x), (-c, u), (-c * dt, v), (-c * dt**2 / 2 * (1 - 2 * be), a)])
            self.v = lincomb([(1.0, v), (dt * (1 - ga), a), (ga * dt, a1)])
            self.u, self.a = x, a1
        elif algo == "midpoint":
            v1 = lincomb([(2 / dt, x), (-2 / dt, u), (-1.0, v)])
            self.a = lincomb([(2 / dt, v1), (-2 / dt, v), (-1.0, a)])
            self.u, self.v = x, v1
        elif algo == "hht":
            c = 1 / (be * dt)
            a1 = lincomb([(c / dt, x), (-c / dt, u), (-c, v), (1 - 1 / (2 * be), a)])
            self.v = lincomb([(dt * (1 - ga), a), (dt * ga, a1), (1.0, v)])
            self.u, self.a = x, a1
        elif algo == "euler_implicit":
            v1 = lincomb([(1 / dt, x), (-1 / dt, u)])
            self.a = lincomb([(1 / dt, v1), (-1 / dt, v)])
            self.u, self.v = x, v1
        elif algo == "euler_explicit":
            self.u = lincomb([(1.0, u), (dt, v)])
            self.v = lincomb([(1.0, v), (dt, x)])
            self.a = x
        else:
            raise NotImplementedError(f"Algo {algo} is not implemented here.")
