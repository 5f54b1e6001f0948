"""Device-resident drivers of the two consumers named by the north star: the linear-elastic solve of config 2
(`Simulations.Elastic`, EasyFEA/Simulations/_elastic.py:123-152 + `_Simu.Solve`) and the staggered phase-field loop of
configs 3/4 (`Simulations.PhaseField.Solve`, EasyFEA/Simulations/_phasefield.py:300-432).

Everything between the mesh upload and the converged nodal fields stays in HBM: element systems (S1/S3/S4), the
deterministic CSR replay (A2) and the Jacobi-PCG.  Dirichlet conditions are applied by masking (projected CG) with the
prescribed values carried in the start vector, which is what `Solvers.__Solver_1` (Solvers.py:502-553) obtains by
slicing `A[unknown][:, unknown]` and `b_u - A_uk x_k` on the host.

The same driver runs single-GPU (`world == 1`: every node is owned) and row-sharded (`easyfea_b200.dist`): vectors are
laid out `[owned | halo]`, matrices hold the owned rows, and after each solve the halo part of the new field is
refreshed so the ghost elements integrate with up-to-date nodal values.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dv
from . import operators as op
from .assembly import Assembler, DeviceCsr
from .solver import pcg


class LocalSystem:
    """Assembly of the owned rows of one element group in local numbering (single GPU: all rows)."""

    def __init__(self, group, part=None, comm_factory=None):
        self.group = group
        self.asm = Assembler()
        self.part = part
        self.n_local = int(group.Ncoords)
        self.n_owned = self.n_local if part is None else int(part.n_owned)
        self._comm_factory = comm_factory
        self._comms = {}

    def comm(self, dof_n: int):
        if self.part is None or self.part.world == 1:
            return None
        if dof_n not in self._comms:
            self._comms[dof_n] = self._comm_factory(self.part, dof_n)
        return self._comms[dof_n]

    def matrix(self, Xe, dof_n: int) -> DeviceCsr:
        d = int(dof_n)
        pat = self.asm.pattern(d, True, self.n_local * d, (self.group,))
        data = pat.replay([Xe], n_nodes=self.n_owned)
        nrows = self.n_owned * d
        if self.n_owned == self.n_local:
            return DeviceCsr(pat.indptr, pat.indices, data, (nrows, nrows), pat.node_graph)
        if not hasattr(pat, "_nnz_owned"):
            pat._nnz_owned = int(pat.indptr[nrows].item())
        nz = pat._nnz_owned
        return DeviceCsr(pat.indptr[:nrows + 1], pat.indices[:nz], data[:nz], (nrows, self.n_local * d), pat.node_graph)

    def vector(self, Fe, dof_n: int) -> torch.Tensor:
        d = int(dof_n)
        pat = self.asm.pattern(d, False, self.n_local * d, (self.group,))
        pat.replay([Fe])
        return pat.last_dense[: self.n_owned * d]

    def refresh_halo(self, x_local: torch.Tensor, dof_n: int) -> None:
        c = self.comm(dof_n)
        if c is not None:
            c.halo_exchange(x_local)


class Dirichlet:
    """Prescribed dofs of one problem in LOCAL numbering (only owned dofs matter for the solve)."""

    def __init__(self, n_dofs_local: int):
        self.n = int(n_dofs_local)
        self.dofs = np.empty(0, dtype=np.int64)
        self.values = np.empty(0)
        self._dev = None

    def add(self, nodes, values, components, dof_n: int):
        """`add_dirichlet(nodes, values, directions)` of the reference (_simu.py `add_dirichlet`): one value per component."""
        nodes = np.asarray(nodes, dtype=np.int64)
        for val, c in zip(values, components):
            self.dofs = np.concatenate([self.dofs, nodes * dof_n + int(c)])
            self.values = np.concatenate([self.values, np.full(nodes.size, float(val))])
        self._dev = None

    def unique(self):
        """(dofs, values) with one entry per dof: the LAST prescription wins, as sequential assignment does in the reference
        (`x[dofs] = values` on the host); a scatter with duplicate indices has no defined winner on the device"""
        rev_dofs, rev_vals = self.dofs[::-1], self.values[::-1]
        u, first = np.unique(rev_dofs, return_index=True)
        return u, rev_vals[first]

    def device_arrays(self, n_owned_dofs: int):
        """(free_mask uint8 (n_owned_dofs), dofs int64 tensor, values tensor), deduplicated and uploaded once per set of
        conditions"""
        if self._dev is None or self._dev[0] != n_owned_dofs:
            dofs, values = self.unique()
            mask = np.ones(n_owned_dofs, dtype=np.uint8)
            mask[dofs[dofs < n_owned_dofs]] = 0
            self._dev = (n_owned_dofs, dv.to_device(mask, np.uint8), dv.to_device(dofs, np.int64), dv.to_device(values))
        return self._dev[1:]


def _apply(x_local: torch.Tensor, dofs: torch.Tensor, values: torch.Tensor) -> None:
    if dofs.numel():
        x_local[dofs] = values  # `dofs` is duplicate-free (Dirichlet.unique)


def _check(info, what: str):
    """a solve that stops at `maxiter` must not pass silently (the reference's direct solvers cannot fail this way)"""
    if not info["converged"]:
        import warnings

        warnings.warn(f"{what}: Jacobi-PCG stopped after {info['iterations']} iterations at relative residual "
                      f"{info['rel_residual']:.3e}", RuntimeWarning, stacklevel=3)


class ElasticSolve:
    """Config 2: K = assemble(LinearizedElasticity(C)) then Jacobi-PCG on the free dofs."""

    def __init__(self, system: LocalSystem, C, thickness: float = 1.0):
        self.sys, self.C, self.thickness = system, np.asarray(C, dtype=np.float64), float(thickness)
        self.dim = int(system.group.dim)
        self.bc = Dirichlet(system.n_local * self.dim)
        self.u = torch.zeros(system.n_local * self.dim, dtype=torch.float64, device=dv.device())
        self.pcg_fused = "auto"  # True / False force the fused peer-memory form / the kernel-per-operation loop (solver.pcg)
        self.pcg_persistent = False  # True: one cooperative kernel per solve instead of three kernels per iteration
        self.pcg_single_reduction = "auto"  # True / False force / forbid the Chronopoulos-Gear form (one all-reduce per iteration)
        self.pcg_precond_degree = "auto"  # m: Chebyshev-Jacobi polynomial preconditioner of degree m - 1 (1 = plain Jacobi), solver.pcg

    def assemble(self) -> DeviceCsr:
        scale = self.thickness if self.dim == 2 else 1.0
        Ke = op.elastic_Ke_dev(self.sys.group, self.C, op.RIGI, scale)
        return self.sys.matrix(Ke, self.dim)

    def solve(self, tol=1e-8, maxiter=None, b=None):
        d = self.dim
        K = self.assemble()
        nown = self.sys.n_owned * d
        mask, dofs, vals = self.bc.device_arrays(nown)
        _apply(self.u, dofs, vals)
        self.sys.refresh_halo(self.u, d)
        rhs = torch.zeros(nown, dtype=torch.float64, device=self.u.device) if b is None else dv.to_device(b)
        x, info = pcg(K, rhs, x0=self.u, free_mask=mask, tol=tol, maxiter=maxiter, comm=self.sys.comm(d), fused=self.pcg_fused,
                      persistent=self.pcg_persistent, single_reduction=self.pcg_single_reduction, precond_degree=self.pcg_precond_degree)
        _check(info, "ElasticSolve")
        self.u[:nown] = x
        self.sys.refresh_halo(self.u, d)
        return self.u, info


class PhaseFieldStaggered:
    """`Simulations.PhaseField` reduced to the staggered loop: damage then displacement, each assembled and solved on
    the device.  `model` is an `easyfea_b200.phasefield.PhaseFieldModel`."""

    def __init__(self, system: LocalSystem, model, pcg_tol: float = 1e-10, pcg_maxiter: int = None):
        self.sys, self.pfm = system, model
        self.dim = int(system.group.dim)
        dev = dv.device()
        self.u = torch.zeros(system.n_local * self.dim, dtype=torch.float64, device=dev)
        self.d = torch.zeros(system.n_local, dtype=torch.float64, device=dev)
        self.psiP = None       # (Ne, nPg_mass) of the last damage assembly, Simulations/_phasefield.py:536
        self.psiP_old = None   # history field, replaced in Save_Iter only (:623-625)
        self.bc_u = Dirichlet(system.n_local * self.dim)
        self.bc_d = Dirichlet(system.n_local)
        self.f_ext = None      # nodal external forces of the displacement problem (owned dofs), `add_neumann`
        self.pcg_tol, self.pcg_maxiter = pcg_tol, pcg_maxiter
        self.pcg_fused, self.pcg_persistent, self.pcg_single_reduction = "auto", False, "auto"
        self.pcg_precond_degree = "auto"
        self._updatedDamage = self._updatedDisplacement = False
        self.info = {}

    def Bc_Init(self):
        self.bc_u = Dirichlet(self.sys.n_local * self.dim)
        self.bc_d = Dirichlet(self.sys.n_local)
        self.f_ext = None

    def add_neumann(self, nodes, values, components):
        """`add_neumann(nodes, values, directions)` of the reference (_simu.py:2476-2510): point loads, one value per component
        applied to every listed node (LOCAL ids; loads on halo nodes belong to their owner rank and are dropped here)"""
        nown = self.sys.n_owned * self.dim
        if self.f_ext is None:
            self.f_ext = torch.zeros(nown, dtype=torch.float64, device=self.u.device)
        nodes = np.asarray(nodes, dtype=np.int64)
        nodes = nodes[nodes < self.sys.n_owned]
        for val, c in zip(values, components):
            dofs = dv.to_device(nodes * self.dim + int(c), np.int64)
            self.f_ext.index_add_(0, dofs, torch.full((dofs.numel(),), float(val), dtype=torch.float64, device=self.u.device))

    def Calc_Psi_Ext(self) -> float:
        """`_Calc_Psi_Ext` (Simulations/_phasefield.py:819-836): u . f over the owned dofs, summed over the ranks"""
        if self.f_ext is None:
            return 0.0
        e = torch.dot(self.u[: self.f_ext.numel()], self.f_ext).reshape(1)
        c = self.sys.comm(self.dim)
        if c is not None:
            c.all_reduce_sum(e)
        return float(e.item())

    def add_dirichlet(self, nodes, values, components, problemType="elastic"):
        if problemType == "damage":
            self.bc_d.add(nodes, values, components, 1)
        else:
            self.bc_u.add(nodes, values, components, self.dim)

    # -- matrices with the reference's "assemble only when the other field changed" flags ----------------------------
    # `PhaseField.Get_K_C_M_F` (Simulations/_phasefield.py:248-269): Kd is rebuilt when the displacement changed since its
    # last assembly, Ku when the damage changed; neither `Bc_Init` nor `Save_Iter` resets the flags.
    def _get_Kd(self):
        """`__Construct_Damage_Matrix` (:540-571) -> (Kd, Fd), assembled from the current displacement and the history"""
        if not self._updatedDamage:
            s = self.sys
            Ke, Fe, self.psiP = self.pfm.damage_system_dev(s.group, self.u, self.psiP_old)
            self._Kd, self._Fd = s.matrix(Ke, 1), s.vector(Fe, 1).clone()  # the pattern reuses its dense vector buffer
            self._updatedDamage = True
        return self._Kd, self._Fd

    def _get_Ku(self):
        """`__Construct_Elastic_Matrix` (:444-482): the split uses the CURRENT displacement, g the current damage"""
        if not self._updatedDisplacement:
            Ke = self.pfm.elastic_Ke_dev(self.sys.group, self.u, self.d)
            self._Ku = self.sys.matrix(Ke, self.dim)
            self._updatedDisplacement = True
        return self._Ku

    def Need_Update(self, value=True):
        self._updatedDamage = self._updatedDisplacement = not value

    def _energy(self, A: DeviceCsr, x_local: torch.Tensor, dof_n: int) -> torch.Tensor:
        """`_Simu.Calc_Energy` (Simulations/_simu.py:177-210): 1/2 x[owned] . (A[owned] @ x), summed over the ranks (1-element
        device tensor); the dot product rides in the SpMV (fixed-order partials)."""
        from . import _lib
        from .solver import spmv

        nrows = A.indptr.numel() - 1
        partials = dv.empty((_lib.load().efb_pcg_partials_size(),))
        y = dv.empty((nrows,))
        spmv(A, x_local, y, 0, None, partials)
        e = torch.zeros(1, dtype=torch.float64, device=y.device)
        _lib.call("efb_pcg_reduce", dv.ptr(partials), 1, dv.ptr(e), dv.stream_ptr())
        c = self.sys.comm(dof_n)
        if c is not None:
            c.all_reduce_sum(e)
        return 0.5 * e

    def Calc_Psi_Crack(self) -> float:
        """`_Calc_Psi_Crack` (:802-823): 1/2 d^T Kd d with the current damage matrix"""
        return float(self._energy(self._get_Kd()[0], self.d, 1).item())

    def Calc_Psi_Elas(self) -> float:
        """`_Calc_Psi_Elas` (:779-800): 1/2 u^T Ku u"""
        return float(self._energy(self._get_Ku(), self.u, self.dim).item())

    # -- the two sub-problems ----------------------------------------------------------------------------------
    def solve_damage(self):
        """`__Solve_damage` (:573-578)"""
        s = self.sys
        K, F = self._get_Kd()
        mask, dofs, vals = self.bc_d.device_arrays(s.n_owned)
        _apply(self.d, dofs, vals)
        s.refresh_halo(self.d, 1)
        x, info = pcg(K, F, x0=self.d, free_mask=mask, tol=self.pcg_tol, maxiter=self.pcg_maxiter, comm=s.comm(1),
                      fused=self.pcg_fused, persistent=self.pcg_persistent, single_reduction=self.pcg_single_reduction, precond_degree=self.pcg_precond_degree)
        _check(info, "PhaseFieldStaggered damage solve")
        self.d[: s.n_owned] = x
        s.refresh_halo(self.d, 1)
        self._updatedDisplacement = False  # new damage -> new displacement matrices (:367-368)
        self.info["damage"] = info
        return self.d

    def solve_elastic(self):
        """`__Solve_elastic`"""
        s, dim = self.sys, self.dim
        K = self._get_Ku()
        nown = s.n_owned * dim
        mask, dofs, vals = self.bc_u.device_arrays(nown)
        _apply(self.u, dofs, vals)
        s.refresh_halo(self.u, dim)
        rhs = torch.zeros(nown, dtype=torch.float64, device=self.u.device) if self.f_ext is None else self.f_ext
        x, info = pcg(K, rhs, x0=self.u, free_mask=mask, tol=self.pcg_tol, maxiter=self.pcg_maxiter, comm=s.comm(dim),
                      fused=self.pcg_fused, persistent=self.pcg_persistent, single_reduction=self.pcg_single_reduction, precond_degree=self.pcg_precond_degree)
        _check(info, "PhaseFieldStaggered displacement solve")
        self.u[:nown] = x
        s.refresh_halo(self.u, dim)
        self._updatedDamage = False  # new displacement -> new damage matrices (:372-373)
        self.info["elastic"] = info
        return self.u

    def _reduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
        c = self.sys.comm(1)
        if c is not None:
            (c.all_reduce_max if op == "max" else c.all_reduce_sum)(t)
        return t

    def iterate(self):
        """one staggered iteration; returns max |d_new - d_old| over the owned nodes (device scalar, all-reduced)"""
        d_n = self.d[: self.sys.n_owned].clone()
        self.solve_damage()
        self.solve_elastic()
        conv = self._reduce((self.d[: self.sys.n_owned] - d_n).abs().max().reshape(1), "max")
        dmax = self._reduce(self.d[: self.sys.n_owned].max().reshape(1), "max")
        return conv, dmax

    def Solve(self, tolConv=1.0, maxIter=500, convOption=0):
        """(u, d, converged) — `Simulations.PhaseField.Solve` (:300-432).  convOption 0: max |d_np1 - d_n|; 1: relative change
        of the crack energy; 2: of the total energy crack + elastic - external work (Ambati 2015, :354-361, 381-384);
        3: summed relative increments of u and d (Pech 2022)."""
        assert 0 < tolConv <= 1, "tolConv must be between 0 and 1."
        assert maxIter > 1, "Must be > 1."
        assert convOption in (0, 1, 2, 3)
        nd, nu = self.sys.n_owned, self.sys.n_owned * self.dim
        Niter, converged, convIter = 0, False, 0.0
        old_damage = self.d[:nd].clone() if self.pfm.solver == "HistoryDamage" else None
        while not converged and Niter < maxIter:
            Niter += 1
            d_n, u_n = self.d[:nd].clone(), self.u[:nu].clone()
            if convOption == 1:
                E_n = self.Calc_Psi_Crack()
            elif convOption == 2:
                E_n = self.Calc_Psi_Crack() + self.Calc_Psi_Elas() - self.Calc_Psi_Ext()
            self.solve_damage()
            self.solve_elastic()
            d1, u1 = self.d[:nd], self.u[:nu]
            if convOption == 0:
                convIter = float(self._reduce((d1 - d_n).abs().max().reshape(1), "max").item())
            elif convOption in (1, 2):
                E1 = self.Calc_Psi_Crack()
                if convOption == 2:
                    E1 += self.Calc_Psi_Elas() - self.Calc_Psi_Ext()
                convIter = abs(E_n - E1) if E1 == 0 else abs((E_n - E1) / E1)
            else:
                def rel_sum(new, old):
                    diff = (new - old).abs()
                    nz = new != 0
                    diff = torch.where(nz, diff * (1 / new.abs().clamp_min(1e-300)), diff)
                    return float(self._reduce(diff.sum().reshape(1), "sum").item())

                convU, convD = rel_sum(u1, u_n), rel_sum(d1, d_n)
                convIter = max(convD, convU)
            dmax = float(self._reduce(d1.max().reshape(1), "max").item())
            if tolConv == 1 or dmax == 0:
                converged = True
            elif convOption == 3:
                converged = convD <= tolConv and convU <= tolConv * 0.999
            else:
                converged = convIter <= tolConv
        if old_damage is not None:
            # HistoryDamage: irreversibility enforced on the nodal field at the end of the solve (:400-404)
            self.d[:nd] = torch.maximum(old_damage, self.d[:nd])
            self.sys.refresh_halo(self.d, 1)
            self._updatedDisplacement = False
        self.Niter, self.convIter = Niter, convIter
        return self.u, self.d, converged

    def Save_Iter(self):
        """history update of the `History` solver, Simulations/_phasefield.py:623-625"""
        if self.pfm.solver == "History":
            self.psiP_old = self.psiP
