"""GPU parity of the phase-field law + simulation-level builders, and of the Jacobi-PCG consumer.  -m gpu"""
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.helpers import make_mesh, rel_err
from tests.test_hostcheck_pf_csr import generic_states

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12


@pytest.fixture(scope="module")
def efb():
    from easyfea_b200 import _lib, assembly, mesh, operators, phasefield, solver

    _lib.require_cuda()

    class NS:
        pass

    ns = NS()
    ns.asm, ns.mesh, ns.op, ns.pf, ns.solver = assembly, mesh, operators, phasefield, solver
    return ns


@pytest.mark.parametrize("dim,planeStress", [(2, False), (2, True), (3, False)])
@pytest.mark.parametrize("split", ["Bourdin", "Amor", "Miehe", "Stress", "He"])
def test_split_generic_states(efb, dim, planeStress, split):
    om = orc.IsoMaterial(dim, 210000.0, 0.3, planeStress)
    pfm = efb.pf.PhaseFieldModel(efb.pf.IsotropicMaterial(dim, 210000.0, 0.3, planeStress), split, "AT2", 2.7, 0.01)
    eps = generic_states(dim, 2000, 4, seed=21)
    cP, cM = pfm.Calc_C(eps, verif=True)
    psiP, psiM = pfm.Calc_psi_e_pg(eps)
    ocP, ocM = orc.calc_C(om, split, eps)
    opP, opM = orc.calc_psi(om, split, eps)
    assert rel_err(cP, ocP) < TOL and rel_err(cM, ocM) < TOL
    assert rel_err(psiP, opP) < TOL and rel_err(psiM, opM) < TOL
    psi = 0.5 * np.einsum("epi,ij,epj->ep", eps, om.C, eps)
    assert rel_err(psiP + psiM, psi) < TOL  # tests/Models/phasefield_test.py:133-137


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("split", ["Amor", "Miehe", "Stress", "He"])
def test_split_degenerate_states(efb, dim, split):
    om = orc.IsoMaterial(dim, 210000.0, 0.3)
    pfm = efb.pf.PhaseFieldModel(efb.pf.IsotropicMaterial(dim, 210000.0, 0.3, False), split, "AT2", 2.7, 0.01)
    ns = 3 if dim == 2 else 6
    rng = np.random.default_rng(2)
    eps = rng.normal(size=(12, 3, ns)) * 1e-3
    eps[0] = 0.0
    eps[1] = 0.0; eps[1, :, 0] = 1e-3
    eps[2] = 0.0; eps[2, :, :dim] = 1e-3
    eps[3, 1] = 0.0
    eps[4, 0] = 0.0; eps[4, 0, :2] = 2e-3
    eps[5] = 0.0; eps[5, :, 0] = 1e-3; eps[5, :, 1:dim] = -0.3e-3
    cP, cM = pfm.Calc_C(eps)
    assert np.isfinite(cP).all() and np.isfinite(cM).all()
    assert rel_err(cP + cM, np.broadcast_to(om.C, cP.shape)) < TOL
    ocP, _ = orc.calc_C(om, split, eps, clamp=True)
    assert rel_err(cP[6:], ocP[6:]) < TOL


@pytest.mark.parametrize("name", ["QUAD9", "HEXA8", "TRI3", "TETRA4", "HEXA27"])
@pytest.mark.parametrize("split", ["Amor", "Miehe", "Stress", "He"])
def test_golden_splits_on_solved_like_fields(efb, name, split):
    d = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    g = efb.mesh.ElemGroup(name, d["connect"], d["coords"])
    pfm = efb.pf.PhaseFieldModel(efb.pf.IsotropicMaterial(g.dim, 210000.0, 0.3, False), split, "AT2", 2.7, 0.01)
    for mt in ("rigi", "mass"):
        cP, cM = pfm.Calc_C(d[f"eps_{mt}"])
        psiP, psiM = pfm.Calc_psi_e_pg(d[f"eps_{mt}"])
        assert rel_err(cP, d[f"cP_{split}_{mt}"]) < TOL and rel_err(cM, d[f"cM_{split}_{mt}"]) < TOL
        assert rel_err(psiP, d[f"psiP_{split}_{mt}"]) < TOL and rel_err(psiM, d[f"psiM_{split}_{mt}"]) < TOL
    assert rel_err(pfm.Get_g_e_pg(d["dmg"], g, "rigi"), d["g_rigi"]) < TOL


@pytest.mark.parametrize("elemType,split,regu", [("TRI3", "Miehe", "AT2"), ("TETRA4", "He", "AT2"), ("QUAD9", "Amor", "AT1"),
                                                 ("HEXA8", "Stress", "AT1")])
def test_simulation_level_builders(efb, elemType, split, regu):
    """S3 / S4 of SURVEY §8a against the oracle compositions (configs 3 and 4 element types + two more)."""
    from easyfea_b200 import elements as el

    rng = np.random.default_rng(8)
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    dim = g.dim
    Nn = coords.shape[0]
    thickness = 0.5 if dim == 2 else 1.0
    om = orc.IsoMaterial(dim, 210e9, 0.3, False)
    pfm = efb.pf.PhaseFieldModel(efb.pf.IsotropicMaterial(dim, 210e9, 0.3, False, thickness), split, regu, 2.7e3, 1e-2)
    u = rng.normal(size=Nn * dim) * 1e-5
    dmg = rng.uniform(0, 0.9, Nn)
    u_e = orc.locate_sol_e(u, connect, dim)
    tr, tm = el.gauss_table(elemType, "rigi"), el.gauss_table(elemType, "mass")
    geo_r = orc.geometry(coords[connect][:, :, :dim], tr.dN_pg, tr.weights)
    geo_m = orc.geometry(coords[connect][:, :, :dim], tm.dN_pg, tm.weights)
    Ke = pfm.elastic_Ke_dev(g, u, dmg).cpu().numpy()  # one-pass kernel for TRI3 / TETRA4, composition otherwise
    Ke4 = pfm.elastic_Ke_dev(g, u, dmg, fused=False).cpu().numpy()  # strain -> degradation -> split -> stiffness
    assert rel_err(Ke, Ke4) < TOL
    ref = thickness * orc.pf_elastic_Ke(geo_r, tr.N_pg, om, split, u_e, dmg[connect], clamp=True)
    assert rel_err(Ke, ref) < TOL  # measured 1.5e-16 .. 1.9e-13 (Stress split in 3D), profiles/r2_dist2_and_pf_errors.log
    old = rng.uniform(0, 1, (g.Ne, tm.nPg)) * float(np.median(orc.calc_psi(om, split, orc.strain(geo_m, u_e), True)[0]))
    Kd, Fd, psiP = pfm.damage_system_dev(g, u, old)
    rK, rF, rpsi = orc.pf_damage_system(geo_m, tm.N_pg, om, split, regu, 2.7e3, 1e-2, u_e, old, clamp=True)
    assert rel_err(psiP.cpu().numpy(), rpsi) < TOL
    assert rel_err(Kd.cpu().numpy(), thickness * rK) < TOL
    assert rel_err(Fd.cpu().numpy(), thickness * rF[..., 0]) < TOL


def test_pcg_elastic_cube(efb):
    """HEXA8 cube clamped at x=0, u_x=0.1 at x=1 (config 2 BCs): ||Ax-b||/||b|| <= 1e-8 and x matches a direct solve."""
    import scipy.sparse.linalg as spla
    import torch

    from easyfea_b200 import meshgen

    n = 12
    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=0.15, seed=0)
    g = efb.mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    Nn = coords.shape[0]
    Ndof = 3 * Nn
    om = orc.IsoMaterial(3, 210000.0, 0.3)
    A = efb.asm.Assembler()
    K = A.Assemble_csr({g: efb.op.elastic_Ke_dev(g, om.C)}, 3, Ndof, True, as_device=True)
    x0 = np.zeros(Ndof)
    known = np.zeros(Ndof, bool)
    left = np.flatnonzero(coords[:, 0] < 1e-9 + coords[:, 0].min())
    right = np.flatnonzero(coords[:, 0] > coords[:, 0].max() - 1e-9)
    # structured ids: clamp the x=0 lattice plane, pull the x=1 plane
    lat = np.arange(Nn) % (n + 1)
    left, right = np.flatnonzero(lat == 0), np.flatnonzero(lat == n)
    for c in range(3):
        known[left * 3 + c] = True
    known[right * 3] = True
    x0[right * 3] = 0.1
    b = np.zeros(Ndof)
    # default: Chebyshev-Jacobi polynomial preconditioner inside the fused iterations (efb_pcg_iterate_cheb)
    xc, infoc = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000)
    assert infoc["converged"] and infoc["fused"] and infoc["precond_degree"] == efb.solver.CHEB_DEGREE, infoc
    # plain Jacobi (degree 1): the reference of the variants below
    x, info = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, precond_degree=1)
    assert info["converged"] and info["precond_degree"] == 1, info
    assert infoc["iterations"] * 2 < info["iterations"], (infoc, info)  # ~3.6x fewer iterations at degree 4
    x = x.cpu().numpy()
    assert np.linalg.norm(xc.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    # the kernel-per-operation loop with the same polynomial: same iterates up to the summation order
    # (FP64 matrix values inside the polynomial there, single precision in the fused kernels: the same iteration counts)
    xcu, infocu = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, fused=False, precond_degree=efb.solver.CHEB_DEGREE)
    assert infocu["converged"] and not infocu["fused"] and abs(infocu["iterations"] - infoc["iterations"]) <= 25, (infoc, infocu)
    assert infoc["precond_fp32"] and not infocu["precond_fp32"]
    assert np.linalg.norm(xcu.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    xc2, infoc2 = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000)
    assert infoc2["iterations"] == infoc["iterations"] and np.array_equal(xc2.cpu().numpy(), xc.cpu().numpy())  # reproducible
    Ks = K.to_scipy()
    free = ~known
    rhs = (b - Ks @ (x0 * known))[free]
    res = np.linalg.norm(Ks[free][:, free] @ x[free] - rhs) / np.linalg.norm(rhs)
    assert res <= 1e-8
    xd = spla.spsolve(Ks[free][:, free].tocsc(), rhs)
    assert np.linalg.norm(x[free] - xd) / np.linalg.norm(xd) < 1e-6
    assert np.array_equal(x[known], x0[known])
    # the fused iterations (3 kernels each, reductions folded by the last CTA) follow the kernel-per-operation loop up to
    # the summation order of the dot products (one-wave grids), and are reproducible run to run
    assert info["fused"]
    x2, info2 = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, fused=False, precond_degree=1)
    assert not info2["fused"] and abs(info2["iterations"] - info["iterations"]) <= 25
    assert np.linalg.norm(x2.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    x2b, info2b = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, precond_degree=1)
    assert info2b["iterations"] == info["iterations"] and np.array_equal(x2b.cpu().numpy(), x)
    # opt-in: ONE persistent cooperative kernel per solve; same iterates as the three-kernel form up to summation order
    assert not info["persistent"]
    x2c, info2c = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, persistent=True)
    assert info2c["fused"] and info2c["persistent"] and abs(info2c["iterations"] - info["iterations"]) <= 25
    assert np.linalg.norm(x2c.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    Kc = efb.asm.DeviceCsr(K.indptr, K.indices, K.data, K.shape)  # generic CSR form (no node graph)
    # single-reduction (Chronopoulos-Gear) form: same Krylov iterates up to rounding, one all-reduce per iteration
    x2e, info2e = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, single_reduction=True)
    assert info2e["converged"] and info2e["single_reduction"] and abs(info2e["iterations"] - info["iterations"]) <= 25, (info, info2e)
    assert np.linalg.norm(x2e.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    res_e = np.linalg.norm(Ks[free][:, free] @ x2e.cpu().numpy()[free] - rhs) / np.linalg.norm(rhs)
    assert res_e <= 2e-8, res_e
    x2f, info2f = efb.solver.pcg(Kc, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, single_reduction=True, check_every=7)
    assert info2f["converged"] and np.linalg.norm(x2f.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5
    # maxiter is honoured exactly by the persistent kernel
    _, info2d = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-30, maxiter=17, persistent=True)
    assert info2d["iterations"] == 17 and not info2d["converged"]
    # generic CSR form of the fused SpMV (no node graph) and a system without a mask
    from easyfea_b200.assembly import DeviceCsr

    Kc = DeviceCsr(K.indptr, K.indices, K.data, K.shape)
    x3, info3 = efb.solver.pcg(Kc, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, check_every=7)
    assert info3["converged"] and np.linalg.norm(x3.cpu().numpy() - x) / np.linalg.norm(x) < 1e-5  # both stop at 1e-8 residual
    import scipy.sparse as sp

    Kreg = (Ks + 1e3 * sp.identity(Ndof, format="csr")).tocsr()
    Kreg.sort_indices()
    Kr = DeviceCsr(dv_t(Kreg.indptr), dv_t(Kreg.indices), dv_t(Kreg.data), Kreg.shape)
    rhs_full = np.random.default_rng(3).standard_normal(Ndof)
    x4, info4 = efb.solver.pcg(Kr, rhs_full, tol=1e-10, maxiter=5000)
    assert info4["converged"]
    assert np.linalg.norm(Kreg @ x4.cpu().numpy() - rhs_full) / np.linalg.norm(rhs_full) <= 1e-9


def test_pcg_polynomial_stall_falls_back_to_jacobi(efb, monkeypatch):
    """a spectral bound far below the largest eigenvalue of D^-1 A makes the Chebyshev polynomial indefinite: the solve must not
    hang or diverge — after CHEB_STALL_ITERS iterations without a new minimum of |r| it finishes with plain Jacobi"""
    import warnings

    from easyfea_b200 import meshgen

    n = 8
    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=0.15, seed=2)
    g = efb.mesh.ElemGroup("HEXA8", connect, coords)
    Nn = coords.shape[0]
    Ndof = 3 * Nn
    mat = orc.IsoMaterial(3, 210000.0, 0.3)
    K = efb.asm.Assembler().Assemble_csr({g: efb.op.elastic_Ke_dev(g, mat.C)}, 3, Ndof, True, as_device=True)
    lat = np.arange(Nn) % (n + 1)
    known = np.zeros(Ndof, bool)
    for c in range(3):
        known[np.flatnonzero(lat == 0) * 3 + c] = True
    known[np.flatnonzero(lat == n) * 3] = True
    x0 = np.zeros(Ndof)
    x0[np.flatnonzero(lat == n) * 3] = 0.1
    b = np.zeros(Ndof)
    xr, infor = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, precond_degree=1)
    monkeypatch.setattr(efb.solver, "CHEB_SAFETY", 0.3)      # lmax used = 0.3 x the estimate: the polynomial changes sign
    monkeypatch.setattr(efb.solver, "CHEB_STALL_ITERS", 50)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        x, info = efb.solver.pcg(K, b, x0=x0, free_mask=~known, tol=1e-8, maxiter=5000, precond_degree=4)
    assert info["converged"], info
    if info.get("fell_back_to_jacobi"):
        assert any("stalled" in str(m.message) for m in w)
    assert np.linalg.norm(x.cpu().numpy() - xr.cpu().numpy()) <= 1e-5 * np.linalg.norm(xr.cpu().numpy())


def dv_t(a):
    from easyfea_b200 import device as dv

    return dv.to_device(a)


@pytest.mark.parametrize("elemType,dof_n", [("HEXA8", 3), ("TETRA4", 3), ("TRI3", 2), ("TRI3", 1), ("QUAD9", 2), ("HEXA27", 1)])
def test_spmv_nodeblock_matches_scipy(efb, elemType, dof_n):
    """The solver's node-block SpMV (column structure read from the node adjacency) against scipy's CSR product, with a
    row mask and the fused x.y partials; also against the library's plain CSR SpMV on the same matrix."""
    import torch

    from easyfea_b200 import _lib
    from easyfea_b200 import device as dv
    from easyfea_b200.assembly import DeviceCsr

    rng = np.random.default_rng(11)
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords, all_nodes_used=True)
    Nn = coords.shape[0]
    ndof = connect.shape[1] * dof_n
    Xe = rng.standard_normal((connect.shape[0], ndof, ndof))
    A = efb.asm.Assembler().Assemble_csr({g: Xe}, dof_n, Nn * dof_n, True, as_device=True)
    assert A.node_graph is not None
    x = rng.standard_normal(Nn * dof_n)
    mask = (rng.uniform(size=Nn * dof_n) > 0.2).astype(np.uint8)
    ref = (A.to_scipy() @ x) * mask
    xd, md = dv.to_device(x), dv.to_device(mask)
    partials = dv.empty((_lib.load().efb_pcg_partials_size(),))
    y = efb.solver.spmv(A, xd, mask=md, partials=partials)
    assert rel_err(y.cpu().numpy(), ref) < 1e-13
    nblk = partials.numel() // 3  # efb_pcg_partials_size() = 3 x the number of reduction CTAs
    assert abs(float(partials[:nblk].sum()) - float(x @ ref)) <= 1e-10 * np.abs(x * ref).sum()
    y_csr = efb.solver.spmv(DeviceCsr(A.indptr, A.indices, A.data, A.shape), xd, mask=md)
    assert rel_err(y.cpu().numpy(), y_csr.cpu().numpy()) < 1e-13
    assert np.array_equal(y.cpu().numpy()[mask == 0], np.zeros(int((mask == 0).sum())))
