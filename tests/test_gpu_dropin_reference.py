"""The drop-in boundary against the LIVE reference, on the GPU.  -m gpu

The unmodified reference (offline install `baseline/_ref`, which travels with the repository snapshot; gmsh stubbed) runs
each simulation twice on the same hand-built mesh: with its own NumPy/SciPy path, and after `easyfea_b200.dropin.install()`
rebinds its entry points to the device.  Compared: CSR structure (`np.array_equal` + dtype), assembled values (bit-identical
when only the assembly is rebound, <= 1e-12 with the device operators), solutions.

Tolerances on solutions: the element matrices differ from the reference's einsum results in the last bits (<= 1e-12
normwise), which a direct solve amplifies by the condition number of the system, so fields are compared at 1e-9 when the
reference's own solver runs on the device-assembled matrix, and at residual 1e-8 / field 1e-6 when the device Jacobi-PCG
(tol 1e-12) replaces it (north star: solutions to a stated 1e-8 relative residual)."""
import numpy as np
import pytest

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_import import import_reference, reference_available

    if not reference_available(travel_only=True):
        pytest.skip("no reference install under baseline/_ref (python -m pip install --no-deps --target baseline/_ref /root/reference)")
    from easyfea_b200 import _lib

    _lib.require_cuda()
    return import_reference(travel_only=True)


@pytest.fixture()
def dropin():
    from easyfea_b200 import dropin as d

    if d.installed():
        d.uninstall()
    yield d
    if d.installed():
        d.uninstall()
    d.config.update(min_dofs=100_000, pcg_tol=1e-10)


# ---------------------------------------------------------------------------------------------------------
# config 1: README cantilever (QUAD9 28 x 3 + SEG3 boundary group, add_surfLoad)
# ---------------------------------------------------------------------------------------------------------
def _cantilever(ref):
    from tests import ref_helpers as rh

    simu, mesh, nodesX0, nodesXL = rh.readme_cantilever(ref)
    u = np.array(simu.Solve())
    K, _, _, F = simu.Get_K_C_M_F()
    return simu, mesh, u, K, F, nodesXL


def test_config1_cantilever_assembly_only_is_bit_identical(ref, dropin):
    _, _, u0, K0, F0, _ = _cantilever(ref)
    patched = dropin.install(ref, levels=(2,))
    assert "_Simu._Simu__Assemble_csr" in patched
    _, _, u1, K1, F1, _ = _cantilever(ref)
    assert K1.indptr.dtype == K0.indptr.dtype and K1.indices.dtype == K0.indices.dtype
    assert np.array_equal(K1.indptr, K0.indptr) and np.array_equal(K1.indices, K0.indices)
    assert np.array_equal(K1.data, K0.data), "deterministic replay must reproduce np.bincount bit for bit"
    assert np.array_equal(F1.toarray(), F0.toarray())
    assert np.array_equal(u1, u0)


def test_config1_cantilever_all_levels(ref, dropin):
    simu0, mesh, u0, K0, F0, nodesXL = _cantilever(ref)
    assert abs(u0.reshape(-1, 2)[nodesXL, 1].mean() + 0.928) < 5e-3  # beam theory -0.92194, SURVEY.md section 8c probe -0.92820
    patched = dropin.install(ref)  # levels 0-4; 798 dofs < min_dofs -> the reference's own spsolve on the device-built K
    assert len(patched) >= 20
    simu1, _, u1, K1, F1, _ = _cantilever(ref)
    assert np.array_equal(K1.indptr, K0.indptr) and np.array_equal(K1.indices, K0.indices)
    assert rel_err(K1.data, K0.data) < 1e-12
    assert rel_err(F1.toarray(), F0.toarray()) < 1e-12
    assert rel_err(u1, u0) < 1e-9
    assert dropin.stats["device_solves"] == 0
    # the reference's own post-processing runs on top of the patched getters (level 0: Get_B_e_pg, Get_leftDispPart_e_pg, ...)
    assert rel_err(simu1.Result("Svm", nodeValues=False), simu0.Result("Svm", nodeValues=False)) < 1e-8
    assert abs(simu1.Result("Wdef") - simu0.Result("Wdef")) < 1e-9 * abs(simu0.Result("Wdef"))


def test_config1_cantilever_device_solver(ref, dropin):
    """level 4: `Solve_simu` goes to the masked device Jacobi-PCG (threshold lowered so that 798 dofs qualify)"""
    _, _, u0, K0, F0, _ = _cantilever(ref)
    dropin.install(ref, min_dofs=1, pcg_tol=1e-12)
    n0 = dropin.stats["device_solves"]
    simu1, _, u1, K1, F1, _ = _cantilever(ref)
    assert dropin.stats["device_solves"] == n0 + 1 and dropin.stats["last_info"]["converged"]
    known, unknown = simu1.Bc_dofs_known_unknown(simu1.problemType)
    b = np.asarray(simu1._Solver_Apply_Neumann(simu1.problemType).toarray()).ravel()
    r = (K0 @ u1 - b)[unknown]
    assert np.linalg.norm(r) / np.linalg.norm(b[unknown]) < 1e-8  # against the REFERENCE's matrix
    assert rel_err(u1, u0) < 1e-6


def test_level4_solve_axb_threshold(ref, dropin):
    """`_Solve_Axb` (Solvers.py:225): large reduced systems on the device, small ones handed to the reference's solver"""
    from EasyFEA.Simulations import Solvers

    simu, _, u0, K0, F0, _ = _cantilever(ref)
    known, unknown = simu.Bc_dofs_known_unknown(simu.problemType)
    Aii = K0[unknown, :][:, unknown].tocsr()
    from scipy import sparse

    bi = np.asarray(simu._Solver_Apply_Neumann(simu.problemType).toarray()).ravel()[unknown]  # the surface load
    assert np.linalg.norm(bi) > 0
    bcol = sparse.csr_matrix(bi.reshape(-1, 1))  # `__Solver_1` hands b over as a (n, 1) sparse column
    x_ref = Solvers._Solve_Axb(simu, simu.problemType, Aii, bcol, np.zeros(unknown.size), [], [])
    dropin.install(ref, levels=(4,), min_dofs=10**9)
    h0 = dropin.stats["host_solves"]
    x_small = Solvers._Solve_Axb(simu, simu.problemType, Aii, bcol, np.zeros(unknown.size), [], [])
    assert dropin.stats["host_solves"] == h0 + 1 and np.array_equal(x_small, x_ref)
    dropin.config["min_dofs"], dropin.config["pcg_tol"] = 1, 1e-12
    d0 = dropin.stats["device_solves"]
    x_dev = Solvers._Solve_Axb(simu, simu.problemType, Aii, bcol, np.zeros(unknown.size), [], [])
    assert dropin.stats["device_solves"] == d0 + 1
    assert np.linalg.norm(Aii @ x_dev - bi) / np.linalg.norm(bi) < 1e-8 and rel_err(x_dev, x_ref) < 1e-6


def test_level4_rejects_nonsymmetric(ref, dropin):
    from scipy import sparse

    from easyfea_b200._lib import EfbError

    A = sparse.csr_matrix(np.array([[4.0, 1.0, 0.0], [0.5, 3.0, 0.2], [0.0, 0.1, 2.0]]))
    with pytest.raises(EfbError):
        dropin.device_solve(A, np.ones(3), np.zeros(3))


# ---------------------------------------------------------------------------------------------------------
# configs 3 / 4 (small): one staggered load step of Simulations.PhaseField
# ---------------------------------------------------------------------------------------------------------
def _pf_step(ref, elemType, n, split, dep):
    from tests import ref_helpers as rh

    simu, sets, dim = rh.phasefield_case(ref, elemType, n, split)
    rh.apply_shear(simu, sets, dim, dep)
    u, d, conv = simu.Solve(1e-3, 50, convOption=0)
    Ku = simu.Get_K_C_M_F(simu.ProblemTypes.elastic)[0]
    Kd, _, _, Fd = simu.Get_K_C_M_F(simu.ProblemTypes.damage)
    return {"u": np.array(u), "d": np.array(d), "Niter": simu._PhaseField__Niter, "Ku": Ku, "Kd": Kd, "Fd": Fd,
            "psiP": np.asarray(simu._PhaseField__psiP_e_pg), "sets": sets}


@pytest.mark.parametrize("elemType,n,split", [("TRI3", (16, 16), "Miehe"), ("TETRA4", (6, 6, 2), "He")])
def test_phasefield_load_step(ref, dropin, elemType, n, split):
    dep = 8e-6
    r0 = _pf_step(ref, elemType, n, split, dep)
    assert r0["Niter"] > 1
    dropin.install(ref)  # reference's own solver on device-built systems (small mesh)
    r1 = _pf_step(ref, elemType, n, split, dep)
    for key in ("Ku", "Kd"):
        assert np.array_equal(r1[key].indptr, r0[key].indptr) and np.array_equal(r1[key].indices, r0[key].indices)
    assert r1["Niter"] == r0["Niter"]
    assert rel_err(r1["d"], r0["d"]) < 1e-8 and rel_err(r1["u"], r0["u"]) < 1e-8
    assert rel_err(r1["psiP"], r0["psiP"]) < 1e-8  # psi+ of the converged state (history applied)
    assert rel_err(r1["Kd"].data, r0["Kd"].data) < 1e-8 and rel_err(r1["Ku"].data, r0["Ku"].data) < 1e-8
    dropin.uninstall()
    dropin.install(ref, min_dofs=1, pcg_tol=1e-12)  # both sub-problems through the device Jacobi-PCG
    d0 = dropin.stats["device_solves"]
    r2 = _pf_step(ref, elemType, n, split, dep)
    assert dropin.stats["device_solves"] >= d0 + 2 * r2["Niter"]
    assert r2["Niter"] == r0["Niter"]
    assert rel_err(r2["d"], r0["d"]) < 1e-6 and rel_err(r2["u"], r0["u"]) < 1e-6


def test_phasefield_law_on_reference_objects(ref, dropin):
    """level 3 on a live `Models.PhaseField`: Calc_C / Calc_psi_e_pg / Calc_Sigma_e_pg / Get_g_e_pg, FeArray results"""
    from EasyFEA import Models
    from tests import ref_helpers as rh

    rng = np.random.default_rng(11)
    for dim, et, n in ((2, "TRI3", (5, 4)), (3, "TETRA4", (3, 2, 2))):
        simu, sets, _ = rh.phasefield_case(ref, et, n, "Miehe")
        g = simu.mesh.Get_list_groupElem()[0]
        ns = 3 if dim == 2 else 6
        eps = ref.FEM.FeArray.asfearray(rng.normal(size=(g.Ne, 3, ns)) * 1e-3)  # the law takes FeArrays (e, p, ns)
        d_n = rng.uniform(0, 1, simu.mesh.Nn)
        for split in ("Amor", "Miehe", "Stress", "He"):
            pfm = Models.PhaseField(simu.phaseFieldModel.material, split, "AT2", 2.7e3, 1e-4)
            want = (pfm.Calc_C(eps), pfm.Calc_psi_e_pg(eps), pfm.Calc_Sigma_e_pg(eps), pfm.Get_g_e_pg(d_n, g, "mass"))
            dropin.install(ref, levels=(3,))
            got = (pfm.Calc_C(eps), pfm.Calc_psi_e_pg(eps), pfm.Calc_Sigma_e_pg(eps), pfm.Get_g_e_pg(d_n, g, "mass"))
            dropin.uninstall()
            assert type(got[0][0]).__name__ == "FeArray"
            for a, b in zip(got[:3], want[:3]):
                assert rel_err(a[0], b[0]) < 1e-12 and rel_err(a[1], b[1]) < 1e-12, (dim, split)
            assert rel_err(got[3], want[3]) < 1e-12


# ---------------------------------------------------------------------------------------------------------
# config 5 (small): one parabolic thermal step and one Newmark step on HEXA27
# ---------------------------------------------------------------------------------------------------------
def _thermal(ref, steps=2):
    from EasyFEA import Models, Simulations
    from easyfea_b200 import meshgen
    from tests import ref_helpers as rh

    lattice, connect = meshgen.structured_mesh("HEXA27", (3, 3, 2))
    coords, _ = meshgen.structured_mesh("HEXA27", (3, 3, 2), jitter=0.12, seed=7)
    mesh = rh.ref_mesh(ref, "HEXA27", coords, connect)
    simu = Simulations.Thermal(mesh, Models.Thermal(k=1.5, c=2.0))
    simu.rho = 1.3
    simu.Solver_Set_Parabolic_Algorithm(dt=0.1, alpha=0.5)
    x = lattice[:, 0]
    lo, hi = np.flatnonzero(x < 1e-12), np.flatnonzero(x > x.max() - 1e-12)
    out = []
    for s in range(steps):
        simu.Bc_Init()
        simu.add_dirichlet(lo, [0.0], ["t"])
        simu.add_dirichlet(hi, [40.0 + 10 * s], ["t"])
        simu.Solve()
        simu.Save_Iter()
        out.append((np.array(simu.thermal), np.array(simu.thermalDot)))
    K, C, _, _ = simu.Get_K_C_M_F()
    return out, K, C


def test_thermal_steps(ref, dropin):
    out0, K0, C0 = _thermal(ref)
    dropin.install(ref)
    out1, K1, C1 = _thermal(ref)
    assert np.array_equal(K1.indices, K0.indices) and np.array_equal(C1.indptr, C0.indptr)
    assert rel_err(K1.data, K0.data) < 1e-12 and rel_err(C1.data, C0.data) < 1e-12
    for (t1, v1), (t0, v0) in zip(out1, out0):
        assert rel_err(t1, t0) < 1e-9 and rel_err(v1, v0) < 1e-8
    dropin.uninstall()
    dropin.install(ref, min_dofs=1, pcg_tol=1e-12)
    out2, _, _ = _thermal(ref)
    for (t2, v2), (t0, v0) in zip(out2, out0):
        assert rel_err(t2, t0) < 1e-6


def test_elastodynamic_newmark_step(ref, dropin):
    from EasyFEA import AlgoType, Models, Simulations
    from easyfea_b200 import meshgen
    from tests import ref_helpers as rh

    def run():
        lattice, connect = meshgen.structured_mesh("HEXA8", (4, 3, 3))
        coords, _ = meshgen.structured_mesh("HEXA8", (4, 3, 3), jitter=0.12, seed=7)
        mesh = rh.ref_mesh(ref, "HEXA8", coords, connect)
        simu = Simulations.Elastic(mesh, Models.Elastic.Isotropic(3, E=210000.0, v=0.3))
        simu.rho = 2.0
        simu.Set_Rayleigh_Damping_Coefs(coefM=0.15, coefK=2e-4)
        simu.Solver_Set_Hyperbolic_Algorithm(dt=0.05, algo=AlgoType.newmark)
        x = lattice[:, 0]
        lo, hi = np.flatnonzero(x < 1e-12), np.flatnonzero(x > x.max() - 1e-12)
        for s in range(2):
            simu.Bc_Init()
            simu.add_dirichlet(lo, [0.0, 0.0, 0.0], ["x", "y", "z"])
            simu.add_dirichlet(hi, [0.01 * (s + 1)], ["x"])
            simu.Solve()
            simu.Save_Iter()
        K, C, M, _ = simu.Get_K_C_M_F()
        return np.array(simu.displacement), np.array(simu.speed), np.array(simu.accel), K, C, M

    u0, v0, a0, K0, C0, M0 = run()
    dropin.install(ref)
    u1, v1, a1, K1, C1, M1 = run()
    for A1, A0 in ((K1, K0), (C1, C0), (M1, M0)):
        assert np.array_equal(A1.indptr, A0.indptr) and np.array_equal(A1.indices, A0.indices)
        assert rel_err(A1.data, A0.data) < 1e-12
    assert rel_err(u1, u0) < 1e-9 and rel_err(v1, v0) < 1e-8 and rel_err(a1, a0) < 1e-7


def test_level1_falls_back_for_groups_outside_the_path(ref, dropin):
    """1D groups (the SEG3 boundary of config 1) are outside the device path: the patched operators and getters hand them to
    the reference's own code instead of raising"""
    from tests import ref_helpers as rh

    simu, mesh, _, _ = rh.readme_cantilever(ref)
    seg = mesh.Get_list_groupElem(1)[0]
    Ops = ref.FEM.Operators
    want = Ops.Bilinear.UV(seg, 2.0, 1)
    want_jac = np.array(seg.Get_jacobian_e_pg("mass"))
    dropin.install(ref)
    seg._InitMatrix()
    assert np.array_equal(Ops.Bilinear.UV(seg, 2.0, 1), want)
    assert np.array_equal(np.array(seg.Get_jacobian_e_pg("mass")), want_jac)


def test_hyperelastic_operator_on_reference_objects(ref, dropin):
    """section 8f rank 3: `Operators.NonLinear.SecondPiolaKirchhoffStressTensor` on live `HyperElasticState` / material objects"""
    from EasyFEA.FEM import MatrixType
    from EasyFEA.Models.HyperElastic import NeoHookean, SaintVenantKirchhoff
    from EasyFEA.Models.HyperElastic._state import HyperElasticState
    from easyfea_b200 import meshgen
    from tests import ref_helpers as rh

    rng = np.random.default_rng(4)
    for et, n, dim in (("TETRA4", (3, 3, 2), 3), ("HEXA8", (3, 2, 2), 3), ("QUAD4", (5, 4), 2)):
        coords, connect = meshgen.structured_mesh(et, n, jitter=0.15, seed=2)
        mesh = rh.ref_mesh(ref, et, coords, connect)
        g = mesh.Get_list_groupElem()[0]
        u = rng.normal(size=mesh.Nn * dim) * 0.03
        for mat in (SaintVenantKirchhoff(dim, lmbda=121.0, mu=81.0, thickness=0.7), NeoHookean(dim, K=3.0, thickness=0.7)):
            Op = ref.FEM.Operators.NonLinear
            K0, R0 = Op.SecondPiolaKirchhoffStressTensor(mat, HyperElasticState(g, u, MatrixType.rigi))
            dropin.install(ref, levels=(1,))
            K1, R1 = Op.SecondPiolaKirchhoffStressTensor(mat, HyperElasticState(g, u, MatrixType.rigi))
            dropin.uninstall()
            assert K1.shape == K0.shape and rel_err(K1, K0) < 1e-12 and rel_err(R1, R0) < 1e-12
