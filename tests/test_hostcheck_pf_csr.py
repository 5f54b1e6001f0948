"""CPU parity of the phase-field law (csrc/pf_math.cuh) and of the CSR pattern/replay bodies (csrc/csr_kernels.cuh)
through the test-only host emulation, against the NumPy oracle."""
import ctypes

import numpy as np
import pytest

from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc
from tests.helpers import host_group, make_mesh, p, rel_err

I = ctypes.c_int
SPLITS = {"Bourdin": 0, "Amor": 1, "Miehe": 2, "Stress": 3, "He": 4}


class CPfMat(ctypes.Structure):
    """mirror of `efb_pf_material`"""

    _fields_ = [("dim", ctypes.c_int32), ("split", ctypes.c_int32), ("planeStress", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("E", ctypes.c_double), ("v", ctypes.c_double), ("lam", ctypes.c_double), ("mu", ctypes.c_double),
                ("bulk", ctypes.c_double), ("C", ctypes.c_double * 36), ("sqrtC", ctypes.c_double * 36),
                ("inv_sqrtC", ctypes.c_double * 36)]


def c_material(mat: orc.IsoMaterial, split: str):
    m = CPfMat(mat.dim, SPLITS[split], int(mat.planeStress), 0, mat.E, mat.v, mat.lam, mat.mu, mat.bulk)
    n = mat.C.size
    m.C[:n] = mat.C.ravel().tolist()
    m.sqrtC[:n] = mat.sqrtC.ravel().tolist()
    m.inv_sqrtC[:n] = mat.inv_sqrtC.ravel().tolist()
    return m


def generic_states(dim, Ne, nPg, seed, min_gap=1e-3):
    """random strain states whose principal values are separated by >= min_gap relative (well-conditioned projectors)"""
    rng = np.random.default_rng(seed)
    ns = 3 if dim == 2 else 6
    eps = rng.normal(size=(4 * Ne, nPg, ns)) * 1e-3
    w = np.linalg.eigvalsh(orc._vec_to_mat(eps))
    gap = np.diff(w, axis=-1).min(-1) / np.abs(w).max(-1)
    good = (gap > min_gap).all(axis=1)
    return np.ascontiguousarray(eps[good][:Ne])


def run_split(hostcheck, mat, split, eps, g=None):
    Ne, nPg, ns = eps.shape
    cP = np.empty((Ne, nPg, ns, ns)); cM = np.empty_like(cP); psiP = np.empty((Ne, nPg)); psiM = np.empty((Ne, nPg))
    Cdeg = np.empty_like(cP) if g is not None else None
    m = c_material(mat, split)
    assert hostcheck.hc_pf_split(ctypes.byref(m), p(eps), ctypes.c_int64(Ne), ctypes.c_int32(nPg), p(cP), p(cM), p(psiP), p(psiM),
                                 p(g), p(Cdeg)) == 0
    return cP, cM, psiP, psiM, Cdeg


@pytest.mark.parametrize("dim,planeStress", [(2, False), (2, True), (3, False)])
@pytest.mark.parametrize("split", list(SPLITS))
def test_split_generic_states(hostcheck, dim, planeStress, split):
    mat = orc.IsoMaterial(dim, 210000.0, 0.3, planeStress)
    eps = generic_states(dim, 300, 4, seed=11)
    g = np.random.default_rng(1).uniform(0, 1, eps.shape[:2])
    cP, cM, psiP, psiM, Cdeg = run_split(hostcheck, mat, split, eps, g)
    ocP, ocM = orc.calc_C(mat, split, eps)
    opP, opM = orc.calc_psi(mat, split, eps)
    assert rel_err(cP, ocP) < 1e-12 and rel_err(cM, ocM) < 1e-12
    assert rel_err(psiP, opP) < 1e-12 and rel_err(psiM, opM) < 1e-12
    assert rel_err(Cdeg, g[..., None, None] * ocP + ocM) < 1e-12
    # the reference's own property checks (tests/Models/phasefield_test.py:90-137): c+ + c- = C, psi+ + psi- = psi
    assert rel_err(cP + cM, np.broadcast_to(mat.C, cP.shape)) < 1e-12
    psi = 0.5 * np.einsum("epi,ij,epj->ep", eps, mat.C, eps)
    assert rel_err(psiP + psiM, psi) < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("split", ["Amor", "Miehe", "Stress", "He"])
def test_split_degenerate_states_repair_policy(hostcheck, dim, split):
    """States where the reference's formulas are NaN / knife-edge: the build must return finite numbers that equal the
    oracle's repair policy (clamp=True) and still satisfy the partition c+ + c- = C."""
    mat = orc.IsoMaterial(dim, 210000.0, 0.3)
    ns = 3 if dim == 2 else 6
    rng = np.random.default_rng(2)
    eps = rng.normal(size=(12, 3, ns)) * 1e-3
    eps[0] = 0.0  # zero state
    eps[1] = 0.0; eps[1, :, 0] = 1e-3  # uniaxial
    eps[2] = 0.0; eps[2, :, :dim] = 1e-3  # isotropic
    eps[3, 1] = 0.0  # one zero point inside a generic element
    eps[4, 0] = 0.0; eps[4, 0, :2] = 2e-3  # equibiaxial point inside a generic element
    eps[5] = 0.0; eps[5, :, 0] = 1e-3; eps[5, :, 1:dim] = -0.3e-3  # (e, -ve, -ve)
    cP, cM, psiP, psiM, _ = run_split(hostcheck, mat, split, np.ascontiguousarray(eps))
    assert np.isfinite(cP).all() and np.isfinite(cM).all() and np.isfinite(psiP).all()
    assert rel_err(cP + cM, np.broadcast_to(mat.C, cP.shape)) < 1e-11
    ocP, ocM = orc.calc_C(mat, split, eps, clamp=True)
    # generic elements (untouched by any special point) must agree tightly
    generic = np.arange(6, 12)
    assert rel_err(cP[generic], ocP[generic]) < 1e-12


def _pattern(hostcheck, connects, nPes, Nn):
    n = len(connects)
    c32 = [np.ascontiguousarray(c, dtype=np.int32) for c in connects]
    ptrs = (ctypes.c_void_p * n)(*[c.ctypes.data for c in c32])
    Ne = (ctypes.c_int64 * n)(*[c.shape[0] for c in c32])
    nPe = (ctypes.c_int32 * n)(*nPes)
    hostcheck.hc_pattern_build.restype = ctypes.c_void_p
    hostcheck.hc_pattern_nnz_node.restype = ctypes.c_int64
    P = ctypes.c_void_p(hostcheck.hc_pattern_build(I(n), ptrs, Ne, nPe, ctypes.c_int64(Nn)))
    return P, (c32, ptrs, Ne, nPe)


@pytest.mark.parametrize("elemType", ["TRI3", "QUAD9", "TETRA4", "HEXA8", "HEXA27"])
@pytest.mark.parametrize("dof_n", [1, "dim"])
def test_pattern_and_replay_bit_exact(hostcheck, elemType, dof_n):
    coords, connect = make_mesh(elemType)
    dim, nPe = el.elem_dim(elemType), el.elem_nPe(elemType)
    d = dim if dof_n == "dim" else 1
    Nn = coords.shape[0] + 2  # two orphan nodes at the end: rows without entries
    Ndof = Nn * d + 3  # plus Lagrange rows (_simu.py:154-158)
    P, (c32, ptrs, Ne, nPeA) = _pattern(hostcheck, [connect], [nPe], Nn)
    nnz = hostcheck.hc_pattern_nnz_node(P) * d * d
    indptr = np.empty(Ndof + 1, np.int32); indices = np.empty(nnz, np.int32)
    hostcheck.hc_pattern_expand(P, ctypes.c_int64(Nn), I(d), ctypes.c_int64(Ndof), p(indptr), p(indices))
    oinv, oind, optr, onnz = orc.csr_map([connect], d, Ndof, True)
    assert onnz == nnz and np.array_equal(indptr, optr) and np.array_equal(indices, oind)
    inv = np.empty(connect.shape[0] * (nPe * d) ** 2, np.int32)
    hostcheck.hc_pattern_inv(P, I(1), ptrs, Ne, nPeA, I(d), p(inv))
    assert np.array_equal(inv, oinv)
    # replay: bit-identical to np.bincount on random element matrices
    rng = np.random.default_rng(9)
    Ke = rng.normal(size=(connect.shape[0], nPe * d, nPe * d))
    dptr = (ctypes.c_void_p * 1)(Ke.ctypes.data)
    out = np.full(nnz, np.nan)
    hostcheck.hc_replay_matrix(P, I(1), dptr, Ne, nPeA, I(d), ctypes.c_int64(Nn), p(out))
    assert np.array_equal(out, orc.assemble_replay([Ke], oinv, onnz))
    # vectors
    Fe = rng.normal(size=(connect.shape[0], nPe * d))
    fptr = (ctypes.c_void_p * 1)(Fe.ctypes.data)
    dense = np.full(Nn * d, np.nan)
    hostcheck.hc_replay_vector(P, I(1), fptr, Ne, nPeA, I(d), ctypes.c_int64(Nn), p(dense))
    vinv, vind, vptr, vnnz = orc.csr_map([connect], d, Ndof, False)
    ref = orc.assemble_replay([Fe], vinv, vnnz)
    has = np.diff(vptr) > 0
    assert np.array_equal(dense[has[: Nn * d]], ref) and np.all(dense[~has[: Nn * d]] == 0)
    hostcheck.hc_pattern_free(P)


def test_pattern_two_groups_mixed(hostcheck):
    """TRI3 + QUAD4 sharing nodes (the cardiac-benchmark situation of _simu.py:1001-1004: several groups feed one matrix)."""
    cq, q = make_mesh("QUAD4", (4, 3))
    tri = np.concatenate([q[:6, [0, 1, 2]], q[:6, [0, 2, 3]]])  # triangles over the first 6 quads
    quad = q[4:]
    Nn = cq.shape[0]
    for d in (1, 2):
        P, (c32, ptrs, Ne, nPeA) = _pattern(hostcheck, [tri, quad], [3, 4], Nn)
        nnz = hostcheck.hc_pattern_nnz_node(P) * d * d
        indptr = np.empty(Nn * d + 1, np.int32); indices = np.empty(nnz, np.int32)
        hostcheck.hc_pattern_expand(P, ctypes.c_int64(Nn), I(d), ctypes.c_int64(Nn * d), p(indptr), p(indices))
        oinv, oind, optr, onnz = orc.csr_map([tri, quad], d, Nn * d, True)
        assert onnz == nnz and np.array_equal(indptr, optr) and np.array_equal(indices, oind)
        inv = np.empty(oinv.size, np.int32)
        hostcheck.hc_pattern_inv(P, I(2), ptrs, Ne, nPeA, I(d), p(inv))
        assert np.array_equal(inv, oinv)
        rng = np.random.default_rng(3)
        K1 = rng.normal(size=(tri.shape[0], 3 * d, 3 * d)); K2 = rng.normal(size=(quad.shape[0], 4 * d, 4 * d))
        dptr = (ctypes.c_void_p * 2)(K1.ctypes.data, K2.ctypes.data)
        out = np.empty(nnz)
        hostcheck.hc_replay_matrix(P, I(2), dptr, Ne, nPeA, I(d), ctypes.c_int64(Nn), p(out))
        assert np.array_equal(out, orc.assemble_replay([K1, K2], oinv, onnz))
        hostcheck.hc_pattern_free(P)


@pytest.mark.parametrize("elemType,dim", [("TRI3", 2), ("TETRA4", 3)])
@pytest.mark.parametrize("split", list(SPLITS))
def test_fused_S3_simplex(hostcheck, elemType, dim, split):
    """`efb_pf_elastic_Ke` body (strain -> split -> g cP + cM -> K_e in one pass, configs 3 and 4) against the oracle composition
    of Simulations/_phasefield.py:444-482"""
    from tests.helpers import make_mesh

    rng = np.random.default_rng(8)
    coords, connect = make_mesh(elemType, (7, 6) if dim == 2 else (4, 3, 3))
    g, keep, tab = host_group(elemType, coords, connect, "rigi")
    Nn = coords.shape[0]
    mat = orc.IsoMaterial(dim, 210e9, 0.3, False)
    u = rng.normal(size=Nn * dim) * 1e-5
    dmg = rng.uniform(0, 0.9, Nn)
    c32 = np.ascontiguousarray(connect, dtype=np.int32)
    Ke = np.empty((g.Ne, g.nPe * dim, g.nPe * dim))
    m = c_material(mat, split)
    assert hostcheck.hc_pf_elastic_Ke(ctypes.byref(m), ctypes.byref(g), p(c32), p(u), p(dmg), ctypes.c_double(1e-12), ctypes.c_double(0.5),
                                      p(Ke)) == 0
    geo = orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights)
    ref = 0.5 * orc.pf_elastic_Ke(geo, tab.N_pg, mat, split, orc.locate_sol_e(u, connect, dim), dmg[connect], clamp=True)
    assert rel_err(Ke, ref) < 1e-12
