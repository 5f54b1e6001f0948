"""Kernel-body parity on the CPU: the phase-structured block bodies of csrc/elem_kernels.cuh, compiled by g++ through
tests/hostcheck (test-only emulation), against the NumPy oracle.  Tolerance 1e-12 relative (north star)."""
import ctypes

import numpy as np
import pytest

from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc
from tests.helpers import ELEM_CASES, host_group, make_mesh, p, rel_err

TOL = 1e-12
D = ctypes.c_double
I = ctypes.c_int


def _geo(elemType, coords, connect, tab):
    dim = el.elem_dim(elemType)
    return orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights)


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
def test_geometry(hostcheck, elemType, mt):
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, mt)
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    ns = 3 if dim == 2 else 6
    F = np.empty((Ne, nPg, dim, dim)); detF = np.empty((Ne, nPg)); jac = np.empty((Ne, nPg)); wJ = np.empty((Ne, nPg))
    invF = np.empty_like(F); dN = np.empty((Ne, nPg, dim, nPe)); B = np.empty((Ne, nPg, ns, nPe * dim))
    assert hostcheck.hc_geometry(ctypes.byref(g), p(F), p(detF), p(jac), p(wJ), p(invF), p(dN), p(B)) == 0
    geo = _geo(elemType, coords, connect, tab)
    for name, arr in (("F", F), ("detF", detF), ("jac", jac), ("wJ", wJ), ("invF", invF), ("dN", dN)):
        assert rel_err(arr, geo[name]) < TOL, name
    assert rel_err(B, orc.B_matrix(geo["dN"])) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
def test_geometry_parts(hostcheck, elemType, mt):
    """G8-G10: Get_leftDispPart / ReactionPart / DiffusePart / SourcePart_e_pg (_group_elem.py:1314-1407)"""
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, mt)
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    ns = 3 if dim == 2 else 6
    geo = _geo(elemType, coords, connect, tab)
    for dof_n in (1, dim):
        nd = nPe * dof_n
        left = np.empty((Ne, nPg, nPe * dim, ns)); reac = np.empty((Ne, nPg, nd, nd))
        diff = np.empty((Ne, nPg, nPe, dim)); src = np.empty((Ne, nPg, nd, dof_n))
        assert hostcheck.hc_geometry_parts(ctypes.byref(g), I(dof_n), p(left), p(reac), p(diff), p(src)) == 0
        ref = orc.geometry_parts(geo, tab.N_pg, dof_n)
        for name, arr in (("leftDisp", left), ("reaction", reac), ("diffuse", diff), ("source", src)):
            assert rel_err(arr, ref[name]) < TOL, (name, dof_n)
    reac = np.empty((Ne, nPg, nPe, nPe))  # outputs are optional: mass-type factors alone need no gradient table
    assert hostcheck.hc_geometry_parts(ctypes.byref(g), I(1), None, p(reac), None, None) == 0
    assert rel_err(reac, orc.geometry_parts(geo, tab.N_pg, 1)["reaction"]) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("C_mode", [0, 1, 2])
def test_elastic_Ke(hostcheck, elemType, C_mode):
    rng = np.random.default_rng(3)
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, "rigi")
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    ns = 3 if dim == 2 else 6
    C0 = orc.IsoMaterial(dim, 210000.0, 0.3).C
    shape = [(ns, ns), (Ne, ns, ns), (Ne, nPg, ns, ns)][C_mode]
    C = np.ascontiguousarray(np.broadcast_to(C0, shape) * rng.uniform(0.5, 2.0, shape))  # non-symmetric on purpose
    out = np.empty((Ne, nPe * dim, nPe * dim))
    assert hostcheck.hc_elastic_Ke(ctypes.byref(g), p(C), I(C_mode), D(1.5), p(out)) == 0
    ref = 1.5 * orc.linearized_elasticity(_geo(elemType, coords, connect, tab), C)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("elemType", ["TRI3", "QUAD4", "TRI6", "QUAD9", "TETRA4", "HEXA8"])
@pytest.mark.parametrize("form", [1, 2, 3])
def test_elastic_Ke_warp_forms(hostcheck, elemType, form):
    """Warp-autonomous homogeneous-C forms: general, symmetric (cyclic block cover + mirrored stores), symmetric + ortho
    (structural zeros of C skipped) — the same bodies the device kernel `k_elastic_w` runs."""
    rng = np.random.default_rng(5)
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, "rigi")
    dim, nPe, Ne = g.dim, g.nPe, g.Ne
    ns = 3 if dim == 2 else 6
    C = orc.IsoMaterial(dim, 210000.0, 0.3).C.copy()
    if form == 1:  # arbitrary, non-symmetric
        C = C * rng.uniform(0.5, 2.0, C.shape) + rng.uniform(-1e4, 1e4, C.shape)
    elif form == 2:  # symmetric, fully populated
        R = rng.uniform(-2e4, 2e4, C.shape)
        C = C + R + R.T
    else:  # orthotropic pattern: distinct normal block and shear moduli
        C[:dim, :dim] *= np.array(rng.uniform(0.5, 2.0, (dim, dim)) + 0.0)
        C[:dim, :dim] = 0.5 * (C[:dim, :dim] + C[:dim, :dim].T)
        C[np.arange(dim, ns), np.arange(dim, ns)] *= rng.uniform(0.5, 2.0, ns - dim)
    C = np.ascontiguousarray(C)
    out = np.full((Ne, nPe * dim, nPe * dim), np.nan)
    assert hostcheck.hc_elastic_Ke_warp(ctypes.byref(g), p(C), I(form), D(1.5), p(out)) == 0
    ref = 1.5 * orc.linearized_elasticity(_geo(elemType, coords, connect, tab), C)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
@pytest.mark.parametrize("body", ["warp", "block"])
def test_scalar_operators(hostcheck, elemType, mt, body, monkeypatch):
    """both bodies of the scalar operators: the warp-autonomous form (csrc/scalar_warp.cuh, elements of at most 8 nodes) and the
    phase-structured block form (every element type); the device picks per operator (elem_kernels.cu launch_scalar)"""
    if body == "block":
        monkeypatch.setenv("HC_SCALAR_BLOCK", "1")
    rng = np.random.default_rng(4)
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, mt)
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    geo = _geo(elemType, coords, connect, tab)
    coefs = {0: 2.5, 1: rng.uniform(1, 2, Ne), 2: rng.uniform(1, 2, nPg), 3: rng.uniform(1, 2, (Ne, nPg))}

    def call(**kw):
        a = dict(r=None, r_mode=0, r_scalar=0.0, has_r=0, A=None, A_mode=0, k=None, k_mode=0, k_scalar=0.0, has_k=0, f=None,
                 f_mode=0, f_scalar=0.0, has_f=0, dof_n=1, scale=1.0, Ke=None, Fe=None, keep=0)
        a.update(kw)
        rc = hostcheck.hc_scalar(ctypes.byref(g), p(a["r"]), I(a["r_mode"]), D(a["r_scalar"]), I(a["has_r"]), p(a["A"]), I(a["A_mode"]),
                                 p(a["k"]), I(a["k_mode"]), D(a["k_scalar"]), I(a["has_k"]), p(a["f"]), I(a["f_mode"]),
                                 D(a["f_scalar"]), I(a["has_f"]), I(a["dof_n"]), D(a["scale"]), p(a["Ke"]), p(a["Fe"]), I(a["keep"]))
        assert rc == 0

    for mode, c in coefs.items():
        arr = None if mode == 0 else np.ascontiguousarray(c)
        sc = c if mode == 0 else 0.0
        for dof_n in (1, dim):
            out = np.empty((Ne, nPe * dof_n, nPe * dof_n))
            call(r=arr, r_mode=mode, r_scalar=sc, has_r=1, dof_n=dof_n, Ke=out, scale=0.5)
            assert rel_err(out, 0.5 * orc.uv(geo, tab.N_pg, c, dof_n)) < TOL, ("uv", mode, dof_n)
            outf = np.empty((Ne, nPe * dof_n, dof_n))
            call(f=arr, f_mode=mode, f_scalar=sc, has_f=1, dof_n=dof_n, Fe=outf, keep=1)
            assert rel_err(outf, orc.source_v(geo, tab.N_pg, c, dof_n)) < TOL, ("V", mode, dof_n)
        out = np.empty((Ne, nPe, nPe))
        call(k=arr, k_mode=mode, k_scalar=sc, has_k=1, Ke=out)
        assert rel_err(out, orc.grad_u_a_grad_v(geo, None, c)) < TOL, ("GradUGradV", mode)
        for A_mode, shape in enumerate([(dim, dim), (Ne, dim, dim), (Ne, nPg, dim, dim)]):
            A = rng.uniform(1, 2, shape)
            call(A=A, A_mode=A_mode, k=arr, k_mode=mode, k_scalar=sc, has_k=1, Ke=out)
            assert rel_err(out, orc.grad_u_a_grad_v(geo, A, c)) < TOL, ("GradU_A_GradV", mode, A_mode)
    # fused damage system: Ke = R + D, Fe
    r = rng.uniform(1, 2, (Ne, nPg)); f = rng.uniform(1, 2, (Ne, nPg)); A = np.eye(dim) + 0.1 * rng.uniform(size=(dim, dim))
    Ke = np.empty((Ne, nPe, nPe)); Fe = np.empty((Ne, nPe))
    call(r=r, r_mode=3, has_r=1, A=A, A_mode=0, k_scalar=0.027, has_k=1, f=f, f_mode=3, has_f=1, Ke=Ke, Fe=Fe, scale=2.0)
    assert rel_err(Ke, 2.0 * (orc.uv(geo, tab.N_pg, r) + orc.grad_u_a_grad_v(geo, A, 0.027))) < TOL
    assert rel_err(Fe, 2.0 * orc.source_v(geo, tab.N_pg, f)[..., 0]) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
def test_strain_internal_force_degradation(hostcheck, elemType, mt):
    rng = np.random.default_rng(5)
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, mt)
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    ns = 3 if dim == 2 else 6
    geo = _geo(elemType, coords, connect, tab)
    c32 = keep[0]
    u = rng.normal(size=coords.shape[0] * dim) * 1e-3
    eps = np.empty((Ne, nPg, ns))
    assert hostcheck.hc_strain(ctypes.byref(g), p(c32), p(u), p(eps)) == 0
    assert rel_err(eps, orc.strain(geo, orc.locate_sol_e(u, connect, dim))) < TOL
    sig = rng.normal(size=(Ne, nPg, ns))
    out = np.empty((Ne, nPe * dim))
    assert hostcheck.hc_internal_force(ctypes.byref(g), p(sig), p(out)) == 0
    assert rel_err(out, orc.internal_force(geo, sig)) < TOL
    d = rng.uniform(0, 1, coords.shape[0])
    gd = np.empty((Ne, nPg))
    assert hostcheck.hc_degradation(ctypes.byref(g), p(c32), p(d), D(1e-12), p(gd)) == 0
    assert rel_err(gd, orc.degradation(d[connect], tab.N_pg)) < TOL


@pytest.mark.parametrize("elemType", ["TRI3", "QUAD4", "TRI6", "QUAD9", "TETRA4", "HEXA8", "TETRA10"])
def test_hyperelastic_Ke_Re(hostcheck, elemType):
    """section 8f rank 3: kernel body of `efb_hyperelastic_Ke_Re` (B = De(u) grad, material + geometric tangent, residual) against
    the oracle restatement of NonLinear.py:37-201, random symmetric d2W and random dW at every Gauss point"""
    rng = np.random.default_rng(12)
    coords, connect = make_mesh(elemType)
    g, keep, tab = host_group(elemType, coords, connect, "rigi")
    dim, nPe, nPg, Ne = g.dim, g.nPe, g.nPg, g.Ne
    ns = 3 if dim == 2 else 6
    Nn = coords.shape[0]
    u = rng.normal(size=Nn * dim) * 0.05
    dW = rng.normal(size=(Ne, nPg, ns))
    d2W = rng.normal(size=(Ne, nPg, ns, ns))
    d2W = d2W + np.swapaxes(d2W, -1, -2)
    c32 = np.ascontiguousarray(connect, dtype=np.int32)
    Ke = np.empty((Ne, nPe * dim, nPe * dim)); Re = np.empty((Ne, nPe * dim))
    assert hostcheck.hc_hyper(ctypes.byref(g), p(c32), p(u), p(dW), p(d2W), D(0.8), p(Ke), p(Re)) == 0
    geo = _geo(elemType, coords, connect, tab)
    K, R = orc.hyper_Ke_Re(geo, orc.locate_sol_e(u, connect, dim), dW, d2W, dim, 0.8)
    assert rel_err(Ke, K) < TOL and rel_err(Re, R) < TOL
