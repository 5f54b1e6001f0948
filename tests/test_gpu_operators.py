"""GPU parity (through the C ABI) of geometry + operators against the oracle and the golden fixtures. -m gpu"""
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.helpers import ELEM_CASES, make_mesh, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12  # north star: element matrices agree to a relative 1e-12 in FP64
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def efb():
    from easyfea_b200 import _lib
    from easyfea_b200 import elements, mesh, operators

    _lib.require_cuda()

    class NS:
        pass

    ns = NS()
    ns.op, ns.mesh, ns.el = operators, mesh, elements
    return ns


def _geo(efb, elemType, coords, connect, mt):
    tab = efb.el.gauss_table(elemType, mt)
    dim = efb.el.elem_dim(elemType)
    return orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights), tab


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
def test_geometry_getters(efb, elemType, mt):
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    geo, tab = _geo(efb, elemType, coords, connect, mt)
    op = efb.op
    assert rel_err(op.Get_F_e_pg(g, mt), geo["F"]) < TOL
    assert rel_err(op.Get_jacobian_e_pg(g, mt), geo["jac"]) < TOL
    assert rel_err(op.Get_jacobian_e_pg(g, mt, absoluteValues=False), geo["detF"]) < TOL
    assert rel_err(op.Get_weightedJacobian_e_pg(g, mt), geo["wJ"]) < TOL
    assert rel_err(op.Get_invF_e_pg(g, mt), geo["invF"]) < TOL
    assert rel_err(op.Get_dN_e_pg(g, mt), geo["dN"]) < TOL
    assert rel_err(op.Get_B_e_pg(g, mt), orc.B_matrix(geo["dN"])) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
@pytest.mark.parametrize("mt", ["rigi", "mass"])
def test_geometry_part_getters(efb, elemType, mt):
    """G8-G10 (_group_elem.py:1314-1407) through the C ABI"""
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    geo, tab = _geo(efb, elemType, coords, connect, mt)
    op = efb.op
    for dof_n in (1, g.dim):
        ref = orc.geometry_parts(geo, tab.N_pg, dof_n)
        assert rel_err(op.Get_ReactionPart_e_pg(g, mt, dof_n), ref["reaction"]) < TOL
        assert rel_err(op.Get_SourcePart_e_pg(g, mt, dof_n), ref["source"]) < TOL
    assert rel_err(op.Get_leftDispPart_e_pg(g, mt), ref["leftDisp"]) < TOL
    assert rel_err(op.Get_DiffusePart_e_pg(g, mt), ref["diffuse"]) < TOL


@pytest.mark.parametrize("elemType", list(ELEM_CASES))
def test_operators_all_broadcast_modes(efb, elemType):
    rng = np.random.default_rng(3)
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    op = efb.op
    dim, nPe, Ne = g.dim, g.nPe, g.Ne
    ns = 3 if dim == 2 else 6
    C0 = orc.IsoMaterial(dim, 210000.0, 0.3).C
    for mt in ("rigi", "mass"):
        geo, tab = _geo(efb, elemType, coords, connect, mt)
        nPg = tab.nPg
        for shape in [(ns, ns), (Ne, ns, ns), (Ne, nPg, ns, ns)]:
            C = np.broadcast_to(C0, shape) * rng.uniform(0.5, 2.0, shape)
            K = op.LinearizedElasticity(g, C, mt)
            assert K.flags.writeable and K.flags.c_contiguous and K.shape == (Ne, nPe * dim, nPe * dim)
            assert rel_err(K, orc.linearized_elasticity(geo, C)) < TOL
        coefs = [2.5, rng.uniform(1, 2, Ne), rng.uniform(1, 2, (Ne, nPg))]
        if nPg != Ne:
            coefs.append(rng.uniform(1, 2, nPg))
        for c in coefs:
            for dof_n in (1, dim):
                assert rel_err(op.UV(g, c, dof_n, mt), orc.uv(geo, tab.N_pg, c, dof_n)) < TOL
                assert rel_err(op.V(g, c, dof_n, mt), orc.source_v(geo, tab.N_pg, c, dof_n)) < TOL
            assert rel_err(op.GradUGradV(g, c, mt), orc.grad_u_a_grad_v(geo, None, c)) < TOL
            for shape in [(dim, dim), (Ne, dim, dim), (Ne, nPg, dim, dim)]:
                A = rng.uniform(1, 2, shape)
                assert rel_err(op.GradU_A_GradV(g, A, c, mt), orc.grad_u_a_grad_v(geo, A, c)) < TOL
        sig = rng.normal(size=(Ne, nPg, ns))
        assert rel_err(op.InternalForce(g, sig, mt), orc.internal_force(geo, sig)) < TOL
        u = rng.normal(size=coords.shape[0] * dim) * 1e-3
        assert rel_err(op.Calc_Epsilon_e_pg(g, u, mt), orc.strain(geo, orc.locate_sol_e(u, connect, dim))) < TOL


def test_argument_errors(efb):
    coords, connect = make_mesh("HEXA8")
    g = efb.mesh.ElemGroup("HEXA8", connect, coords)
    with pytest.raises(ValueError):
        efb.op.LinearizedElasticity(g, np.ones((5, 5)))
    with pytest.raises(ValueError):
        efb.op.LinearizedElasticity(g, np.ones((g.Ne + 1, 6, 6)))
    with pytest.raises(ValueError):
        efb.op.UV(g, np.ones(g.Ne + 3))


def test_group_with_unused_coordinates_and_cache_invalidation(efb):
    """a group that uses a subset of the mesh nodes (local coord rows != global ids) and a coordinate edit"""
    coords, connect = make_mesh("TETRA4")
    extra = np.concatenate([np.full((3, 3), 9.0), coords])  # three orphan nodes in front
    g = efb.mesh.ElemGroup("TETRA4", connect + 3, extra)
    geo, tab = _geo(efb, "TETRA4", coords, connect, "rigi")
    C = orc.IsoMaterial(3, 210000.0, 0.3).C
    assert rel_err(efb.op.LinearizedElasticity(g, C), orc.linearized_elasticity(geo, C)) < TOL
    moved = extra.copy()
    moved[3:, 0] *= 1.5
    g.coord = moved  # must drop the device mirror (same invalidation point as the reference, _group_elem.py:290-295)
    geo2 = orc.geometry(moved[3:][connect], tab.dN_pg, tab.weights)
    assert rel_err(efb.op.LinearizedElasticity(g, C), orc.linearized_elasticity(geo2, C)) < TOL


@pytest.mark.parametrize("name", ["QUAD9", "HEXA8", "TRI3", "TETRA4", "HEXA27"])
def test_golden_fixtures(efb, name):
    """the five config element types against outputs of the live reference (tests/golden/make_golden.py)"""
    d = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    g = efb.mesh.ElemGroup(name, d["connect"], d["coords"])
    op = efb.op
    dim = g.dim
    nPg = d["w_pg_rigi"].size
    for mt in ("rigi", "mass"):
        assert rel_err(op.Get_jacobian_e_pg(g, mt), d[f"jac_{mt}"]) < TOL
        assert rel_err(op.Get_dN_e_pg(g, mt), d[f"dN_e_pg_{mt}"]) < TOL
        assert rel_err(op.Calc_Epsilon_e_pg(g, d["u"], mt), d[f"eps_{mt}"]) < TOL
    assert rel_err(op.Get_B_e_pg(g, "rigi"), d["B_rigi"]) < TOL
    assert rel_err(op.LinearizedElasticity(g, d["C"]), d["Ke"]) < TOL
    assert rel_err(op.LinearizedElasticity(g, d["C_e_pg"]), d["Ke_epg"]) < TOL
    assert rel_err(op.UV(g, d["rho_e_pg"], dim), d["Me"]) < TOL
    assert rel_err(op.UV(g, 2.0, 1), d["Me1"]) < TOL
    assert rel_err(op.GradU_A_GradV(g, d["A"], 3.0), d["De"]) < TOL
    assert rel_err(op.GradUGradV(g, d["rho_e_pg"][:, :1].repeat(nPg, 1)), d["De0"]) < TOL
    assert rel_err(op.V(g, d["rho_e_pg"], 1), d["Fe"]) < TOL


@pytest.mark.parametrize("elemType", ["TRI3", "QUAD4", "QUAD9", "TETRA4", "HEXA8", "TETRA10"])
def test_hyperelastic_Ke_Re(efb, elemType):
    """section 8f rank 3 (`efb_hyperelastic_Ke_Re`): material + geometric tangent and residual vs the oracle restatement of
    Operators/NonLinear.py:37-201, random symmetric d2W / random dW"""
    rng = np.random.default_rng(12)
    coords, connect = make_mesh(elemType)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    dim, Ne = g.dim, g.Ne
    ns = 3 if dim == 2 else 6
    geo, tab = _geo(efb, elemType, coords, connect, "rigi")
    u = rng.normal(size=coords.shape[0] * dim) * 0.05
    dW = rng.normal(size=(Ne, tab.nPg, ns))
    d2W = rng.normal(size=(Ne, tab.nPg, ns, ns))
    d2W = d2W + np.swapaxes(d2W, -1, -2)
    Ke, Re = efb.op.hyperelastic_Ke_Re_dev(g, u, dW, d2W, "rigi", 0.8)
    K, R = orc.hyper_Ke_Re(geo, orc.locate_sol_e(u, connect, dim), dW, d2W, dim, 0.8)
    assert rel_err(Ke.cpu().numpy(), K) < TOL and rel_err(Re.cpu().numpy(), R) < TOL
