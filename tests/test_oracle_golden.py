"""Pin the NumPy oracle against the golden fixtures minted from the live reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["QUAD9", "HEXA8", "TRI3", "TETRA4", "HEXA27"]
TOL = 1e-12


def load(name):
    return dict(np.load(os.path.join(GOLD, f"{name}.npz")))


@pytest.mark.parametrize("name", CASES)
def test_tables_match_reference(name):
    from easyfea_b200 import elements as el

    d = load(name)
    for mt in ("rigi", "mass"):
        tab = el.gauss_table(name, mt)
        assert np.array_equal(tab.weights, d[f"w_pg_{mt}"])
        assert np.abs(tab.N_pg - d[f"N_pg_{mt}"]).max() < 1e-15
        assert np.abs(tab.dN_pg - d[f"dN_pg_{mt}"]).max() < 2e-15


@pytest.mark.parametrize("name", CASES)
def test_geometry_and_operators(name):
    d = load(name)
    co, cn = d["coords"], d["connect"]
    dim = d["dN_pg_rigi"].shape[1]
    geo = {mt: orc.geometry(co[cn][:, :, :dim], d[f"dN_pg_{mt}"], d[f"w_pg_{mt}"]) for mt in ("rigi", "mass")}
    for mt in ("rigi", "mass"):
        assert rel_err(geo[mt]["jac"], d[f"jac_{mt}"]) < TOL
        assert rel_err(geo[mt]["dN"], d[f"dN_e_pg_{mt}"]) < TOL
    assert rel_err(orc.B_matrix(geo["rigi"]["dN"]), d["B_rigi"]) < TOL
    assert rel_err(orc.linearized_elasticity(geo["rigi"], d["C"]), d["Ke"]) < TOL
    assert rel_err(orc.linearized_elasticity(geo["rigi"], d["C_e_pg"]), d["Ke_epg"]) < TOL
    assert rel_err(orc.uv(geo["mass"], d["N_pg_mass"], d["rho_e_pg"], dim), d["Me"]) < TOL
    assert rel_err(orc.uv(geo["mass"], d["N_pg_mass"], 2.0, 1), d["Me1"]) < TOL
    assert rel_err(orc.grad_u_a_grad_v(geo["rigi"], d["A"], 3.0), d["De"]) < TOL
    nPg = d["w_pg_rigi"].size
    assert rel_err(orc.grad_u_a_grad_v(geo["rigi"], None, d["rho_e_pg"][:, :1].repeat(nPg, 1)), d["De0"]) < TOL
    assert rel_err(orc.source_v(geo["mass"], d["N_pg_mass"], d["rho_e_pg"], 1), d["Fe"]) < TOL
    for mt in ("rigi", "mass"):
        assert rel_err(orc.strain(geo[mt], orc.locate_sol_e(d["u"], cn, dim)), d[f"eps_{mt}"]) < TOL
    assert rel_err(orc.degradation(d["dmg"][cn], d["N_pg_rigi"]), d["g_rigi"]) < TOL


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("split", ["Amor", "Miehe", "Stress", "He"])
def test_splits(name, split):
    d = load(name)
    dim = d["dN_pg_rigi"].shape[1]
    mat = orc.IsoMaterial(dim, 210000.0, 0.3, planeStress=False)
    assert np.array_equal(mat.C, d["C"])
    for mt in ("rigi", "mass"):
        eps = d[f"eps_{mt}"]
        cP, cM = orc.calc_C(mat, split, eps)
        pP, pM = orc.calc_psi(mat, split, eps)
        assert rel_err(cP, d[f"cP_{split}_{mt}"]) < TOL and rel_err(cM, d[f"cM_{split}_{mt}"]) < TOL
        assert rel_err(pP, d[f"psiP_{split}_{mt}"]) < TOL and rel_err(pM, d[f"psiM_{split}_{mt}"]) < TOL
        cPc, _ = orc.calc_C(mat, split, eps, clamp=True)  # the repair policy must not touch finite reference points
        assert np.array_equal(cPc, cP)


@pytest.mark.parametrize("name", CASES)
def test_csr_map_and_replay(name):
    d = load(name)
    cn = d["connect"]
    dim = d["dN_pg_rigi"].shape[1]
    Nn = d["coords"].shape[0]
    for dof_n in (1, dim):
        inv, indices, indptr, nnz = orc.csr_map([cn], dof_n, Nn * dof_n, True)
        for got, key in ((inv, "inv"), (indices, "indices"), (indptr, "indptr")):
            ref = d[f"{key}_{dof_n}"]
            assert got.dtype == ref.dtype and np.array_equal(got, ref), key
    inv, indices, indptr, nnz = orc.csr_map([cn], dim, Nn * dim, True)
    assert np.array_equal(orc.assemble_replay([d["Ke"]], inv, nnz), d["K_data"])  # bit-exact ordered sum
    vinv, vind, vptr, vnnz = orc.csr_map([cn], 1, Nn, False)
    assert np.array_equal(vptr, d["F_indptr"])
    assert np.array_equal(orc.assemble_replay([d["Fe"]], vinv, vnnz), d["F_data"])
