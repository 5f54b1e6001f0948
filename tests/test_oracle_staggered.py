"""Pin the oracle's restatement of the staggered phase-field loop against fixtures minted from the live reference
(tests/golden/make_golden_staggered.py) — same direct solver, so fields agree to rounding-amplified 1e-8."""
import os

import numpy as np
import pytest

from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,elemType,split", [("TRI3_Miehe", "TRI3", "Miehe"), ("TETRA4_He", "TETRA4", "He"),
                                                 ("QUAD9_Amor", "QUAD9", "Amor")])
def test_staggered_oracle_matches_reference(name, elemType, split):
    d = dict(np.load(os.path.join(GOLD, f"staggered_{name}.npz")))
    L, l0, E, v, Gc = d["params"]
    dim = el.elem_dim(elemType)
    tr, tm = el.gauss_table(elemType, "rigi"), el.gauss_table(elemType, "mass")
    simu = orc.StaggeredOracle(d["coords"], d["connect"], tr.dN_pg, tr.weights, tr.N_pg, tm.dN_pg, tm.weights, tm.N_pg,
                               orc.IsoMaterial(dim, E, v, False), split, "AT2", Gc, l0)
    for k, dep in enumerate(d["loads"]):
        simu.Bc_Init()
        simu.add_dirichlet(d["crack"], [1], [0], problemType="damage")
        simu.add_dirichlet(d["top"], [dep, 0.5 * dep] + [0] * (dim - 2), list(range(dim)))
        simu.add_dirichlet(d["bot"], [0] * dim, list(range(dim)))
        u, dmg, conv = simu.Solve(1e-3, 50)
        assert conv and simu.Niter == int(d[f"Niter_{k}"])
        print(name, k, rel_err(dmg, d[f"d_{k}"]), rel_err(u, d[f"u_{k}"]), rel_err(simu.psiP, d[f"psiP_{k}"]))
        assert rel_err(dmg, d[f"d_{k}"]) < 1e-8
        assert rel_err(u, d[f"u_{k}"]) < 1e-8
        assert rel_err(simu.psiP, d[f"psiP_{k}"]) < 1e-8
        simu.Save_Iter()
