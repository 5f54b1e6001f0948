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


@pytest.mark.parametrize("opt", [1, 2, 3])
@pytest.mark.parametrize("name,elemType,split", [("TRI3_Miehe", "TRI3", "Miehe"), ("TETRA4_He", "TETRA4", "He")])
def test_staggered_convergence_options_match_reference(name, elemType, split, opt):
    """convOption 1 (crack energy), 2 (total energy), 3 (summed relative increments) of `Simulations.PhaseField.Solve`
    (Simulations/_phasefield.py:354-397), with the reference's "assemble only when the other field changed" flags."""
    d = dict(np.load(os.path.join(GOLD, f"staggered_conv{opt}_{name}.npz")))
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
        u, dmg, conv = simu.Solve(float(d["tolConv"]), 60, convOption=int(d["convOption"]))
        assert conv and simu.Niter == int(d[f"Niter_{k}"]), (simu.Niter, int(d[f"Niter_{k}"]))
        assert abs(simu.convIter - float(d[f"convIter_{k}"])) <= 1e-6 * abs(float(d[f"convIter_{k}"]))
        assert rel_err(dmg, d[f"d_{k}"]) < 1e-8 and rel_err(u, d[f"u_{k}"]) < 1e-8
        assert rel_err(simu.psiP, d[f"psiP_{k}"]) < 1e-8
        # the generator evaluated both energies after the solve (which refreshes Kd, hence the history, for convOption 3)
        assert abs(simu.Psi_Crack() - float(d[f"Psi_Crack_{k}"])) <= 1e-8 * abs(float(d[f"Psi_Crack_{k}"]))
        assert abs(simu.Psi_Elas() - float(d[f"Psi_Elas_{k}"])) <= 1e-8 * abs(float(d[f"Psi_Elas_{k}"]))
        simu.Save_Iter()
