"""Post-processing restatement (oracle.field_result_e / psi_elas_e / node_values) against fixtures minted from the live
reference (`Simulations.Elastic.Result`, tests/golden/make_golden_results.py)."""
import os

import numpy as np
import pytest

from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["TRI3", "QUAD9", "TETRA4", "HEXA8", "HEXA27"]


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def result_names(dim):
    comps = ["xx", "yy", "xy"] if dim == 2 else ["xx", "yy", "zz", "yz", "xz", "xy"]
    return ["S" + c for c in comps] + ["E" + c for c in comps] + ["Svm", "Evm", "Stress", "Strain"]


def load(et):
    d = np.load(os.path.join(GOLD, f"results_{et}.npz"))
    dim = el.elem_dim(et)
    tab = el.gauss_table(et, "rigi")
    coords, connect = d["coords"], d["connect"]
    geo = orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights)
    eps = orc.strain(geo, orc.locate_sol_e(d["u"], connect, dim))
    return d, dim, geo, eps


@pytest.mark.parametrize("et", CASES)
def test_results_match_reference(et):
    d, dim, geo, eps = load(et)
    Nn = d["coords"].shape[0]
    sig = orc.hooke(eps, d["C"])
    for name in result_names(dim):
        field = eps if (name.startswith("E") or name == "Strain") else sig
        val_e = orc.field_result_e(field, name)
        assert rel(val_e, d[name + "_e"]) < 1e-12, name
        assert rel(orc.node_values(d["connect"], Nn, val_e), d[name + "_n"]) < 1e-12, name
    Wdef_e = orc.psi_elas_e(eps, d["C"], geo["wJ"], float(d["thickness"]))
    assert rel(Wdef_e, d["Wdef_e"]) < 1e-12 and abs(Wdef_e.sum() - float(d["Wdef"])) <= 1e-12 * abs(float(d["Wdef"]))
