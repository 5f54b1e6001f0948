"""GPU: post-processing fields (easyfea_b200.results) against fixtures minted from `Simulations.Elastic.Result` of the live
reference and against the NumPy oracle: element and node values of every strain/stress result, von Mises, Wdef (<= 1e-12)."""
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.test_oracle_results import CASES, GOLD, rel, result_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def efb():
    import types

    from easyfea_b200 import _lib, mesh, results

    _lib.require_cuda()
    return types.SimpleNamespace(mesh=mesh, results=results)


@pytest.mark.parametrize("et", CASES)
def test_results_match_reference(efb, et):
    d = np.load(os.path.join(GOLD, f"results_{et}.npz"))
    g = efb.mesh.ElemGroup(et, d["connect"], d["coords"], all_nodes_used=True)
    R = efb.results.ElasticResults(g, d["C"], float(d["thickness"]))
    for name in result_names(g.dim):
        assert rel(R.Result(d["u"], name, nodeValues=False), d[name + "_e"]) < 1e-12, name
        assert rel(R.Result(d["u"], name, nodeValues=True), d[name + "_n"]) < 1e-12, name
    assert rel(R.Result(d["u"], "Wdef_e", nodeValues=False), d["Wdef_e"]) < 1e-12
    assert abs(R.Result(d["u"], "Wdef") - float(d["Wdef"])) <= 1e-12 * abs(float(d["Wdef"]))
    with pytest.raises(Exception):
        R.Result(d["u"], "Sab")


def test_node_values_bitwise_vs_scipy_order(efb):
    """element -> node averaging sums in ascending element order, like scipy's `connect_n_e @ values_e`"""
    d = np.load(os.path.join(GOLD, "results_HEXA8.npz"))
    g = efb.mesh.ElemGroup("HEXA8", d["connect"], d["coords"], all_nodes_used=True)
    rng = np.random.default_rng(5)
    vals = rng.standard_normal((d["connect"].shape[0], 4))
    got = efb.results.Get_Node_Values(g, vals)
    assert np.array_equal(got, orc.node_values(d["connect"], d["coords"].shape[0], vals))
    got1 = efb.results.Get_Node_Values(g, vals[:, 0].copy())
    assert got1.shape == (d["coords"].shape[0],) and np.array_equal(got1, got[:, 0])
