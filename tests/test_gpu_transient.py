"""GPU: the device time stepping (easyfea_b200.transient) against fixtures minted from the live reference and against the
NumPy oracle — thermal parabolic steps and elastodynamic steps with every linear scheme (BASELINE config 5).
The device path solves with Jacobi-PCG to 1e-11 where the reference solves directly: fields agree to <= 1e-7 relative."""
import glob
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.test_oracle_transient import GOLD, elastic_bcs, elastic_system, rel, thermal_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def efb():
    import types

    from easyfea_b200 import _lib, mesh, phasefield, staggered, transient

    _lib.require_cuda()
    return types.SimpleNamespace(mesh=mesh, staggered=staggered, transient=transient, phasefield=phasefield)


def test_lincomb(efb):
    import torch

    rng = np.random.default_rng(0)
    vs = [rng.standard_normal(10007) for _ in range(4)]
    cs = [0.5, -2.0, 0.0, 3.25]
    tv = [torch.from_numpy(v).cuda() for v in vs]
    out = efb.transient.lincomb(list(zip(cs, tv)))
    assert np.allclose(out.cpu().numpy(), 0.5 * vs[0] + -2.0 * vs[1] + 3.25 * vs[3], rtol=1e-14, atol=1e-14)  # FMA vs mul+add
    efb.transient.lincomb([(1.0, tv[0]), (1.0, tv[1])], out=tv[0])  # in place
    assert np.allclose(tv[0].cpu().numpy(), vs[0] + vs[1], rtol=1e-15, atol=1e-15)


@pytest.mark.parametrize("et", ["HEXA27", "HEXA8"])
def test_thermal_parabolic_steps(efb, et):
    d = np.load(os.path.join(GOLD, f"transient_thermal_{et}.npz"))
    k, c, rho, dt, alpha = d["params"]
    g = efb.mesh.ElemGroup(et, d["connect"], d["coords"], all_nodes_used=True)
    ts = efb.transient.TransientSolve.thermal(efb.staggered.LocalSystem(g), k, rho * c)
    K, C, _, _ = thermal_system(d, et)
    assert rel(ts.K.to_scipy().toarray(), K.toarray()) < 1e-12 and rel(ts.C.to_scipy().toarray(), C.toarray()) < 1e-12
    ts.Solver_Set_Parabolic_Algorithm(dt, alpha)
    ts.pcg_tol = 1e-12
    for s in range(3):
        ts.bc = efb.staggered.Dirichlet(g.Ncoords)
        ts.bc.add(d["lo"], [0.0], [0], 1)
        ts.bc.add(d["hi"], [float(d["hi_values"][s])], [0], 1)
        ts.Solve()
        assert ts.info["converged"]
        assert rel(ts.u.cpu().numpy(), d[f"u_{s}"]) < 1e-8 and rel(ts.v.cpu().numpy(), d[f"v_{s}"]) < 1e-7


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "transient_elastic_*.npz"))), ids=os.path.basename)
def test_elastodynamic_steps(efb, path):
    name = os.path.basename(path)[len("transient_elastic_"):-4]
    et, algo = name.split("_", 1)
    d = np.load(path)
    E, v, rho, cM, cK, dt = d["params"]
    beta, gamma, alpha = d["scheme"]
    g = efb.mesh.ElemGroup(et, d["connect"], d["coords"], all_nodes_used=True)
    Nn = g.Ncoords
    Cmat = orc.IsoMaterial(3, E, v).C
    ts = efb.transient.TransientSolve.elastodynamic(efb.staggered.LocalSystem(g), Cmat, rho, coefM=cM, coefK=cK)
    K, C, M, _ = elastic_system(d, et)
    for ours, ref in ((ts.K, K), (ts.C, C), (ts.M, M)):
        assert rel(ours.to_scipy().toarray(), ref.toarray()) < 1e-12
    ts.Solver_Set_Hyperbolic_Algorithm(dt, algo, beta, gamma, alpha)
    ts.pcg_tol = 1e-12
    o = orc.TransientOracle(K, C, M)
    o.set_hyperbolic(dt, algo, beta, gamma, alpha)
    tol = 1e-5 if algo == "euler_explicit" else 1e-7  # the explicit run is unstable at this dt: errors are amplified
    for s in range(3):
        dofs, vals, F = elastic_bcs(d, s, Nn)
        ts.bc = efb.staggered.Dirichlet(Nn * 3)
        ts.bc.add(d["lo"], [0.0, 0.0, 0.0], [0, 1, 2], 3)
        ts.bc.add(d["hi"], [float(d["hi_values"][s])], [0], 3)
        ts.Solve(F)
        o.Solve(F, dofs, vals)
        assert ts.info["converged"], ts.info
        for ours, key, oref in ((ts.u, "u", o.u), (ts.v, "v", o.v), (ts.a, "a", o.a)):
            got = ours.cpu().numpy()
            assert rel(got, d[f"{key}_{s}"]) < tol, (algo, s, key, rel(got, d[f"{key}_{s}"]))
            assert rel(got, oref) < tol
