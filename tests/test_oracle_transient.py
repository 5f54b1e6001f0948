"""The NumPy restatement of the time-scheme system build (oracle.TransientOracle) against fixtures minted from the live
reference (tests/golden/make_golden_transient.py): thermal parabolic steps and elastodynamic steps with every linear scheme."""
import glob
import os

import numpy as np
import pytest

from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def csr(Xe, connect, dof_n, Nn):
    import scipy.sparse as sp

    inv, indices, indptr, nnz = orc.csr_map([connect], dof_n, Nn * dof_n, True)
    return sp.csr_matrix((orc.assemble_replay([Xe], inv, nnz), indices, indptr), shape=(Nn * dof_n, Nn * dof_n))


def thermal_system(d, et):
    k, c, rho, dt, alpha = d["params"]
    coords, connect = d["coords"], d["connect"]
    Nn = coords.shape[0]
    tr, tm = el.gauss_table(et, "rigi"), el.gauss_table(et, "mass")
    Ke = orc.grad_u_a_grad_v(orc.geometry(coords[connect], tr.dN_pg, tr.weights), None, k)
    Ce = orc.uv(orc.geometry(coords[connect], tm.dN_pg, tm.weights), tm.N_pg, rho * c, 1)
    return csr(Ke, connect, 1, Nn), csr(Ce, connect, 1, Nn), dt, alpha


def elastic_system(d, et):
    E, v, rho, cM, cK, dt = d["params"]
    coords, connect = d["coords"], d["connect"]
    Nn = coords.shape[0]
    tr, tm = el.gauss_table(et, "rigi"), el.gauss_table(et, "mass")
    Ke = orc.linearized_elasticity(orc.geometry(coords[connect], tr.dN_pg, tr.weights), orc.IsoMaterial(3, E, v).C)
    Me = orc.uv(orc.geometry(coords[connect], tm.dN_pg, tm.weights), tm.N_pg, rho, 3)
    K, M = csr(Ke, connect, 3, Nn), csr(Me, connect, 3, Nn)
    return K, cK * K + cM * M, M, dt


def elastic_bcs(d, s, Nn):
    lo, hi, mid = d["lo"], d["hi"], d["mid"]
    dofs = np.concatenate([lo * 3, lo * 3 + 1, lo * 3 + 2, hi * 3])
    vals = np.concatenate([np.zeros(3 * lo.size), np.full(hi.size, d["hi_values"][s])])
    F = np.zeros(Nn * 3)  # `add_neumann` spreads the given total over the nodes (`__Bc_pointLoad`, _simu.py:2759-2762)
    F[mid * 3], F[mid * 3 + 1], F[mid * 3 + 2] = 3.0 * (s + 1) / mid.size, -2.0 / mid.size, 1.0 / mid.size
    return dofs, vals, F


@pytest.mark.parametrize("et", ["HEXA27", "HEXA8"])
def test_thermal_parabolic_steps(et):
    d = np.load(os.path.join(GOLD, f"transient_thermal_{et}.npz"))
    K, C, dt, alpha = thermal_system(d, et)
    o = orc.TransientOracle(K, C)
    o.set_parabolic(dt, alpha)
    dofs = np.concatenate([d["lo"], d["hi"]])
    for s in range(3):
        vals = np.concatenate([np.zeros(d["lo"].size), np.full(d["hi"].size, d["hi_values"][s])])
        o.Solve(np.zeros(K.shape[0]), dofs, vals)
        assert rel(o.u, d[f"u_{s}"]) < 1e-11 and rel(o.v, d[f"v_{s}"]) < 1e-10


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "transient_elastic_*.npz"))), ids=os.path.basename)
def test_elastodynamic_steps(path):
    name = os.path.basename(path)[len("transient_elastic_"):-4]
    et, algo = name.split("_", 1)
    d = np.load(path)
    K, C, M, dt = elastic_system(d, et)
    beta, gamma, alpha = d["scheme"]
    o = orc.TransientOracle(K, C, M)
    o.set_hyperbolic(dt, algo, beta, gamma, alpha)
    Nn = d["coords"].shape[0]
    for s in range(3):
        dofs, vals, F = elastic_bcs(d, s, Nn)
        o.Solve(F, dofs, vals)
        tol = 1e-7 if algo == "euler_explicit" else 1e-9  # the explicit run is unstable at this dt: errors are amplified
        assert rel(o.u, d[f"u_{s}"]) < tol, (s, rel(o.u, d[f"u_{s}"]))
        assert rel(o.v, d[f"v_{s}"]) < tol and rel(o.a, d[f"a_{s}"]) < tol
