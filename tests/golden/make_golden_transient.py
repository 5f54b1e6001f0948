"""Mint transient fixtures (BASELINE config 5) from the LIVE reference (EasyFEA v3.5.1 at /root/reference, gmsh stubbed).

Run in the authoring container only:  python tests/golden/make_golden_transient.py
Thermal parabolic steps (`Simulations.Thermal`, `Solver_Set_Parabolic_Algorithm`) and elastodynamic steps
(`Simulations.Elastic`, `Solver_Set_Hyperbolic_Algorithm` with every linear time scheme, Rayleigh damping, nodal loads) on
small jittered HEXA8 / HEXA27 meshes, default direct solver; u, v, a are stored after every `Solve()`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

import_reference()
from EasyFEA import AlgoType, Models, Simulations  # noqa: E402
from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh  # noqa: E402

from easyfea_b200 import meshgen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NSTEPS = 3


def build(et, n):
    lattice, connect = meshgen.structured_mesh(et, n)
    coords, _ = meshgen.structured_mesh(et, n, jitter=0.12, seed=7)
    g = GroupElemFactory.Create(ElemType(et), connect, coords)
    return Mesh({ElemType(et): g}), lattice, coords, connect


def thermal(et, n):
    mesh, lattice, coords, connect = build(et, n)
    k, c, rho, dt, alpha = 1.5, 2.0, 1.3, 0.1, 0.5
    simu = Simulations.Thermal(mesh, Models.Thermal(k=k, c=c))
    simu.rho = rho
    simu.Solver_Set_Parabolic_Algorithm(dt=dt, alpha=alpha)
    x = lattice[:, 0]
    lo, hi = np.flatnonzero(x < 1e-12), np.flatnonzero(x > x.max() - 1e-12)
    d = {"coords": coords, "connect": connect, "lo": lo, "hi": hi, "params": np.array([k, c, rho, dt, alpha]),
         "hi_values": np.array([40.0, 55.0, 30.0])}
    for s in range(NSTEPS):
        simu.Bc_Init()
        simu.add_dirichlet(lo, [0.0], ["t"])
        simu.add_dirichlet(hi, [float(d["hi_values"][s])], ["t"])
        simu.Solve()
        simu.Save_Iter()
        d[f"u_{s}"], d[f"v_{s}"] = np.array(simu.thermal), np.array(simu.thermalDot)
    np.savez_compressed(os.path.join(OUT, f"transient_thermal_{et}.npz"), **d)
    print("thermal", et, "max T", float(d[f"u_{NSTEPS - 1}"].max()))


def elastic(et, n, algo, **kw):
    mesh, lattice, coords, connect = build(et, n)
    E, v, rho, cM, cK, dt = 210000.0, 0.3, 2.0, 0.15, 2e-4, 0.05
    simu = Simulations.Elastic(mesh, Models.Elastic.Isotropic(3, E=E, v=v))
    simu.rho = rho
    simu.Set_Rayleigh_Damping_Coefs(coefM=cM, coefK=cK)
    simu.Solver_Set_Hyperbolic_Algorithm(dt=dt, algo=algo, **kw)
    x = lattice[:, 0]
    lo, hi = np.flatnonzero(x < 1e-12), np.flatnonzero(x > x.max() - 1e-12)
    mid = np.flatnonzero(np.abs(x - 0.5 * x.max()) < 0.26 * x.max())
    mid = np.setdiff1d(mid, np.concatenate([lo, hi]))
    d = {"coords": coords, "connect": connect, "lo": lo, "hi": hi, "mid": mid, "params": np.array([E, v, rho, cM, cK, dt]),
         "hi_values": np.array([0.01, 0.025, 0.02]), "load": np.array([3.0, -2.0, 1.0]),
         "scheme": np.array([kw.get("beta", 0.25), kw.get("gamma", 0.5), kw.get("alpha", 0.5)])}
    for s in range(NSTEPS):
        simu.Bc_Init()
        simu.add_dirichlet(lo, [0.0, 0.0, 0.0], ["x", "y", "z"])
        simu.add_dirichlet(hi, [float(d["hi_values"][s])], ["x"])
        simu.add_neumann(mid, [3.0 * (s + 1), -2.0, 1.0], ["x", "y", "z"])  # nodal loads, scaled per step in x
        simu.Solve()
        simu.Save_Iter()
        d[f"u_{s}"], d[f"v_{s}"], d[f"a_{s}"] = np.array(simu.displacement), np.array(simu.speed), np.array(simu.accel)
    np.savez_compressed(os.path.join(OUT, f"transient_elastic_{et}_{algo.name if hasattr(algo, 'name') else algo}.npz"), **d)
    print("elastic", et, algo, "max |u|", float(np.abs(d[f"u_{NSTEPS - 1}"]).max()), "max |a|", float(np.abs(d[f"a_{NSTEPS - 1}"]).max()))


if __name__ == "__main__":
    thermal("HEXA27", (3, 3, 2))
    thermal("HEXA8", (5, 4, 3))
    elastic("HEXA27", (2, 2, 2), AlgoType.newmark)
    elastic("HEXA8", (4, 3, 3), AlgoType.newmark)
    elastic("HEXA8", (4, 3, 3), AlgoType.midpoint)
    elastic("HEXA8", (4, 3, 3), AlgoType.hht, alpha=0.1)
    elastic("HEXA8", (4, 3, 3), AlgoType.hht_newmark, alpha=1 / 6)
    elastic("HEXA8", (4, 3, 3), AlgoType.euler_implicit)
    elastic("HEXA8", (4, 3, 3), AlgoType.euler_explicit)
