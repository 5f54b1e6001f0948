#!/usr/bin/env python
"""dev: Chebyshev-Jacobi PCG on the phase-field configs (bench.phase_field_leg set-up) — iterations and seconds per staggered
iteration for several polynomial degrees, power-iteration estimates of lmax.   python scripts/cheb_probe.py [--cfg 3] [--n 300]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import mesh, meshgen, phasefield, solver, staggered  # noqa: E402


def build(cfg, n):
    cfg4 = cfg == 4
    et, dim, split = ("TETRA4", 3, "He") if cfg4 else ("TRI3", 2, "Miehe")
    L, l0 = 1e-3, 1e-5 * max(1.0, 1000.0 / n)
    lattice, connect = meshgen.structured_mesh(et, n, lengths=(L,) * dim)
    coords, _ = meshgen.structured_mesh(et, n, lengths=(L,) * dim, jitter=0.15, seed=1)
    g = mesh.ElemGroup(et, connect, coords, all_nodes_used=True)
    sysm = staggered.LocalSystem(g)
    pfm = phasefield.PhaseFieldModel(phasefield.IsotropicMaterial(dim, 210e9, 0.3, planeStress=False, thickness=1.0), split, "AT2", 2.7e3, l0)
    simu = staggered.PhaseFieldStaggered(sysm, pfm, pcg_tol=1e-8, pcg_maxiter=30000)
    ix, iy = np.rint(lattice[:, 0] / L * n).astype(np.int64), np.rint(lattice[:, 1] / L * n).astype(np.int64)
    loc = np.arange(coords.shape[0])
    comps = list(range(dim))
    simu.add_dirichlet(loc[(iy == n // 2) & (ix <= n // 2)], [1], [0], problemType="damage")
    simu.add_dirichlet(loc[iy == n], [8e-6, 4e-6, 0.0][:dim], comps)
    simu.add_dirichlet(loc[iy == 0], [0.0] * dim, comps)
    return simu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--degrees", default="1,4")
    ap.add_argument("--safety", default="1.25")
    ap.add_argument("--fp32", default="1")
    args = ap.parse_args()
    for deg in [int(v) for v in args.degrees.split(",")]:
        for saf in [float(v) for v in args.safety.split(",")] if deg > 1 else [1.1]:
            solver.CHEB_SAFETY = saf
            solver.CHEB_FP32 = bool(int(args.fp32))
            simu = build(args.cfg, args.n)
            simu.pcg_precond_degree = deg
            out = {"cfg": args.cfg, "n": args.n, "degree": deg, "safety": saf, "fp32": solver.CHEB_FP32, "iters": []}
            for k in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                simu.iterate()
                torch.cuda.synchronize()
                out["iters"].append({"s": round(time.perf_counter() - t0, 4), "damage": simu.info["damage"]["iterations"],
                                     "elastic": simu.info["elastic"]["iterations"], "conv": bool(simu.info["elastic"]["converged"]),
                                     "lmax": simu.info["elastic"].get("lmax")})
            print(json.dumps(out), flush=True)
            del simu
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
