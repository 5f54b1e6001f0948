#!/usr/bin/env python
"""dev: the scalar operator kernels (UV, GradUGradV, damage system) alone — ms and GB/s of gather + output against the HBM copy peak.
EFB_SCALAR_BLOCK=1 forces the phase-structured block form, =-1 the warp-autonomous form (default: per-operator rule).    python scripts/scalar_probe.py [ELEM n]..."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import mesh, meshgen, operators  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402


def timeit(fn, reps=None):
    reps = int(os.environ.get("SCALAR_PROBE_REPS", "10")) if reps is None else reps
    fn()
    torch.cuda.synchronize()
    if reps <= 0:  # one launch per operator (ncu captures)
        return float("nan")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cases = sys.argv[1:] or ["HEXA8", "128", "TETRA4", "100", "TRI3", "1000", "QUAD4", "1000"]
    for et, n in zip(cases[0::2], cases[1::2]):
        coords, connect = meshgen.structured_mesh(et, int(n), jitter=0.2, seed=0)
        g = mesh.ElemGroup(et, connect, coords, all_nodes_used=True)
        dim, nPe, Ne = g.dim, g.nPe, g.Ne
        out = {"elem": et, "Ne": Ne, "form": {"1": "block", "-1": "warp"}.get(os.environ.get("EFB_SCALAR_BLOCK", ""), "auto")}
        gather = Ne * nPe * (4 + 8 * dim)
        Ms = dv.empty((Ne, nPe, nPe))
        Mv = dv.empty((Ne, nPe * dim, nPe * dim))
        for name, fn, nbytes in (
            ("UV_scalar", lambda: operators.mass_Me_dev(g, 2.0, 1, "mass", 1.0, out=Ms), gather + Ne * nPe * nPe * 8),
            ("UV_vector", lambda: operators.mass_Me_dev(g, 2.0, dim, "mass", 1.0, out=Mv), gather + Ne * (nPe * dim) ** 2 * 8),
            ("GradUGradV", lambda: operators.diffusion_Ke_dev(g, None, 1.5, "rigi", 1.0, out=Ms), gather + Ne * nPe * nPe * 8),
        ):
            t = timeit(fn)
            out[name] = {"ms": round(t, 4), "GBps": round(nbytes / t / 1e6, 1)}
        print(json.dumps(out), flush=True)
        del Ms, Mv, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
