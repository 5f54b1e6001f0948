#!/usr/bin/env python
"""Timing probe of the fused assembly kernel against the two-kernel path (dev tool, run on the GPU box).
    python scripts/fused_probe.py [--n 128] [--S 8,16,32] [--reps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import assembly, mesh, meshgen, operators  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--S", default="8,16,32")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--elem", default="HEXA8")
    ap.add_argument("--no-two", action="store_true")
    ap.add_argument("--no-mma", action="store_true")
    args = ap.parse_args()
    et = args.elem
    coords, connect = meshgen.structured_mesh(et, args.n, jitter=0.2, seed=0)
    g = mesh.ElemGroup(et, connect, coords, all_nodes_used=True)
    dim = g.dim
    Nn = coords.shape[0]
    lam, mu = 210000.0 * 0.3 / (1.3 * 0.4), 210000.0 / 2.6
    ns = 3 if dim == 2 else 6
    I = np.zeros(ns)
    I[:dim] = 1
    C = lam * np.outer(I, I) + 2 * mu * np.eye(ns)
    pat = assembly.Assembler().pattern(dim, True, Nn * dim, (g,))
    out = {"elem": et, "n": args.n, "Ne": g.Ne, "Nn": Nn, "nnz": pat.nnz}
    data = dv.empty((pat.nnz,))
    ref = None
    if not args.no_two:
        ndof = g.nPe * dim
        Ke = dv.empty((g.Ne, ndof, ndof))
        out["two_kernel_ms"] = timed(lambda: (operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke), pat.replay([Ke], out=data)), args.reps)
        ref = data.clone()
        del Ke
    for S in [int(s) for s in args.S.split(",") if s]:
        try:
            torch.cuda.synchronize()
            import time

            t0 = time.perf_counter()
            sched = assembly.FusedSchedule(pat.graph, S=S)
            torch.cuda.synchronize()
            t_sched = time.perf_counter() - t0
            smem = assembly.smem_bytes(dim, g.nPe, sched.nPg, S, sched.cap_e, sched.max_deg)
            ms = timed(lambda: assembly.assemble_elastic_fused(sched, C, out=data), args.reps)
            err = float((data - ref).norm() / ref.norm()) if ref is not None else None
            out[f"fused_S{S}"] = {"ms": ms, "cap_e": sched.cap_e, "redundancy": sched.redundancy(), "smem": smem,
                                  "n_clusters": sched.n_clusters, "schedule_s": t_sched, "rel_err_vs_two_kernel": err}
            del sched
        except Exception as exc:
            out[f"fused_S{S}"] = {"error": repr(exc)[:300]}
    if et == "HEXA8" and not args.no_mma:
        import time

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms_ = assembly.MmaSchedule(pat.graph)
        torch.cuda.synchronize()
        t_sched = time.perf_counter() - t0
        t = timed(lambda: assembly.assemble_elastic_mma(ms_, C, out=data), args.reps)
        err = float((data - ref).norm() / ref.norm()) if ref is not None else None
        out["mma"] = {"ms": t, "t_cap": ms_.t_cap, "cap4": ms_.cap4, "rmax": ms_.rmax, "redundancy": ms_.redundancy(), "smem": ms_.smem_bytes(), "rec_words": ms_.rec_words, "pw_max": ms_.pw_max,
                      "n_clusters": ms_.n_clusters, "rounds_per_cluster": ms_.n_rounds / max(ms_.n_clusters, 1),
                      "row_tiles_per_element": ms_.n_row_tiles / g.Ne, "prog_bytes": ms_.prog.numel() * 4,
                      "schedule_s": t_sched, "rel_err_vs_two_kernel": err}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
