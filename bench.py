#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json config 2: 3D linear-elastic HEXA8 cube, element integration (K_e at 8
Gauss points per element) + deterministic CSR replay assembly, FP64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CELLS] [--impl reference]

One "step" = one pass of the hot path over the whole mesh: the K_e kernel over all elements, then the CSR replay.
`value` is Gauss-point evaluations per second with everything resident in HBM; `e2e` is the same metric measured
through the host-buffer boundary (pinned connect/coords in, CSR data out, copies inside the timed region).
`--impl reference` times the CPU path (the NumPy oracle port of the reference, oracle/easyfea_oracle.py) on the box's
host cores on a bounded sample of the same workload.
Multi-GPU (torchrun): the mesh is extended along z, one slab of n^3 owned elements (+ one ghost layer) per rank; no
data-path collective (weak scaling); time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "element Gauss-point evaluations/s (HEXA8 elastic K_e + CSR replay assembly, FP64)"
E_MOD, NU = 210000.0, 0.3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks line')."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# -----------------------------------------------------------------------------------------------------------------
def slab_mesh(n: int, rank: int, world: int):
    """Owned slab of n^3 HEXA8 cells of the n x n x (world*n) cube + one ghost layer of cells on each interior side."""
    from easyfea_b200 import meshgen

    lo = 1 if rank > 0 else 0
    hi = 1 if rank < world - 1 else 0
    nz = n + lo + hi
    coords, connect = meshgen.structured_mesh("HEXA8", (n, n, nz), lengths=(1.0, 1.0, nz / n), jitter=0.2, seed=rank)
    return coords, connect


class CpuSample:
    """The reference's CPU path (NumPy restatement, oracle/) on a bounded sample of the workload."""

    def __init__(self, n_sample: int):
        from easyfea_b200 import elements as el
        from easyfea_b200 import meshgen
        from oracle import easyfea_oracle as orc

        self.orc = orc
        self.coords, self.connect = meshgen.structured_mesh("HEXA8", n_sample, jitter=0.2, seed=0)
        self.tab = el.gauss_table("HEXA8", "rigi")
        self.C = orc.IsoMaterial(3, E_MOD, NU).C
        Nn = self.coords.shape[0]
        t0 = time.perf_counter()
        self.inv, _, _, self.nnz = orc.csr_map([self.connect], 3, Nn * 3, True)  # pattern: one-time, reported apart
        self.pattern_s = time.perf_counter() - t0
        self.units = self.connect.shape[0] * self.tab.nPg

    def step(self) -> float:
        """one pass: cold geometry + K_e einsum + bincount replay (what a first `Assembly` does) -> seconds"""
        orc = self.orc
        t0 = time.perf_counter()
        geo = orc.geometry(self.coords[self.connect], self.tab.dN_pg, self.tab.weights)
        Ke = orc.linearized_elasticity(geo, self.C)
        orc.assemble_replay([Ke], self.inv, self.nnz)
        return time.perf_counter() - t0


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all0 = time.perf_counter()
    n_sample = args.cpu_sample
    smp = CpuSample(n_sample)
    for _ in range(args.warmup):
        smp.step()
    secs = [smp.step() for _ in range(args.steps)]
    dt = float(np.mean(secs))
    value = smp.units / dt
    sample = f"HEXA8 {n_sample}^3 = {n_sample**3} elements x 8 Gauss points per step (same jittered-cube family)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE config 2: HEXA8 elastic cube, K_e (8 GP) + CSR replay, CPU sample of {n_sample}^3 "
                                   f"elements per step, E={E_MOD}, v={NU}"},
            "cpu_baseline": {"value": value, "unit": "GP/s", "cores": blas_threads(), "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count(), "pattern_build_s": smp.pattern_s},
            "e2e": {"value": value, "unit": "GP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from easyfea_b200 import _lib, assembly, mesh, operators
    from easyfea_b200 import device as dv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.require_cuda()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.n
    K, W = args.steps, args.warmup

    coords, connect = slab_mesh(n, rank, world)
    Ne, Nn = connect.shape[0], coords.shape[0]
    nPg, nPe, ndof = 8, 8, 24
    g = mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    C = np.asarray(_material_C())

    # ---- one-time: device mirror + CSR pattern ----
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dg = mesh.device_group(g)
    A = assembly.Assembler()
    pat = A.pattern(3, True, Nn * 3, (g,))
    torch.cuda.synchronize()
    t_pattern = time.perf_counter() - t0
    nnz = pat.nnz
    n_entries = Ne * ndof * ndof

    Ke = dv.empty((Ne, ndof, ndof))
    data = dv.empty((nnz,))
    Cd = np.ascontiguousarray(C)  # homogeneous C goes by value through the kernel arguments

    def step():
        operators.elastic_Ke_dev(g, Cd, "rigi", 1.0, out=Ke)
        pat.replay([Ke], out=data)

    for _ in range(max(W, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * K + 1)]
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        ev[0].record()
        for k in range(K):
            operators.elastic_Ke_dev(g, Cd, "rigi", 1.0, out=Ke)
            ev[3 * k + 1].record()
            pat.replay([Ke], out=data)
            ev[3 * k + 2].record()
            ev[3 * k + 3].record()
        torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[3 * K])
    t_ke = np.mean([ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(K)])
    t_rp = np.mean([ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(K)])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    # units processed: owned + ghost elements are all integrated (ghost work is real work of the sharded path)
    units = torch.tensor([Ne * nPg], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(units)
    value = float(units.item()) * K / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (per launch, this rank) ----
    peak, peak_src = measured_peaks()
    bytes_replay = n_entries * 8 + Ne * nPe * nPe * 4 + Ne * nPe * 8 + nnz * 8 + (Nn + 1) * 16
    bytes_ke = Ne * (nPe * (4 + 24) + ndof * ndof * 8)
    flops_ke = Ne * nPg * 4695
    if t_rp >= t_ke:
        roof = {"kernel": "k_replay_matrix", "bound": "hbm", "achieved": bytes_replay / (t_rp * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": None}
    else:
        roof = {"kernel": "k_elastic<3,8>", "bound": "hbm", "achieved": bytes_ke / (t_ke * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": None}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = peak_src

    line = {"metric": METRIC, "value": value, "unit": "GP/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config 2: HEXA8 elastic cube {n}^3 = {n**3} owned elements per GPU"
                                   f"{' (+ghost layer)' if world > 1 else ''}, K_e (8 GP) + CSR replay, E={E_MOD}, v={NU}",
                       "elements_per_gpu": Ne, "nodes_per_gpu": Nn, "nnz_per_gpu": nnz, "jitter": 0.2,
                       "l2": "inputs larger than L2 (K_e array >> 126 MB)" if n_entries * 8 > 4 * 126e6 else
                             "working set may fit L2: use --n >= 60"},
            "roofline": roof,
            "kernels": {"Ke_ms": float(t_ke), "Ke_GPps": Ne * nPg / (t_ke * 1e-3), "Ke_TFLOPs": flops_ke / (t_ke * 1e-3) / 1e12,
                        "Ke_GBps": bytes_ke / (t_ke * 1e-3) / 1e9, "replay_ms": float(t_rp),
                        "replay_GBps": bytes_replay / (t_rp * 1e-3) / 1e9,
                        "replay_GBps_survey_formula": (n_entries * 12 + nnz * 8) / (t_rp * 1e-3) / 1e9,
                        "pattern_build_s": t_pattern},
            "gpu_launches": 2 * K, "clocks": clk.summary()}

    if rank == 0:
        # ---- e2e through the host-buffer boundary (pinned in, pinned out) ----
        line["e2e"] = e2e_leg(args, g, coords, connect, C, pat, Ke, data, world)
        if world == 1 and not args.no_cpu:
            smp = CpuSample(args.cpu_sample)
            smp.step()
            secs = [smp.step() for _ in range(3)]
            line["cpu_baseline"] = {"value": smp.units / float(np.mean(secs)), "unit": "GP/s", "cores": blas_threads(),
                                    "kind": "port", "host_cores": os.cpu_count(), "pattern_build_s": smp.pattern_s,
                                    "sample": f"HEXA8 {args.cpu_sample}^3 = {args.cpu_sample**3} elements, cold geometry + K_e "
                                              f"einsum + bincount replay, mean of 3 passes of {np.mean(secs):.2f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _material_C():
    lam = E_MOD * NU / ((1 + NU) * (1 - 2 * NU))
    mu = E_MOD / (2 * (1 + NU))
    I = np.array([1.0, 1, 1, 0, 0, 0])
    return lam * np.outer(I, I) + 2 * mu * np.eye(6)


def e2e_leg(args, g, coords, connect, C, pat, Ke, data, world):
    """Same step through host buffers: pinned (connect int32, coords) -> device, K_e + replay, CSR data -> pinned host."""
    import torch

    from easyfea_b200 import mesh, operators

    steps = max(1, min(args.steps, args.e2e_steps))
    h_conn = torch.from_numpy(connect.astype(np.int32)).pin_memory()
    h_coord = torch.from_numpy(coords).pin_memory()
    h_out = torch.empty(data.numel(), dtype=torch.float64).pin_memory()
    dg = mesh.device_group(g)
    Cd = np.ascontiguousarray(C)

    def one():
        dg.connect.copy_(h_conn.view_as(dg.connect), non_blocking=True)  # the element kernel reads these buffers
        dg.coord.copy_(h_coord, non_blocking=True)
        operators.elastic_Ke_dev(g, Cd, "rigi", 1.0, out=Ke)
        pat.replay([Ke], out=data)
        h_out.copy_(data, non_blocking=True)

    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    chk = float(h_out[:1000].sum())  # the result is really on the host
    return {"value": connect.shape[0] * 8 * world / dt, "unit": "GP/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(h_conn.numel() * 4 + h_coord.numel() * 8), "d2h_bytes_per_step": int(h_out.numel() * 8),
            "steps": steps, "checksum_head": chk,
            "note": "rank 0's slab timed alone; value scaled by n_gpus (shards are independent)" if world > 1 else
                    "pinned host buffers; H2D of connectivity+coordinates and D2H of the CSR data inside the timed region"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=200, help="cells per side of the owned HEXA8 cube (200 -> 8.0 M elements)")
    ap.add_argument("--cpu-sample", type=int, default=32, help="cells per side of the CPU sample (32 -> 32 768 elements)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
