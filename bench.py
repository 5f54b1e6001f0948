#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json config 2: 3D linear-elastic HEXA8 cube, element integration (K_e at 8
Gauss points per element) + deterministic CSR replay assembly, FP64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n CELLS] [--impl reference]

One "step" = one pass of the hot path over the whole mesh: element integration at 8 Gauss points per element + assembly of
the owned CSR rows, by the fused kernel `efb_assemble_elastic` (K_e never materialised); the two-kernel path (K_e kernel,
then the deterministic CSR replay) is timed beside it (`extras.two_kernel`).
`value` is Gauss-point evaluations per second with everything resident in HBM; `e2e` is the same metric measured
through the host-buffer boundary (pinned connect/coords in, CSR data out, copies inside the timed region) on every rank at
once (max over ranks).
`--impl reference` times the CPU path — the UNMODIFIED reference (EasyFEA, offline install under baseline/_ref, its own
`Operators.Bilinear.LinearizedElasticity` + `_Simu.__Assemble_csr`) when it is importable, else the NumPy oracle port of it
(oracle/easyfea_oracle.py), stated in `cpu_baseline.kind` — on the box's host cores on a bounded sample of the same workload.
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

METRIC = "element Gauss-point evaluations/s (HEXA8 elastic element integration + CSR assembly, FP64)"
E_MOD, NU = 210000.0, 0.3


def ncu_traffic(kernel: str, units: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/traffic.json), scaled from
    the captured launch to `units` (elements / nodes) of the launch timed here; None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)[kernel]
        return (t["read_bytes"] + t["write_bytes"]) / t["units"] * units
    except Exception:
        return None


_STDOUT_FD = None


def quiet_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries write there too (NCCL prints its version banner on the first
    communicator, whatever NCCL_DEBUG says in some builds): file descriptor 1 is pointed at stderr for the whole run and the
    line is written to a saved copy of the original descriptor by `emit`."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_STDOUT_FD, data)


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
def slab_system(n: int, rank: int, world: int):
    """Rank `rank`'s shard of the n x n x (world*n) HEXA8 cube (z-slabs): its own n^3 elements + the
    ghost layer touching its owned nodes, in local numbering [owned | halo].  Returns (group, partition)."""
    from easyfea_b200 import dist as efd
    from easyfea_b200 import mesh, meshgen

    connect, elem_ids, owner_of, coords_of = meshgen.hexa8_slab(n, rank, world, jitter=0.2, seed=0)
    part = efd.Partition.from_candidates(connect, elem_ids, owner_of, rank, world,
                                         own_chunk=(rank * n**3, (rank + 1) * n**3))
    g = mesh.ElemGroup("HEXA8", part.connect, coords_of(part.nodes), all_nodes_used=True)
    return g, part


def workload_config(n: int, world: int):
    """`config` of BOTH arms (the reference arm times a bounded sample of this workload, described in its cpu_baseline.sample)"""
    n_own = n**3
    return {"workload": f"BASELINE config 2: HEXA8 elastic cube, {n}^3 = {n_own} elements per GPU"
                        f"{' (z-slabs of one ' + str(n) + 'x' + str(n) + 'x' + str(world * n) + ' mesh, + ghost layer)' if world > 1 else ''}"
                        f", element integration (8 Gauss points) + assembly of the owned CSR rows, E={E_MOD}, v={NU}",
            "elements_per_gpu": n_own, "jitter": 0.2,
            "l2": "inputs and outputs larger than L2 (CSR data >> 126 MB)" if n >= 60 else "working set may fit L2: use --n >= 60"}


class CpuSample:
    """The CPU path on a bounded sample of the workload: the UNMODIFIED reference (kind "reference": EasyFEA from
    baseline/_ref, `Operators.Bilinear.LinearizedElasticity` on a group with cold caches + `_Simu.__Assemble_csr`), or its
    NumPy restatement (kind "port", oracle/easyfea_oracle.py) when the reference cannot be imported."""

    def __init__(self, n_sample: int, kind: str):
        from easyfea_b200 import elements as el
        from easyfea_b200 import meshgen

        self.kind = kind
        t0 = time.perf_counter()
        if kind == "reference":
            from oracle.ref_cases import hexa8_elastic
            from oracle.ref_import import import_reference

            E = import_reference(travel_only=True)
            self.simu, self.g, self.C, self.Ndof, _, connect = hexa8_elastic(E, n_sample, E_MOD, NU)
            self.Ops = E.FEM.Operators
            self.simu._Simu__Get_csr_map(3, True, self.Ndof, (self.g,))  # pattern: one-time, cached by the reference
            self.units = connect.shape[0] * 8
        else:
            from oracle import easyfea_oracle as orc

            self.orc = orc
            self.coords, self.connect = meshgen.structured_mesh("HEXA8", n_sample, jitter=0.2, seed=0)
            self.tab = el.gauss_table("HEXA8", "rigi")
            self.C = orc.IsoMaterial(3, E_MOD, NU).C
            Nn = self.coords.shape[0]
            self.inv, _, _, self.nnz = orc.csr_map([self.connect], 3, Nn * 3, True)
            self.units = self.connect.shape[0] * self.tab.nPg
        self.pattern_s = time.perf_counter() - t0

    def step(self) -> float:
        """one pass: cold geometry + K_e einsum + bincount replay (what the GPU step recomputes every time) -> seconds"""
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.g._InitMatrix()  # drop the cached B / wJ.B^T of the previous pass: geometry is part of the step
            Ke = self.Ops.Bilinear.LinearizedElasticity(self.g, self.C)
            self.simu._Simu__Assemble_csr({self.g: Ke}, 3, self.Ndof, True)
        else:
            orc = self.orc
            geo = orc.geometry(self.coords[self.connect], self.tab.dN_pg, self.tab.weights)
            Ke = orc.linearized_elasticity(geo, self.C)
            orc.assemble_replay([Ke], self.inv, self.nnz)
        return time.perf_counter() - t0


def cpu_kind() -> str:
    """"reference" when the unmodified reference is importable from baseline/_ref (it travels with the snapshot), else "port" """
    try:
        from oracle.ref_import import reference_available

        return "reference" if reference_available(travel_only=True) else "port"
    except Exception:
        return "port"


def _cpu_worker(kind, n_sample, warm, passes, barrier, out):
    """one host process of the CPU arm: its own sample mesh, `passes` timed steps after a common barrier"""
    try:
        smp = CpuSample(n_sample, kind)
        for _ in range(warm):
            smp.step()
        barrier.wait(timeout=900)
        t0 = time.perf_counter()
        for _ in range(passes):
            smp.step()
        out.put((time.perf_counter() - t0, smp.units, smp.pattern_s))
    except Exception as exc:  # pragma: no cover
        out.put((float("nan"), 0, repr(exc)))
        try:
            barrier.abort()
        except Exception:
            pass


def cpu_arm(n_sample: int, warm: int, passes: int, procs: int = 0, kind: str = None):
    """The CPU path on ALL host cores: `procs` processes (default: one per core), each integrating and assembling its own
    n_sample^3-element chunk — what an MPI-partitioned EasyFEA run does for this path (docs/howto/use_mpi.md); NumPy's
    einsum/bincount are single-threaded.  Returns (GP/s over all processes, seconds per step, processes, pattern seconds, kind)."""
    import multiprocessing as mp

    kind = kind or cpu_kind()
    procs = procs or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context
    barrier, out = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_cpu_worker, args=(kind, n_sample, warm, passes, barrier, out)) for _ in range(procs)]
    for p in ps:
        p.start()
    res = [out.get(timeout=1800) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    bad = [r for r in res if not r[1]]
    if bad:
        if kind == "reference":  # the live reference failed on this box: say so and time the port instead
            sys.stderr.write(f"bench: reference arm failed ({bad[0][2]}); timing the NumPy port\n")
            return cpu_arm(n_sample, warm, passes, procs, "port")
        raise RuntimeError(f"CPU arm worker failed: {bad[0][2]}")
    wall = max(r[0] for r in res)
    return sum(r[1] for r in res) * passes / wall, wall / passes, procs, float(np.mean([r[2] for r in res])), kind


def cpu_sample_text(kind, procs, n_sample, dt, passes):
    what = ("the unmodified reference (EasyFEA 3.5.1, baseline/_ref): Operators.Bilinear.LinearizedElasticity with cold group caches + "
            "_Simu.__Assemble_csr" if kind == "reference" else "NumPy port of the reference (oracle/): cold geometry + K_e einsum + bincount replay")
    return (f"{procs} host processes x HEXA8 {n_sample}^3 = {n_sample**3} elements x 8 Gauss points per step, {what}; {passes} timed "
            f"passes of {dt:.2f} s each (one process per core, like an MPI-partitioned run of the reference)")


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
    value, dt, procs, pattern_s, kind = cpu_arm(n_sample, max(args.warmup, 1), args.steps, args.cpu_procs, args.cpu_kind or None)
    units_step = procs * n_sample**3 * 8
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n, args.gpus),
            "cpu_baseline": {"value": value, "unit": "GP/s", "cores": procs, "kind": kind,
                             "sample": cpu_sample_text(kind, procs, n_sample, dt, args.steps),
                             "host_cores": os.cpu_count(), "pattern_build_s": pattern_s, "units_per_step": units_step},
            "e2e": {"value": value, "unit": "GP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    emit(line)


# -----------------------------------------------------------------------------------------------------------------
FP64_PEAK_TFLOPS = 36.9  # measured on this pool with scripts/micro/dfma_peak.cu (pure DFMA, 32 warps/SM); nominal 37.2
FP64_DMMA_PEAK_TFLOPS = 37.1  # measured on this pool with scripts/micro/dmma_peak.cu (mma.sync m8n8k4 f64: same pipe, same rate as DFMA)
FLOPS_PER_GP = 4695      # HEXA8 structure-aware count of the general contraction, SURVEY.md section 8d
FLOPS_PER_GP_EXECUTED = 2150  # what k_elastic_w<3,8,sym,ortho> issues: 8 lanes x (15 DMUL + 105 DFMA) + geometry (~380)


def run_ours(args):
    import torch

    from easyfea_b200 import _lib, assembly, mesh, operators
    from easyfea_b200 import device as dv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.require_cuda()
    dist = None
    if world > 1:
        import torch.distributed as dist

        # stdout carries exactly ONE JSON line: NCCL's own "NCCL version ..." banner (NCCL_DEBUG=VERSION) goes to stdout too
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.n
    K, W = args.steps, max(args.warmup, 3)

    g, part = slab_system(n, rank, world)
    Ne, Nn = g.Ne, g.Ncoords
    n_own = n**3
    nPg, nPe, ndof = 8, 8, 24
    C = np.ascontiguousarray(_material_C())  # homogeneous C goes by value through the kernel arguments

    # ---- one-time: device mirror, CSR pattern of the local nodes (owned rows are a prefix), node schedule of the fused kernel ----
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mesh.device_group(g)
    A = assembly.Assembler()
    pat = A.pattern(3, True, Nn * 3, (g,))
    torch.cuda.synchronize()
    t_pattern = time.perf_counter() - t0
    t0 = time.perf_counter()
    sched = assembly.FusedSchedule(pat.graph, n_nodes=part.n_owned)
    torch.cuda.synchronize()
    t_sched = time.perf_counter() - t0
    t0 = time.perf_counter()
    msched = assembly.MmaSchedule(pat.graph, n_nodes=part.n_owned, base=sched if sched.S == 16 else None)
    torch.cuda.synchronize()
    t_msched = time.perf_counter() - t0
    if not msched.fits():
        msched = None
    nnz_owned = int(pat.indptr[part.n_owned * 3].item())
    n_entries = Ne * ndof * ndof
    data = dv.empty((pat.nnz,))

    def step_fused():
        assembly.assemble_elastic_fused(sched, C, "rigi", 1.0, out=data)

    def step_mma():
        assembly.assemble_elastic_mma(msched, C, "rigi", 1.0, out=data)

    Ke_box = {}

    def step_two():
        if "Ke" not in Ke_box:
            Ke_box["Ke"] = dv.empty((Ne, ndof, ndof))
        operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke_box["Ke"])
        pat.replay([Ke_box["Ke"]], out=data, n_nodes=part.n_owned)

    def time_steps(fn, k, split=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        ev[0].record()
        for i in range(k):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[k]), [ev[i].elapsed_time(ev[i + 1]) for i in range(k)]

    # which path is the step: --path mma | fused | two | auto (auto: the fastest on this mesh, decided in the warm-up)
    path = args.path
    paths = {"fused": step_fused, "two": step_two}
    if msched is not None:
        paths["mma"] = step_mma
    elif path == "mma":
        raise SystemExit("--path mma: the mesh does not fit efb_assemble_elastic_mma")
    for _ in range(2):
        for fn in paths.values():
            fn()
    torch.cuda.synchronize()
    probe = {name: time_steps(fn, 3)[0] / 3 for name, fn in paths.items()}
    if path == "auto":
        names = sorted(paths)
        path = min(probe, key=probe.get)
        if world > 1:  # every rank takes the same decision: rank 0's
            t = torch.tensor([names.index(path)], device=dev)
            dist.broadcast(t, 0)
            path = names[int(t.item())]
    step = paths[path]
    single = path in ("fused", "mma")  # one kernel per step

    # the clock sampler starts before the warm-up (nvidia-smi needs ~1 s to come up on an 8-GPU box) and keeps sampling
    # through warm-up + timed steps: the same kernels, the same load
    with ClockSampler(local_rank) as clk:
        t_w = time.perf_counter()
        n_w = 0
        while n_w < W or (time.perf_counter() - t_w < 1.5 and n_w < 200):  # >= W warm-up steps, and until the sampler has samples
            step()
            n_w += 1
            if n_w >= W:
                torch.cuda.synchronize()
                if clk.lines:
                    break
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if single:
            total_ms, per = time_steps(step, K)
            t_single = float(np.mean(per))
        else:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
            ev[0].record()
            for k in range(K):
                operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke_box["Ke"])
                ev[2 * k + 1].record()
                pat.replay([Ke_box["Ke"]], out=data, n_nodes=part.n_owned)
                ev[2 * k + 2].record()
            torch.cuda.synchronize()
            total_ms = ev[0].elapsed_time(ev[2 * K])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    # units processed = elements of the ranks' OWN chunks (ghost elements are integrated too, but not counted twice)
    value = float(world * n_own * nPg) * K / (total_ms * 1e-3)

    # ---- the other paths, timed beside the headline (fewer steps) ----
    k_other = max(2, min(K, 5))
    step_two()
    torch.cuda.synchronize()
    if single:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * k_other + 1)]
        ev[0].record()
        for k in range(k_other):
            operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke_box["Ke"])
            ev[2 * k + 1].record()
            pat.replay([Ke_box["Ke"]], out=data, n_nodes=part.n_owned)
            ev[2 * k + 2].record()
        torch.cuda.synchronize()
        Kt = k_other
    else:
        Kt = K
    t_fused = t_single if path == "fused" else time_steps(step_fused, k_other)[0] / k_other
    t_mma = None
    if msched is not None:
        t_mma = t_single if path == "mma" else time_steps(step_mma, k_other)[0] / k_other
    t_ke = float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(Kt)]))
    t_rp = float(np.mean([ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(Kt)]))
    Ke_box.clear()  # 36.9 GB at 8 M elements: the fused paths never allocate it
    (step_mma if path == "mma" else step_fused)()    # `data` = the fused assembly for the legs below
    torch.cuda.synchronize()

    # ---- roofline of the dominant kernel (per launch, this rank) ----
    peak, peak_src = measured_peaks()
    bytes_replay = n_entries * 8 + Ne * nPe * nPe * 4 + Ne * nPe * 8 + nnz_owned * 8 + (part.n_owned + 1) * 16
    bytes_ke = Ne * (nPe * (4 + 24) + ndof * ndof * 8)
    # fused kernel: CSR data once + schedule (4 B descriptor + nPe slot positions per (node, element) pair, the clusters'
    # connectivity slabs, the node records) + the coordinates of the mesh once
    bytes_fused = (nnz_owned * 8 + sched.n_tasks * (4 + nPe * 4) + sched.cl_conn.numel() * 4 + sched.cl_nodes.numel() * 8 + Nn * 24)
    ke_tflops_exec = Ne * nPg * FLOPS_PER_GP_EXECUTED / (t_ke * 1e-3) / 1e12
    # executed FP64 work of the fused kernel per Gauss point of an OWNED element: 64 node-pair blocks x 9 FMA + geometry of
    # `redundancy` elements (~330 flop each: F, det, inverse, gradients)
    fused_flops_gp = 2 * 64 * 9 + sched.redundancy() * 330
    # MMA form: algorithmic FP64 work per element = 8 Gauss points x (64 node-pair blocks x 9 FMA + 204 FMA of geometry: F, det,
    # inverse, gradients); executed = the DMMA instructions the schedule issues (6 per needed row tile, 12 per pass of 4
    # elements, 512 flop each) + ~200 scalar FP64 instructions per lane and pass
    mma_flops_alg = n_own * nPg * (64 * 9 + 204) * 2.0
    mma_flops_exec = None
    bytes_mma = None
    if msched is not None:
        n_pass = msched.n_clusters * (msched.cap4 // 4)
        mma_flops_exec = (msched.n_row_tiles * 6 + n_pass * 12) * 512.0 + n_pass * 200 * 64.0
        bytes_mma = nnz_owned * 8 + msched.recs.numel() * 4 + msched.prog.numel() * 4 + Nn * 24
    if path == "mma":
        roof = {"kernel": "k_assemble_hexa8_mma<ortho>", "bound": "tensor", "achieved": mma_flops_alg / (t_mma * 1e-3) / 1e12,
                "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "traffic": ncu_traffic("k_assemble_hexa8_mma<1>", part.n_owned),
                "executed_TFLOPs": mma_flops_exec / (t_mma * 1e-3) / 1e12,
                "hbm_GBps": bytes_mma / (t_mma * 1e-3) / 1e9, "hbm_frac": bytes_mma / (t_mma * 1e-3) / 1e9 / peak,
                "note": "fused integration + assembly on the FP64 MMA instruction (DMMA.8x8x4, FP64 pipe): peak = the DMMA rate measured "
                        "on this pool's B200 (scripts/micro/dmma_peak.cu: 37.1 TFLOP/s, the same as DFMA); achieved = ALGORITHMIC flops "
                        "(12 480 per element); the node-owned form executes ~2.4x that (elements are integrated once per cluster that "
                        "touches them) and is bound by latency at 16 warps per SM (DESIGN.md section 4.7)"}
    elif path == "fused":
        roof = {"kernel": "k_assemble_elastic<3,8,8,ortho>", "bound": "hbm", "achieved": bytes_fused / (t_fused * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": ncu_traffic("k_assemble_elastic<3,8,8,1>", part.n_owned),
                "note": "fused integration + assembly: DRAM traffic is the CSR data once; the kernel is bound by instruction issue / "
                        "shared-memory bandwidth at 16 warps per SM, not by HBM (DESIGN.md section 4.7)"}
    elif t_rp >= t_ke:
        roof = {"kernel": "k_replay_tma<3,8>", "bound": "hbm", "achieved": bytes_replay / (t_rp * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": ncu_traffic("k_replay_tma<3,8>", part.n_owned)}
    else:
        roof = {"kernel": "k_elastic_w<3,8,sym,ortho>", "bound": "hbm", "achieved": bytes_ke / (t_ke * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "traffic": ncu_traffic("k_elastic_w<3,8,1,1>", Ne)}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = peak_src if roof["bound"] == "hbm" else "measured FP64 DMMA rate (scripts/micro/dmma_peak.cu, profiles/r2_dmma_peak.log)"

    cfg = workload_config(n, world)
    line = {"metric": METRIC, "value": value, "unit": "GP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "config_detail": {"path": path, "path_probe_ms": probe, "elements_integrated_rank0": Ne, "nodes_rank0": Nn,
                              "nnz_owned_rank0": nnz_owned, "fused_cluster_nodes": sched.S, "fused_clusters": sched.n_clusters,
                              "fused_elements_per_cluster_max": sched.cap_e, "fused_geometry_redundancy": sched.redundancy()},
            "roofline": roof,
            "kernels": {"fused_kernel": "k_assemble_elastic<3,8,8,ortho> (K_e never materialised; outer-product form, Dt applied per CSR block)",
                        "fused_ms": t_fused, "fused_GPps": n_own * nPg / (t_fused * 1e-3), "fused_GBps": bytes_fused / (t_fused * 1e-3) / 1e9,
                        "fused_hbm_frac": bytes_fused / (t_fused * 1e-3) / 1e9 / peak,
                        "fused_TFLOPs_executed": n_own * nPg * fused_flops_gp / (t_fused * 1e-3) / 1e12,
                        "fused_fp64_frac_executed": n_own * nPg * fused_flops_gp / (t_fused * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                        "fused_bytes_per_launch": bytes_fused, "two_kernel_bytes_per_launch": bytes_ke + bytes_replay,
                        "two_kernel_ms": t_ke + t_rp,
                        "Ke_kernel": "k_elastic_w<3,8,sym,ortho> (isotropic C: symmetric + structural-zero form)",
                        "Ke_ms": t_ke, "Ke_GPps": Ne * nPg / (t_ke * 1e-3), "Ke_GBps": bytes_ke / (t_ke * 1e-3) / 1e9,
                        "Ke_hbm_frac": bytes_ke / (t_ke * 1e-3) / 1e9 / peak,
                        "Ke_TFLOPs_executed": ke_tflops_exec,
                        "Ke_fp64_frac_executed": ke_tflops_exec / FP64_PEAK_TFLOPS, "fp64_peak_tflops": FP64_PEAK_TFLOPS,
                        "replay_kernel": "k_replay_tma<3,8>",
                        "replay_ms": t_rp, "replay_GBps": bytes_replay / (t_rp * 1e-3) / 1e9,
                        "replay_hbm_frac": bytes_replay / (t_rp * 1e-3) / 1e9 / peak,
                        "replay_GBps_survey_formula": (n_entries * 12 + nnz_owned * 8) / (t_rp * 1e-3) / 1e9,
                        "mma_kernel": "k_assemble_hexa8_mma<ortho> (T_e = G^T G as FP64 m8n8k4 MMAs, staged T rows + gather programs, "
                                      "warp-specialised persistent CTA)",
                        "mma_ms": t_mma, "mma_GPps": None if t_mma is None else n_own * nPg / (t_mma * 1e-3),
                        "mma_TFLOPs_algorithmic": None if t_mma is None else mma_flops_alg / (t_mma * 1e-3) / 1e12,
                        "mma_TFLOPs_executed": None if t_mma is None else mma_flops_exec / (t_mma * 1e-3) / 1e12,
                        "mma_bytes_per_launch": bytes_mma, "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                        "pattern_build_s": t_pattern, "fused_schedule_build_s": t_sched, "mma_schedule_build_s": t_msched},
            "gpu_launches": (1 if single else 2) * K, "clocks": clk.summary()}

    extras = {}
    if not args.no_solve:
        try:
            extras["pcg"] = pcg_leg(args, g, part, pat, data, world, dist)
        except Exception as exc:  # the headline must survive a failing extra
            extras["pcg"] = {"error": repr(exc)[:300]}
    # ---- e2e through the host-buffer boundary (pinned in, pinned out), every rank at once ----
    try:
        line["e2e"] = e2e_leg(args, g, part, C, pat, sched, data, world, dist, path, msched)
    except Exception as exc:
        line["e2e"] = {"error": repr(exc)[:300]}
    del data, sched, msched, pat, A
    torch.cuda.empty_cache()
    if not args.no_pf:
        for cfg_id in ([3, 4] if args.pf_config == 0 else [args.pf_config]):
            try:
                extras[f"phase_field_config{cfg_id}"] = phase_field_leg(args, rank, world, dist, cfg_id)
            except Exception as exc:
                extras[f"phase_field_config{cfg_id}"] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()
    if not args.no_transient:
        try:
            extras["transient"] = transient_leg(args, rank, world, dist)
        except Exception as exc:
            extras["transient"] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_parity:
        try:
            line["parity"] = parity_leg(args)
        except Exception as exc:
            line["parity"] = {"error": repr(exc)[:300]}
        try:
            extras["e2e_solve"] = e2e_solve_leg(args)
        except Exception as exc:
            extras["e2e_solve"] = {"error": repr(exc)[:300]}
        try:
            extras["config1"] = config1_leg(args)
        except Exception as exc:
            extras["config1"] = {"error": repr(exc)[:300]}
    line["extras"] = extras

    if rank == 0:
        if world == 1 and not args.no_cpu:
            v, dt, procs, pattern_s, kind = cpu_arm(args.cpu_sample, 1, 3, args.cpu_procs, args.cpu_kind or None)
            line["cpu_baseline"] = {"value": v, "unit": "GP/s", "cores": procs, "kind": kind, "host_cores": os.cpu_count(),
                                    "pattern_build_s": pattern_s, "sample": cpu_sample_text(kind, procs, args.cpu_sample, dt, 3)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _material_C():
    lam = E_MOD * NU / ((1 + NU) * (1 - 2 * NU))
    mu = E_MOD / (2 * (1 + NU))
    I = np.array([1.0, 1, 1, 0, 0, 0])
    return lam * np.outer(I, I) + 2 * mu * np.eye(6)


def pcg_leg(args, g, part, pat, data, world, dist):
    """The consumer of config 2: Jacobi-PCG on the assembled (row-sharded) K with the x=0 face clamped and u_z pulled on the
    far face; `iters` iterations timed on the device (halo exchange + 2 all-reduces per iteration when sharded)."""
    import torch

    from easyfea_b200 import dist as efd
    from easyfea_b200 import solver
    from easyfea_b200.assembly import DeviceCsr

    n, d = args.n, 3
    nrows = part.n_owned * d
    nz = int(pat.indptr[nrows].item())
    K = DeviceCsr(pat.indptr[:nrows + 1], pat.indices[:nz], data[:nz], (nrows, part.n_local * d), pat.node_graph)
    comm = None
    if world > 1:
        part.plan_exchange()
        comm = efd.RowComm(part, d)
    P = (n + 1) * (n + 1)
    plane = part.nodes // P
    x0 = np.zeros(part.n_local * d)
    free = np.ones(nrows, dtype=np.uint8)
    loc = np.arange(part.n_owned)
    for c in range(3):
        free[loc[plane[:part.n_owned] == 0] * 3 + c] = 0
    top = np.flatnonzero(plane == world * n)
    x0[top * 3 + 2] = 0.01
    free[top[top < part.n_owned] * 3 + 2] = 0
    b = torch.zeros(nrows, dtype=torch.float64, device=data.device)
    iters = args.pcg_iters
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fused, degree=1, its=iters):
        solver.pcg(K, b, x0=x0, free_mask=free, tol=1e-30, maxiter=5, check_every=5, comm=comm, fused=fused, precond_degree=degree)  # warm-up
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        x, info = solver.pcg(K, b, x0=x0, free_mask=free, tol=1e-30, maxiter=its, check_every=its, comm=comm, fused=fused,
                             precond_degree=degree)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=data.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / info["iterations"], info

    per, info = timed(True)
    per_unfused, info_u = timed(False)
    # the polynomial form: one outer iteration = `degree` products (degree - 1 of them on single-precision matrix values)
    deg = solver.CHEB_DEGREE
    per_cheb, info_c = timed(True, deg, max(iters // deg, 5))
    bytes_it = nz * 8 + (nz // 9) * 4 + nrows * (4 + 12 * 8)  # node-block SpMV: one int32 per 3x3 block
    return {"iterations_timed": info["iterations"], "ms_per_iter": per, "rel_residual_after": info["rel_residual"],
            "ms_per_iter_unfused": per_unfused, "rel_residual_after_unfused": info_u["rel_residual"],
            "GBps_per_gpu": bytes_it / (per * 1e-3) / 1e9, "dofs_per_gpu": nrows, "nnz_per_gpu": nz,
            "halo_bytes_per_exchange": 0 if comm is None else comm.bytes_per_exchange,
            "chebyshev": {"degree": deg, "outer_iterations_timed": info_c["iterations"], "ms_per_outer_iter": per_cheb,
                          "rel_residual_after": info_c["rel_residual"], "fp32_inner_products": bool(info_c.get("precond_fp32")),
                          "note": "plain Jacobi needs ~3.6 iterations for the progress of one outer iteration at degree 4 "
                                  "(efb_pcg_iterate_cheb: the default of pcg() inside the fused path; set-up — power iteration, "
                                  "single-precision copy — is inside the timed region)"},
            "note": "plain Jacobi (precond_degree=1); fused: 3 kernels per iteration, reductions and halo pushes through peer memory inside the kernels "
                    "(efb_pcg_iterate); unfused: one kernel per vector operation and, when sharded, NCCL send/recv + 2 all-reduces "
                    "per iteration.  Setup (diagonal, reference norm, initial residual: 3 extra SpMVs) is inside the timed region"}


def phase_field_leg(args, rank, world, dist, cfg_id):
    """BASELINE config 3 (default): 2D shear test, TRI3 2*n^2 elements, Miehe split, AT2; `--pf-config 4`: 3D notched specimen,
    TETRA4 6*n^3 elements (Kuhn split of n^3 cubes), He split, AT2 — seconds per staggered iteration (damage assembly + solve,
    displacement assembly + solve) with the mesh partitioned element-wise over the ranks."""
    import torch

    from easyfea_b200 import dist as efd
    from easyfea_b200 import mesh, meshgen, phasefield, staggered

    cfg4 = cfg_id == 4
    et, dim, split = ("TETRA4", 3, "He") if cfg4 else ("TRI3", 2, "Miehe")
    n = (args.pf_n4 if cfg4 else args.pf_n) or (119 if cfg4 else 1000)
    L, l0 = 1e-3, 1e-5 * max(1.0, 1000.0 / n)  # l0 = 2 h like examples/PhaseField/Shear.py (clC = l0/2 there)
    lengths = (L,) * dim
    lattice, connect = meshgen.structured_mesh(et, n, lengths=lengths)
    coords, _ = meshgen.structured_mesh(et, n, lengths=lengths, jitter=0.15, seed=1)
    Nn, Ne = coords.shape[0], connect.shape[0]
    if world > 1:
        part = efd.Partition.from_global(connect, Nn, world, rank)
        g = mesh.ElemGroup(et, part.connect, coords[part.nodes], all_nodes_used=True)
        sysm = staggered.LocalSystem(g, part, lambda p, d: efd.RowComm(p, d))
        nodes = part.nodes
    else:
        g = mesh.ElemGroup(et, connect, coords, all_nodes_used=True)
        sysm = staggered.LocalSystem(g)
        nodes = np.arange(Nn)
    del connect
    pfm = phasefield.PhaseFieldModel(phasefield.IsotropicMaterial(dim, 210e9, 0.3, planeStress=False, thickness=1.0), split,
                                     "AT2", 2.7e3, l0)
    simu = staggered.PhaseFieldStaggered(sysm, pfm, pcg_tol=1e-8, pcg_maxiter=args.pf_maxiter)
    simu.pcg_fused = False if args.pf_unfused else "auto"
    simu.pcg_single_reduction = {"auto": "auto", "on": True, "off": False}[args.pf_single_reduction]
    simu.pcg_precond_degree = "auto" if args.pf_degree == "auto" else int(args.pf_degree)
    ix, iy = np.rint(lattice[nodes, 0] / L * n).astype(np.int64), np.rint(lattice[nodes, 1] / L * n).astype(np.int64)
    loc = np.arange(nodes.size)
    comps = list(range(dim))
    simu.add_dirichlet(loc[(iy == n // 2) & (ix <= n // 2)], [1], [0], problemType="damage")  # crack / notch as d = 1
    simu.add_dirichlet(loc[iy == n], [8e-6, 4e-6, 0.0][:dim], comps)
    simu.add_dirichlet(loc[iy == 0], [0.0] * dim, comps)
    simu.iterate()  # warm-up iteration (pattern build, first solves from zero fields)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = args.pf_iters
    e0.record()
    for _ in range(its):
        conv, dmax = simu.iterate()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=conv.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"workload": f"BASELINE config {4 if cfg4 else 3}: {et} {'notched specimen' if cfg4 else 'shear test'}, {Ne} elements, "
                        f"{Nn} nodes, {split}, AT2, strong-scaled over {world} GPU(s)", "s_per_iter": ms / its / 1e3,
            "iterations_timed": its, "pcg_iters_damage": simu.info["damage"]["iterations"],
            "pcg_iters_elastic": simu.info["elastic"]["iterations"],
            "pcg_converged": bool(simu.info["damage"]["converged"] and simu.info["elastic"]["converged"]),
            "pcg_fused": bool(simu.info["elastic"].get("fused", False)), "pcg_single_reduction": bool(simu.info["elastic"].get("single_reduction", False)),
            "pcg_precond_degree": {"damage": simu.info["damage"].get("precond_degree", 1), "elastic": simu.info["elastic"].get("precond_degree", 1)},
            "last_damage_increment": float(conv.item()), "max_damage": float(dmax.item())}


def transient_leg(args, rank, world, dist):
    """BASELINE config 5: HEXA27 mesh, K + M assembly (elastodynamics: 81x81 element matrices at 27 Gauss points) and Newmark
    steps, then thermal K + C assembly and parabolic steps — seconds per assembly and per time step, element-partitioned."""
    import torch

    from easyfea_b200 import dist as efd
    from easyfea_b200 import mesh, meshgen, staggered, transient

    n = args.tr_n
    lattice, connect = meshgen.structured_mesh("HEXA27", n)
    coords, _ = meshgen.structured_mesh("HEXA27", n, jitter=0.1, seed=3)
    Nn = coords.shape[0]
    if world > 1:
        part = efd.Partition.from_global(connect, Nn, world, rank)
        g = mesh.ElemGroup("HEXA27", part.connect, coords[part.nodes], all_nodes_used=True)
        sysm = staggered.LocalSystem(g, part, lambda p, d: efd.RowComm(p, d))
        nodes = part.nodes
    else:
        g = mesh.ElemGroup("HEXA27", connect, coords, all_nodes_used=True)
        sysm = staggered.LocalSystem(g)
        nodes = np.arange(Nn)
    x = lattice[nodes, 0]
    loc = np.arange(nodes.size)
    lo, hi = loc[x < 1e-12], loc[x > 1 - 1e-12]

    def timed(fn, reps):
        fn()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    out = {"workload": f"BASELINE config 5: HEXA27 {n}^3 = {n**3} elements, {Nn} nodes, over {world} GPU(s)"}
    box = {}

    def build_dyn():
        box.pop("dyn", None)  # free the previous system first: K, C, M of this mesh are 3 x 3 GB
        box["dyn"] = transient.TransientSolve.elastodynamic(sysm, _material_C(), 7.8e-3, coefM=0.1, coefK=1e-5)

    out["elastodynamic_assembly_ms"] = timed(build_dyn, 2)  # K_e, M_e (27 Gauss points), two CSR replays, Rayleigh C
    dyn = box["dyn"]
    dyn.Solver_Set_Hyperbolic_Algorithm(dt=1e-3)  # Newmark average acceleration
    dyn.pcg_tol = 1e-8
    dyn.bc.add(lo, [0.0, 0.0, 0.0], [0, 1, 2], 3)
    dyn.bc.add(hi, [1e-3], [0], 3)
    out["newmark_step_ms"] = timed(lambda: dyn.Solve(), args.tr_steps)
    out["newmark_pcg_iters"] = dyn.info["iterations"]
    out["newmark_pcg_converged"] = bool(dyn.info["converged"])
    del dyn, box["dyn"]

    def build_th():
        box.pop("th", None)
        box["th"] = transient.TransientSolve.thermal(sysm, 1.0, 1.0)

    out["thermal_assembly_ms"] = timed(build_th, 2)
    th = box["th"]
    th.Solver_Set_Parabolic_Algorithm(dt=0.1, alpha=0.5)
    th.pcg_tol = 1e-8
    th.bc.add(lo, [0.0], [0], 1)
    th.bc.add(hi, [40.0], [0], 1)
    out["parabolic_step_ms"] = timed(lambda: th.Solve(), args.tr_steps)
    out["parabolic_pcg_iters"] = th.info["iterations"]
    out["parabolic_pcg_converged"] = bool(th.info["converged"])
    return out


def e2e_leg(args, g, part, C, pat, sched, data, world, dist, path, msched=None):
    """Same step through host buffers, on EVERY rank at once: pinned (connect int32, coords) -> device, integration + assembly,
    owned CSR data -> pinned host; wall time between barriers, max over ranks."""
    import torch

    from easyfea_b200 import assembly, mesh, operators
    from easyfea_b200 import device as dv

    steps = max(1, min(args.steps, args.e2e_steps))
    dg = mesh.device_group(g)
    h_conn = dg.connect.cpu().pin_memory()
    h_coord = dg.coord.cpu().pin_memory()
    nz = int(pat.indptr[part.n_owned * 3].item())
    h_out = torch.empty(nz, dtype=torch.float64).pin_memory()
    Ke = None
    if path == "two":
        Ke = dv.empty((g.Ne, 24, 24))

    def one():
        dg.connect.copy_(h_conn, non_blocking=True)  # the element kernel reads these buffers
        dg.coord.copy_(h_coord, non_blocking=True)
        if path == "mma":  # the schedules hold their own copy of the connectivity (built once, like the CSR pattern)
            assembly.assemble_elastic_mma(msched, C, "rigi", 1.0, out=data)
        elif path == "fused":
            assembly.assemble_elastic_fused(sched, C, "rigi", 1.0, out=data)
        else:
            operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke)
            pat.replay([Ke], out=data, n_nodes=part.n_owned)
        h_out.copy_(data[:nz], non_blocking=True)

    one()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=data.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    chk = float(h_out[:1000].sum())  # the result is really on the host
    return {"value": args.n**3 * 8 * world / dt, "unit": "GP/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(h_conn.numel() * 4 + h_coord.numel() * 8), "d2h_bytes_per_step": int(h_out.numel() * 8),
            "steps": steps, "checksum_head": chk, "path": path,
            "note": "pinned host buffers; H2D of connectivity+coordinates and D2H of the owned CSR data inside the timed region, on "
                    "every rank concurrently (max over ranks); the D2H of the assembled matrix at PCIe speed dominates"}


def parity_leg(args):
    """Parity numbers of this build on a common-size mesh (HEXA8 `--parity-n`^3, jittered): CSR structure and values of the
    device paths against the CPU path — the LIVE reference when baseline/_ref is importable, else the NumPy oracle."""
    import torch

    from easyfea_b200 import assembly, mesh, meshgen, operators

    n = args.parity_n
    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=0.2, seed=0)
    Nn = coords.shape[0]
    C = _material_C()
    against = cpu_kind()
    t0 = time.perf_counter()
    if against == "reference":
        from oracle.ref_cases import hexa8_elastic
        from oracle.ref_import import import_reference

        E = import_reference(travel_only=True)
        simu, gref, Cref, Ndof, _, _ = hexa8_elastic(E, n, E_MOD, NU)
        Ke_ref = E.FEM.Operators.Bilinear.LinearizedElasticity(gref, Cref)
        Kref = simu._Simu__Assemble_csr({gref: Ke_ref}, 3, Ndof, True)
        C = np.asarray(Cref)
    else:
        from easyfea_b200 import elements as el
        from oracle import easyfea_oracle as orc
        from scipy import sparse

        tab = el.gauss_table("HEXA8", "rigi")
        Ke_ref = orc.linearized_elasticity(orc.geometry(coords[connect], tab.dN_pg, tab.weights), C)
        inv, indices, indptr, nnz = orc.csr_map([connect], 3, Nn * 3, True)
        Kref = sparse.csr_matrix((orc.assemble_replay([Ke_ref], inv, nnz), indices, indptr), shape=(Nn * 3, Nn * 3))
    t_cpu = time.perf_counter() - t0
    g = mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    pat = assembly.Assembler().pattern(3, True, Nn * 3, (g,))
    Ke = operators.elastic_Ke_dev(g, C)
    two = pat.replay([Ke]).cpu().numpy()
    fused = assembly.assemble_elastic_fused(assembly.FusedSchedule(pat.graph), C).cpu().numpy()
    mma = assembly.assemble_elastic_mma(assembly.MmaSchedule(pat.graph), C).cpu().numpy()
    replay_ref = pat.replay([torch.from_numpy(np.ascontiguousarray(Ke_ref))]).cpu().numpy()  # the CPU path's own K_e, replayed on the device
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    return {"against": against, "mesh": f"HEXA8 {n}^3 = {n**3} elements, {Nn} nodes", "cpu_seconds": t_cpu,
            "indptr_equal": bool(np.array_equal(pat.indptr.cpu().numpy(), Kref.indptr) and pat.indptr.cpu().numpy().dtype == Kref.indptr.dtype),
            "indices_equal": bool(np.array_equal(pat.indices.cpu().numpy(), Kref.indices)),
            "Ke_rel_err": rel(Ke.cpu().numpy(), np.asarray(Ke_ref)),
            "data_rel_err_two_kernel": rel(two, Kref.data), "data_rel_err_fused": rel(fused, Kref.data),
            "data_rel_err_mma": rel(mma, Kref.data),
            "replay_of_cpu_Ke_bit_identical": bool(np.array_equal(replay_ref, Kref.data)), "tolerance": 1e-12}


def e2e_solve_leg(args):
    """mesh in -> displacement out: BASELINE config 2 at `--solve-n`^3 elements from HOST buffers (pinned connectivity and
    coordinates), fused assembly + Jacobi-PCG to a 1e-8 relative residual on the device, solution back to the host; the CSR
    pattern / node schedule (one-time per mesh, like the reference's cached csr map) are timed apart."""
    import torch

    from easyfea_b200 import assembly, mesh, meshgen, solver

    n = args.solve_n
    lattice, connect = meshgen.structured_mesh("HEXA8", n)
    coords, _ = meshgen.structured_mesh("HEXA8", n, jitter=0.2, seed=0)
    Nn = coords.shape[0]
    C = _material_C()
    g = mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dg = mesh.device_group(g)
    pat = assembly.Assembler().pattern(3, True, Nn * 3, (g,))
    sched = assembly.FusedSchedule(pat.graph)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    h_conn, h_coord = dg.connect.cpu().pin_memory(), dg.coord.cpu().pin_memory()
    x = lattice[:, 0]
    lo, hi = np.flatnonzero(x < 1e-12), np.flatnonzero(x > 1 - 1e-12)
    free = np.ones(Nn * 3, dtype=np.uint8)
    x0 = np.zeros(Nn * 3)
    for c in range(3):
        free[lo * 3 + c] = 0
    free[hi * 3] = 0
    x0[hi * 3] = 0.1  # u_x = 0.1 on the far face, clamped near face
    h_free, h_x0 = torch.from_numpy(free).pin_memory(), torch.from_numpy(x0).pin_memory()
    h_u = torch.empty(Nn * 3, dtype=torch.float64).pin_memory()
    data = torch.empty(pat.nnz, dtype=torch.float64, device=dg.coord.device)
    b = torch.zeros(Nn * 3, dtype=torch.float64, device=data.device)

    def one():
        dg.connect.copy_(h_conn, non_blocking=True)
        dg.coord.copy_(h_coord, non_blocking=True)
        d_free, d_x0 = h_free.to(data.device, non_blocking=True), h_x0.to(data.device, non_blocking=True)
        assembly.assemble_elastic_fused(sched, C, "rigi", 1.0, out=data)
        K = assembly.DeviceCsr(pat.indptr, pat.indices, data, pat.shape, pat.node_graph)
        u, info = solver.pcg(K, b, x0=d_x0, free_mask=d_free, tol=1e-8, check_every=50)
        h_u.copy_(u, non_blocking=True)
        torch.cuda.synchronize()
        return info

    one()
    t0 = time.perf_counter()
    info = one()
    dt = time.perf_counter() - t0
    return {"workload": f"HEXA8 {n}^3 = {n**3} elements, {Nn * 3} dofs: H2D mesh + fused assembly + Jacobi-PCG (tol 1e-8) + D2H u",
            "seconds": dt, "GP_per_s": n**3 * 8 / dt, "pcg_iterations": info["iterations"], "rel_residual": info["rel_residual"],
            "converged": bool(info["converged"]), "one_time_setup_s": t_setup, "u_max": float(h_u.abs().max()),
            "h2d_bytes": int(h_conn.numel() * 4 + h_coord.numel() * 8 + free.size + x0.size * 8), "d2h_bytes": int(h_u.numel() * 8)}


def config1_leg(args):
    """BASELINE config 1 (README cantilever, QUAD9 28 x 3, `Simulations.Elastic.Solve`): the unmodified reference with and without
    `easyfea_b200.dropin.install()` — parity of K and u and both wall times (84 elements: a functional check, not a speed claim)."""
    if cpu_kind() != "reference":
        return {"skipped": "no reference install under baseline/_ref"}
    from easyfea_b200 import dropin
    from oracle.ref_cases import readme_cantilever
    from oracle.ref_import import import_reference

    E = import_reference(travel_only=True)

    def run():
        t0 = time.perf_counter()
        simu, mesh_, n0, nL = readme_cantilever(E)
        u = np.array(simu.Solve())
        K = simu.Get_K_C_M_F()[0]
        return u, K, time.perf_counter() - t0, float(u.reshape(-1, 2)[nL, 1].mean())

    u0, K0, t_ref, tip0 = run()
    dropin.install(E)
    try:
        run()  # first patched run builds the device mirrors / pattern
        u1, K1, t_dev, tip1 = run()
    finally:
        dropin.uninstall()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    return {"workload": "BASELINE config 1: README cantilever, QUAD9 28x3 (+SEG3 boundary group), plane stress, add_surfLoad, Solve()",
            "reference_s": t_ref, "with_dropin_s": t_dev, "tip_uy_reference": tip0, "tip_uy_dropin": tip1,
            "indptr_indices_equal": bool(np.array_equal(K0.indptr, K1.indptr) and np.array_equal(K0.indices, K1.indices)),
            "K_data_rel_err": rel(K1.data, K0.data), "u_rel_err": rel(u1, u0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--cells", dest="n", type=int, default=200, help="cells per side of the owned HEXA8 cube (200 -> 8.0 M elements)")
    ap.add_argument("--cpu-sample", type=int, default=24, help="cells per side of the CPU sample (24 -> 13 824 elements per process)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="host processes of the CPU arm (0 = one per core)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the Jacobi-PCG extra")
    ap.add_argument("--pcg-iters", type=int, default=50)
    ap.add_argument("--no-pf", action="store_true", help="skip the phase-field staggered-iteration extra")
    ap.add_argument("--pf-config", type=int, default=0, choices=[0, 3, 4],
                    help="phase-field extras: 0 = both (default), 3 = config 3 only (TRI3, Miehe), 4 = config 4 only (TETRA4, He)")
    ap.add_argument("--pf-n", type=int, default=0, help="config 3 cells per side (default 1000 -> 2.0 M TRI3)")
    ap.add_argument("--pf-n4", type=int, default=0, help="config 4 cells per side (default 119 -> 10.1 M TETRA4)")
    ap.add_argument("--path", default="auto", choices=["auto", "mma", "fused", "two"],
                    help="the step: fused integration+assembly on the FP64 MMA instruction, the DFMA fused kernel, the two-kernel path "
                         "(K_e, then replay), or the fastest of them")
    ap.add_argument("--cpu-kind", default="", choices=["", "reference", "port"], help="CPU arm: live reference or NumPy port (default: reference if importable)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity / e2e_solve / config 1 legs")
    ap.add_argument("--parity-n", type=int, default=16, help="cells per side of the common-size parity mesh")
    ap.add_argument("--solve-n", type=int, default=100, help="cells per side of the mesh-in -> u-out leg (100 -> 1.0 M elements)")
    ap.add_argument("--pf-iters", type=int, default=2)
    ap.add_argument("--pf-maxiter", type=int, default=20000)
    ap.add_argument("--no-transient", action="store_true", help="skip the HEXA27 transient extra (config 5)")
    ap.add_argument("--tr-n", type=int, default=40, help="HEXA27 cells per side (40 -> 64 000 elements, 531 441 nodes)")
    ap.add_argument("--tr-steps", type=int, default=3)
    ap.add_argument("--pf-single-reduction", default="auto", choices=["auto", "on", "off"],
                    help="phase-field solves with the Chronopoulos-Gear PCG form (auto: small shards)")
    ap.add_argument("--pf-degree", default="auto", help="phase-field solves: Chebyshev-Jacobi polynomial degree m (1 = plain Jacobi, auto = by shard size)")
    ap.add_argument("--pf-unfused", action="store_true", help="phase-field solves with the NCCL/kernel-per-operation PCG loop")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
