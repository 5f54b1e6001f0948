"""dev: fused vs NCCL-loop PCG on z-slabs of a HEXA8 cube under torchrun (one rank per GPU).  env: PROBE_N, PROBE_ITERS"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from easyfea_b200 import assembly, mesh, operators, solver  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402
from easyfea_b200 import dist as efd  # noqa: E402
from easyfea_b200.assembly import DeviceCsr  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
n, iters = int(os.environ.get("PROBE_N", "128")), int(os.environ.get("PROBE_ITERS", "50"))
g, part = bench.slab_system(n, rank, world)
pat = assembly.Assembler().pattern(3, True, g.Ncoords * 3, (g,))
Ke = operators.elastic_Ke_dev(g, np.ascontiguousarray(bench._material_C()), "rigi", 1.0)
data = pat.replay([Ke], n_nodes=part.n_owned)
del Ke
nrows = part.n_owned * 3
nz = int(pat.indptr[nrows].item())
K = DeviceCsr(pat.indptr[:nrows + 1], pat.indices[:nz], data[:nz], (nrows, part.n_local * 3), pat.node_graph)
comm = None
if world > 1:
    part.plan_exchange()
    comm = efd.RowComm(part, 3)
P = (n + 1) * (n + 1)
plane = part.nodes // P
x0 = np.zeros(part.n_local * 3)
free = np.ones(nrows, dtype=np.uint8)
loc = np.arange(part.n_owned)
for c in range(3):
    free[loc[plane[:part.n_owned] == 0] * 3 + c] = 0
top = np.flatnonzero(plane == world * n)
x0[top * 3 + 2] = 0.01
free[top[top < part.n_owned] * 3 + 2] = 0
b = torch.zeros(nrows, dtype=torch.float64, device="cuda")
for rep in range(2):
    for fused, persistent, sr in ((True, False, True), (True, False, False), (False, False, False)):
        solver.pcg(K, b, x0=x0, free_mask=free, tol=1e-30, maxiter=3, check_every=3, comm=comm, fused=fused, persistent=persistent, single_reduction=sr)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, info = solver.pcg(K, b, x0=x0, free_mask=free, tol=1e-30, maxiter=iters, check_every=iters, comm=comm, fused=fused, persistent=persistent, single_reduction=sr)
        e1.record()
        torch.cuda.synchronize()
        print(f"rank {rank}/{world} n={n} fused={fused} single_reduction={sr} rep={rep}: {e0.elapsed_time(e1) / iters:.4f} ms/iter rel {info['rel_residual']:.6e}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
