"""dev: a few fused and unfused PCG iterations on a HEXA8 cube / TRI3 square, for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import assembly, mesh, meshgen, operators, solver  # noqa: E402

elem = os.environ.get("PROBE_ELEM", "HEXA8")
n = int(os.environ.get("PROBE_N", "100"))
iters = int(os.environ.get("PROBE_ITERS", "10"))
coords, connect = meshgen.structured_mesh(elem, n, jitter=0.15, seed=0)
g = mesh.ElemGroup(elem, connect, coords, all_nodes_used=True)
dim = g.dim
lam, mu = 121153.8, 80769.2
ns = 3 if dim == 2 else 6
I = np.array([1.0] * dim + [0.0] * (ns - dim))
C = lam * np.outer(I, I) + 2 * mu * np.eye(ns)
Ke = operators.elastic_Ke_dev(g, C, "rigi", 1.0)
Nn = coords.shape[0]
K = assembly.Assembler().Assemble_csr({g: Ke}, dim, Nn * dim, True, as_device=True)
free = np.ones(Nn * dim, dtype=np.uint8)
x0 = np.zeros(Nn * dim)
lo = np.flatnonzero(coords[:, 0] < 1e-9)
hi = np.flatnonzero(coords[:, 0] > coords[:, 0].max() - 1e-9)
for c in range(dim):
    free[lo * dim + c] = 0
free[hi * dim] = 0
x0[hi * dim] = 0.01
b = torch.zeros(Nn * dim, dtype=torch.float64, device="cuda")
for fused, persistent, sr in ((True, False, True), (True, False, False), (False, False, False)):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, info = solver.pcg(K, b, x0=x0, free_mask=free, tol=1e-30, maxiter=iters, check_every=iters, fused=fused, persistent=persistent, single_reduction=sr)
        e1.record()
        torch.cuda.synchronize()
        print(f"{elem} n={n} fused={fused} single_reduction={sr} rep={rep}: {e0.elapsed_time(e1) / iters:.4f} ms/iter, rel {info['rel_residual']:.6e}", flush=True)
