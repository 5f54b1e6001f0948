"""dev: time the node-block SpMV alone (CUDA events) on assembled matrices.  env: PROBE_ELEM, PROBE_N, PROBE_DOF"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import _lib, assembly, mesh, meshgen, solver  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402

elem = os.environ.get("PROBE_ELEM", "HEXA8")
n = int(os.environ.get("PROBE_N", "100"))
coords, connect = meshgen.structured_mesh(elem, n, jitter=0.15, seed=0)
g = mesh.ElemGroup(elem, connect, coords, all_nodes_used=True)
d = int(os.environ.get("PROBE_DOF", str(g.dim)))
Nn = coords.shape[0]
ndof = connect.shape[1] * d
Xe = torch.randn(connect.shape[0], ndof, ndof, dtype=torch.float64, device="cuda")
A = assembly.Assembler().Assemble_csr({g: Xe}, d, Nn * d, True, as_device=True)
del Xe
x = torch.randn(A.shape[1], dtype=torch.float64, device="cuda")
y = dv.empty((A.shape[0],))
mask = torch.ones(A.shape[0], dtype=torch.uint8, device="cuda")
partials = dv.empty((_lib.load().efb_pcg_partials_size(),))
ref = None
reps = 20
for _ in range(3):
    solver.spmv(A, x, y, 0, mask, partials)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    solver.spmv(A, x, y, 0, mask, partials)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
gb = (A.nnz * 8 + A.nnz // (d * d) * 4 + A.shape[0] * 17 + Nn * 8) / 1e9
chk = float((A.to_scipy() @ x.cpu().numpy() - y.cpu().numpy()).__abs__().max()) if A.nnz < 3e8 else -1.0
print(f"{elem} n={n} d={d} body={os.environ.get('EFB_SPMV_BODY', '-')} stream={os.environ.get('EFB_SPMV_STREAM', '0')}: "
      f"{ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s  maxerr {chk:.2e}", flush=True)
