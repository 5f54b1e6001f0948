#!/usr/bin/env python
"""dev tool: time the element-stiffness kernel (3 C modes) and the CSR replay of one library build.
EASYFEA_B200_LIB selects the .so; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import assembly, mesh, meshgen, operators  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def main():
    elem = sys.argv[1] if len(sys.argv) > 1 else "HEXA8"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    coords, connect = meshgen.structured_mesh(elem, n, jitter=0.2, seed=0)
    g = mesh.ElemGroup(elem, connect, coords, all_nodes_used=True)
    dg = mesh.device_group(g)
    dim = dg.dim
    ns = 3 if dim == 2 else 6
    nPg = dg.nPg("rigi")
    lam, mu = 121153.8, 80769.2
    I = np.zeros(ns)
    I[:dim] = 1
    C = lam * np.outer(I, I) + 2 * mu * np.eye(ns)
    ndof = dg.nPe * dim
    Ne = dg.Ne
    Ke = dv.empty((Ne, ndof, ndof))
    out = {"lib": os.path.basename(os.environ.get("EASYFEA_B200_LIB", "default")), "elem": elem, "Ne": Ne, "nPg": nPg}
    t = timeit(lambda: operators.elastic_Ke_dev(g, C, "rigi", 1.0, out=Ke), reps)
    out["Ke_const_ms"] = t
    out["Ke_const_GPps"] = Ne * nPg / t / 1e-3
    ref = Ke[:1000].clone()
    Ce = torch.from_numpy(C).cuda().expand(Ne, ns, ns).contiguous()
    t = timeit(lambda: operators.elastic_Ke_dev(g, Ce, "rigi", 1.0, out=Ke), reps)
    out["Ke_e_ms"] = t
    out["err_e"] = float((Ke[:1000] - ref).norm() / ref.norm())
    del Ce
    Cep = torch.from_numpy(C).cuda().expand(Ne, nPg, ns, ns).contiguous()
    t = timeit(lambda: operators.elastic_Ke_dev(g, Cep, "rigi", 1.0, out=Ke), reps)
    out["Ke_epg_ms"] = t
    out["err_epg"] = float((Ke[:1000] - ref).norm() / ref.norm())
    del Cep
    if os.environ.get("TUNE_REPLAY", "1") == "1":
        pat = assembly.Assembler().pattern(dim, True, coords.shape[0] * dim, (g,))
        data = dv.empty((pat.nnz,))
        t = timeit(lambda: pat.replay([Ke], out=data), reps)
        out["replay_ms"] = t
        out["replay_GBps"] = (Ne * ndof * ndof * 8 + Ne * dg.nPe**2 * 4 + pat.nnz * 8) / t / 1e6
    if os.environ.get("TUNE_MASS", "0") == "1":
        Me = dv.empty((Ne, ndof, ndof))
        out["Me_vec_ms"] = timeit(lambda: operators.mass_Me_dev(g, 2.0, dim, "mass", 1.0, out=Me), reps)
        del Me
        Ms = dv.empty((Ne, dg.nPe, dg.nPe))
        out["Me_scalar_ms"] = timeit(lambda: operators.mass_Me_dev(g, 2.0, 1, "mass", 1.0, out=Ms), reps)
        out["Kdiff_scalar_ms"] = timeit(lambda: operators.diffusion_Ke_dev(g, None, 1.5, "rigi", 1.0, out=Ms), reps)
        pat1 = assembly.Assembler().pattern(1, True, coords.shape[0], (g,))
        d1 = dv.empty((pat1.nnz,))
        out["replay_scalar_ms"] = timeit(lambda: pat1.replay([Ms], out=d1), reps)
    if os.environ.get("TUNE_SPMV", "0") == "1":
        from easyfea_b200 import solver
        from easyfea_b200.assembly import DeviceCsr
        A = pat.assemble([Ke])
        x = torch.randn(A.shape[1], dtype=torch.float64, device="cuda")
        y = dv.empty((A.shape[0],))
        mask = torch.ones(A.shape[0], dtype=torch.uint8, device="cuda")
        partials = dv.empty((4096,))
        out["spmv_node_ms"] = timeit(lambda: solver.spmv(A, x, y, 0, mask, partials), reps)
        B = DeviceCsr(A.indptr, A.indices, A.data, A.shape)
        out["spmv_csr_ms"] = timeit(lambda: solver.spmv(B, x, y, 0, mask, partials), reps)
        out["spmv_bytes_GB"] = (A.nnz * 8 + A.nnz // dim**2 * 4 + A.shape[0] * 24) / 1e9
        out["spmv_node_GBps"] = out["spmv_bytes_GB"] / out["spmv_node_ms"] * 1e3
        diag = dv.empty((A.shape[0],))
        from easyfea_b200 import _lib
        out["diag_ms"] = timeit(lambda: _lib.call("efb_csr_diagonal", A.shape[0], 0, A.index_bytes, dv.ptr(A.indptr), dv.ptr(A.indices),
                                                  dv.ptr(A.data), dv.ptr(diag), dv.stream_ptr()), 3)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
