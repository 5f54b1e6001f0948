"""Jacobi-preconditioned conjugate gradient on a device CSR matrix — the large-mesh consumer of the assembled system
(north star).  The reference solves through `Solvers._Solve_Axb` (Simulations/Solvers.py:225-394) after slicing
`A[dofsUnknown,:][:,dofsUnknown]` on the host (:526-530); here Dirichlet dofs are masked instead (projected CG), all
scalars stay on the device and every dot product is a fixed-order reduction, so a solve is reproducible run to run.

Multi-GPU: rows are sharded in contiguous blocks; `p` is kept at global length on every rank and only its interface
entries are exchanged per iteration (see `easyfea_b200.dist`).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from . import device as dv
from .assembly import DeviceCsr


def lanes_per_row(nnz: int, nrows: int) -> int:
    avg = nnz / max(nrows, 1)
    for lanes in (4, 8, 16):
        if avg <= 2 * lanes:
            return lanes
    return 32


def spmv(A: DeviceCsr, x: torch.Tensor, y: torch.Tensor = None, row_offset: int = 0, mask=None, partials=None):
    nrows = A.indptr.numel() - 1
    if y is None:
        y = dv.empty((nrows,))
    ng = getattr(A, "node_graph", None)
    if ng is not None and ng[2] in (1, 2, 3) and nrows % ng[2] == 0:
        adjptr, adj, d = ng  # node-block product: the column structure is read from the node adjacency
        n_nodes = nrows // d
        _lib.call("efb_spmv_nodeblock", n_nodes, d, dv.ptr(adjptr), dv.ptr(adj), dv.ptr(A.data), dv.ptr(x), int(row_offset),
                  dv.ptr(mask), dv.ptr(y), dv.ptr(partials), lanes_per_row(A.nnz, nrows), dv.stream_ptr())
        return y
    _lib.call("efb_spmv_csr", nrows, A.index_bytes, dv.ptr(A.indptr), dv.ptr(A.indices), dv.ptr(A.data), dv.ptr(x), int(row_offset),
              dv.ptr(mask), dv.ptr(y), dv.ptr(partials), lanes_per_row(A.nnz, nrows), dv.stream_ptr())
    return y


def pcg(A: DeviceCsr, b, x0=None, free_mask=None, tol: float = 1e-8, maxiter: int = None, check_every: int = 25,
        row_offset: int = 0, comm=None):
    """Solve A x = b on the free dofs (free_mask True / 1 = unknown; other entries of x keep the values of x0).

    A holds the local row block [row_offset, row_offset + nrows) with GLOBAL column indices.  Stops when
    ||r|| <= tol * ||b - A x_known|| (both restricted to free dofs).  Returns (x_local, info dict).
    `comm` (easyfea_b200.dist.RowComm) supplies the halo exchange and scalar all-reduces for row-sharded runs.
    """
    dev = A.data.device
    nrows = A.indptr.numel() - 1
    n_glob = A.shape[1]
    st = dv.stream_ptr
    b = dv.to_device(b)
    assert b.numel() == nrows
    mask = None
    if free_mask is not None:
        mask = dv.to_device(free_mask).to(torch.uint8).contiguous()
        assert mask.numel() == nrows
    maxiter = int(maxiter if maxiter is not None else 10 * n_glob)
    P = _lib.load().efb_pcg_partials_size()
    partials = dv.empty((P,))
    scal = torch.zeros(6, dtype=torch.float64, device=dev)  # [rz, pAp, rz_new, rr, -, -]
    s_ptr = lambda i: ctypes.c_void_p(scal.data_ptr() + 8 * i)  # noqa: E731

    # vectors: p at global length (halo entries live outside the owned block), the rest local
    x_full = torch.zeros(n_glob, dtype=torch.float64, device=dev)
    x = x_full[row_offset:row_offset + nrows]
    if x0 is not None:
        x.copy_(dv.to_device(x0).reshape(-1)[:nrows])  # a local [owned | halo] vector may be passed: the owned part counts
    p_full = torch.zeros(n_glob, dtype=torch.float64, device=dev)
    p = p_full[row_offset:row_offset + nrows]
    r, z, Ap = dv.empty((nrows,)), dv.empty((nrows,)), dv.empty((nrows,))

    diag = dv.empty((nrows,))
    _lib.call("efb_csr_diagonal", nrows, int(row_offset), A.index_bytes, dv.ptr(A.indptr), dv.ptr(A.indices), dv.ptr(A.data),
              dv.ptr(diag), st())
    inv_diag = dv.empty((nrows,))
    _lib.call("efb_pcg_inv_diag", nrows, dv.ptr(diag), dv.ptr(mask), dv.ptr(inv_diag), st())

    def reduce_to(slot, m):
        _lib.call("efb_pcg_reduce", dv.ptr(partials), m, s_ptr(slot), st())
        if comm is not None:
            comm.all_reduce_sum(scal[slot:slot + m])

    # reference norm: || b - A x_known || on the free dofs (x_known = x0 on constrained dofs, 0 elsewhere)
    if mask is not None:
        xk = torch.zeros_like(x_full)
        xk[row_offset:row_offset + nrows] = x * (1 - mask.to(torch.float64))
        if comm is not None:
            comm.halo_exchange(xk)
        spmv(A, xk, Ap, row_offset)
        _lib.call("efb_pcg_init", nrows, dv.ptr(b), dv.ptr(Ap), dv.ptr(inv_diag), dv.ptr(mask), dv.ptr(r), dv.ptr(z), dv.ptr(p),
                  dv.ptr(partials), st())
        reduce_to(2, 2)
        bnorm2 = float(scal[3].item())
    else:
        _lib.call("efb_pcg_dot", nrows, dv.ptr(b), dv.ptr(b), dv.ptr(partials), st())
        reduce_to(3, 1)
        bnorm2 = float(scal[3].item())
    if bnorm2 == 0.0:
        return x.clone(), {"iterations": 0, "rel_residual": 0.0, "converged": True}

    # initial residual with the actual start vector
    if comm is not None:
        comm.halo_exchange(x_full)
    spmv(A, x_full, Ap, row_offset)
    _lib.call("efb_pcg_init", nrows, dv.ptr(b), dv.ptr(Ap), dv.ptr(inv_diag), dv.ptr(mask), dv.ptr(r), dv.ptr(z), dv.ptr(p),
              dv.ptr(partials), st())
    reduce_to(2, 2)  # scal[2] = r.z, scal[3] = r.r
    scal[0:1].copy_(scal[2:3])
    it = 0
    rr = float(scal[3].item())
    target = tol * tol * bnorm2
    while rr > target and it < maxiter:
        for _ in range(check_every):
            if comm is not None:
                comm.halo_exchange(p_full)
            spmv(A, p_full, Ap, row_offset, mask, partials)
            reduce_to(1, 1)  # pAp
            _lib.call("efb_pcg_update_xr", nrows, s_ptr(0), s_ptr(1), dv.ptr(p), dv.ptr(Ap), dv.ptr(x), dv.ptr(r), dv.ptr(inv_diag),
                      dv.ptr(mask), dv.ptr(z), dv.ptr(partials), st())
            reduce_to(2, 2)  # rz_new, rr
            _lib.call("efb_pcg_update_p", nrows, s_ptr(2), s_ptr(0), dv.ptr(z), dv.ptr(mask), dv.ptr(p), st())
            scal[0:1].copy_(scal[2:3])
            it += 1
        rr = float(scal[3].item())
        if rr != rr:
            raise _lib.EfbError("PCG broke down (NaN residual): matrix not SPD on the free dofs?")
    rel = (rr / bnorm2) ** 0.5
    return x.clone(), {"iterations": it, "rel_residual": rel, "converged": rel <= tol}
