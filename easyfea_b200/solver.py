"""Jacobi-preconditioned conjugate gradient on a device CSR matrix — the large-mesh consumer of the assembled system
(north star).  The reference solves through `Solvers._Solve_Axb` (Simulations/Solvers.py:225-394) after slicing
`A[dofsUnknown,:][:,dofsUnknown]` on the host (:526-530); here Dirichlet dofs are masked instead (projected CG), all
scalars stay on the device and every dot product is a fixed-order reduction, so a solve is reproducible run to run.

Multi-GPU: rows are sharded in contiguous blocks; `p` is kept over `[owned | halo]` on every rank and only its interface
entries are exchanged per iteration — by stores into the neighbours' peer-mapped buffers from inside the iteration
kernels (`efb_pcg_iterate`, `easyfea_b200.dist.PeerWorkspace`), not by collective calls.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from . import device as dv
from .assembly import DeviceCsr


def lanes_per_row(nnz: int, nrows: int) -> int:
    """lanes cooperating on one row of the generic CSR product"""
    avg = nnz / max(nrows, 1)
    for lanes in (4, 8, 16):
        if avg <= 2 * lanes:
            return lanes
    return 32


def lanes_per_node(nnz: int, nrows: int, dof_n: int = 1) -> int:
    """lanes cooperating on one node of the node-block product (block form: a lane takes whole neighbour blocks, so the count
    follows the number of NEIGHBOURS, not the row length).  Measured on B200 at the phase-field sizes (profiles/README.md):
    TRI3 d=2 (7 neighbours) 69-72 us with 2 or 4 lanes, 88-97 with 8; TETRA4 d=3 (15) 331 us with 4 or 8, 424 with 2;
    HEXA8 d=3 (27): 4 beats 8 and 2 on the polynomial steps."""
    import os

    if os.environ.get("EFB_SPMV_LANES"):  # dev/tuning knob
        return int(os.environ["EFB_SPMV_LANES"])
    deg = nnz / max(nrows, 1) / max(dof_n, 1)
    if deg <= 10:
        return 2
    if deg <= 20:
        return 4
    return 16  # HEXA8 d=3 (27 neighbours, 24 M dofs): 3.78 ms per PCG iteration with 16 lanes, 3.85 with 8, 4.36 with 4


def cheb_lanes_per_node(nnz: int, nrows: int, dof_n: int) -> int:
    """lanes per node of the BLOCK form of the node product (the polynomial steps on single-precision values: a lane takes whole
    neighbour blocks).  Measured on B200 at the phase-field sizes: TRI3 d=2 (7 neighbours) 63 us per step with 2 lanes, 74 with
    4, 120 with 8; TETRA4 d=3 (15 neighbours) 288 us with 4, 321 with 2, 325 with 8."""
    deg = nnz / max(nrows, 1) / max(dof_n, 1)
    if deg <= 10:
        return 2
    if deg <= 40:
        return 4
    return 8


def spmv(A: DeviceCsr, x: torch.Tensor, y: torch.Tensor = None, row_offset: int = 0, mask=None, partials=None):
    nrows = A.indptr.numel() - 1
    if y is None:
        y = dv.empty((nrows,))
    ng = getattr(A, "node_graph", None)
    if ng is not None and ng[2] in (1, 2, 3) and nrows % ng[2] == 0:
        adjptr, adj, d, max_deg = ng  # node-block product: the column structure is read from the node adjacency
        n_nodes = nrows // d
        _lib.call("efb_spmv_nodeblock", n_nodes, d, dv.ptr(adjptr), dv.ptr(adj), dv.ptr(A.data), dv.ptr(x), int(row_offset),
                  dv.ptr(mask), dv.ptr(y), dv.ptr(partials), lanes_per_node(A.nnz, nrows, d), dv.stream_ptr())
        return y
    _lib.call("efb_spmv_csr", nrows, A.index_bytes, dv.ptr(A.indptr), dv.ptr(A.indices), dv.ptr(A.data), dv.ptr(x), int(row_offset),
              dv.ptr(mask), dv.ptr(y), dv.ptr(partials), lanes_per_row(A.nnz, nrows), dv.stream_ptr())
    return y


def _system_struct(A: DeviceCsr, nrows, mask, inv_diag, x, r, z, Ap, partials) -> _lib.EfbPcgSystem:
    S = _lib.EfbPcgSystem()
    S.nrows = nrows
    ng = getattr(A, "node_graph", None)
    if ng is not None and ng[2] in (1, 2, 3) and nrows % ng[2] == 0:
        adjptr, adj, d, max_deg = ng
        S.kind, S.dof_n, S.index_bytes = 1, d, 0
        S.indptr, S.indices = adjptr.data_ptr(), adj.data_ptr()
        S.lanes = lanes_per_node(A.nnz, nrows, d)
    else:
        S.kind, S.dof_n, S.index_bytes = 0, 0, A.index_bytes
        S.indptr, S.indices = A.indptr.data_ptr(), A.indices.data_ptr()
        S.lanes = lanes_per_row(A.nnz, nrows)
    S.data = A.data.data_ptr()
    S.free_mask = mask.data_ptr() if mask is not None else None
    S.inv_diag, S.x, S.r, S.z, S.Ap, S.partials = (t.data_ptr() for t in (inv_diag, x, r, z, Ap, partials))
    return S


# Above this many owned dofs per rank the kernel-per-operation loop (NCCL exchanges when sharded) is the faster one: the
# iterations are bandwidth-bound and the fused kernels' last-CTA folds and one-wave grids cost 2 % (1 GPU, 24 M dofs) to 12 %
# (8 GPUs, 24 M dofs each; profiles/r1_bench_n200_8gpu.json).  Below it the sync points dominate and the fused peer-memory
# form wins (2 GPUs, 1 M dofs each: 0.82 vs 1.69 s per staggered iteration).
FUSED_MAX_DOFS = 8_000_000
# Below this many owned dofs per rank an iteration is bound by its kernel launches and grid-wide folds, not by bandwidth: the
# single-reduction (Chronopoulos-Gear) form runs two kernels and one all-reduce per iteration instead of three and two —
# 29.6 vs 60.6 us per iteration at 0.25 M dofs on one B200, equal at 2 M dofs (profiles/r2_pcg_small_systems.log).  Its
# recurrences carry one more rounding error per step, so `single_reduction="auto"` keeps the classic form for tol < 1e-10.
SINGLE_REDUCTION_MAX_DOFS = 1_500_000
# Chebyshev-Jacobi polynomial preconditioner (`precond_degree`): z = q(D^-1 A) D^-1 r with q the Chebyshev polynomial of degree
# m - 1 on [lmax / CHEB_RATIO, lmax], lmax = CHEB_SAFETY x a power-iteration estimate of the largest eigenvalue of D^-1 A.
# m = 4 cuts the iterations ~3.6x for 1.1x the products (TRI3 / TETRA4 / HEXA8 elasticity with a degraded band,
# profiles/README.md): fewer all-reduces and grid-wide sync points, which is what bounds small and strong-scaled systems.
CHEB_DEGREE = 4
CHEB_RATIO = 30.0
CHEB_SAFETY = 1.25      # on the power-iteration estimate (converges from below; a stale or low lmax makes the polynomial indefinite)
CHEB_POWER_ITERS = 10   # per solve: the matrix of a staggered / Newton loop changes between solves, and so does lmax
# "auto" picks the polynomial for shards of at most this many matrix entries (a product of <= ~70 us): above, an iteration is
# bound by the product itself and the polynomial's 1.1x products + heavier epilogue cost more than the saved reductions
# (HEXA8 3 M dofs on one B200: 0.60 s against 0.49 s with plain Jacobi; TETRA4 2.6 M dofs per rank on two: 0.69 against 0.62 s)
CHEB_MAX_NNZ = 32_000_000
# the m - 1 products INSIDE the polynomial read the matrix values in single precision (efb_cast_f32 once per solve): a polynomial
# in fl32(A) is still a fixed symmetric operator, i.e. a valid preconditioner, and those products move half the bytes; the outer
# product, the residual, every vector and every dot product stay FP64.  With it the polynomial pays at every size.
CHEB_FP32 = True
CHEB_STALL_ITERS = 400  # outer iterations in which |r| has not halved: the polynomial is dropped for plain Jacobi


def cheb_coefficients(degree: int, lmin: float, lmax: float):
    """(theta, [(c1_k, c2_k) for k = 1 .. degree-1]) of the recurrence d_k = c1_k d_{k-1} + c2_k D^-1 (r - A z_k)"""
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho, out = 1.0 / sigma, []
    for _ in range(1, degree):
        rho_n = 1.0 / (2.0 * sigma - rho)
        out.append((rho_n * rho, 2.0 * rho_n / delta))
        rho = rho_n
    return theta, out


def estimate_lmax(A: DeviceCsr, inv_diag, mask, comm=None, iters: int = CHEB_POWER_ITERS) -> float:
    """largest eigenvalue of D^-1 A on the free dofs by power iteration (set-up of the polynomial preconditioner; the estimate
    converges from below — the caller applies CHEB_SAFETY).  Everything stays on the device until the final read."""
    dev = A.data.device
    nrows = A.indptr.numel() - 1
    n_glob = A.shape[1]
    rank = 0 if comm is None else int(comm.part.rank)
    i = torch.arange(nrows, dtype=torch.float64, device=dev)
    v = torch.frac(torch.sin(i * 12.9898 + 78.233 * (rank + 1)) * 43758.5453) * 2.0 - 1.0  # deterministic, rank-dependent
    if mask is not None:
        v = v * mask.to(torch.float64)
    full = torch.zeros(n_glob, dtype=torch.float64, device=dev)
    w = dv.empty((nrows,))
    t = torch.zeros(1, dtype=torch.float64, device=dev)
    lam = torch.ones(1, dtype=torch.float64, device=dev)

    def norm(x):
        t[0] = x @ x
        if comm is not None:
            comm.all_reduce_sum(t)
        return torch.sqrt(t)

    nv = norm(v)
    for _ in range(iters):
        full[:nrows] = v / torch.clamp(nv, min=1e-300)
        if comm is not None:
            comm.halo_exchange(full)
        spmv(A, full, w, 0, mask)
        v = w * inv_diag
        nv = norm(v)
        lam = nv.clone()  # |D^-1 A u| with |u| = 1
    out = float(lam.item())
    return out if out > 0.0 and out == out else 1.0


def pcg(A: DeviceCsr, b, x0=None, free_mask=None, tol: float = 1e-8, maxiter: int = None, check_every: int = 25, comm=None,
        fused="auto", persistent: bool = False, single_reduction="auto", precond_degree="auto", reuse_setup: bool = False):
    """Solve A x = b on the free dofs (free_mask True / 1 = unknown; other entries of x keep the values of x0).

    A holds the owned rows in LOCAL numbering `[owned | halo]` columns.  Stops when ||r|| <= tol * ||b - A x_known||
    (both restricted to free dofs).  Returns (x_owned, info dict).  `comm` (easyfea_b200.dist.RowComm) supplies the halo
    exchange and scalar all-reduces for row-sharded runs.

    `fused="auto"` (default): fused below `FUSED_MAX_DOFS` owned dofs per rank, the kernel-per-operation loop above.
    `fused=True`: the iterations run on the device, reductions and the halo exchange go through peer memory
    inside the kernels: three kernels per iteration, enqueued `check_every` at a time (`efb_pcg_iterate`).  With
    `persistent=True` (matrices assembled by this library) ONE cooperative kernel iterates until convergence instead
    (`efb_pcg_solve_persistent`: grid barriers between the steps, every rank leaves in the same iteration, no host round
    trip at all); measured equal on small systems and ~8 % slower on large ones (profiles/README.md), hence opt-in.
    `single_reduction="auto"` picks that form for small shards and tol >= 1e-10 (see SINGLE_REDUCTION_MAX_DOFS).
    `single_reduction=True`: the Chronopoulos-Gear form of the same iteration (`efb_pcg_iterate_cg2`) — one all-reduce, two
    kernels and two cross-GPU sync points per iteration instead of two, three and three; one more vector pass.  Opt-in:
    measured 13 % / 8 % faster at 0.2 M dofs on 1 / 2 GPUs and 1-3 % slower at >= 2 M dofs per rank, identical iteration
    counts on the phase-field systems, but 18 % SLOWER on config 4 over 8 GPUs (0.33 vs 0.28 s per staggered iteration,
    one measurement each, profiles/README.md) — not understood yet, so the classic form stays the default.
    `precond_degree` = m: Chebyshev-Jacobi polynomial preconditioner of degree m - 1 in D^-1 A (m = 1: plain Jacobi); "auto" =
    CHEB_DEGREE unless the single-reduction or persistent form is forced.  Inside the fused iterations
    (`efb_pcg_iterate_cheb`) the m - 1 extra products cost one neighbour halo flag each and no all-reduce.
    `reuse_setup=True`: the caller guarantees that `A` and the mask are the ones of its previous call with this flag (time
    stepping on a constant matrix): the spectral bound and the single-precision copy of the values are taken from `A`.
    `fused=False` keeps
    one collective call per exchange (NCCL through torch.distributed) and one kernel per vector operation — the baseline
    the fused path is measured against (bench.py) and checked against (tests).
    """
    dev = A.data.device
    nrows = A.indptr.numel() - 1
    n_glob = A.shape[1]
    big = None

    def largest_shard():
        nonlocal big
        if big is None:  # the same decision on every rank: the largest shard decides
            big = torch.tensor([float(nrows)], dtype=torch.float64, device=dev)
            if comm is not None:
                comm.all_reduce_max(big)
            big = float(big.item())
        return big

    if fused == "auto":
        fused = bool(single_reduction is True or persistent) or largest_shard() <= FUSED_MAX_DOFS
    if precond_degree == "auto":  # an explicitly requested single-reduction / persistent form keeps plain Jacobi
        if single_reduction is True or persistent or not fused:
            precond_degree = 1
        elif CHEB_FP32:  # inner products on single-precision matrix values: a gain at every size (half the matrix traffic)
            precond_degree = CHEB_DEGREE
        else:
            t = torch.tensor([float(A.nnz)], dtype=torch.float64, device=dev)
            if comm is not None:
                comm.all_reduce_max(t)
            precond_degree = CHEB_DEGREE if float(t.item()) <= CHEB_MAX_NNZ else 1
    degree = max(1, int(precond_degree))
    if degree > 1:
        single_reduction, persistent = False, False
    if single_reduction == "auto":
        single_reduction = bool(fused) and not persistent and tol >= 1e-10 and largest_shard() <= SINGLE_REDUCTION_MAX_DOFS
    single_reduction = bool(single_reduction) and bool(fused)
    st = dv.stream_ptr
    b = dv.to_device(b)
    assert b.numel() == nrows
    mask = None
    if free_mask is not None:
        mask = dv.to_device(free_mask).to(torch.uint8).contiguous()
        assert mask.numel() == nrows
    if maxiter is None:
        # rank-independent default: every rank must leave the loop in the same iteration (collectives / peer flags pair up)
        if comm is not None:
            t = torch.tensor([float(n_glob)], dtype=torch.float64, device=dev)
            comm.all_reduce_max(t)
            maxiter = 10 * int(t.item())
        else:
            maxiter = 10 * n_glob
    maxiter = int(maxiter)
    P = _lib.load().efb_pcg_partials_size()
    partials = dv.empty((P,))
    scal = torch.zeros(6, dtype=torch.float64, device=dev)  # [rz, pAp, rz_new, rr, -, -]
    s_ptr = lambda i: ctypes.c_void_p(scal.data_ptr() + 8 * i)  # noqa: E731

    # vectors: p over [owned | halo] (the halo entries belong to other ranks), the rest owned only
    ws = None
    if fused:
        if comm is not None:
            ws = comm.pcg_workspace()
            assert ws.n == n_glob, "the communicator was planned for another vector length"
        else:
            from .dist import LocalWorkspace

            ws = LocalWorkspace(n_glob)
    x_full = torch.zeros(n_glob, dtype=torch.float64, device=dev)
    x = x_full[:nrows]
    if x0 is not None:
        x.copy_(dv.to_device(x0).reshape(-1)[:nrows])  # a local [owned | halo] vector may be passed: the owned part counts
    p_full = ws.p[0] if ws is not None else torch.zeros(n_glob, dtype=torch.float64, device=dev)
    p = p_full[:nrows]
    r, z, Ap = dv.empty((nrows,)), dv.empty((nrows,)), dv.empty((nrows,))

    diag = dv.empty((nrows,))
    _lib.call("efb_csr_diagonal", nrows, 0, A.index_bytes, dv.ptr(A.indptr), dv.ptr(A.indices), dv.ptr(A.data), dv.ptr(diag), st())
    inv_diag = dv.empty((nrows,))
    _lib.call("efb_pcg_inv_diag", nrows, dv.ptr(diag), dv.ptr(mask), dv.ptr(inv_diag), st())

    def reduce_to(slot, m):
        _lib.call("efb_pcg_reduce", dv.ptr(partials), m, s_ptr(slot), st())
        if comm is not None:
            comm.all_reduce_sum(scal[slot:slot + m])

    lmin = lmax = None
    if degree > 1:
        cache = getattr(A, "_cheb_setup", None) if reuse_setup else None
        if cache is not None and cache[0] == (A.data.data_ptr(), nrows):
            lmax = cache[1]
        else:
            lmax = CHEB_SAFETY * estimate_lmax(A, inv_diag, mask, comm)
            cache = None
        lmin = lmax / CHEB_RATIO
        theta, coefs = cheb_coefficients(degree, lmin, lmax)
        d_vec = torch.zeros(nrows, dtype=torch.float64, device=dev)
        data32 = None
        if CHEB_FP32 and ws is not None:
            if cache is not None and cache[2] is not None:
                data32 = cache[2]
            else:
                data32 = torch.empty(A.data.numel(), dtype=torch.float32, device=dev)
                _lib.call("efb_cast_f32", A.data.numel(), dv.ptr(A.data), dv.ptr(data32), st())
        if reuse_setup:
            A._cheb_setup = ((A.data.data_ptr(), nrows), lmax, data32)
        z_full = ws.zb[0] if ws is not None else torch.zeros(n_glob, dtype=torch.float64, device=dev)

        def cheb_apply(r_vec):
            """z_full[:nrows] = q(D^-1 A) D^-1 r (kernel-per-operation form: set-up of the fused loop, body of the unfused one)"""
            z_full.zero_()
            d_vec.zero_()
            zo = z_full[:nrows]
            _lib.call("efb_pcg_cheb_update", nrows, dv.ptr(r_vec), None, dv.ptr(inv_diag), 0.0, 1.0 / theta, dv.ptr(d_vec), dv.ptr(zo), st())
            for c1, c2 in coefs:
                if comm is not None:
                    comm.halo_exchange(z_full)
                spmv(A, z_full, Ap, 0, mask)
                _lib.call("efb_pcg_cheb_update", nrows, dv.ptr(r_vec), dv.ptr(Ap), dv.ptr(inv_diag), c1, c2, dv.ptr(d_vec), dv.ptr(zo), st())
            return zo

    # reference norm: || b - A x_known || on the free dofs (x_known = x0 on constrained dofs, 0 elsewhere)
    if mask is not None:
        xk = torch.zeros_like(x_full)
        xk[:nrows] = x * (1 - mask.to(torch.float64))
        if comm is not None:
            comm.halo_exchange(xk)
        spmv(A, xk, Ap)
        _lib.call("efb_pcg_init", nrows, dv.ptr(b), dv.ptr(Ap), dv.ptr(inv_diag), dv.ptr(mask), dv.ptr(r), dv.ptr(z), dv.ptr(p),
                  dv.ptr(partials), st())
        reduce_to(2, 2)
        bnorm2 = float(scal[3].item())
    else:
        _lib.call("efb_pcg_dot", nrows, dv.ptr(b), dv.ptr(b), dv.ptr(partials), st())
        reduce_to(3, 1)
        bnorm2 = float(scal[3].item())
    if bnorm2 == 0.0:
        return x.clone(), {"iterations": 0, "rel_residual": 0.0, "converged": True, "rhs_norm": 0.0}

    # initial residual with the actual start vector
    if comm is not None:
        comm.halo_exchange(x_full)
    spmv(A, x_full, Ap)
    _lib.call("efb_pcg_init", nrows, dv.ptr(b), dv.ptr(Ap), dv.ptr(inv_diag), dv.ptr(mask), dv.ptr(r), dv.ptr(z), dv.ptr(p),
              dv.ptr(partials), st())
    reduce_to(2, 2)  # scal[2] = r.z, scal[3] = r.r
    if degree > 1:  # z_0 = q(D^-1 A) D^-1 r_0, p_0 = z_0, r.z with the polynomial preconditioner
        z0 = cheb_apply(r)
        z.copy_(z0)
        p.copy_(z0)
        _lib.call("efb_pcg_dot", nrows, dv.ptr(r), dv.ptr(z), dv.ptr(partials), st())
        reduce_to(2, 1)
    scal[0:1].copy_(scal[2:3])
    it = 0
    rr = float(scal[3].item())
    target = tol * tol * bnorm2
    if ws is not None:
        # p_0 complete on [owned | halo] in buffer 0, r.z of iteration 0 in the control block; then only kernels
        if comm is not None:
            comm.halo_exchange(p_full)
        ws.ctrl[ws.rz_off:ws.rz_off + 1].copy_(scal[2:3])
        ws.ctrl[ws.rr_off:ws.rr_off + 1].copy_(scal[3:4])
        S = _system_struct(A, nrows, mask, inv_diag, x, r, z, Ap, partials)
        use_persistent = bool(persistent) and S.kind == 1 and not single_reduction
        if single_reduction:
            # z_0 already sits in buffer 0 (the classic start has p_0 = z_0) with its halo; w_0 = A z_0 with z_0.w_0, p = s = 0
            spmv(A, p_full, Ap, 0, mask, partials)
            reduce_to(1, 1)  # scal[1] = z.Az
            ws.ctrl[ws.cg2_init_off:ws.cg2_init_off + 1].copy_(scal[2:3])
            ws.ctrl[ws.cg2_init_off + 1:ws.cg2_init_off + 2].copy_(scal[1:2])
            ws.ctrl[ws.cg2_init_off + 2:ws.cg2_init_off + 3].copy_(scal[3:4])
            z.zero_()  # the `z` vector of the system holds p in this form
            s_vec = torch.zeros(nrows, dtype=torch.float64, device=dev)
            S.s = s_vec.data_ptr()
        best_rr, best_it, stalled = rr, 0, False
        while rr > target and it < maxiter:
            if single_reduction:
                k = min(int(check_every), maxiter - it)
                _lib.call("efb_pcg_iterate_cg2", ctypes.byref(S), ctypes.byref(ws.peer), k, it, st())
                rr, err, _ = ws.status()
                ws.advance(k, 1)
                it += k
                if err:
                    raise _lib.EfbError("PCG: a wait on a neighbour rank timed out (peer process lost?)")
                if rr != rr:
                    raise _lib.EfbError("PCG broke down (NaN residual): matrix not SPD on the free dofs?")
                continue
            if degree > 1:
                k = min(int(check_every), maxiter - it)
                _lib.call("efb_pcg_iterate_cheb", ctypes.byref(S), ctypes.byref(ws.peer), k, it, degree, float(lmin), float(lmax), dv.ptr(d_vec),
                          dv.ptr(data32), cheb_lanes_per_node(A.nnz, nrows, S.dof_n) if (S.kind == 1 and data32 is not None) else 0, st())
                rr, err, _ = ws.status()
                ws.advance(k, 2, degree)
                it += k
                if err:
                    raise _lib.EfbError("PCG: a wait on a neighbour rank timed out (peer process lost?)")
                if rr == rr and rr < 0.25 * best_rr:  # |r| halved since the last mark
                    best_rr, best_it = rr, it
                if rr != rr or it - best_it >= CHEB_STALL_ITERS:  # lmax underestimated: the polynomial is not positive
                    stalled = True
                    break
                continue
            if use_persistent:
                _lib.call("efb_pcg_solve_persistent", ctypes.byref(S), ctypes.byref(ws.peer), it, maxiter - it, float(target), st())
                rr, err, k = ws.status()
                if k == 0 and not err:
                    raise _lib.EfbError("PCG: the persistent kernel made no progress")
            else:
                k = min(int(check_every), maxiter - it)
                _lib.call("efb_pcg_iterate", ctypes.byref(S), ctypes.byref(ws.peer), k, it, st())
                rr, err, _ = ws.status()
            ws.advance(k)
            it += k
            if err:
                flags = ws.ctrl[:16].cpu().numpy().view("uint64").tolist()
                raise _lib.EfbError(f"PCG: a wait on a neighbour rank timed out (peer process lost?) at iteration <= {it}: reduction "
                                    f"flags {flags[:8]}, halo flags {flags[8:]}, expected sequence numbers {ws.peer.ar_seq} / "
                                    f"{ws.peer.halo_seq}")
            if rr != rr:
                raise _lib.EfbError("PCG broke down (NaN residual): matrix not SPD on the free dofs?")
    else:
        best_rr, best_it, stalled = rr, 0, False
        while rr > target and it < maxiter:
            for _ in range(min(int(check_every), maxiter - it)):
                if comm is not None:
                    comm.halo_exchange(p_full)
                spmv(A, p_full, Ap, 0, mask, partials)
                reduce_to(1, 1)  # pAp
                _lib.call("efb_pcg_update_xr", nrows, s_ptr(0), s_ptr(1), dv.ptr(p), dv.ptr(Ap), dv.ptr(x), dv.ptr(r), dv.ptr(inv_diag),
                          dv.ptr(mask), dv.ptr(z), dv.ptr(partials), st())
                reduce_to(2, 2)  # rz_new, rr
                if degree > 1:
                    z.copy_(cheb_apply(r))
                    _lib.call("efb_pcg_dot", nrows, dv.ptr(r), dv.ptr(z), dv.ptr(partials), st())
                    reduce_to(2, 1)  # rz_new with the polynomial preconditioner
                _lib.call("efb_pcg_update_p", nrows, s_ptr(2), s_ptr(0), dv.ptr(z), dv.ptr(mask), dv.ptr(p), st())
                scal[0:1].copy_(scal[2:3])
                it += 1
            rr = float(scal[3].item())
            if degree > 1:
                if rr == rr and rr < 0.25 * best_rr:
                    best_rr, best_it = rr, it
                if rr != rr or it - best_it >= CHEB_STALL_ITERS:
                    stalled = True
                    break
            if rr != rr:
                raise _lib.EfbError("PCG broke down (NaN residual): matrix not SPD on the free dofs?")
    if degree > 1 and stalled:
        # continue from the best we have with plain Jacobi (x holds the current iterate; a NaN iterate restarts from x0)
        import warnings

        warnings.warn(f"PCG: the Chebyshev-Jacobi preconditioner stalled after {it} iterations (lmax = {lmax:.3g} too small?); "
                      "continuing with plain Jacobi", RuntimeWarning)
        xs = x if bool(torch.isfinite(x).all()) else (dv.to_device(x0).reshape(-1)[:nrows] if x0 is not None else torch.zeros_like(x))
        # (the reference norm depends on the constrained entries of the start vector only: the same tolerance applies)
        x2, info2 = pcg(A, b, x0=xs.clone(), free_mask=free_mask, tol=tol, maxiter=max(maxiter - it, 1), check_every=check_every, comm=comm,
                        fused=fused, precond_degree=1)
        info2["iterations"] += it
        info2["precond_degree"] = degree
        info2["fell_back_to_jacobi"] = True
        return x2, info2
    rel = (rr / bnorm2) ** 0.5
    return x.clone(), {"iterations": it, "rel_residual": rel, "converged": rel <= tol, "rhs_norm": bnorm2 ** 0.5, "fused": ws is not None,
                       "persistent": ws is not None and use_persistent, "single_reduction": ws is not None and bool(single_reduction),
                       "precond_degree": degree, "lmax": lmax, "precond_fp32": bool(degree > 1 and ws is not None and CHEB_FP32)}
