// Jacobi-preconditioned CG building blocks on a (row-sharded) CSR matrix — the consumer of the assembled system
// named by the north star (the reference dispatches to scipy/pypardiso/PETSc in Simulations/Solvers.py:225-394).
//
// All scalars (r.z, p.Ap, ...) stay on the device; every dot product is a fixed-order two-stage reduction
// (per-CTA partials in a fixed grid, then one CTA), hence deterministic.  Dirichlet dofs are handled by masking
// (projected CG): the search direction is zero on constrained rows, so A_UU is never extracted.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// neighbour blocks prefetched per lane in the block form of the node product, by dof_n (A/B on B200, profiles/README.md)
#ifndef EFB_BLK_PF64
#define EFB_BLK_PF64(D) ((D) == 3 ? 1 : ((D) == 2 ? EFB_BLK_PF64_D2 : EFB_BLK_PF_D1))
#endif
#ifndef EFB_BLK_PF32
#define EFB_BLK_PF32(D) ((D) == 3 ? 2 : ((D) == 2 ? EFB_BLK_PF32_D2 : EFB_BLK_PF_D1))
#endif
#ifndef EFB_BLK_PF64_D2
#define EFB_BLK_PF64_D2 3  // TRI3 d=2, 2 lanes: 70.0 -> 66.7 us (4: 77 us)
#endif
#ifndef EFB_BLK_PF32_D2
#define EFB_BLK_PF32_D2 4  // single precision: 62.1 -> 59.2 us
#endif
#ifndef EFB_BLK_PF_D1
#define EFB_BLK_PF_D1 4    // scalar systems (damage): 36.1 / 42.6 -> 35.2 / 36.0 us
#endif
#ifndef EFB_SPMV_MINB
#define EFB_SPMV_MINB 4  // resident CTAs per SM of the product kernels: the block form wants 64 registers
#endif
#ifndef EFB_SPMV_REST_UNROLL
#define EFB_SPMV_REST_UNROLL 2  // columns of a row beyond the prefetched ones: loads of this many steps are in flight together
#endif
#define EFB_STR_(x) #x
#define EFB_UNROLL_N(n) _Pragma(EFB_STR_(unroll n))

namespace efb {

constexpr int kRedBlocks = 1184;  // 148 SMs x 8: fixed grid for every reduction producer
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
    // fixed-order tree over 256 threads
    red[threadIdx.x] = v;
    __syncthreads();
    for (int w = kRedThreads / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    const double s = red[0];
    __syncthreads();
    return s;
}

// what happens to a finished row of the product (called by the lane that holds the row sum; masked rows arrive as 0):
// the default stores it and returns the row's part of x_row . y
struct EpiStore {
    const double* __restrict__ x;
    long long x_row_offset;
    double* __restrict__ y;
    bool want_dot;
    __device__ __forceinline__ double operator()(long long r, double v) const {
        y[r] = v;
        return want_dot ? x[x_row_offset + r] * v : 0.0;
    }
};

// y[r] = sum_k data[k] x[indices[k]] for local rows; LPR lanes cooperate on one row; returns the thread's part of x_row . y
template <class IDX, int LPR, class EPI, class VT = double>
__device__ __forceinline__ double spmv_rows_epi(long long nrows, const IDX* __restrict__ indptr, const IDX* __restrict__ indices,
                                                const VT* __restrict__ data, const double* __restrict__ x,
                                                const unsigned char* __restrict__ row_mask, EPI& epi) {
    const int lane = threadIdx.x % LPR;
    constexpr int RPB = kRedThreads / LPR;  // rows per CTA per pass
    double local = 0.0;
    // `base` is CTA-uniform, so every lane of a row group executes the same number of shuffle steps
    for (long long base = (long long)blockIdx.x * RPB; base < nrows; base += (long long)gridDim.x * RPB) {
        const long long r = base + threadIdx.x / LPR;
        const bool live = r < nrows;
        double s = 0.0;
        if (live && (!row_mask || row_mask[r])) {
            const long long k0 = indptr[r], k1 = indptr[r + 1];
            for (long long k = k0 + lane; k < k1; k += LPR) s += (double)data[k] * x[indices[k]];
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off, LPR);
        if (live && lane == 0) local += epi(r, s);
    }
    return local;
}

template <class IDX, int LPR>
__device__ __forceinline__ double spmv_rows(long long nrows, const IDX* __restrict__ indptr, const IDX* __restrict__ indices,
                                            const double* __restrict__ data, const double* __restrict__ x, long long x_row_offset,
                                            const unsigned char* __restrict__ row_mask, double* __restrict__ y, bool want_dot) {
    EpiStore epi{x, x_row_offset, y, want_dot};
    return spmv_rows_epi<IDX, LPR>(nrows, indptr, indices, data, x, row_mask, epi);
}

template <class IDX, int LPR>
__global__ void __launch_bounds__(kRedThreads)
    k_spmv(long long nrows, const IDX* __restrict__ indptr, const IDX* __restrict__ indices, const double* __restrict__ data,
           const double* __restrict__ x, long long x_row_offset, const unsigned char* __restrict__ row_mask, double* __restrict__ y,
           double* __restrict__ dot_partials) {
    __shared__ double red[kRedThreads];
    const double local = spmv_rows<IDX, LPR>(nrows, indptr, indices, data, x, x_row_offset, row_mask, y, dot_partials != nullptr);
    if (dot_partials) {
        const double tot = block_sum(local, red);
        if (threadIdx.x == 0) dot_partials[blockIdx.x] = tot;
    }
}

// Node-block form of the SpMV for matrices assembled by this library: the dof rows of node n are stored as one
// contiguous D x (D*deg) block (csr_kernels.cuh), so the column structure is the NODE adjacency (adjptr/adj, one int32 per
// D x D block instead of one index per coefficient: 9x less index traffic in 3D) and every gathered x value serves the D
// rows of the block.  LPN lanes cooperate on one node; lane l takes columns l, l + LPN, ... of all D rows (coalesced).
// Software-pipelined form: the adjacency pointers and the first PF column ids of a lane's NEXT node are loaded while the
// current node is processed, so the only dependent load left on a node's critical path is the gather of x (the plain
// form pays adjptr -> adj -> x, three DRAM latencies, per node and is latency-bound at ~55 % of HBM bandwidth).
template <int D, int LPN, int PF, class EPI, class VT = double>
__device__ __forceinline__ double spmv_nodes_pipe(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj,
                                                  const VT* __restrict__ data, const double* __restrict__ x,
                                                  const unsigned char* __restrict__ row_mask, EPI& epi) {
    const int lane = threadIdx.x % LPN;
    constexpr int NPB = kRedThreads / LPN;  // nodes per CTA per pass
    const long long stride = (long long)gridDim.x * NPB;
    double local = 0.0;
    long long n = (long long)blockIdx.x * NPB + threadIdx.x / LPN;
    long long a0 = 0, a1 = 0;
    if (n < n_nodes) {
        a0 = adjptr[n];
        a1 = adjptr[n + 1];
    }
    int col[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int l = lane + LPN * k;
        col[k] = (l < D * (int)(a1 - a0)) ? adj[a0 + l / D] : 0;
    }
    // CTA-uniform trip count: every lane of a group runs the same number of shuffle steps
    for (long long base = (long long)blockIdx.x * NPB; base < n_nodes; base += stride) {
        const long long nn = n + stride;
        long long na0 = 0, na1 = 0;
        if (nn < n_nodes) {
            na0 = adjptr[nn];
            na1 = adjptr[nn + 1];
        }
        const int rowlen = D * (int)(a1 - a0);
        const VT* blk = data + (long long)D * D * a0;
        double xv[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int l = lane + LPN * k;
            xv[k] = (l < rowlen) ? x[(long long)col[k] * D + (l % D)] : 0.0;
        }
        int ncol[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int l = lane + LPN * k;
            ncol[k] = (l < D * (int)(na1 - na0)) ? adj[na0 + l / D] : 0;
        }
        double s[D];
#pragma unroll
        for (int i = 0; i < D; ++i) s[i] = 0.0;
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int l = lane + LPN * k;
            if (l < rowlen) {
#pragma unroll
                for (int i = 0; i < D; ++i) s[i] += (double)blk[i * rowlen + l] * xv[k];
            }
        }
        if (rowlen > LPN * PF) {  // longer rows than the prefetch depth: the rest the plain way
            const int* cols = adj + a0;
            EFB_UNROLL_N(EFB_SPMV_REST_UNROLL)
            for (int l = lane + LPN * PF; l < rowlen; l += LPN) {
                const int c = l / D, j = l - c * D;
                const double xr = x[(long long)cols[c] * D + j];
#pragma unroll
                for (int i = 0; i < D; ++i) s[i] += (double)blk[i * rowlen + l] * xr;
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int off = LPN / 2; off > 0; off >>= 1) s[i] += __shfl_down_sync(0xffffffffu, s[i], off, LPN);
        if (n < n_nodes && lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const long long r = n * D + i;
                const double v = (!row_mask || row_mask[r]) ? s[i] : 0.0;
                local += epi(r, v);
            }
        }
        n = nn;
        a0 = na0;
        a1 = na1;
#pragma unroll
        for (int k = 0; k < PF; ++k) col[k] = ncol[k];
    }
    return local;
}

// Block form of the same product: a lane takes whole NEIGHBOUR NODES (c = lane, lane + LPN, ...), i.e. the D x D block of the
// matrix and the D entries of x of that neighbour: D (D + 1) loads per step instead of D + 1 — more bytes in flight per lane and
// wider contiguous pieces per row, which is what the single-precision values of the polynomial steps need (with one column per
// lane a 4-lane group reads 16 bytes of a row per instruction: the step is bound by requests, not bytes, and the halved matrix
// traffic buys nothing).  Same software pipeline: the neighbour ids of the lane's NEXT node travel while this one is processed.
template <int D, int LPN, int PF, class EPI, class VT>
__device__ __forceinline__ double spmv_nodes_blk(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj,
                                                 const VT* __restrict__ data, const double* __restrict__ x,
                                                 const unsigned char* __restrict__ row_mask, EPI& epi) {
    const int lane = threadIdx.x % LPN;
    constexpr int NPB = kRedThreads / LPN;  // nodes per CTA per pass
    const long long stride = (long long)gridDim.x * NPB;
    double local = 0.0;
    long long n = (long long)blockIdx.x * NPB + threadIdx.x / LPN;
    long long a0 = 0, a1 = 0;
    if (n < n_nodes) {
        a0 = adjptr[n];
        a1 = adjptr[n + 1];
    }
    int col[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int c = lane + LPN * k;
        col[k] = (c < (int)(a1 - a0)) ? adj[a0 + c] : 0;
    }
    for (long long base = (long long)blockIdx.x * NPB; base < n_nodes; base += stride) {  // CTA-uniform trip count
        const long long nn = n + stride;
        long long na0 = 0, na1 = 0;
        if (nn < n_nodes) {
            na0 = adjptr[nn];
            na1 = adjptr[nn + 1];
        }
        const int deg = (int)(a1 - a0), rowlen = D * deg;
        const VT* blk = data + (long long)D * D * a0;
        double xv[PF][D];
        VT bv[PF][D][D];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int c = lane + LPN * k;
            const bool on = c < deg;
            const double* xp = x + (long long)col[k] * D;
#pragma unroll
            for (int j = 0; j < D; ++j) xv[k][j] = on ? xp[j] : 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) bv[k][i][j] = on ? blk[i * rowlen + c * D + j] : (VT)0;
        }
        int ncol[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int c = lane + LPN * k;
            ncol[k] = (c < (int)(na1 - na0)) ? adj[na0 + c] : 0;
        }
        double s[D];
#pragma unroll
        for (int i = 0; i < D; ++i) s[i] = 0.0;
#pragma unroll
        for (int k = 0; k < PF; ++k)
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) s[i] += (double)bv[k][i][j] * xv[k][j];
        if (deg > LPN * PF) {  // more neighbours than the prefetch depth: the rest the plain way
            const int* cols = adj + a0;
#pragma unroll 2
            for (int c = lane + LPN * PF; c < deg; c += LPN) {
                const double* xp = x + (long long)cols[c] * D;
                double xr[D];
#pragma unroll
                for (int j = 0; j < D; ++j) xr[j] = xp[j];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) s[i] += (double)blk[i * rowlen + c * D + j] * xr[j];
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int off = LPN / 2; off > 0; off >>= 1) s[i] += __shfl_down_sync(0xffffffffu, s[i], off, LPN);
        if (n < n_nodes && lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const long long r = n * D + i;
                const double v = (!row_mask || row_mask[r]) ? s[i] : 0.0;
                local += epi(r, v);
            }
        }
        n = nn;
        a0 = na0;
        a1 = na1;
#pragma unroll
        for (int k = 0; k < PF; ++k) col[k] = ncol[k];
    }
    return local;
}

// prefetched steps per lane (A/B on B200, profiles/README.md: 2 for short rows, 3 when a whole warp works on one node; 4 and 8
// lose to register pressure)
#ifndef EFB_SPMV_PF
#define EFB_SPMV_PF(LPN) ((LPN) == 32 ? 3 : 2)
#endif

// the node product every kernel uses: block form (a lane takes whole neighbour blocks), one step prefetched for D = 3, two for
// D <= 2 (A/B on B200 at 64 registers / 4 CTAs per SM, profiles/README.md: TETRA4 d=3 454 -> 331 us = 6.1 TB/s, TRI3 d=2 76 -> 69 us)
template <int D, int LPN>
__device__ __forceinline__ double spmv_nodes(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj,
                                             const double* __restrict__ data, const double* __restrict__ x, long long x_row_offset,
                                             const unsigned char* __restrict__ row_mask, double* __restrict__ y, bool want_dot) {
    EpiStore epi{x, x_row_offset, y, want_dot};
#ifdef EFB_SPMV_COLUMN_FORM
    return spmv_nodes_pipe<D, LPN, EFB_SPMV_PF(LPN)>(n_nodes, adjptr, adj, data, x, row_mask, epi);
#else
    return spmv_nodes_blk<D, LPN, EFB_BLK_PF64(D), EpiStore, double>(n_nodes, adjptr, adj, data, x, row_mask, epi);
#endif
}

template <int D, int LPN>
__global__ void __launch_bounds__(kRedThreads, EFB_SPMV_MINB)  // block form: 64 registers (4 CTAs per SM) hold a whole neighbour block per step
    k_spmv_node(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj, const double* __restrict__ data,
                const double* __restrict__ x, long long x_row_offset, const unsigned char* __restrict__ row_mask, double* __restrict__ y,
                double* __restrict__ dot_partials) {
    __shared__ double red[kRedThreads];
    const double local = spmv_nodes<D, LPN>(n_nodes, adjptr, adj, data, x, x_row_offset, row_mask, y, dot_partials != nullptr);
    if (dot_partials) {
        const double tot = block_sum(local, red);
        if (threadIdx.x == 0) dot_partials[blockIdx.x] = tot;
        if (blockIdx.x == 0)  // one-wave grid: efb_pcg_reduce folds the whole partials array, absent CTAs count as zero
            for (int b = gridDim.x + threadIdx.x; b < kRedBlocks; b += kRedThreads) dot_partials[b] = 0.0;
    }
}


template <int D>
static int launch_spmv_node(long long n_nodes, const long long* adjptr, const int* adj, const double* data, const double* x,
                            long long x_row_offset, const unsigned char* row_mask, double* y, double* partials, int lpn, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define EFB_SPMVN(L)                                                                                                   \
    {                                                                                                                  \
        int per_sm = 1;                                                                                                \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spmv_node<D, L>, kRedThreads, 0);                     \
        const int g = min(kRedBlocks, sms * max(per_sm, 1));                                                           \
        k_spmv_node<D, L><<<g, kRedThreads, 0, st>>>(n_nodes, adjptr, adj, data, x, x_row_offset, row_mask, y, partials); \
    }
    switch (lpn) {
        case 2: EFB_SPMVN(2); break;
        case 4: EFB_SPMVN(4); break;
        case 8: EFB_SPMVN(8); break;
        case 16: EFB_SPMVN(16); break;
        default: EFB_SPMVN(32); break;
    }
#undef EFB_SPMVN
    return check_launch("efb_spmv_nodeblock");
}

__global__ void __launch_bounds__(kRedThreads) k_dot(long long n, const double* __restrict__ a, const double* __restrict__ b,
                                                     double* __restrict__ partials) {
    __shared__ double red[kRedThreads];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) s += a[i] * b[i];
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// out[0..m) = sum over the kRedBlocks partials of m interleaved quantities (partials[q*kRedBlocks + b])
__global__ void __launch_bounds__(kRedThreads) k_reduce_final(const double* __restrict__ partials, int m, double* __restrict__ out) {
    __shared__ double red[kRedThreads];
    for (int q = 0; q < m; ++q) {
        double s = 0.0;
        for (int b = threadIdx.x; b < kRedBlocks; b += kRedThreads) s += partials[q * kRedBlocks + b];
        const double tot = block_sum(s, red);
        if (threadIdx.x == 0) out[q] = tot;
    }
}

__global__ void k_diagonal(long long nrows, long long row_offset, int index_bytes, const void* indptr_, const void* indices_,
                           const double* __restrict__ data, double* __restrict__ diag) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double d = 0.0;
    if (index_bytes == 4) {
        const int* indptr = (const int*)indptr_;
        const int* indices = (const int*)indices_;
        for (long long k = indptr[r]; k < indptr[r + 1]; ++k)
            if (indices[k] == row_offset + r) d = data[k];
    } else {
        const long long* indptr = (const long long*)indptr_;
        const long long* indices = (const long long*)indices_;
        for (long long k = indptr[r]; k < indptr[r + 1]; ++k)
            if (indices[k] == row_offset + r) d = data[k];
    }
    diag[r] = d;
}

// x += alpha p ; r -= alpha Ap ; z = M^-1 r (masked) ; partials of (r.z, r.r)      alpha = rz / pAp (device scalars)
__global__ void __launch_bounds__(kRedThreads)
    k_update_xr(long long n, const double* __restrict__ rz, const double* __restrict__ pAp, const double* __restrict__ p,
                const double* __restrict__ Ap, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ inv_diag,
                const unsigned char* __restrict__ mask, double* __restrict__ z, double* __restrict__ partials) {
    __shared__ double red[kRedThreads];
    const double alpha = rz[0] / pAp[0];
    double s_rz = 0.0, s_rr = 0.0;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        if (mask && !mask[i]) continue;
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        const double zi = ri * inv_diag[i];
        z[i] = zi;
        s_rz += ri * zi;
        s_rr += ri * ri;
    }
    const double t0 = block_sum(s_rz, red);
    const double t1 = block_sum(s_rr, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t0;
        partials[kRedBlocks + blockIdx.x] = t1;
    }
}

// p = z + (rz_new / rz_old) p   (masked rows stay 0)
__global__ void k_update_p(long long n, const double* __restrict__ rz_new, const double* __restrict__ rz_old,
                           const double* __restrict__ z, const unsigned char* __restrict__ mask, double* __restrict__ p) {
    const double beta = rz_new[0] / rz_old[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        p[i] = z[i] + beta * p[i];
    }
}

// r = (b - y) masked ; z = M^-1 r ; p = z ; partials of (r.z, r.r)
__global__ void __launch_bounds__(kRedThreads)
    k_init_residual(long long n, const double* __restrict__ b, const double* __restrict__ Ax, const double* __restrict__ inv_diag,
                    const unsigned char* __restrict__ mask, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                    double* __restrict__ partials) {
    __shared__ double red[kRedThreads];
    double s_rz = 0.0, s_rr = 0.0;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        double ri = 0.0, zi = 0.0;
        if (!mask || mask[i]) {
            ri = b[i] - Ax[i];
            zi = ri * inv_diag[i];
        }
        r[i] = ri;
        z[i] = zi;
        p[i] = zi;
        s_rz += ri * zi;
        s_rr += ri * ri;
    }
    const double t0 = block_sum(s_rz, red);
    const double t1 = block_sum(s_rr, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t0;
        partials[kRedBlocks + blockIdx.x] = t1;
    }
}

__global__ void k_inv_diag(long long n, const double* __restrict__ diag, const unsigned char* __restrict__ mask, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (mask && !mask[i]) ? 0.0 : 1.0 / diag[i];
}

// send-buffer packing of the halo exchange: dst[i] = src[idx[i]]
__global__ void k_pack(long long n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

template <class IDX>
static int launch_spmv(long long nrows, const void* indptr, const void* indices, const double* data, const double* x,
                       long long x_row_offset, const unsigned char* row_mask, double* y, double* partials, int lpr, cudaStream_t st) {
    const IDX* ip = (const IDX*)indptr;
    const IDX* ix = (const IDX*)indices;
#define EFB_SPMV(L) k_spmv<IDX, L><<<kRedBlocks, kRedThreads, 0, st>>>(nrows, ip, ix, data, x, x_row_offset, row_mask, y, partials)
    switch (lpr) {
        case 4: EFB_SPMV(4); break;
        case 8: EFB_SPMV(8); break;
        case 16: EFB_SPMV(16); break;
        default: EFB_SPMV(32); break;
    }
#undef EFB_SPMV
    return check_launch("efb_spmv_csr");
}

// ---------------------------------------------------------------------------------------------------------------
// Fused PCG iterations with peer-memory communication (include/easyfea_b200.h, efb_pcg_iterate)
// ---------------------------------------------------------------------------------------------------------------
struct PcgCtrl {
    unsigned long long ar_flag[EFB_MAX_RANKS];    // [src] = number of reductions rank `src` has published here
    unsigned long long halo_flag[EFB_MAX_RANKS];  // [src] = number of halo pushes rank `src` has completed into this region
    double ar_val[2][EFB_MAX_RANKS][4];           // [seq & 1][src][quantity]
    unsigned int ticket[4];                       // local: last-CTA detection, one counter per kernel
    unsigned int error;                           // local: 1 = a wait timed out (peer lost); drains every later wait
    unsigned int pad_[3];
    double rz[2];                                 // local: r.z of iteration it in rz[it & 1]
    double pAp;
    double rr;                                    // local: r.r after the last finished iteration
    unsigned long long gbar;                      // local: grid-barrier arrivals of the persistent kernel (zeroed before each launch)
    unsigned long long iters_done;                // local: iterations the last persistent launch ran
    double cg2_gam[2], cg2_alp[2];                // local, single-reduction form: r.z and alpha of iteration it-1 in [it & 1]
    double cg2_init[3];                           // local, single-reduction form: (r.z, z.Az, r.r) of the start vector
};
constexpr int kCtrlBytes = 1024;
static_assert(sizeof(PcgCtrl) <= kCtrlBytes, "control block");
constexpr long long kSpinTimeoutCycles = 20000000000LL;  // ~10 s at 1.9 GHz: a lost peer must not hang the GPU, a rank delayed
                                                         // by a host-side pause (page-in, GC) must not be mistaken for one

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
// programmatic dependent launch (no-ops unless the launch carries the stream-serialisation attribute): the next kernel of the
// stream may start its CTAs while this one runs; it must not touch data of this one before pdl_wait()
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one thread: wait until *flag >= want (flags are monotonic); false when the wait was abandoned
__device__ bool spin_until(const unsigned long long* flag, unsigned long long want, PcgCtrl* own) {
    if (ld_acquire_sys(flag) >= want) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (*(volatile unsigned int*)&own->error) return false;
        if (clock64() - t0 > kSpinTimeoutCycles) {
            atomicExch(&own->error, 1u);
            return false;
        }
        __nanosleep(40);
    }
    return true;
}

// every CTA: wait for the `world` contributions of reduction number `seq` (1-based) and add them in rank order
// (lane q of warp 0 waits for rank q; the sum over ranks is a fixed-order loop, identical on every rank)
template <int M>
__device__ __forceinline__ void gather_reduction(const efb_pcg_peer& P, PcgCtrl* own, unsigned long long seq, double (&out)[M],
                                                 double* sh) {
    if (threadIdx.x < EFB_MAX_RANKS) {
        const int q = threadIdx.x;
        if (q < P.world) {
            spin_until(&own->ar_flag[q], seq, own);
            const int buf = (int)((seq - 1) & 1);
#pragma unroll
            for (int m = 0; m < M; ++m) sh[q * M + m] = *(volatile double*)&own->ar_val[buf][q][m];
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double acc = 0.0;
        for (int q = 0; q < P.world; ++q) acc += sh[q * M + m];
        out[m] = acc;
    }
    __syncthreads();
}

// every CTA hands in its partial(s); the last one to arrive folds all of them in fixed order and stores the result into
// every rank's control block, then raises the flags (reduction number `seq`, 1-based).  Returns true in the last CTA.
template <int M>
__device__ __forceinline__ bool publish_reduction(const efb_pcg_peer& P, PcgCtrl* own, double* __restrict__ partials,
                                                  const double (&mine)[M], int ticket_id, unsigned long long seq, double* red,
                                                  double (&total)[M]) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) partials[m * kRedBlocks + blockIdx.x] = mine[m];
        __threadfence();
        is_last = atomicAdd(&own->ticket[ticket_id], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kRedThreads) s += __ldcg(&partials[m * kRedBlocks + b]);
        total[m] = block_sum(s, red);
    }
    if (threadIdx.x == 0) {
        const int buf = (int)((seq - 1) & 1);
        for (int q = 0; q < P.world; ++q) {
            PcgCtrl* c = (PcgCtrl*)P.base[q];
#pragma unroll
            for (int m = 0; m < M; ++m) *(volatile double*)&c->ar_val[buf][P.rank][m] = total[m];
        }
        fence_sys();
        for (int q = 0; q < P.world; ++q) st_release_sys(&((PcgCtrl*)P.base[q])->ar_flag[P.rank], seq);
        own->ticket[ticket_id] = 0;
    }
    return true;
}

#ifndef EFB_PCG_SPMV_MINB
#define EFB_PCG_SPMV_MINB EFB_SPMV_MINB  // like the standalone kernel
#endif
struct SpmvArgs {  // host-side bundle only; the kernels take plain __restrict__ pointers (alias analysis, load batching)
    long long n;  // rows (CSR) or nodes (node blocks)
    const void* indptr;
    const void* indices;
    const double* data;
    const unsigned char* mask;
    double* Ap;
    double* partials;
};

// (1) Ap = A p_it with the p.Ap partials; waits for the neighbours' halo entries of p_it first
template <int KIND, int A, int B>  // KIND 0: CSR, A = index bytes, B = lanes per row; KIND 1: node blocks, A = dof_n, B = lanes per node
__global__ void __launch_bounds__(kRedThreads, EFB_PCG_SPMV_MINB)
    k_pcg_spmv(long long n, const void* __restrict__ indptr, const void* __restrict__ indices, const double* __restrict__ data,
               const double* __restrict__ p, const unsigned char* __restrict__ mask, double* __restrict__ Ap,
               double* __restrict__ partials, efb_pcg_peer P, unsigned long long ar_done, unsigned long long halo_done) {
    // ar_done / halo_done: reductions / halo pushes published before this iteration (P.ar_seq + 2 k, P.halo_seq + k)
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    pdl_launch_dependents();
    if (P.n_recv > 0) {
        if (threadIdx.x == 0)
            for (int i = 0; i < P.n_recv; ++i) spin_until(&own->halo_flag[P.recv_rank[i]], halo_done, own);
        __syncthreads();
    }
    pdl_wait();
    double local;
    if constexpr (KIND == 0) {
        if constexpr (A == 4)
            local = spmv_rows<int, B>(n, (const int*)indptr, (const int*)indices, data, p, 0, mask, Ap, true);
        else
            local = spmv_rows<long long, B>(n, (const long long*)indptr, (const long long*)indices, data, p, 0, mask, Ap, true);
    } else {
        local = spmv_nodes<A, B>(n, (const long long*)indptr, (const int*)indices, data, p, 0, mask, Ap, true);
    }
    double mine[1] = {block_sum(local, red)}, total[1];
    publish_reduction<1>(P, own, partials, mine, 0, ar_done + 1ull, red, total);
}

// (2) alpha = r.z / p.Ap ; x += alpha p ; r -= alpha Ap ; z = M^-1 r ; publishes (r.z, r.r)
__global__ void __launch_bounds__(kRedThreads)
    k_pcg_update_xr(long long n, const double* __restrict__ p, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                    const double* __restrict__ Ap, const double* __restrict__ inv_diag, const unsigned char* __restrict__ mask,
                    double* __restrict__ partials, efb_pcg_peer P, long long it, unsigned long long ar_done) {
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    double pAp[1];
    gather_reduction<1>(P, own, ar_done + 1ull, pAp, red);
    const double alpha = own->rz[it & 1] / pAp[0];
    double s_rz = 0.0, s_rr = 0.0;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        if (mask && !mask[i]) continue;
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        const double zi = ri * inv_diag[i];
        z[i] = zi;
        s_rz += ri * zi;
        s_rr += ri * ri;
    }
    double mine[2], total[2];
    mine[0] = block_sum(s_rz, red);
    mine[1] = block_sum(s_rr, red);
    publish_reduction<2>(P, own, partials, mine, 1, ar_done + 2ull, red, total);
}

// (3) beta = (r.z)_new / (r.z)_old ; p_{it+1} = z + beta p_it into the other p buffer, interface entries also into the
// neighbours' halo segments; the last CTA records the scalars and raises the neighbours' halo flags
__global__ void __launch_bounds__(kRedThreads)
    k_pcg_update_p(long long n, const double* __restrict__ z, const unsigned char* __restrict__ mask, const double* __restrict__ p,
                   double* __restrict__ pn, efb_pcg_peer P, long long it, unsigned long long ar_done, unsigned long long halo_done) {
    __shared__ double red[kRedThreads];
    __shared__ bool is_last;
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    double t[2];
    pdl_launch_dependents();
    gather_reduction<2>(P, own, ar_done + 2ull, t, red);
    pdl_wait();
    const double beta = t[0] / own->rz[it & 1];
    const int nxt = (int)(it & 1) ^ 1;
    const long long stride = (long long)gridDim.x * kRedThreads, first = (long long)blockIdx.x * kRedThreads + threadIdx.x;
    // interface entries first: the neighbours wait for them
    for (int s = 0; s < P.n_send; ++s) {
        const int q = P.send_rank[s];
        double* dst = (double*)((char*)P.base[q] + P.pbuf_off[q][nxt]) + P.send_dst[s];
        const long long j0 = P.send_ptr[s], cnt = P.send_ptr[s + 1] - j0;
        for (long long j = first; j < cnt; j += stride) {
            const int i = P.send_idx[j0 + j];
            dst[j] = (mask && !mask[i]) ? 0.0 : z[i] + beta * p[i];
        }
    }
    for (long long i = first; i < n; i += stride) pn[i] = (mask && !mask[i]) ? 0.0 : z[i] + beta * p[i];
    __syncthreads();  // every thread's stores into the neighbours happen-before thread 0's fence (cumulative), hence before the flag
    if (threadIdx.x == 0) {
        if (P.n_send > 0) fence_sys();
        else __threadfence();
        is_last = atomicAdd(&own->ticket[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        own->rz[nxt] = t[0];
        own->rr = t[1];
        if (P.n_send > 0) {
            fence_sys();
            for (int s = 0; s < P.n_send; ++s)
                st_release_sys(&((PcgCtrl*)P.base[P.send_rank[s]])->halo_flag[P.rank], halo_done + 1ull);
        }
        own->ticket[2] = 0;
    }
}

using SpmvKernel = void (*)(long long, const void*, const void*, const double*, const double*, const unsigned char*, double*, double*,
                            efb_pcg_peer, unsigned long long, unsigned long long);

// ---------------------------------------------------------------------------------------------------------------
// Chebyshev-Jacobi polynomial preconditioner inside the fused iterations (efb_pcg_iterate_cheb): z = q(D^-1 A) D^-1 r with q the
// degree m-1 Chebyshev polynomial of the interval [lmin, lmax] of D^-1 A.  Same CG, m-1 more products per iteration, ~m
// times fewer iterations (profiles/README.md): the iteration count is what bounds strong-scaled shards, because every
// iteration costs two all-reduces and three grid-wide sync points whatever the shard size; the inner products need neither
// (one neighbour-to-neighbour halo flag each).  Recurrence (theta = (lmax+lmin)/2, delta = (lmax-lmin)/2, sigma = theta/delta,
// rho_0 = 1/sigma):  d_0 = D^-1 r / theta, z_1 = d_0;  rho_k = 1/(2 sigma - rho_{k-1}),
// d_k = rho_k rho_{k-1} d_{k-1} + (2 rho_k / delta) D^-1 (r - A z_k),  z_{k+1} = z_k + d_k,  k = 1 .. m-1.
// z_k lives over [owned | halo] in two more peer buffers (pbuf_off[.][2 + (k-1 & 1)]); the thread that produces an interface
// entry stores it into the neighbours' halo segments (row-wise push plan), the last CTA raises their halo flags.
// ---------------------------------------------------------------------------------------------------------------
// interface entry i of a new vector into the neighbours' buffers `which`
__device__ __forceinline__ void push_entry(const efb_pcg_peer& P, const int* __restrict__ push_id, int which, long long i, double v) {
    const int c = push_id[i];
    if (c >= 0)
        for (long long t = P.push_ptr[c]; t < P.push_ptr[c + 1]; ++t) {
            const int q = P.send_rank[P.push_nbr[t]];
            ((double*)((char*)P.base[q] + P.pbuf_off[q][which]))[P.push_pos[t]] = v;
        }
}

// the last CTA of a kernel that pushed interface entries raises the neighbours' halo flags to `value`
__device__ __forceinline__ void raise_halo_flags(const efb_pcg_peer& P, PcgCtrl* own, int ticket_id, unsigned long long value) {
    __shared__ bool last_push;
    __syncthreads();  // every thread's peer stores happen-before thread 0's fence (cumulative), hence before the flag
    if (threadIdx.x == 0) {
        fence_sys();
        last_push = atomicAdd(&own->ticket[ticket_id], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_push && threadIdx.x == 0) {
        fence_sys();
        for (int sidx = 0; sidx < P.n_send; ++sidx) st_release_sys(&((PcgCtrl*)P.base[P.send_rank[sidx]])->halo_flag[P.rank], value);
        own->ticket[ticket_id] = 0;
    }
}

// (2c) alpha = r.z / p.Ap ; x += alpha p ; r -= alpha Ap ; d_0 = D^-1 r / theta ; z_1 = d_0 (owned rows + the neighbours' halos)
__global__ void __launch_bounds__(kRedThreads)
    k_pcg_update_xr_cheb(long long n, const double* __restrict__ p, double* __restrict__ x, double* __restrict__ r, double* __restrict__ d,
                         double* __restrict__ z1, const double* __restrict__ Ap, const double* __restrict__ inv_diag,
                         const unsigned char* __restrict__ mask, double inv_theta, efb_pcg_peer P, long long it, unsigned long long ar_done,
                         unsigned long long halo_done) {
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    double pAp[1];
    pdl_launch_dependents();
    gather_reduction<1>(P, own, ar_done + 1ull, pAp, red);
    pdl_wait();
    const double alpha = own->rz[it & 1] / pAp[0];
    const int* __restrict__ push_id = P.n_send > 0 ? P.push_id : nullptr;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        double di = 0.0;
        if (!mask || mask[i]) {
            x[i] += alpha * p[i];
            const double ri = r[i] - alpha * Ap[i];
            r[i] = ri;
            di = ri * inv_diag[i] * inv_theta;
        }
        d[i] = di;
        z1[i] = di;
        if (push_id) push_entry(P, push_id, 2, i, di);
    }
    if (P.n_send > 0) raise_halo_flags(P, own, 1, halo_done + 1ull);
}

// epilogue of a Chebyshev step: row i of t = A z_k arrives; d, z_{k+1} and (last step) the partial sums of r.z and r.r
struct EpiCheb {
    const double* __restrict__ r;
    const double* __restrict__ inv_diag;
    const double* __restrict__ zin;
    double* __restrict__ d;
    double* __restrict__ zout;
    const int* __restrict__ push_id;  // nullptr: nothing to push (no neighbours, or the last step)
    const efb_pcg_peer* P;
    double c1, c2;
    int which;  // peer buffer of z_{k+1}
    double s_rz, s_rr;
    __device__ __forceinline__ double operator()(long long i, double t) {
        const double ri = r[i];
        const double di = c1 * d[i] + c2 * inv_diag[i] * (ri - t);  // inv_diag is 0 on constrained rows: d and z stay 0 there
        const double zi = zin[i] + di;
        d[i] = di;
        zout[i] = zi;
        if (push_id) push_entry(*P, push_id, which, i, zi);
        s_rz += ri * zi;
        s_rr += ri * ri;
        return 0.0;
    }
};

// (2d) Chebyshev step k: waits for the neighbours' halo entries of z_k, t = A z_k, d_k, z_{k+1}; the last step publishes (r.z, r.r)
// polynomial steps in block form at 64 registers (4 CTAs per SM) — A/B on B200 at the phase-field sizes (profiles/README.md):
// TETRA4 d=3 420 -> 288 us per step, TRI3 d=2 77 -> 59 us
#ifndef EFB_PCG_CHEB_MINB
#define EFB_PCG_CHEB_MINB 4
#endif
#ifndef EFB_PCG_CHEB_MINB
#define EFB_PCG_CHEB_MINB EFB_PCG_SPMV_MINB
#endif
template <int KIND, int A, int B, class VT>
__global__ void __launch_bounds__(kRedThreads, EFB_PCG_CHEB_MINB)
    k_pcg_cheb(long long n, const void* __restrict__ indptr, const void* __restrict__ indices, const VT* __restrict__ data,
               const double* __restrict__ zin, double* __restrict__ zout, const double* __restrict__ r, const double* __restrict__ inv_diag,
               double* __restrict__ d, const unsigned char* __restrict__ mask, double* __restrict__ partials, double c1, double c2, int which,
               int last, efb_pcg_peer P, unsigned long long ar_done, unsigned long long halo_wait) {
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    pdl_launch_dependents();
    if (P.n_recv > 0) {
        if (threadIdx.x == 0)
            for (int i = 0; i < P.n_recv; ++i) spin_until(&own->halo_flag[P.recv_rank[i]], halo_wait, own);
        __syncthreads();
    }
    pdl_wait();
    EpiCheb epi{r, inv_diag, zin, d, zout, (P.n_send > 0 && !last) ? P.push_id : nullptr, &P, c1, c2, which, 0.0, 0.0};
    if constexpr (KIND == 0) {
        if constexpr (A == 4)
            spmv_rows_epi<int, B, EpiCheb, VT>(n, (const int*)indptr, (const int*)indices, data, zin, mask, epi);
        else
            spmv_rows_epi<long long, B, EpiCheb, VT>(n, (const long long*)indptr, (const long long*)indices, data, zin, mask, epi);
    } else {
        if constexpr (sizeof(VT) == 4)  // single-precision values: block form, EFB_BLK_PF32(dof_n) neighbour blocks prefetched per lane
            spmv_nodes_blk<A, B, EFB_BLK_PF32(A), EpiCheb, VT>(n, (const long long*)indptr, (const int*)indices, data, zin, mask, epi);
        else
            spmv_nodes_pipe<A, B, EFB_SPMV_PF(B), EpiCheb, VT>(n, (const long long*)indptr, (const int*)indices, data, zin, mask, epi);
    }
    if (last) {
        double mine[2], total[2];
        mine[0] = block_sum(epi.s_rz, red);
        mine[1] = block_sum(epi.s_rr, red);
        publish_reduction<2>(P, own, partials, mine, 1, ar_done + 2ull, red, total);
    } else if (P.n_send > 0) {
        raise_halo_flags(P, own, 1, halo_wait + 1ull);
    }
}

template <class VT>
using ChebKernel = void (*)(long long, const void*, const void*, const VT*, const double*, double*, const double*, const double*, double*,
                            const unsigned char*, double*, double, double, int, int, efb_pcg_peer, unsigned long long, unsigned long long);
template <int KIND, int A, class VT>
static ChebKernel<VT> pcg_cheb_kernel(int lanes) {
    switch (lanes) {
        case 2: return k_pcg_cheb<KIND, A, 2, VT>;
        case 4: return k_pcg_cheb<KIND, A, 4, VT>;
        case 8: return k_pcg_cheb<KIND, A, 8, VT>;
        case 16: return k_pcg_cheb<KIND, A, 16, VT>;
        default: return k_pcg_cheb<KIND, A, 32, VT>;
    }
}

// matrix values in single precision for the products INSIDE the polynomial preconditioner (any fixed symmetric operator is a
// valid preconditioner; the outer product, the residual and every vector stay FP64)
__global__ void k_cast_f32(long long n, const double* __restrict__ src, float* __restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}

// unfused building block of the same recurrence: d = c1 d + c2 D^-1 (r - t), z += d  (t = A z of the caller's product)
__global__ void k_cheb_update(long long n, const double* __restrict__ r, const double* __restrict__ t, const double* __restrict__ inv_diag,
                              double c1, double c2, double* __restrict__ d, double* __restrict__ z) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double di = c1 * d[i] + c2 * inv_diag[i] * (r[i] - (t ? t[i] : 0.0));
        d[i] = di;
        z[i] += di;
    }
}


template <int KIND, int A>
static SpmvKernel pcg_spmv_kernel(int lanes) {
    switch (lanes) {
        case 2: return k_pcg_spmv<KIND, A, 2>;
        case 4: return k_pcg_spmv<KIND, A, 4>;
        case 8: return k_pcg_spmv<KIND, A, 8>;
        case 16: return k_pcg_spmv<KIND, A, 16>;
        default: return k_pcg_spmv<KIND, A, 32>;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent form: ONE kernel runs the iterations until convergence.  The three steps are separated by grid barriers
// (monotonic arrival counter in the control block; the grid is one resident wave, launched cooperatively), after a barrier
// EVERY CTA folds the per-CTA partials itself in the same fixed order, so no CTA waits for a designated reducer, and
// every CTA of every rank sees the same r.r and leaves the loop in the same iteration: no host round trip, no kernel
// launch and no launch-boundary drain per iteration (3 x ~13 us in the three-kernel form, which dominates small systems
// and strong-scaled shards).  Multi-GPU: CTA 0 stores the rank's sums into the peers' control blocks and raises their
// flags; the interface entries of p' go straight into the neighbours' buffers, as in the three-kernel form.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs arrive, all CTAs leave; `target` = arrivals expected so far (monotonic).  False when the wait was abandoned.
__device__ __forceinline__ bool grid_barrier(PcgCtrl* own, unsigned long long target) {
    __shared__ int ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&own->gbar, 1ull);
        int good = 1;
        const long long t0 = clock64();
        while (ld_acquire_gpu(&own->gbar) < target) {
            if (*(volatile unsigned int*)&own->error || clock64() - t0 > kSpinTimeoutCycles) {
                atomicExch(&own->error, 1u);
                good = 0;
                break;
            }
        }
        ok = good;
    }
    __syncthreads();
    return ok != 0;
}

// after a grid barrier: every CTA folds the G per-CTA partials of M quantities in fixed order; with several ranks CTA 0
// publishes the rank's sums (reduction number `seq`) and every CTA adds the ranks' sums in rank order
template <int M>
__device__ __forceinline__ void fold_partials(const efb_pcg_peer& P, PcgCtrl* own, const double* __restrict__ partials, int G,
                                              unsigned long long seq, double (&out)[M], double* red) {
    double mine[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double s = 0.0;
        for (int b = threadIdx.x; b < G; b += kRedThreads) s += __ldcg(&partials[m * kRedBlocks + b]);
        const double tot = block_sum(s, red);  // red[0] read by every thread inside block_sum
        mine[m] = tot;
    }
    if (P.world == 1) {
#pragma unroll
        for (int m = 0; m < M; ++m) out[m] = mine[m];
        return;
    }
    const int buf = (int)((seq - 1) & 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int q = 0; q < P.world; ++q) {
            if (q == P.rank) continue;
            PcgCtrl* c = (PcgCtrl*)P.base[q];
#pragma unroll
            for (int m = 0; m < M; ++m) *(volatile double*)&c->ar_val[buf][P.rank][m] = mine[m];
        }
        fence_sys();
        for (int q = 0; q < P.world; ++q)
            if (q != P.rank) st_release_sys(&((PcgCtrl*)P.base[q])->ar_flag[P.rank], seq);
    }
    if (threadIdx.x < EFB_MAX_RANKS) {
        const int q = threadIdx.x;
        if (q < P.world) {
            if (q == P.rank) {
#pragma unroll
                for (int m = 0; m < M; ++m) red[q * M + m] = mine[m];
            } else {
                spin_until(&own->ar_flag[q], seq, own);
#pragma unroll
                for (int m = 0; m < M; ++m) red[q * M + m] = *(volatile double*)&own->ar_val[buf][q][m];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double acc = 0.0;
        for (int q = 0; q < P.world; ++q) acc += red[q * M + m];
        out[m] = acc;
    }
    __syncthreads();
}

template <int D, int LPN>
__global__ void __launch_bounds__(kRedThreads, EFB_PCG_SPMV_MINB)
    k_pcg_persistent(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj, const double* __restrict__ data,
                     const unsigned char* __restrict__ mask, const double* __restrict__ inv_diag, double* __restrict__ x,
                     double* __restrict__ r, double* __restrict__ z, double* __restrict__ Ap, double* __restrict__ partials,
                     efb_pcg_peer P, long long it0, long long max_iters, double target_rr) {
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    const int G = (int)gridDim.x;
    const long long n = n_nodes * D;
    const long long first = (long long)blockIdx.x * kRedThreads + threadIdx.x, stride = (long long)G * kRedThreads;
    double* pa = partials;                   // p.Ap partials
    double* pb = partials + kRedBlocks;      // (r.z, r.r) partials: a CTA may still fold `pa` while a faster one writes these
    double rz_old = own->rz[it0 & 1], rr = own->rr;
    unsigned long long nbar = 0;
    long long k = 0;
    bool alive = true;
    for (; k < max_iters && alive; ++k) {
        const long long it = it0 + k;
        const double* p = (const double*)((const char*)P.base[P.rank] + P.pbuf_off[P.rank][it & 1]);
        double* pn = (double*)((char*)P.base[P.rank] + P.pbuf_off[P.rank][(it & 1) ^ 1]);
        // (1) Ap = A p, p.Ap
        if (P.n_recv > 0) {
            if (threadIdx.x == 0)
                for (int i = 0; i < P.n_recv; ++i) spin_until(&own->halo_flag[P.recv_rank[i]], P.halo_seq + (unsigned long long)k, own);
            __syncthreads();
        }
        const double local = spmv_nodes<D, LPN>(n_nodes, adjptr, adj, data, p, 0, mask, Ap, true);
        const double tot = block_sum(local, red);
        if (threadIdx.x == 0) pa[blockIdx.x] = tot;
        nbar += G;
        alive = grid_barrier(own, nbar);
        double pAp[1];
        fold_partials<1>(P, own, pa, G, P.ar_seq + 2ull * (unsigned long long)k + 1ull, pAp, red);
        // (2) x, r, z and (r.z, r.r)
        const double alpha = rz_old / pAp[0];
        double s_rz = 0.0, s_rr = 0.0;
        for (long long i = first; i < n; i += stride) {
            if (mask && !mask[i]) continue;
            x[i] += alpha * p[i];
            const double ri = r[i] - alpha * Ap[i];
            r[i] = ri;
            const double zi = ri * inv_diag[i];
            z[i] = zi;
            s_rz += ri * zi;
            s_rr += ri * ri;
        }
        const double t0 = block_sum(s_rz, red), t1 = block_sum(s_rr, red);
        if (threadIdx.x == 0) {
            pb[blockIdx.x] = t0;
            pb[kRedBlocks + blockIdx.x] = t1;
        }
        nbar += G;
        alive = grid_barrier(own, nbar) && alive;
        double t[2];
        fold_partials<2>(P, own, pb, G, P.ar_seq + 2ull * (unsigned long long)k + 2ull, t, red);
        // (3) p' = z + beta p into the other buffer, interface entries also into the neighbours' halo segments
        const double beta = t[0] / rz_old;
        const int nxt = (int)(it & 1) ^ 1;
        for (int s = 0; s < P.n_send; ++s) {
            const int q = P.send_rank[s];
            double* dst = (double*)((char*)P.base[q] + P.pbuf_off[q][nxt]) + P.send_dst[s];
            const long long j0 = P.send_ptr[s], cnt = P.send_ptr[s + 1] - j0;
            for (long long j = first; j < cnt; j += stride) {
                const int i = P.send_idx[j0 + j];
                dst[j] = (mask && !mask[i]) ? 0.0 : z[i] + beta * p[i];
            }
        }
        for (long long i = first; i < n; i += stride) pn[i] = (mask && !mask[i]) ? 0.0 : z[i] + beta * p[i];
        if (P.n_send > 0) fence_sys();  // this thread's stores into the neighbours are visible before it arrives
        nbar += G;
        alive = grid_barrier(own, nbar) && alive;
        if (P.n_send > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
            fence_sys();
            for (int s = 0; s < P.n_send; ++s)
                st_release_sys(&((PcgCtrl*)P.base[P.send_rank[s]])->halo_flag[P.rank], P.halo_seq + (unsigned long long)k + 1ull);
        }
        rz_old = t[0];
        rr = t[1];
        if (!(rr > target_rr)) {  // converged, or NaN: every CTA of every rank sees the same bits and leaves together
            ++k;
            break;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        own->rz[(it0 + k) & 1] = rz_old;
        own->rr = rr;
        own->iters_done = (unsigned long long)k;
    }
}

using PersistentKernel = void (*)(long long, const long long*, const int*, const double*, const unsigned char*, const double*, double*,
                                  double*, double*, double*, double*, efb_pcg_peer, long long, long long, double);
template <int D>
static PersistentKernel pcg_persistent_kernel(int lanes) {
    switch (lanes) {
        case 2: return k_pcg_persistent<D, 2>;
        case 4: return k_pcg_persistent<D, 4>;
        case 8: return k_pcg_persistent<D, 8>;
        case 16: return k_pcg_persistent<D, 16>;
        default: return k_pcg_persistent<D, 32>;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-reduction form (Chronopoulos & Gear 1989): the same Krylov iterates with ONE all-reduce per iteration.
//   p = z + beta p ; s = w + beta s (= A p) ; x += alpha p ; r -= alpha s ; z = M^-1 r ; w = A z ;
//   gamma' = r.z, delta = z.w (one fused reduction) ; beta = gamma'/gamma ; alpha = gamma' / (delta - beta gamma'/alpha)
// Two kernels per iteration (vector update, SpMV) and two cross-GPU sync points (the reduction, the halo of z) instead
// of three and three: strong-scaled shards are bound by those sync points.  The halo travels with z.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
    k_cg2_update(long long n, const double* __restrict__ zc, double* __restrict__ zn, const double* __restrict__ w, double* __restrict__ p,
                 double* __restrict__ s, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ inv_diag,
                 const unsigned char* __restrict__ mask, double* __restrict__ partials, efb_pcg_peer P, long long it,
                 unsigned long long ar_done, unsigned long long halo_done) {
    __shared__ double red[kRedThreads];
    __shared__ bool is_last;
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    const int nxt = (int)(it & 1) ^ 1;
    const int* __restrict__ push_id = P.n_send > 0 ? P.push_id : nullptr;
    // the thread that produces an interface entry of the new z also stores it into the halo segments of the neighbours
    auto push = [&](long long i, double zi) {
        const int c = push_id[i];
        if (c >= 0)
            for (long long t = P.push_ptr[c]; t < P.push_ptr[c + 1]; ++t) {
                const int q = P.send_rank[P.push_nbr[t]];
                ((double*)((char*)P.base[q] + P.pbuf_off[q][nxt]))[P.push_pos[t]] = zi;
            }
    };
    double v[3];  // gamma = r.z, delta = z.Az, r.r of the current iterate
    if (it == 0) {
        v[0] = own->cg2_init[0];
        v[1] = own->cg2_init[1];
        v[2] = own->cg2_init[2];
    } else {
        gather_reduction<3>(P, own, ar_done, v, red);
    }
    double beta = 0.0, alpha = v[0] / v[1];
    if (it > 0) {
        beta = v[0] / own->cg2_gam[it & 1];
        alpha = v[0] / (v[1] - beta * v[0] / own->cg2_alp[it & 1]);
    }
    double s_rz = 0.0, s_rr = 0.0;
    for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
        if (mask && !mask[i]) {
            zn[i] = 0.0;
            if (push_id) push(i, 0.0);
            continue;
        }
        const double pi = zc[i] + beta * p[i];
        const double si = w[i] + beta * s[i];
        p[i] = pi;
        s[i] = si;
        x[i] += alpha * pi;
        const double ri = r[i] - alpha * si;
        r[i] = ri;
        const double zi = ri * inv_diag[i];
        zn[i] = zi;
        if (push_id) push(i, zi);
        s_rz += ri * zi;
        s_rr += ri * ri;
    }
    const double t0 = block_sum(s_rz, red), t1 = block_sum(s_rr, red);  // (block_sum ends with a CTA barrier: every thread's
                                                                       // peer stores happen-before thread 0's fence below)
    if (threadIdx.x == 0) {
        partials[kRedBlocks + blockIdx.x] = t0;      // folded by the SpMV kernel's last CTA together with its z.Az partials
        partials[2 * kRedBlocks + blockIdx.x] = t1;
        if (blockIdx.x == 0) {
            own->cg2_gam[(it + 1) & 1] = v[0];
            own->cg2_alp[(it + 1) & 1] = alpha;
            own->rr = v[2];
        }
    }
    if (P.n_send > 0) {  // the last CTA raises the neighbours' halo flags: every interface entry of z is in place
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_sys();
            is_last = atomicAdd(&own->ticket[2], 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (is_last && threadIdx.x == 0) {
            fence_sys();
            for (int sidx = 0; sidx < P.n_send; ++sidx) st_release_sys(&((PcgCtrl*)P.base[P.send_rank[sidx]])->halo_flag[P.rank], halo_done + 1ull);
            own->ticket[2] = 0;
        }
    }
}

// interface entries of the new z into the neighbours' halo segments; the last CTA raises their flags
__global__ void __launch_bounds__(kRedThreads) k_cg2_push(const double* __restrict__ zn, efb_pcg_peer P, int nxt, unsigned long long halo_done) {
    __shared__ bool is_last;
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    const long long stride = (long long)gridDim.x * kRedThreads, first = (long long)blockIdx.x * kRedThreads + threadIdx.x;
    for (int sidx = 0; sidx < P.n_send; ++sidx) {
        const int q = P.send_rank[sidx];
        double* dst = (double*)((char*)P.base[q] + P.pbuf_off[q][nxt]) + P.send_dst[sidx];
        const long long j0 = P.send_ptr[sidx], cnt = P.send_ptr[sidx + 1] - j0;
        for (long long j = first; j < cnt; j += stride) dst[j] = zn[P.send_idx[j0 + j]];
    }
    __syncthreads();  // every thread's stores happen-before thread 0's fence (cumulative), hence before the flag
    if (threadIdx.x == 0) {
        fence_sys();
        is_last = atomicAdd(&own->ticket[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        fence_sys();
        for (int sidx = 0; sidx < P.n_send; ++sidx) st_release_sys(&((PcgCtrl*)P.base[P.send_rank[sidx]])->halo_flag[P.rank], halo_done + 1ull);
        own->ticket[2] = 0;
    }
}

// w = A z (waits for the neighbours' halo entries of z), then ONE reduction of (r.z, z.Az, r.r): the z.Az partials of this
// kernel and the (r.z, r.r) partials the update kernel left behind (g_update CTAs)
template <int KIND, int A, int B>
__global__ void __launch_bounds__(kRedThreads, EFB_PCG_SPMV_MINB)
    k_cg2_spmv(long long n, const void* __restrict__ indptr, const void* __restrict__ indices, const double* __restrict__ data,
               const double* __restrict__ z, const unsigned char* __restrict__ mask, double* __restrict__ w, double* __restrict__ partials,
               int g_update, efb_pcg_peer P, unsigned long long ar_done, unsigned long long halo_done) {
    __shared__ double red[kRedThreads];
    __shared__ bool is_last;
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    if (P.n_recv > 0) {
        if (threadIdx.x == 0)
            for (int i = 0; i < P.n_recv; ++i) spin_until(&own->halo_flag[P.recv_rank[i]], halo_done, own);
        __syncthreads();
    }
    double local;
    if constexpr (KIND == 0) {
        if constexpr (A == 4)
            local = spmv_rows<int, B>(n, (const int*)indptr, (const int*)indices, data, z, 0, mask, w, true);
        else
            local = spmv_rows<long long, B>(n, (const long long*)indptr, (const long long*)indices, data, z, 0, mask, w, true);
    } else {
        local = spmv_nodes<A, B>(n, (const long long*)indptr, (const int*)indices, data, z, 0, mask, w, true);
    }
    const double mine = block_sum(local, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = mine;
        __threadfence();
        is_last = atomicAdd(&own->ticket[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double tot[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int cnt = m == 1 ? (int)gridDim.x : g_update;               // m: 0 r.z, 1 z.Az, 2 r.r
        const double* src = partials + (m == 1 ? 0 : (m == 0 ? 1 : 2)) * kRedBlocks;
        double acc = 0.0;
        for (int b = threadIdx.x; b < cnt; b += kRedThreads) acc += __ldcg(&src[b]);
        tot[m] = block_sum(acc, red);
    }
    if (threadIdx.x == 0) {
        const unsigned long long seq = ar_done + 1ull;
        const int buf = (int)((seq - 1) & 1);
        for (int q = 0; q < P.world; ++q) {
            PcgCtrl* c = (PcgCtrl*)P.base[q];
#pragma unroll
            for (int m = 0; m < 3; ++m) *(volatile double*)&c->ar_val[buf][P.rank][m] = tot[m];
        }
        fence_sys();
        for (int q = 0; q < P.world; ++q) st_release_sys(&((PcgCtrl*)P.base[q])->ar_flag[P.rank], seq);
        own->ticket[0] = 0;
    }
}

// end of a call: r.r of the last iterate into the control block (the host reads it between calls)
__global__ void __launch_bounds__(kRedThreads) k_cg2_finish(efb_pcg_peer P, unsigned long long seq) {
    __shared__ double red[kRedThreads];
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    double v[3];
    gather_reduction<3>(P, own, seq, v, red);
    if (threadIdx.x == 0) own->rr = v[2];
}

using Cg2SpmvKernel = void (*)(long long, const void*, const void*, const double*, const double*, const unsigned char*, double*, double*, int,
                               efb_pcg_peer, unsigned long long, unsigned long long);
template <int KIND, int A>
static Cg2SpmvKernel cg2_spmv_kernel(int lanes) {
    switch (lanes) {
        case 2: return k_cg2_spmv<KIND, A, 2>;
        case 4: return k_cg2_spmv<KIND, A, 4>;
        case 8: return k_cg2_spmv<KIND, A, 8>;
        case 16: return k_cg2_spmv<KIND, A, 16>;
        default: return k_cg2_spmv<KIND, A, 32>;
    }
}

// one wave: every CTA of the three kernels is resident at once (grid-stride bodies, no tail wave), capped by the partials layout
template <class K>
static int resident_blocks_per_sm(K kernel) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, kRedThreads, 0) != cudaSuccess || b < 1) b = 1;
    return b;
}

// out = sum_k c[k] v[k], k < m <= 4 (the time-scheme combinations of CSR value arrays and of nodal vectors; `out` may be one
// of the inputs).  _simu.py:1777-1853 (right-hand sides), :1890-1894 (A = coefK K + coefC C + coefM M), :1552-1657 (correctors)
struct LinComb {
    const double* v[4];
    double c[4];
    int m;
};
__global__ void k_lincomb(long long n, LinComb L, double* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < L.m) s += L.c[k] * L.v[k][i];
        out[i] = s;
    }
}

}  // namespace efb

using namespace efb;

extern "C" int efb_pcg_partials_size(void) { return 3 * kRedBlocks; }

extern "C" int efb_spmv_csr(int64_t nrows, int index_bytes, const void* indptr, const void* indices, const double* data,
                            const double* x, int64_t x_row_offset, const uint8_t* row_mask, double* y, double* dot_partials,
                            int lanes_per_row, void* stream) {
    if (nrows == 0) return 0;
    if (index_bytes == 4)
        return launch_spmv<int>(nrows, indptr, indices, data, x, x_row_offset, row_mask, y, dot_partials, lanes_per_row, as_stream(stream));
    if (index_bytes == 8)
        return launch_spmv<long long>(nrows, indptr, indices, data, x, x_row_offset, row_mask, y, dot_partials, lanes_per_row,
                                      as_stream(stream));
    set_error("efb_spmv_csr: index_bytes must be 4 or 8");
    return 1;
}

extern "C" int efb_spmv_nodeblock(int64_t n_nodes, int dof_n, const int64_t* adjptr, const int32_t* adj, const double* data,
                                  const double* x, int64_t x_row_offset, const uint8_t* row_mask, double* y, double* dot_partials,
                                  int lanes_per_node, void* stream) {
    if (n_nodes == 0) return 0;
    const long long* ap = (const long long*)adjptr;
    switch (dof_n) {
        case 1: return launch_spmv_node<1>(n_nodes, ap, adj, data, x, x_row_offset, row_mask, y, dot_partials, lanes_per_node, as_stream(stream));
        case 2: return launch_spmv_node<2>(n_nodes, ap, adj, data, x, x_row_offset, row_mask, y, dot_partials, lanes_per_node, as_stream(stream));
        case 3: return launch_spmv_node<3>(n_nodes, ap, adj, data, x, x_row_offset, row_mask, y, dot_partials, lanes_per_node, as_stream(stream));
        default: break;
    }
    set_error("efb_spmv_nodeblock: dof_n must be 1, 2 or 3, got %d", dof_n);
    return 1;
}

extern "C" int efb_csr_diagonal(int64_t nrows, int64_t row_offset, int index_bytes, const void* indptr, const void* indices,
                                const double* data, double* diag, void* stream) {
    if (nrows == 0) return 0;
    k_diagonal<<<(unsigned)((nrows + 255) / 256), 256, 0, as_stream(stream)>>>(nrows, row_offset, index_bytes, indptr, indices, data, diag);
    return check_launch("efb_csr_diagonal");
}

extern "C" int efb_pcg_inv_diag(int64_t n, const double* diag, const uint8_t* free_mask, double* out, void* stream) {
    if (n == 0) return 0;
    k_inv_diag<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(n, diag, free_mask, out);
    return check_launch("efb_pcg_inv_diag");
}

extern "C" int efb_pcg_dot(int64_t n, const double* a, const double* b, double* partials, void* stream) {
    k_dot<<<kRedBlocks, kRedThreads, 0, as_stream(stream)>>>(n, a, b, partials);
    return check_launch("efb_pcg_dot");
}

extern "C" int efb_pcg_reduce(const double* partials, int m, double* out, void* stream) {
    if (m < 1 || m > 2) {
        set_error("efb_pcg_reduce: m must be 1 or 2");
        return 1;
    }
    k_reduce_final<<<1, kRedThreads, 0, as_stream(stream)>>>(partials, m, out);
    return check_launch("efb_pcg_reduce");
}

extern "C" int efb_pcg_init(int64_t n, const double* b, const double* Ax, const double* inv_diag, const uint8_t* free_mask,
                            double* r, double* z, double* p, double* partials, void* stream) {
    k_init_residual<<<kRedBlocks, kRedThreads, 0, as_stream(stream)>>>(n, b, Ax, inv_diag, free_mask, r, z, p, partials);
    return check_launch("efb_pcg_init");
}

extern "C" int efb_pcg_update_xr(int64_t n, const double* rz, const double* pAp, const double* p, const double* Ap, double* x,
                                 double* r, const double* inv_diag, const uint8_t* free_mask, double* z, double* partials,
                                 void* stream) {
    k_update_xr<<<kRedBlocks, kRedThreads, 0, as_stream(stream)>>>(n, rz, pAp, p, Ap, x, r, inv_diag, free_mask, z, partials);
    return check_launch("efb_pcg_update_xr");
}

extern "C" int efb_pcg_update_p(int64_t n, const double* rz_new, const double* rz_old, const double* z, const uint8_t* free_mask,
                                double* p, void* stream) {
    k_update_p<<<kRedBlocks, 256, 0, as_stream(stream)>>>(n, rz_new, rz_old, z, free_mask, p);
    return check_launch("efb_pcg_update_p");
}

extern "C" int efb_pcg_ctrl_bytes(void) { return kCtrlBytes; }

extern "C" int efb_pcg_ctrl_layout(int32_t* out3) {  // 5 entries
    out3[0] = (int32_t)(offsetof(PcgCtrl, rz) / 8);
    out3[1] = (int32_t)(offsetof(PcgCtrl, rr) / 8);
    out3[2] = (int32_t)(offsetof(PcgCtrl, error) / 4);
    out3[3] = (int32_t)(offsetof(PcgCtrl, iters_done) / 8);
    out3[4] = (int32_t)(offsetof(PcgCtrl, cg2_init) / 8);
    return 0;
}

extern "C" int efb_pcg_iterate(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, void* stream) {
    const efb_pcg_peer& P = *peer;
    if (P.world < 1 || P.world > EFB_MAX_RANKS || P.rank < 0 || P.rank >= P.world || P.n_send < 0 || P.n_send > EFB_MAX_RANKS ||
        P.n_recv < 0 || P.n_recv > EFB_MAX_RANKS) {
        set_error("efb_pcg_iterate: bad communicator (world %d, rank %d, %d send / %d recv neighbours)", P.world, P.rank, P.n_send, P.n_recv);
        return 1;
    }
    if (sys->nrows == 0 && P.world == 1) return 0;
    cudaStream_t st = as_stream(stream);
    SpmvArgs a{0, sys->indptr, sys->indices, sys->data, sys->free_mask, sys->Ap, sys->partials};
    if (sys->kind == 1) {
        if (sys->dof_n < 1 || sys->dof_n > 3 || sys->nrows % sys->dof_n) {
            set_error("efb_pcg_iterate: node-block systems need dof_n in 1..3 dividing nrows (dof_n %d, nrows %lld)", sys->dof_n, (long long)sys->nrows);
            return 1;
        }
        a.n = sys->nrows / sys->dof_n;
    } else if (sys->kind == 0) {
        if (sys->index_bytes != 4 && sys->index_bytes != 8) {
            set_error("efb_pcg_iterate: index_bytes must be 4 or 8");
            return 1;
        }
        a.n = sys->nrows;
    } else {
        set_error("efb_pcg_iterate: kind must be 0 (CSR) or 1 (node blocks)");
        return 1;
    }
    SpmvKernel spmv_k;
    if (sys->kind == 1)
        spmv_k = sys->dof_n == 1 ? pcg_spmv_kernel<1, 1>(sys->lanes) : sys->dof_n == 2 ? pcg_spmv_kernel<1, 2>(sys->lanes) : pcg_spmv_kernel<1, 3>(sys->lanes);
    else
        spmv_k = sys->index_bytes == 4 ? pcg_spmv_kernel<0, 4>(sys->lanes) : pcg_spmv_kernel<0, 8>(sys->lanes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int g_spmv = min(kRedBlocks, sms * resident_blocks_per_sm(spmv_k));
    int g_xr = min(kRedBlocks, sms * resident_blocks_per_sm(k_pcg_update_xr));
    int g_p = min(kRedBlocks, sms * resident_blocks_per_sm(k_pcg_update_p));
    // small systems: a grid-wide fold costs time proportional to the number of CTAs that take part, so the grid shrinks with
    // the work (at least kMinRowsPerThread rows per thread, never below one CTA per SM)
    {
        static const int rows_per_thread = [] { const char* e = getenv("EFB_PCG_ROWS_PER_THREAD"); return e ? atoi(e) : 4; }();
        if (rows_per_thread > 0) {
            const long long want = (sys->nrows + (long long)kRedThreads * rows_per_thread - 1) / ((long long)kRedThreads * rows_per_thread);
            const int cap = (int)max((long long)sms, min((long long)kRedBlocks, want));
            g_xr = min(g_xr, cap);
            g_p = min(g_p, cap);
            g_spmv = min(g_spmv, max(cap, (int)min((long long)kRedBlocks, (a.n * sys->lanes + kRedThreads - 1) / kRedThreads / 2)));
        }
    }
    // dev: EFB_PCG_TIMING=1 prints the mean duration of the three kernels of this call (CUDA events between the launches)
    static const bool timing = [] { const char* e = getenv("EFB_PCG_TIMING"); return e && atoi(e) > 0; }();
    cudaEvent_t* ev = nullptr;
    if (timing) {
        ev = new cudaEvent_t[3 * n_iters + 1];
        for (int i = 0; i <= 3 * n_iters; ++i) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[0], st);
    }
    for (int k = 0; k < n_iters; ++k) {
        const long long it = it0 + k;
        const double* p = (const double*)((const char*)P.base[P.rank] + P.pbuf_off[P.rank][it & 1]);
        double* pn = (double*)((char*)P.base[P.rank] + P.pbuf_off[P.rank][(it & 1) ^ 1]);
        const unsigned long long ar_done = P.ar_seq + 2ull * (unsigned long long)k, halo_done = P.halo_seq + (unsigned long long)k;
        spmv_k<<<g_spmv, kRedThreads, 0, st>>>(a.n, a.indptr, a.indices, a.data, p, a.mask, a.Ap, a.partials, P, ar_done, halo_done);
        if (timing) cudaEventRecord(ev[3 * k + 1], st);
        k_pcg_update_xr<<<g_xr, kRedThreads, 0, st>>>(sys->nrows, p, sys->x, sys->r, sys->z, sys->Ap, sys->inv_diag, sys->free_mask,
                                                       sys->partials, P, it, ar_done);
        if (timing) cudaEventRecord(ev[3 * k + 2], st);
        k_pcg_update_p<<<g_p, kRedThreads, 0, st>>>(sys->nrows, sys->z, sys->free_mask, p, pn, P, it, ar_done, halo_done);
        if (timing) cudaEventRecord(ev[3 * k + 3], st);
    }
    if (timing) {
        cudaEventSynchronize(ev[3 * n_iters]);
        float t[3] = {0, 0, 0}, ms;
        for (int i = 0; i < 3 * n_iters; ++i) {
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            t[i % 3] += ms;
        }
        fprintf(stderr, "[efb_pcg_iterate rank %d] %d iterations, grids %d/%d/%d: spmv %.1f us, update_xr %.1f us, update_p %.1f us\n", P.rank,
                n_iters, g_spmv, g_xr, g_p, 1e3 * t[0] / n_iters, 1e3 * t[1] / n_iters, 1e3 * t[2] / n_iters);
        for (int i = 0; i <= 3 * n_iters; ++i) cudaEventDestroy(ev[i]);
        delete[] ev;
    }
    return check_launch("efb_pcg_iterate");
}

extern "C" int efb_pcg_cheb_update(int64_t n, const double* r, const double* t, const double* inv_diag, double c1, double c2, double* d,
                                   double* z, void* stream) {
    if (n == 0) return 0;
    const long long want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < kRedBlocks ? want : kRedBlocks);
    k_cheb_update<<<grid, 256, 0, as_stream(stream)>>>(n, r, t, inv_diag, c1, c2, d, z);
    return check_launch("efb_pcg_cheb_update");
}

extern "C" int efb_cast_f32(int64_t n, const double* src, float* dst, void* stream) {
    if (n == 0) return 0;
    const long long want = (n + 255) / 256;
    k_cast_f32<<<(unsigned)(want < 4 * kRedBlocks ? want : 4 * kRedBlocks), 256, 0, as_stream(stream)>>>(n, src, dst);
    return check_launch("efb_cast_f32");
}

template <class K, class... Args>
static void launch_pdl(K kernel, int grid, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kRedThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class VT>
static int pcg_iterate_cheb(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, int degree, double lmin,
                            double lmax, double* d_vec, const VT* cheb_data, int cheb_lanes, void* stream) {
    const efb_pcg_peer& P = *peer;
    if (P.world < 1 || P.world > EFB_MAX_RANKS || P.rank < 0 || P.rank >= P.world || P.n_send < 0 || P.n_send > EFB_MAX_RANKS ||
        P.n_recv < 0 || P.n_recv > EFB_MAX_RANKS) {
        set_error("efb_pcg_iterate_cheb: bad communicator (world %d, rank %d)", P.world, P.rank);
        return 1;
    }
    if (degree < 2 || degree > 16 || !(lmin > 0.0) || !(lmax > lmin) || !d_vec) {
        set_error("efb_pcg_iterate_cheb: degree in 2..16 and 0 < lmin < lmax (degree %d, [%g, %g])", degree, lmin, lmax);
        return 1;
    }
    if (P.n_send > 0 && !P.push_id) {
        set_error("efb_pcg_iterate_cheb: the communicator has no row-wise push plan");
        return 1;
    }
    if (sys->nrows == 0 && P.world == 1) return 0;
    cudaStream_t st = as_stream(stream);
    SpmvArgs a{0, sys->indptr, sys->indices, sys->data, sys->free_mask, sys->Ap, sys->partials};
    if (sys->kind == 1) {
        if (sys->dof_n < 1 || sys->dof_n > 3 || sys->nrows % sys->dof_n) {
            set_error("efb_pcg_iterate_cheb: node-block systems need dof_n in 1..3 dividing nrows");
            return 1;
        }
        a.n = sys->nrows / sys->dof_n;
    } else if (sys->kind == 0 && (sys->index_bytes == 4 || sys->index_bytes == 8)) {
        a.n = sys->nrows;
    } else {
        set_error("efb_pcg_iterate_cheb: kind must be 0 (CSR, 4- or 8-byte indices) or 1 (node blocks)");
        return 1;
    }
    SpmvKernel spmv_k;
    ChebKernel<VT> cheb_k;
    if (sys->kind == 1) {
        spmv_k = sys->dof_n == 1 ? pcg_spmv_kernel<1, 1>(sys->lanes) : sys->dof_n == 2 ? pcg_spmv_kernel<1, 2>(sys->lanes) : pcg_spmv_kernel<1, 3>(sys->lanes);
        static const int cheb_lanes_env = [] { const char* e = getenv("EFB_CHEB_LANES"); return e ? atoi(e) : 0; }();  // dev/tuning knob
        // block form (single-precision values): few lanes per node — A/B on B200: TRI3 d=2 (7 neighbours) 63 us with 2 lanes, 74 with
        // 4, 120 with 8; TETRA4 d=3 (15 neighbours) 288 us with 4, 321 with 2, 325 with 8, 521 with 16 (solver.cheb_lanes_per_node)
        const int cl = cheb_lanes_env > 0 ? cheb_lanes_env : (cheb_lanes > 0 ? cheb_lanes : (sizeof(VT) == 4 ? 4 : sys->lanes));
        cheb_k = sys->dof_n == 1 ? pcg_cheb_kernel<1, 1, VT>(cl) : sys->dof_n == 2 ? pcg_cheb_kernel<1, 2, VT>(cl) : pcg_cheb_kernel<1, 3, VT>(cl);
    } else {
        spmv_k = sys->index_bytes == 4 ? pcg_spmv_kernel<0, 4>(sys->lanes) : pcg_spmv_kernel<0, 8>(sys->lanes);
        cheb_k = sys->index_bytes == 4 ? pcg_cheb_kernel<0, 4, VT>(sys->lanes) : pcg_cheb_kernel<0, 8, VT>(sys->lanes);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int g_spmv = min(kRedBlocks, sms * resident_blocks_per_sm(spmv_k));
    int g_cheb = min(kRedBlocks, sms * resident_blocks_per_sm(cheb_k));
    int g_xr = min(kRedBlocks, sms * resident_blocks_per_sm(k_pcg_update_xr_cheb));
    int g_p = min(kRedBlocks, sms * resident_blocks_per_sm(k_pcg_update_p));
    {
        static const int rows_per_thread = [] { const char* e = getenv("EFB_PCG_ROWS_PER_THREAD"); return e ? atoi(e) : 4; }();
        if (rows_per_thread > 0) {
            const long long want = (sys->nrows + (long long)kRedThreads * rows_per_thread - 1) / ((long long)kRedThreads * rows_per_thread);
            const int cap = (int)max((long long)sms, min((long long)kRedBlocks, want));
            g_xr = min(g_xr, cap);
            g_p = min(g_p, cap);
            static const int spmv_div = [] { const char* e = getenv("EFB_PCG_SPMV_DIV"); return e ? max(atoi(e), 1) : 2; }();
            const int cap_s = max(cap, (int)min((long long)kRedBlocks, (a.n * sys->lanes + kRedThreads - 1) / kRedThreads / spmv_div));
            g_spmv = min(g_spmv, cap_s);
            g_cheb = min(g_cheb, cap_s);
        }
    }
    // coefficients of the recurrence
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double c1[16], c2[16];
    {
        double rho = 1.0 / sigma;
        for (int k = 1; k < degree; ++k) {
            const double rho_n = 1.0 / (2.0 * sigma - rho);
            c1[k] = rho_n * rho;
            c2[k] = 2.0 * rho_n / delta;
            rho = rho_n;
        }
    }
    char* own = (char*)P.base[P.rank];
    double* zb[2] = {(double*)(own + P.pbuf_off[P.rank][2]), (double*)(own + P.pbuf_off[P.rank][3])};
    // programmatic dependent launch (opt-in, EFB_PCG_PDL=1): the CTAs of kernel N+1 start (and poll their cross-GPU flags) while kernel
    // N drains; every kernel of the sequence executes griddepcontrol.wait before it touches anything kernel N wrote or reads.
    // Measured neutral on 2 x 0.25 M dofs (0.150 s per staggered iteration either way: the one-wave grids leave no room for the
    // next kernel's CTAs before this one's exit), hence off by default.
    static const bool pdl = [] { const char* e = getenv("EFB_PCG_PDL"); return e && atoi(e) != 0; }();
    // dev: EFB_PCG_TIMING=1 prints the mean duration of the kernels of this call (CUDA events between the launches)
    static const bool timing = [] { const char* e = getenv("EFB_PCG_TIMING"); return e && atoi(e) > 0; }();
    const int per_it = degree + 2;
    cudaEvent_t* ev = nullptr;
    int ne = 0;
    if (timing) {
        ev = new cudaEvent_t[per_it * n_iters + 1];
        for (int i = 0; i <= per_it * n_iters; ++i) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[ne++], st);
    }
    for (int k = 0; k < n_iters; ++k) {
        const long long it = it0 + k;
        const double* p = (const double*)(own + P.pbuf_off[P.rank][it & 1]);
        double* pn = (double*)(own + P.pbuf_off[P.rank][(it & 1) ^ 1]);
        const unsigned long long ar_done = P.ar_seq + 2ull * (unsigned long long)k;
        const unsigned long long halo_done = P.halo_seq + (unsigned long long)degree * (unsigned long long)k;
        launch_pdl(spmv_k, g_spmv, st, pdl, a.n, a.indptr, a.indices, a.data, p, a.mask, a.Ap, a.partials, P, ar_done, halo_done);
        if (timing) cudaEventRecord(ev[ne++], st);
        launch_pdl(k_pcg_update_xr_cheb, g_xr, st, pdl, (long long)sys->nrows, p, sys->x, sys->r, d_vec, zb[0], (const double*)sys->Ap,
                   sys->inv_diag, sys->free_mask, 1.0 / theta, P, it, ar_done, halo_done);
        if (timing) cudaEventRecord(ev[ne++], st);
        for (int j = 1; j < degree; ++j) {  // z_j in zb[(j-1) & 1] -> z_{j+1} in zb[j & 1]
            launch_pdl(cheb_k, g_cheb, st, pdl, a.n, a.indptr, a.indices, cheb_data, (const double*)zb[(j - 1) & 1], zb[j & 1], (const double*)sys->r,
                       sys->inv_diag, d_vec, a.mask, a.partials, c1[j], c2[j], 2 + (j & 1), j == degree - 1 ? 1 : 0, P, ar_done,
                       halo_done + (unsigned long long)j);
            if (timing) cudaEventRecord(ev[ne++], st);
        }
        // p' = z_m + beta p ; its halo push is push number `degree` of this iteration
        launch_pdl(k_pcg_update_p, g_p, st, pdl, (long long)sys->nrows, (const double*)zb[(degree - 1) & 1], sys->free_mask, p, pn, P, it, ar_done,
                   halo_done + (unsigned long long)(degree - 1));
        if (timing) cudaEventRecord(ev[ne++], st);
    }
    if (timing) {
        cudaEventSynchronize(ev[ne - 1]);
        float t[18] = {0}, ms;
        for (int i = 0; i + 1 < ne; ++i) {
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            t[i % per_it] += ms;
        }
        fprintf(stderr, "[efb_pcg_iterate_cheb rank %d] %d iterations, degree %d, grids %d/%d/%d/%d us per kernel: spmv %.1f, xr %.1f, cheb", P.rank,
                n_iters, degree, g_spmv, g_xr, g_cheb, g_p, 1e3 * t[0] / n_iters, 1e3 * t[1] / n_iters);
        for (int j = 1; j < degree; ++j) fprintf(stderr, " %.1f", 1e3 * t[1 + j] / n_iters);
        fprintf(stderr, ", update_p %.1f\n", 1e3 * t[degree + 1] / n_iters);
        for (int i = 0; i < ne; ++i) cudaEventDestroy(ev[i]);
        delete[] ev;
    }
    return check_launch("efb_pcg_iterate_cheb");
}

extern "C" int efb_pcg_iterate_cheb(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, int degree, double lmin,
                                    double lmax, double* d_vec, const float* data32, int cheb_lanes, void* stream) {
    if (cheb_lanes != 0 && cheb_lanes != 2 && cheb_lanes != 4 && cheb_lanes != 8 && cheb_lanes != 16 && cheb_lanes != 32) {
        set_error("efb_pcg_iterate_cheb: cheb_lanes must be 0 (default), 2, 4, 8, 16 or 32");
        return 1;
    }
    if (data32) return pcg_iterate_cheb<float>(sys, peer, n_iters, it0, degree, lmin, lmax, d_vec, data32, cheb_lanes, stream);
    return pcg_iterate_cheb<double>(sys, peer, n_iters, it0, degree, lmin, lmax, d_vec, sys->data, 0, stream);
}

extern "C" int efb_pcg_iterate_cg2(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, void* stream) {
    const efb_pcg_peer& P = *peer;
    if (P.world < 1 || P.world > EFB_MAX_RANKS || P.rank < 0 || P.rank >= P.world || P.n_send < 0 || P.n_send > EFB_MAX_RANKS ||
        P.n_recv < 0 || P.n_recv > EFB_MAX_RANKS) {
        set_error("efb_pcg_iterate_cg2: bad communicator (world %d, rank %d)", P.world, P.rank);
        return 1;
    }
    if (!sys->s) {
        set_error("efb_pcg_iterate_cg2: the system needs the extra vector s");
        return 1;
    }
    if (n_iters <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    long long n_spmv;
    Cg2SpmvKernel spmv_k;
    if (sys->kind == 1) {
        if (sys->dof_n < 1 || sys->dof_n > 3 || sys->nrows % sys->dof_n) {
            set_error("efb_pcg_iterate_cg2: node-block systems need dof_n in 1..3 dividing nrows");
            return 1;
        }
        n_spmv = sys->nrows / sys->dof_n;
        spmv_k = sys->dof_n == 1 ? cg2_spmv_kernel<1, 1>(sys->lanes) : sys->dof_n == 2 ? cg2_spmv_kernel<1, 2>(sys->lanes) : cg2_spmv_kernel<1, 3>(sys->lanes);
    } else if (sys->kind == 0 && (sys->index_bytes == 4 || sys->index_bytes == 8)) {
        n_spmv = sys->nrows;
        spmv_k = sys->index_bytes == 4 ? cg2_spmv_kernel<0, 4>(sys->lanes) : cg2_spmv_kernel<0, 8>(sys->lanes);
    } else {
        set_error("efb_pcg_iterate_cg2: kind must be 0 (CSR, index_bytes 4 or 8) or 1 (node blocks)");
        return 1;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int g_spmv = min(kRedBlocks, sms * resident_blocks_per_sm(spmv_k));
    const int g_upd = min(kRedBlocks, sms * resident_blocks_per_sm(k_cg2_update));
    long long n_push = P.n_send > 0 ? P.send_ptr[P.n_send] : 0;
    int g_push = (int)((n_push + kRedThreads - 1) / kRedThreads);
    if (g_push > sms) g_push = sms;
    if (g_push < 1) g_push = 1;
    for (int k = 0; k < n_iters; ++k) {
        const long long it = it0 + k;
        const int cur = (int)(it & 1), nxt = cur ^ 1;
        const double* zc = (const double*)((const char*)P.base[P.rank] + P.pbuf_off[P.rank][cur]);
        double* zn = (double*)((char*)P.base[P.rank] + P.pbuf_off[P.rank][nxt]);
        const unsigned long long ar_done = P.ar_seq + (unsigned long long)k, halo_done = P.halo_seq + (unsigned long long)k;
        // p lives in sys->z (z itself lives in the two peer buffers), w in sys->Ap
        const bool rowwise = P.n_send == 0 || P.push_id != nullptr;  // the update kernel pushes the interface entries itself
        efb_pcg_peer Pu = P;
        if (!rowwise) Pu.n_send = 0;  // no row-wise plan given: separate push kernel below
        k_cg2_update<<<g_upd, kRedThreads, 0, st>>>(sys->nrows, zc, zn, sys->Ap, sys->z, sys->s, sys->x, sys->r, sys->inv_diag, sys->free_mask,
                                                    sys->partials, Pu, it, ar_done, halo_done);
        if (!rowwise) k_cg2_push<<<g_push, kRedThreads, 0, st>>>(zn, P, nxt, halo_done);
        spmv_k<<<g_spmv, kRedThreads, 0, st>>>(n_spmv, sys->indptr, sys->indices, sys->data, zn, sys->free_mask, sys->Ap, sys->partials, g_upd, P,
                                               ar_done, halo_done + 1ull);
    }
    k_cg2_finish<<<1, kRedThreads, 0, st>>>(P, P.ar_seq + (unsigned long long)n_iters);
    return check_launch("efb_pcg_iterate_cg2");
}

extern "C" int efb_pcg_solve_persistent(const efb_pcg_system* sys, const efb_pcg_peer* peer, int64_t it0, int64_t max_iters,
                                        double target_rr, void* stream) {
    const efb_pcg_peer& P = *peer;
    if (P.world < 1 || P.world > EFB_MAX_RANKS || P.rank < 0 || P.rank >= P.world) {
        set_error("efb_pcg_solve_persistent: bad communicator (world %d, rank %d)", P.world, P.rank);
        return 1;
    }
    if (sys->kind != 1 || sys->dof_n < 1 || sys->dof_n > 3 || sys->nrows % sys->dof_n) {
        set_error("efb_pcg_solve_persistent: node-block systems only (kind 1, dof_n 1..3 dividing nrows)");
        return 1;
    }
    cudaStream_t st = as_stream(stream);
    PersistentKernel kern = sys->dof_n == 1 ? pcg_persistent_kernel<1>(sys->lanes)
                            : sys->dof_n == 2 ? pcg_persistent_kernel<2>(sys->lanes) : pcg_persistent_kernel<3>(sys->lanes);
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) {
        set_error("efb_pcg_solve_persistent: the device does not support cooperative launches");
        return 1;
    }
    const int G = min(kRedBlocks, sms * resident_blocks_per_sm(kern));
    PcgCtrl* own = (PcgCtrl*)P.base[P.rank];
    cudaError_t err = cudaMemsetAsync(&own->gbar, 0, 2 * sizeof(unsigned long long), st);  // gbar, iters_done
    if (err != cudaSuccess) {
        set_error("efb_pcg_solve_persistent: %s", cudaGetErrorString(err));
        return 1;
    }
    long long n_nodes = sys->nrows / sys->dof_n, it0_ = it0, max_ = max_iters;
    const long long* adjptr = (const long long*)sys->indptr;
    const int* adj = (const int*)sys->indices;
    const double* data = sys->data;
    const unsigned char* mask = sys->free_mask;
    const double* inv_diag = sys->inv_diag;
    double *x = sys->x, *r = sys->r, *z = sys->z, *Ap = sys->Ap, *partials = sys->partials;
    efb_pcg_peer Pv = P;
    void* args[] = {&n_nodes, &adjptr, &adj, &data, &mask, &inv_diag, &x, &r, &z, &Ap, &partials, &Pv, &it0_, &max_, &target_rr};
    err = cudaLaunchCooperativeKernel((const void*)kern, dim3(G), dim3(kRedThreads), args, 0, st);
    if (err != cudaSuccess) {
        set_error("efb_pcg_solve_persistent: cooperative launch of %d CTAs failed: %s", G, cudaGetErrorString(err));
        cudaGetLastError();
        return 1;
    }
    return 0;
}

extern "C" int efb_peer_alloc(int64_t bytes, void** ptr) {
    cudaError_t err = cudaMalloc(ptr, (size_t)bytes);
    if (err == cudaSuccess) err = cudaMemset(*ptr, 0, (size_t)bytes);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        set_error("efb_peer_alloc(%lld): %s", (long long)bytes, cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int efb_peer_free(void* ptr) {
    cudaError_t err = cudaFree(ptr);
    if (err != cudaSuccess) {
        set_error("efb_peer_free: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int efb_peer_export(void* ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    cudaError_t err = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, ptr);
    if (err != cudaSuccess) {
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int efb_peer_open(const void* handle64, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t err = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle: %s (GPUs without peer access?)", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int efb_peer_close(void* ptr) {
    cudaError_t err = cudaIpcCloseMemHandle(ptr);
    if (err != cudaSuccess) {
        set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int efb_lincomb(int64_t n, int m, const double* coefs, const double* const* vecs_host, double* out, void* stream) {
    if (m < 1 || m > 4) {
        set_error("efb_lincomb: m must be 1..4, got %d", m);
        return 1;
    }
    if (n == 0) return 0;
    LinComb L;
    L.m = m;
    for (int k = 0; k < 4; ++k) {
        L.v[k] = k < m ? vecs_host[k] : nullptr;
        L.c[k] = k < m ? coefs[k] : 0.0;
    }
    long long nblk = (n + 255) / 256;
    if (nblk > 148 * 16) nblk = 148 * 16;
    k_lincomb<<<(unsigned)nblk, 256, 0, as_stream(stream)>>>(n, L, out);
    return check_launch("efb_lincomb");
}

extern "C" int efb_pack_f64(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream) {
    if (n == 0) return 0;
    long long nblk = (n + 255) / 256;
    if (nblk > 148 * 16) nblk = 148 * 16;
    k_pack<<<(unsigned)nblk, 256, 0, as_stream(stream)>>>(n, idx, src, dst);
    return check_launch("efb_pack_f64");
}
