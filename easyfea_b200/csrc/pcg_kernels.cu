// Jacobi-preconditioned CG building blocks on a (row-sharded) CSR matrix — the consumer of the assembled system
// named by the north star (the reference dispatches to scipy/pypardiso/PETSc in Simulations/Solvers.py:225-394).
//
// All scalars (r.z, p.Ap, ...) stay on the device; every dot product is a fixed-order two-stage reduction
// (per-CTA partials in a fixed grid, then one CTA), hence deterministic.  Dirichlet dofs are handled by masking
// (projected CG): the search direction is zero on constrained rows, so A_UU is never extracted.
#include "common.cuh"

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

// y[r] = sum_k data[k] x[indices[k]] for local rows; LPR lanes cooperate on one row; optional partial of x_row . y
template <class IDX, int LPR>
__global__ void __launch_bounds__(kRedThreads)
    k_spmv(long long nrows, const IDX* __restrict__ indptr, const IDX* __restrict__ indices, const double* __restrict__ data,
           const double* __restrict__ x, long long x_row_offset, const unsigned char* __restrict__ row_mask, double* __restrict__ y,
           double* __restrict__ dot_partials) {
    __shared__ double red[kRedThreads];
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
            for (long long k = k0 + lane; k < k1; k += LPR) s += data[k] * x[indices[k]];
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off, LPR);
        if (live && lane == 0) {
            y[r] = s;
            if (dot_partials) local += x[x_row_offset + r] * s;
        }
    }
    if (dot_partials) {
        const double tot = block_sum(local, red);
        if (threadIdx.x == 0) dot_partials[blockIdx.x] = tot;
    }
}

// Node-block form of the SpMV for matrices assembled by this library: the dof rows of node n are stored as one
// contiguous D x (D*deg) block (csr_kernels.cuh), so the column structure is the NODE adjacency (adjptr/adj, one int32 per
// D x D block instead of one index per coefficient: 9x less index traffic in 3D) and every gathered x value serves the D
// rows of the block.  LPN lanes cooperate on one node; lane l takes columns l, l + LPN, ... of all D rows (coalesced).
template <int D, int LPN>
__global__ void __launch_bounds__(kRedThreads)
    k_spmv_node(long long n_nodes, const long long* __restrict__ adjptr, const int* __restrict__ adj, const double* __restrict__ data,
                const double* __restrict__ x, long long x_row_offset, const unsigned char* __restrict__ row_mask, double* __restrict__ y,
                double* __restrict__ dot_partials) {
    __shared__ double red[kRedThreads];
    const int lane = threadIdx.x % LPN;
    constexpr int NPB = kRedThreads / LPN;  // nodes per CTA per pass
    double local = 0.0;
    for (long long base = (long long)blockIdx.x * NPB; base < n_nodes; base += (long long)gridDim.x * NPB) {
        const long long n = base + threadIdx.x / LPN;
        const bool live = n < n_nodes;
        double s[D];
#pragma unroll
        for (int i = 0; i < D; ++i) s[i] = 0.0;
        if (live) {
            const long long a0 = adjptr[n];
            const int rowlen = D * (int)(adjptr[n + 1] - a0);
            const double* blk = data + (long long)D * D * a0;
            const int* cols = adj + a0;
#pragma unroll 3
            for (int l = lane; l < rowlen; l += LPN) {
                const int c = l / D, j = l - c * D;
                const double xv = x[(long long)cols[c] * D + j];
#pragma unroll
                for (int i = 0; i < D; ++i) s[i] += blk[i * rowlen + l] * xv;
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int off = LPN / 2; off > 0; off >>= 1) s[i] += __shfl_down_sync(0xffffffffu, s[i], off, LPN);
        if (live && lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const long long r = n * D + i;
                const double v = (!row_mask || row_mask[r]) ? s[i] : 0.0;
                y[r] = v;
                if (dot_partials) local += x[x_row_offset + r] * v;
            }
        }
    }
    if (dot_partials) {
        const double tot = block_sum(local, red);
        if (threadIdx.x == 0) dot_partials[blockIdx.x] = tot;
    }
}

template <int D>
static int launch_spmv_node(long long n_nodes, const long long* adjptr, const int* adj, const double* data, const double* x,
                            long long x_row_offset, const unsigned char* row_mask, double* y, double* partials, int lpn, cudaStream_t st) {
#define EFB_SPMVN(L) k_spmv_node<D, L><<<kRedBlocks, kRedThreads, 0, st>>>(n_nodes, adjptr, adj, data, x, x_row_offset, row_mask, y, partials)
    switch (lpn) {
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

}  // namespace efb

using namespace efb;

extern "C" int efb_pcg_partials_size(void) { return 2 * kRedBlocks; }

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

extern "C" int efb_pack_f64(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream) {
    if (n == 0) return 0;
    long long nblk = (n + 255) / 256;
    if (nblk > 148 * 16) nblk = 148 * 16;
    k_pack<<<(unsigned)nblk, 256, 0, as_stream(stream)>>>(n, idx, src, dst);
    return check_launch("efb_pack_f64");
}
