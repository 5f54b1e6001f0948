// C-ABI entry points of the CSR pattern build (A1) and the deterministic replay (A2).
#include "common.cuh"
#include "csr_kernels.cuh"

namespace efb {

static int make_table(GroupTable& T, int n_groups, const int32_t* const* connect, const double* const* data,
                      const int64_t* Ne, const int32_t* nPe, int dof_n) {
    if (n_groups < 1 || n_groups > kMaxGroups) {
        set_error("number of element groups must be in [1, %d], got %d", kMaxGroups, n_groups);
        return 1;
    }
    T.n = n_groups;
    T.qoff[0] = T.poff[0] = T.koff[0] = 0;
    for (int g = 0; g < n_groups; ++g) {
        T.connect[g] = connect ? connect[g] : nullptr;
        T.data[g] = data ? data[g] : nullptr;
        T.Ne[g] = Ne[g];
        T.nPe[g] = nPe[g];
        const long long ndof = (long long)nPe[g] * dof_n;
        T.qoff[g + 1] = T.qoff[g] + Ne[g] * nPe[g];
        T.poff[g + 1] = T.poff[g] + Ne[g] * nPe[g] * nPe[g];
        T.koff[g + 1] = T.koff[g] + Ne[g] * ndof * ndof;
    }
    return 0;
}

static inline unsigned blocks_for(long long n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

__global__ void k_count_node_rows(const int* connect, long long n, int* cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) count_node_rows_item(connect, i, cnt);
}

__global__ void k_fill_node_rows(const int* connect, long long n, long long qoff, const long long* rowptr, int* cursor,
                                 long long* qlist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fill_node_rows_item(connect, i, qoff, rowptr, cursor, qlist);
}

__global__ void k_sort_node_rows(long long Nn, const long long* rowptr, long long* qlist) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < Nn) sort_node_rows_item(n, rowptr, qlist);
}

__global__ void __launch_bounds__(128) k_count_adj(GroupTable T, long long Nn, const long long* rowptr, const long long* qlist,
                                                   int* deg, int* err_flag) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nn) return;
    int buf[kAdjCap];
    const int len = gather_neighbours(T, n, rowptr, qlist, buf);
    if (len < 0) {
        *err_flag = 1;
        deg[n] = 0;
    } else
        deg[n] = len;
}

__global__ void __launch_bounds__(128) k_fill_adj(GroupTable T, long long Nn, const long long* rowptr, const long long* qlist,
                                                  const long long* adjptr, int* adj) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nn) return;
    int buf[kAdjCap];
    const int len = gather_neighbours(T, n, rowptr, qlist, buf);
    int* dst = adj + adjptr[n];
    for (int k = 0; k < len; ++k) dst[k] = buf[k];
}

template <class IDX>
__global__ void k_expand_indptr(long long Ndof, long long Nn, int d, const long long* adjptr, IDX* indptr) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= Ndof) expand_indptr_item<IDX>(r, Nn, d, adjptr, indptr);
}

// one thread per adjacency entry; the owning node is found by binary search in adjptr
template <class IDX>
__global__ void k_expand_indices(long long nnz_node, long long Nn, int d, const long long* adjptr, const int* adj, IDX* indices) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz_node) return;
    long long lo = 0, hi = Nn;  // largest n with adjptr[n] <= t
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (adjptr[mid] <= t)
            lo = mid;
        else
            hi = mid;
    }
    expand_indices_item<IDX>(t, lo, d, adjptr, adj, indices);
}

__global__ void k_slot_map(const int* connect, int nPe, long long n, const long long* adjptr, const int* adj, int* pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_map_item(connect, nPe, i, adjptr, adj, pos);
}

__global__ void k_inv_map(const int* connect, int nPe, int d, long long n, const long long* adjptr, const int* pos, int* inv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv_map_item(connect, nPe, d, i, adjptr, pos, inv);
}

__global__ void k_row_has_entry(long long Ndof, long long Nn, int d, const long long* rowptr, int* has) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Ndof) return;
    const long long n = r / d;
    has[r] = (n < Nn && rowptr[n + 1] > rowptr[n]) ? 1 : 0;
}

__global__ void k_compact(long long n, const long long* ptr, const double* dense, double* out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && ptr[r + 1] > ptr[r]) out[ptr[r]] = dense[r];
}

// ---- exclusive scan int32 -> int64 (three passes, 1024 items per CTA) ------------------------------------
constexpr int kScanTile = 1024;

__global__ void __launch_bounds__(256) k_scan_tile_sums(const int* in, long long n, long long* tile_sums) {
    __shared__ long long red[256];
    const long long base = (long long)blockIdx.x * kScanTile;
    long long s = 0;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        if (i < n) s += in[i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(1024) k_scan_tiles(long long* tile_sums, long long ntiles) {
    // single CTA: exclusive scan of the tile sums in place, chunks of 1024 with a running carry
    __shared__ long long buf[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < ntiles; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < ntiles ? tile_sums[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const long long add = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < ntiles) tile_sums[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_scan_final(const int* in, long long n, const long long* tile_offsets, long long* out) {
    __shared__ long long part[256];
    const long long base = (long long)blockIdx.x * kScanTile;
    long long v[4], s = 0;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const long long add = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    long long run = tile_offsets[blockIdx.x] + part[threadIdx.x] - s;
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        if (i < n) out[i] = run;
        run += v[k];
        if (i == n - 1) out[n] = run;
    }
}

// ---- replay ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_replay_matrix(GroupTable T, int d, long long Nn, const long long* rowptr,
                                                       const long long* qlist, const long long* adjptr, const int* pos,
                                                       int acc_per_warp, double* out) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (n >= Nn) return;  // whole warp exits together
    replay_node(T, d, n, rowptr, qlist, adjptr, pos, smem + (size_t)warp * acc_per_warp, out);
}

template <int D, int NPE>
__global__ void __launch_bounds__(256) k_replay_fast(const double* data, long long Nn, const long long* rowptr, const long long* qlist,
                                                     const long long* adjptr, const int* pos, int acc_per_warp, double* out) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (n >= Nn) return;
    replay_node_fast<D, NPE, 4>(data, n, rowptr, qlist, adjptr, pos, smem + (size_t)warp * acc_per_warp, out);
}

template <int D, int NPE>
static int launch_replay_fast(const double* data, long long Nn, const long long* rowptr, const long long* qlist,
                              const long long* adjptr, const int* pos, int max_deg, double* out, cudaStream_t st) {
    const int warps = 8;
    const int acc_per_warp = D * D * max_deg;
    const size_t bytes = sizeof(double) * (size_t)acc_per_warp * warps;
    if (ensure_smem(k_replay_fast<D, NPE>, bytes)) return 1;
    k_replay_fast<D, NPE><<<blocks_for(Nn, warps), warps * 32, bytes, st>>>(data, Nn, rowptr, qlist, adjptr, pos, acc_per_warp, out);
    return check_launch("efb_csr_replay_matrix");
}

__global__ void k_replay_vector(GroupTable T, int d, long long nrows, const long long* rowptr, const long long* qlist, double* out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows) replay_vector_item(T, d, r, rowptr, qlist, out);
}

}  // namespace efb

using namespace efb;

extern "C" int efb_csr_count_node_rows(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                                       const int32_t* nPe_host, int64_t Nn, int32_t* cnt, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, 1)) return 1;
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)Nn, as_stream(stream));
    for (int g = 0; g < T.n; ++g) {
        const long long n = T.Ne[g] * T.nPe[g];
        if (n) k_count_node_rows<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(T.connect[g], n, cnt);
    }
    return check_launch("efb_csr_count_node_rows");
}

extern "C" int efb_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* workspace, void* stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(int64_t), st);
        return check_launch("efb_exclusive_scan_i32");
    }
    const long long ntiles = (n + kScanTile - 1) / kScanTile;
    long long* tiles = reinterpret_cast<long long*>(workspace);
    k_scan_tile_sums<<<(unsigned)ntiles, 256, 0, st>>>(in, n, tiles);
    k_scan_tiles<<<1, 1024, 0, st>>>(tiles, ntiles);
    k_scan_final<<<(unsigned)ntiles, 256, 0, st>>>(in, n, tiles, reinterpret_cast<long long*>(out));
    return check_launch("efb_exclusive_scan_i32");
}

extern "C" int efb_csr_fill_node_rows(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                                      const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, int32_t* cursor,
                                      int64_t* qlist, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, 1)) return 1;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)Nn, st);
    for (int g = 0; g < T.n; ++g) {
        const long long n = T.Ne[g] * T.nPe[g];
        if (n)
            k_fill_node_rows<<<blocks_for(n, 256), 256, 0, st>>>(T.connect[g], n, T.qoff[g], (const long long*)rowptr, cursor,
                                                                  (long long*)qlist);
    }
    if (Nn) k_sort_node_rows<<<blocks_for(Nn, 128), 128, 0, st>>>(Nn, (const long long*)rowptr, (long long*)qlist);
    return check_launch("efb_csr_fill_node_rows");
}

extern "C" int efb_csr_count_adj(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                                 const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, const int64_t* qlist, int32_t* deg,
                                 int32_t* err_flag, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, 1)) return 1;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(err_flag, 0, sizeof(int), st);
    if (Nn) k_count_adj<<<blocks_for(Nn, 128), 128, 0, st>>>(T, Nn, (const long long*)rowptr, (const long long*)qlist, deg, err_flag);
    return check_launch("efb_csr_count_adj");
}

extern "C" int efb_csr_fill_adj(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                                const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, const int64_t* qlist,
                                const int64_t* adjptr, int32_t* adj, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, 1)) return 1;
    if (Nn)
        k_fill_adj<<<blocks_for(Nn, 128), 128, 0, as_stream(stream)>>>(T, Nn, (const long long*)rowptr, (const long long*)qlist,
                                                                        (const long long*)adjptr, adj);
    return check_launch("efb_csr_fill_adj");
}

extern "C" int efb_csr_expand(int64_t Nn, int dof_n, int64_t Ndof, int64_t nnz_node, const int64_t* adjptr, const int32_t* adj,
                              int index_bytes, void* indptr, void* indices, void* stream) {
    cudaStream_t st = as_stream(stream);
    if (index_bytes != 4 && index_bytes != 8) {
        set_error("efb_csr_expand: index_bytes must be 4 or 8");
        return 1;
    }
    if (Ndof < Nn * dof_n) {
        set_error("efb_csr_expand: Ndof < Nn*dof_n");
        return 1;
    }
    if (index_bytes == 4) {
        k_expand_indptr<int><<<blocks_for(Ndof + 1, 256), 256, 0, st>>>(Ndof, Nn, dof_n, (const long long*)adjptr, (int*)indptr);
        if (nnz_node)
            k_expand_indices<int><<<blocks_for(nnz_node, 256), 256, 0, st>>>(nnz_node, Nn, dof_n, (const long long*)adjptr, adj,
                                                                              (int*)indices);
    } else {
        k_expand_indptr<long long><<<blocks_for(Ndof + 1, 256), 256, 0, st>>>(Ndof, Nn, dof_n, (const long long*)adjptr,
                                                                               (long long*)indptr);
        if (nnz_node)
            k_expand_indices<long long><<<blocks_for(nnz_node, 256), 256, 0, st>>>(nnz_node, Nn, dof_n, (const long long*)adjptr,
                                                                                    adj, (long long*)indices);
    }
    return check_launch("efb_csr_expand");
}

extern "C" int efb_csr_slot_map(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                                const int32_t* nPe_host, const int64_t* adjptr, const int32_t* adj, int32_t* pos, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, 1)) return 1;
    for (int g = 0; g < T.n; ++g) {
        const long long n = T.Ne[g] * T.nPe[g] * T.nPe[g];
        if (n)
            k_slot_map<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(T.connect[g], T.nPe[g], n, (const long long*)adjptr, adj,
                                                                           pos + T.poff[g]);
    }
    return check_launch("efb_csr_slot_map");
}

extern "C" int efb_csr_inv_map(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host, const int32_t* nPe_host,
                               int dof_n, const int64_t* adjptr, const int32_t* pos, int32_t* inv, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, connect_host, nullptr, Ne_host, nPe_host, dof_n)) return 1;
    for (int g = 0; g < T.n; ++g) {
        const long long n = T.koff[g + 1] - T.koff[g];
        if (n)
            k_inv_map<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(T.connect[g], T.nPe[g], dof_n, n, (const long long*)adjptr,
                                                                          pos + T.poff[g], inv + T.koff[g]);
    }
    return check_launch("efb_csr_inv_map");
}

extern "C" int efb_csr_row_has_entry(int64_t Nn, int dof_n, int64_t Ndof, const int64_t* rowptr, int32_t* has, void* stream) {
    if (Ndof) k_row_has_entry<<<blocks_for(Ndof, 256), 256, 0, as_stream(stream)>>>(Ndof, Nn, dof_n, (const long long*)rowptr, has);
    return check_launch("efb_csr_row_has_entry");
}

extern "C" int efb_csr_compact_rows(int64_t Ndof, const int64_t* indptr64, const double* dense, double* out, void* stream) {
    if (Ndof) k_compact<<<blocks_for(Ndof, 256), 256, 0, as_stream(stream)>>>(Ndof, (const long long*)indptr64, dense, out);
    return check_launch("efb_csr_compact_rows");
}

extern "C" int efb_csr_replay_matrix(int n_groups, const double* const* data_host, const int64_t* Ne_host, const int32_t* nPe_host,
                                     int dof_n, int64_t Nn, const int64_t* rowptr, const int64_t* qlist, const int64_t* adjptr,
                                     const int32_t* pos, int32_t max_deg, double* data_out, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, nullptr, data_host, Ne_host, nPe_host, dof_n)) return 1;
    if (Nn == 0) return 0;
    if (n_groups == 1) {  // single group with a compiled (dof_n, nPe): hoisted index math, batched loads
#define EFB_FAST(D, N)                                                                                                 \
    if (dof_n == D && nPe_host[0] == N)                                                                                \
        return launch_replay_fast<D, N>(data_host[0], Nn, (const long long*)rowptr, (const long long*)qlist,           \
                                        (const long long*)adjptr, pos, max_deg, data_out, as_stream(stream));
        EFB_FAST(3, 8) EFB_FAST(3, 4) EFB_FAST(3, 10) EFB_FAST(3, 27) EFB_FAST(3, 20) EFB_FAST(3, 6)
        EFB_FAST(2, 3) EFB_FAST(2, 4) EFB_FAST(2, 6) EFB_FAST(2, 8) EFB_FAST(2, 9)
        EFB_FAST(1, 3) EFB_FAST(1, 4) EFB_FAST(1, 6) EFB_FAST(1, 8) EFB_FAST(1, 9) EFB_FAST(1, 10) EFB_FAST(1, 27) EFB_FAST(1, 20)
#undef EFB_FAST
    }
    const int warps = 8;
    const int acc_per_warp = dof_n * dof_n * max_deg;
    const size_t bytes = sizeof(double) * (size_t)acc_per_warp * warps;
    if (ensure_smem(k_replay_matrix, bytes)) return 1;
    k_replay_matrix<<<blocks_for(Nn, warps), warps * 32, bytes, as_stream(stream)>>>(
        T, dof_n, Nn, (const long long*)rowptr, (const long long*)qlist, (const long long*)adjptr, pos, acc_per_warp, data_out);
    return check_launch("efb_csr_replay_matrix");
}

extern "C" int efb_csr_replay_vector(int n_groups, const double* const* data_host, const int64_t* Ne_host, const int32_t* nPe_host,
                                     int dof_n, int64_t Nn, const int64_t* rowptr, const int64_t* qlist, double* out, void* stream) {
    GroupTable T;
    if (make_table(T, n_groups, nullptr, data_host, Ne_host, nPe_host, dof_n)) return 1;
    const long long nrows = (long long)Nn * dof_n;
    if (nrows) k_replay_vector<<<blocks_for(nrows, 256), 256, 0, as_stream(stream)>>>(T, dof_n, nrows, (const long long*)rowptr,
                                                                                      (const long long*)qlist, out);
    return check_launch("efb_csr_replay_vector");
}
