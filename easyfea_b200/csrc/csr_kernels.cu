// C-ABI entry points of the CSR pattern build (A1) and the deterministic replay (A2).
#include "common.cuh"
#include "csr_kernels.cuh"
#include "tma.cuh"

#include <stdlib.h>
#include <string.h>

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

// Pipelined form of the single-group replay: persistent warps walk nodes gw, gw + GW, ...; the row/adjacency pointers of
// node k+2 and the element-row list of node k+1 are already in flight (in registers) while node k is accumulated, so the
// only exposed latency per node is that of the K_e rows themselves, of which up to SB (all 8 of an interior HEXA8 node)
// are requested at once.  The slot positions of a source are loaded once (lanes 0..NPE-1) and handed out by shuffles.
// Accumulation order per slot is still ascending element-row index: bit-identical to np.bincount.
template <int D, int NPE, int SB, int MINB>
__global__ void __launch_bounds__(256, MINB)
    k_replay_pipe(const double* __restrict__ data, long long Nn, const long long* __restrict__ rowptr,
                  const long long* __restrict__ qlist, const long long* __restrict__ adjptr, const int* __restrict__ pos,
                  int acc_per_warp, double* __restrict__ out) {
    extern __shared__ double smem[];
    constexpr int NDOF = D * NPE, NV = D * NDOF, VPL = (NV + 31) / 32;
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* acc = smem + (size_t)warp * acc_per_warp;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp, GW = (long long)gridDim.x * (blockDim.x >> 5);

    // lanes 0,1: rowptr[n], rowptr[n+1]; lanes 2,3: adjptr[n], adjptr[n+1]
    auto load_meta = [&](long long n) -> long long {
        if (n >= Nn || lane >= 4) return 0;
        return lane < 2 ? rowptr[n + lane] : adjptr[n + lane - 2];
    };
    // static decomposition of this lane's values of a source: i = lane + 32k -> (row ii, column node b, component jj)
    int v_row[VPL], v_b[VPL], v_jj[VPL];
    EFB_UNROLL
    for (int k = 0; k < VPL; ++k) {
        const int i = lane + 32 * k, ii = i / NDOF, j = i - ii * NDOF;
        v_row[k] = ii;
        v_b[k] = (j / D) % NPE;  // % NPE keeps the shuffle source lane in range for the unused tail values
        v_jj[k] = j % D;
    }

    long long m_cur = load_meta(gw);
    long long m_nxt = load_meta(gw + GW);
    long long s_begin = __shfl_sync(FULL, m_cur, 0), s_end = __shfl_sync(FULL, m_cur, 1);
    long long a0 = __shfl_sync(FULL, m_cur, 2), a1 = __shfl_sync(FULL, m_cur, 3);
    long long q_cur = (s_begin + lane < s_end) ? qlist[s_begin + lane] : 0;

    for (long long n = gw; n < Nn; n += GW) {
        // ---- prefetch: pointers of node k+2, element rows of node k+1
        const long long m_fut = load_meta(n + 2 * GW);
        const long long nb = __shfl_sync(FULL, m_nxt, 0), ne = __shfl_sync(FULL, m_nxt, 1);
        const long long na0 = __shfl_sync(FULL, m_nxt, 2), na1 = __shfl_sync(FULL, m_nxt, 3);
        const long long q_nxt = (nb + lane < ne) ? qlist[nb + lane] : 0;

        // ---- node k
        const int deg = (int)(a1 - a0), rowlen = D * deg, blk = D * rowlen;
        const int cnt = (int)(s_end - s_begin);
        for (int i = lane; i < blk; i += 32) acc[i] = 0.0;
        __syncwarp();
        for (int s0 = 0; s0 < cnt; s0 += SB) {
            double v[SB][VPL];
            int prl[SB];
            EFB_UNROLL
            for (int u = 0; u < SB; ++u) {
                if (s0 + u < cnt) {  // warp-uniform
                    const long long q = (s0 + u < 32) ? __shfl_sync(FULL, q_cur, s0 + u) : qlist[s_begin + s0 + u];
                    const double* src = data + q * (long long)NV;
                    EFB_UNROLL
                    for (int k = 0; k < VPL; ++k)
                        if (lane + 32 * k < NV) v[u][k] = src[lane + 32 * k];
                    prl[u] = pos[q * NPE + (lane < NPE ? lane : 0)];
                }
            }
            EFB_UNROLL
            for (int u = 0; u < SB; ++u) {
                if (s0 + u < cnt) {
                    EFB_UNROLL
                    for (int k = 0; k < VPL; ++k) {
                        const int pb = __shfl_sync(FULL, prl[u], v_b[k]);
                        if (lane + 32 * k < NV) acc[v_row[k] * rowlen + pb * D + v_jj[k]] += v[u][k];
                    }
                    __syncwarp();  // the next source may hit the same slots from other lanes
                }
            }
        }
        double* dst = out + (long long)D * D * a0;
        for (int i = lane; i < blk; i += 32) dst[i] = acc[i];
        __syncwarp();

        // ---- rotate the pipeline
        m_nxt = m_fut;
        s_begin = nb; s_end = ne; a0 = na0; a1 = na1;
        q_cur = q_nxt;
    }
}

template <int D, int NPE, int SB, int MINB>
static int launch_replay_pipe(const double* data, long long Nn, const long long* rowptr, const long long* qlist,
                              const long long* adjptr, const int* pos, int max_deg, double* out, cudaStream_t st) {
    const int warps = 8;
    const int acc_per_warp = D * D * max_deg;
    const size_t bytes = sizeof(double) * (size_t)acc_per_warp * warps;
    auto kern = k_replay_pipe<D, NPE, SB, MINB>;
    if (ensure_smem(kern, bytes)) return 1;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (Nn + warps - 1) / warps;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, warps * 32, bytes, st>>>(data, Nn, rowptr, qlist, adjptr, pos, acc_per_warp, out);
    return check_launch("efb_csr_replay_matrix");
}

// Ring form of the single-group replay (the default): a warp owns kRingNodes consecutive nodes, hence one contiguous
// range of the element-row list, and streams those K_e rows (+ their slot positions) through a ring of R slots in shared
// memory with asynchronous copies (cp.async / LDGSTS: no registers held while the data is in flight).  One commit group
// per source, `wait_group R-1` before a source is consumed and one refill issued after it, so R sources (R * 608 B for
// HEXA8) are in flight per warp at all times, across node boundaries.  Accumulation order per slot is ascending
// element-row index: bit-identical to np.bincount.
#ifndef EFB_RING_NODES
#define EFB_RING_NODES 64
#endif
#ifndef EFB_RING_BYTES
#define EFB_RING_BYTES 5000
#endif
constexpr int kRingNodes = EFB_RING_NODES;

template <int BYTES>
__device__ __forceinline__ void cp_async_piece(void* sdst, const void* gsrc) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "n"(BYTES)
                     : "memory");
}

template <int D, int NPE>
struct ReplayRing {
    static constexpr int NDOF = D * NPE, NV = D * NDOF, VPL = (NV + 31) / 32;
    static constexpr int PIECE = (NV % 2 == 0) ? 2 : 1;            // doubles per asynchronous copy
    static constexpr int NP = NV / PIECE, PPL = (NP + 31) / 32;    // pieces per source, per lane
    static constexpr int SLOT = (NV + (NPE + 1) / 2 + 1) & ~1;     // doubles per ring slot: values | positions (int32)
    static constexpr int R0 = EFB_RING_BYTES / (SLOT * 8);
    static constexpr int R = R0 < 3 ? 3 : (R0 > 24 ? 24 : R0);     // slots per warp (~5 KB: 8 HEXA8 rows; a deeper ring costs occupancy and is slower, profiles/README.md)
};

template <int D, int NPE>
__global__ void __launch_bounds__(256)
    k_replay_ring(const double* __restrict__ data, long long Nn, const long long* __restrict__ rowptr,
                  const long long* __restrict__ qlist, const long long* __restrict__ adjptr, const int* __restrict__ pos,
                  int acc_per_warp, double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    using RR = ReplayRing<D, NPE>;
    constexpr int NDOF = RR::NDOF, NV = RR::NV, VPL = RR::VPL, PIECE = RR::PIECE, NP = RR::NP, PPL = RR::PPL, SLOT = RR::SLOT, R = RR::R;
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = R * SLOT + ((acc_per_warp + 1) & ~1);
    double* ring = smem + (size_t)warp * per_warp;
    double* acc = ring + R * SLOT;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long n_lo = gw * kRingNodes;
    if (n_lo >= Nn) return;
    const long long n_hi = (n_lo + kRingNodes < Nn) ? n_lo + kRingNodes : Nn;
    const long long s_lo = rowptr[n_lo], s_hi = rowptr[n_hi];

    // static decomposition of this lane's values of a source: i = lane + 32k -> (row ii, column node b, component jj)
    int v_row[VPL], v_b[VPL], v_jj[VPL];
    EFB_UNROLL
    for (int k = 0; k < VPL; ++k) {
        const int i = lane + 32 * k, ii = i / NDOF, j = i - ii * NDOF;
        v_row[k] = ii;
        v_b[k] = (j / D) % NPE;
        v_jj[k] = j % D;
    }

    // ---- producer side: element rows s_lo.. are requested in order; their ids travel in two 32-wide register chunks
    long long si = s_lo, qbase = s_lo;
    long long qa = (qbase + lane < s_hi) ? qlist[qbase + lane] : 0;
    long long qb = (qbase + 32 + lane < s_hi) ? qlist[qbase + 32 + lane] : 0;
    int islot = 0;
    auto issue = [&]() {
        if (si < s_hi) {  // warp-uniform
            const long long q = __shfl_sync(FULL, qa, (int)(si - qbase));
            double* slot = ring + islot * SLOT;
            const double* src = data + q * (long long)NV;
            EFB_UNROLL
            for (int k = 0; k < PPL; ++k) {
                const int piece = lane + 32 * k;
                if (piece < NP) cp_async_piece<PIECE * 8>(slot + piece * PIECE, src + piece * PIECE);
            }
            if (lane < NPE) cp_async_piece<4>(reinterpret_cast<int*>(slot + NV) + lane, pos + q * NPE + lane);
            ++si;
            islot = (islot + 1 == R) ? 0 : islot + 1;
            if (si - qbase == 32) {
                qbase += 32;
                qa = qb;
                qb = (qbase + 32 + lane < s_hi) ? qlist[qbase + 32 + lane] : 0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    EFB_UNROLL
    for (int r = 0; r < R; ++r) issue();

    // ---- consumer side: nodes in chunks of 31 (32 pointer entries per chunk, the next chunk's already in flight)
    long long sc = s_lo;
    int cslot = 0;
    auto ptr_chunk = [&](const long long* ptr, long long nb) { return ptr[(nb + lane < n_hi) ? nb + lane : n_hi]; };
    long long rp_n = ptr_chunk(rowptr, n_lo), ap_n = ptr_chunk(adjptr, n_lo);
    for (long long nb = n_lo; nb < n_hi; nb += 31) {
        const long long rp = rp_n, ap = ap_n;
        if (nb + 31 < n_hi) {
            rp_n = ptr_chunk(rowptr, nb + 31);
            ap_n = ptr_chunk(adjptr, nb + 31);
        }
        const int nn = (n_hi - nb < 31) ? (int)(n_hi - nb) : 31;
        for (int jn = 0; jn < nn; ++jn) {
            const long long s_e = __shfl_sync(FULL, rp, jn + 1);
            const long long a0 = __shfl_sync(FULL, ap, jn), a1 = __shfl_sync(FULL, ap, jn + 1);
            const int deg = (int)(a1 - a0), rowlen = D * deg, blk = D * rowlen;
            for (int i = lane; i < blk; i += 32) acc[i] = 0.0;
            __syncwarp();
            for (; sc < s_e; ++sc) {
                asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
                __syncwarp();  // every lane's pieces of this source have landed
                const double* slot = ring + cslot * SLOT;
                const int* prow = reinterpret_cast<const int*>(slot + NV);
                EFB_UNROLL
                for (int k = 0; k < VPL; ++k)
                    if (lane + 32 * k < NV) acc[v_row[k] * rowlen + prow[v_b[k]] * D + v_jj[k]] += slot[lane + 32 * k];
                __syncwarp();  // slot consumed by all lanes; the next source may hit the same accumulators
                cslot = (cslot + 1 == R) ? 0 : cslot + 1;
                issue();
            }
            double* dst = out + (long long)D * D * a0;
            for (int i = lane; i < blk; i += 32) dst[i] = acc[i];
            __syncwarp();
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int D, int NPE>
static int launch_replay_ring(const double* data, long long Nn, const long long* rowptr, const long long* qlist,
                              const long long* adjptr, const int* pos, int max_deg, double* out, cudaStream_t st) {
    using RR = ReplayRing<D, NPE>;
    const int warps = 8;
    const int acc_per_warp = D * D * max_deg;
    const size_t bytes = sizeof(double) * (size_t)(RR::R * RR::SLOT + ((acc_per_warp + 1) & ~1)) * warps;
    if (ensure_smem(k_replay_ring<D, NPE>, bytes)) return 1;
    const long long nwarps = (Nn + kRingNodes - 1) / kRingNodes;
    k_replay_ring<D, NPE><<<blocks_for(nwarps, warps), warps * 32, bytes, st>>>(data, Nn, rowptr, qlist, adjptr, pos, acc_per_warp, out);
    return check_launch("efb_csr_replay_matrix");
}

// TMA form of the ring replay (element rows whose values and slot positions are both 16-byte multiples: HEXA8, TETRA4,
// QUAD4/8, HEXA20): the ring is refilled by bulk asynchronous copies (cp.async.bulk, one for the K_e row and one for its
// slot positions) issued by a single lane and completed on a per-slot mbarrier, so keeping R element rows in flight per
// warp costs a handful of instructions per row instead of one cp.async per 16 bytes and lane.
template <int D, int NPE>
struct ReplayTma {
    static constexpr int NDOF = D * NPE, NV = D * NDOF, VPL = (NV + 31) / 32;
    static constexpr bool ok = (NV * 8) % 16 == 0 && (NPE * 4) % 16 == 0;
    static constexpr int SLOT = NV + NPE / 2;                      // doubles per ring slot: values | positions (int32)
    static constexpr unsigned TX = NV * 8 + NPE * 4;               // bytes landing per slot
    static constexpr int R0 = EFB_RING_BYTES / (SLOT * 8);
    static constexpr int R = R0 < 3 ? 3 : (R0 > 24 ? 24 : R0);     // slots per warp (~5 KB: 8 HEXA8 rows; a deeper ring costs occupancy and is slower, profiles/README.md)
};

#ifndef EFB_REPLAY_ELECT
#define EFB_REPLAY_ELECT elect_one()  // (lane == 0) is the form that makes ptxas uniformise every bulk-copy operand
#endif
#ifndef EFB_REPLAY_MINB
#define EFB_REPLAY_MINB 4
#endif
template <int D, int NPE>
__global__ void __launch_bounds__(256, EFB_REPLAY_MINB)
    k_replay_tma(const double* __restrict__ data, long long Nn, const long long* __restrict__ rowptr,
                 const long long* __restrict__ qlist, const long long* __restrict__ adjptr, const int* __restrict__ pos,
                 int acc_per_warp, double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    using RT = ReplayTma<D, NPE>;
    constexpr int NDOF = RT::NDOF, NV = RT::NV, VPL = RT::VPL, SLOT = RT::SLOT, R = RT::R;
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = R * SLOT + ((acc_per_warp + 1) & ~1) + ((R + 1) & ~1);  // ring | accumulators | mbarriers
    double* ring = smem + (size_t)warp * per_warp;
    double* acc = ring + R * SLOT;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(acc + ((acc_per_warp + 1) & ~1));
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long n_lo = gw * kRingNodes;
    if (n_lo >= Nn) return;
    const long long n_hi = (n_lo + kRingNodes < Nn) ? n_lo + kRingNodes : Nn;
    const long long s_lo = rowptr[n_lo], s_hi = rowptr[n_hi];

    if (lane < R) mbar_init(bars + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // this lane's values of a source: i = lane + 32k -> offset of (row ii, component jj) in a node block of row length 1,
    // and the column node b whose slot position selects the column
    int v_row[VPL], v_b[VPL], v_jj[VPL];
    EFB_UNROLL
    for (int k = 0; k < VPL; ++k) {
        const int i = lane + 32 * k, ii = i / NDOF, j = i - ii * NDOF;
        v_row[k] = ii;
        v_b[k] = (j / D) % NPE;
        v_jj[k] = j % D;
    }

    // ---- producer side (lane 0 issues): element rows s_lo.. in order, ids in two 32-wide register chunks
    long long si = s_lo, qbase = s_lo;
    long long qa = (qbase + lane < s_hi) ? qlist[qbase + lane] : 0;
    long long qb = (qbase + 32 + lane < s_hi) ? qlist[qbase + 32 + lane] : 0;
    int islot = 0;
    auto issue = [&]() {
        if (si < s_hi) {  // warp-uniform
            const long long q = __shfl_sync(FULL, qa, (int)(si - qbase));
            if (EFB_REPLAY_ELECT) {
                double* slot = ring + islot * SLOT;
                mbar_expect_tx(bars + islot, RT::TX);
                bulk_load(slot, data + q * (long long)NV, NV * 8, bars + islot);
                bulk_load(slot + NV, pos + q * NPE, NPE * 4, bars + islot);
            }
            ++si;
            islot = (islot + 1 == R) ? 0 : islot + 1;
            if (si - qbase == 32) {
                qbase += 32;
                qa = qb;
                qb = (qbase + 32 + lane < s_hi) ? qlist[qbase + 32 + lane] : 0;
            }
        }
    };
    EFB_UNROLL
    for (int r = 0; r < R; ++r) issue();

    // ---- consumer side
    long long sc = s_lo;
    int cslot = 0;
    unsigned cphase = 0;
    auto ptr_chunk = [&](const long long* ptr, long long nb) { return ptr[(nb + lane < n_hi) ? nb + lane : n_hi]; };
    long long rp_n = ptr_chunk(rowptr, n_lo), ap_n = ptr_chunk(adjptr, n_lo);
    for (long long nb = n_lo; nb < n_hi; nb += 31) {
        const long long rp = rp_n, ap = ap_n;
        if (nb + 31 < n_hi) {
            rp_n = ptr_chunk(rowptr, nb + 31);
            ap_n = ptr_chunk(adjptr, nb + 31);
        }
        const int nn = (n_hi - nb < 31) ? (int)(n_hi - nb) : 31;
        for (int jn = 0; jn < nn; ++jn) {
            const long long s_e = __shfl_sync(FULL, rp, jn + 1);
            const long long a0 = __shfl_sync(FULL, ap, jn), a1 = __shfl_sync(FULL, ap, jn + 1);
            const int deg = (int)(a1 - a0), rowlen = D * deg, blk = D * rowlen;
            int base[VPL];  // accumulator offset of (row, component) for this node
            EFB_UNROLL
            for (int k = 0; k < VPL; ++k) base[k] = v_row[k] * rowlen + v_jj[k];
            for (int i = lane; i < blk; i += 32) acc[i] = 0.0;
            __syncwarp();
            for (; sc < s_e; ++sc) {
                mbar_wait(bars + cslot, cphase);  // the row and its positions have landed
                const double* slot = ring + cslot * SLOT;
                const int* prow = reinterpret_cast<const int*>(slot + NV);
                EFB_UNROLL
                for (int k = 0; k < VPL; ++k)
                    if (lane + 32 * k < NV) acc[base[k] + prow[v_b[k]] * D] += slot[lane + 32 * k];
                __syncwarp();  // slot consumed by all lanes; the next source may hit the same accumulators
                if (++cslot == R) {
                    cslot = 0;
                    cphase ^= 1u;
                }
                issue();
            }
            double* dst = out + (long long)D * D * a0;
            for (int i = lane; i < blk; i += 32) dst[i] = acc[i];
            __syncwarp();
        }
    }
}

template <int D, int NPE>
static int launch_replay_tma(const double* data, long long Nn, const long long* rowptr, const long long* qlist,
                             const long long* adjptr, const int* pos, int max_deg, double* out, cudaStream_t st) {
    using RT = ReplayTma<D, NPE>;
    const int warps = 8;
    const int acc_per_warp = D * D * max_deg;
    const size_t bytes = sizeof(double) * (size_t)(RT::R * RT::SLOT + ((acc_per_warp + 1) & ~1) + ((RT::R + 1) & ~1)) * warps;
    if (ensure_smem(k_replay_tma<D, NPE>, bytes)) return 1;
    const long long nwarps = (Nn + kRingNodes - 1) / kRingNodes;
    k_replay_tma<D, NPE><<<blocks_for(nwarps, warps), warps * 32, bytes, st>>>(data, Nn, rowptr, qlist, adjptr, pos, acc_per_warp, out);
    return check_launch("efb_csr_replay_matrix");
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
        // EFB_REPLAY_KERNEL = fast | pipe | pipe8 | ring | tma (dev/tuning knob)
        const char* env = getenv("EFB_REPLAY_KERNEL");
        // default by the size of an element row (NV = dof_n^2 * nPe doubles), measured on B200 (profiles/README.md):
        // small rows (TRI3, TETRA4) -> fast, mid-size rows (HEXA8) -> ring, large rows (HEXA20/27) -> pipe
        const int nv = dof_n * dof_n * nPe_host[0];
        const int form = env ? (!strcmp(env, "fast") ? 0 : (!strcmp(env, "pipe8") ? 2 : (!strcmp(env, "pipe") ? 1 : (!strcmp(env, "ring") ? 3 : 4))))
                             : (nv < 48 ? 0 : (nv <= 128 ? 4 : 1));
        const bool aligned = (reinterpret_cast<uintptr_t>(data_host[0]) & 15) == 0;
#define EFB_PIPE(D, N, SB, MINB)                                                                                        \
    return launch_replay_pipe<D, N, SB, MINB>(data_host[0], Nn, (const long long*)rowptr, (const long long*)qlist,     \
                                              (const long long*)adjptr, pos, max_deg, data_out, as_stream(stream));
#define EFB_FAST(D, N)                                                                                                 \
    if (dof_n == D && nPe_host[0] == N) {                                                                              \
        if (form == 0)                                                                                                 \
            return launch_replay_fast<D, N>(data_host[0], Nn, (const long long*)rowptr, (const long long*)qlist,       \
                                            (const long long*)adjptr, pos, max_deg, data_out, as_stream(stream));      \
        if constexpr (ReplayTma<D, N>::ok) {                                                                           \
            if (form == 4 && aligned && (reinterpret_cast<uintptr_t>(pos) & 15) == 0)                                  \
                return launch_replay_tma<D, N>(data_host[0], Nn, (const long long*)rowptr, (const long long*)qlist,    \
                                               (const long long*)adjptr, pos, max_deg, data_out, as_stream(stream));   \
        }                                                                                                              \
        if ((form == 3 || form == 4) && aligned)                                                                       \
            return launch_replay_ring<D, N>(data_host[0], Nn, (const long long*)rowptr, (const long long*)qlist,       \
                                            (const long long*)adjptr, pos, max_deg, data_out, as_stream(stream));      \
        if constexpr (D * D * N > 128) {                                                                               \
            EFB_PIPE(D, N, 2, 3)                                                                                       \
        } else if constexpr (D * D * N <= 72) {                                                                        \
            if (form == 2) { EFB_PIPE(D, N, 8, 2) }                                                                    \
            EFB_PIPE(D, N, 4, 3)                                                                                       \
        } else {                                                                                                       \
            EFB_PIPE(D, N, 4, 3)                                                                                       \
        }                                                                                                              \
    }
        EFB_FAST(3, 8) EFB_FAST(3, 4) EFB_FAST(3, 10) EFB_FAST(3, 27) EFB_FAST(3, 20) EFB_FAST(3, 6)
        EFB_FAST(2, 3) EFB_FAST(2, 4) EFB_FAST(2, 6) EFB_FAST(2, 8) EFB_FAST(2, 9)
        EFB_FAST(1, 3) EFB_FAST(1, 4) EFB_FAST(1, 6) EFB_FAST(1, 8) EFB_FAST(1, 9) EFB_FAST(1, 10) EFB_FAST(1, 27) EFB_FAST(1, 20)
#undef EFB_FAST
#undef EFB_PIPE
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
