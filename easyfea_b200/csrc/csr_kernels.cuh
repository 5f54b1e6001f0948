// CSR pattern from the node graph (A1) and deterministic, atomic-free replay (A2) — bodies shared by the device
// kernels and the host emulation of tests/hostcheck.
//
// Reference semantics (EasyFEA/Simulations/_simu.py:989-1102): the pattern is scipy's sorted/unique COO->CSR of
// rows = repeat(assembly_e, ndof), cols = tile(assembly_e, ndof); the data of a slot is np.bincount's ordered sum of
// the element entries k = (group, e, i, j) that land on it.
//
// Here: a dof_n-block problem's pattern is the block expansion of the NODE adjacency graph, so it is built from
// node pairs (Ne*nPe^2 of them instead of Ne*ndof^2 keys) with per-node sorted lists:
//   rowptr/qlist : for node n the ascending list of "element rows" q = qoff_g + e*nPe_g + a that touch it
//   adjptr/adj   : ascending distinct neighbour nodes of n (the block columns of its dof rows)
//   pos          : for every (q, b) the index of node connect[e][b] inside adj(n)
// CSR row n*d+i then starts at d*d*adjptr[n] + i*d*deg(n) and holds columns {m*d+j : m in adj(n), j<d} ascending —
// exactly the canonical pattern.  Replay walks each node's q-list in ascending order, which is ascending k for
// every slot, so the floating-point sums are bit-identical to np.bincount.
#pragma once
#include "frame.cuh"

#ifdef __CUDACC__
#define EFB_ATOMIC_ADD_I32(ptr, v) atomicAdd((ptr), (v))
#define EFB_LANES(lane) for (int lane = threadIdx.x & 31, _efb_l1 = 1; _efb_l1; _efb_l1 = 0, __syncwarp())
#define EFB_LANE_COPIES 1
#define EFB_LANE_SLOT(lane) 0
#else
#define EFB_LANE_COPIES 32
#define EFB_LANE_SLOT(lane) (lane)
static inline int efb_host_fetch_add(int* p, int v) {
    int old = *p;
    *p += v;
    return old;
}
#define EFB_ATOMIC_ADD_I32(ptr, v) efb_host_fetch_add((ptr), (v))
#define EFB_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#endif

namespace efb {

constexpr int kMaxGroups = 8;
constexpr int kAdjCap = 512;  // max distinct neighbour nodes of one node

struct GroupTable {
    int n;
    const int* connect[kMaxGroups];
    const double* data[kMaxGroups];
    long long Ne[kMaxGroups];
    int nPe[kMaxGroups];
    long long qoff[kMaxGroups + 1];  // element-row offsets
    long long poff[kMaxGroups + 1];  // offsets into `pos` (sum of Ne*nPe*nPe)
    long long koff[kMaxGroups + 1];  // element-entry offsets for a given dof_n (inv map)
};

EFB_HD int group_of_q(const GroupTable& T, long long q) {
    int g = 0;
    while (g + 1 < T.n && q >= T.qoff[g + 1]) ++g;
    return g;
}

// ---- stage 1/3: node -> element rows -------------------------------------------------------------------
EFB_D void count_node_rows_item(const int* connect, long long i, int* cnt) { EFB_ATOMIC_ADD_I32(cnt + connect[i], 1); }

EFB_D void fill_node_rows_item(const int* connect, long long i, long long qoff, const long long* rowptr, int* cursor,
                                long long* qlist) {
    const int n = connect[i];
    const int slot = EFB_ATOMIC_ADD_I32(cursor + n, 1);
    qlist[rowptr[n] + slot] = qoff + i;
}

// ascending order restores determinism after the atomic fill
EFB_HD void sort_node_rows_item(long long n, const long long* rowptr, long long* qlist) {
    long long* L = qlist + rowptr[n];
    const int len = (int)(rowptr[n + 1] - rowptr[n]);
    for (int i = 1; i < len; ++i) {
        const long long key = L[i];
        int j = i - 1;
        while (j >= 0 && L[j] > key) {
            L[j + 1] = L[j];
            --j;
        }
        L[j + 1] = key;
    }
}

// ---- stage 4/6: distinct sorted neighbours of node n ----------------------------------------------------
// returns the count, or -1 if it exceeds kAdjCap; `buf` is thread-private scratch of kAdjCap ints
EFB_HD int gather_neighbours(const GroupTable& T, long long n, const long long* rowptr, const long long* qlist, int* buf) {
    int len = 0;
    for (long long s = rowptr[n]; s < rowptr[n + 1]; ++s) {
        const long long q = qlist[s];
        const int g = group_of_q(T, q);
        const int nPe = T.nPe[g];
        const long long e = (q - T.qoff[g]) / nPe;
        const int* row = T.connect[g] + e * nPe;
        for (int b = 0; b < nPe; ++b) {
            const int m = row[b];
            // binary search for the insertion point
            int lo = 0, hi = len;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (buf[mid] < m)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < len && buf[lo] == m) continue;
            if (len == kAdjCap) return -1;
            for (int k = len; k > lo; --k) buf[k] = buf[k - 1];
            buf[lo] = m;
            ++len;
        }
    }
    return len;
}

// ---- stage 7: block expansion ------------------------------------------------------------------------------
template <class IDX>
EFB_HD void expand_indptr_item(long long r, long long Nn, int d, const long long* adjptr, IDX* indptr) {
    // r in [0, Ndof]; rows beyond Nn*d (Lagrange rows, _simu.py:154-158) are empty
    const long long n = r / d;
    const int i = (int)(r % d);
    long long v;
    if (n >= Nn)
        v = (long long)d * d * adjptr[Nn];
    else
        v = (long long)d * d * adjptr[n] + (long long)i * d * (adjptr[n + 1] - adjptr[n]);
    indptr[r] = (IDX)v;
}

template <class IDX>
EFB_HD void expand_indices_item(long long t, long long n, int d, const long long* adjptr, const int* adj, IDX* indices) {
    // t indexes adj; node n owns t in [adjptr[n], adjptr[n+1])
    const long long deg = adjptr[n + 1] - adjptr[n];
    const long long base = (long long)d * d * adjptr[n];
    const long long c = t - adjptr[n];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) indices[base + i * d * deg + c * d + j] = (IDX)((long long)adj[t] * d + j);
}

// ---- stage 8/9: slot maps ------------------------------------------------------------------------------------
EFB_HD int find_in_adj(const long long* adjptr, const int* adj, int n, int m) {
    long long lo = adjptr[n], hi = adjptr[n + 1];
    const long long base = lo;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (adj[mid] < m)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (int)(lo - base);
}

// i over Ne*nPe*nPe of one group: (e, a, b)
EFB_HD void slot_map_item(const int* connect, int nPe, long long i, const long long* adjptr, const int* adj, int* pos) {
    const int b = (int)(i % nPe);
    const long long ea = i / nPe;  // e*nPe + a
    const long long e = ea / nPe;
    pos[i] = find_in_adj(adjptr, adj, connect[ea], connect[e * nPe + b]);
}

// i over Ne*ndof*ndof of one group: the reference's inv[k], _simu.py:1098-1101
EFB_HD void inv_map_item(const int* connect, int nPe, int d, long long i, const long long* adjptr, const int* pos, int* inv) {
    const int ndof = nPe * d;
    const int j = (int)(i % ndof);
    const long long ei = i / ndof;
    const int il = (int)(ei % ndof);
    const long long e = ei / ndof;
    const int a = il / d, ii = il % d, b = j / d, jj = j % d;
    const int n = connect[e * nPe + a];
    const long long deg = adjptr[n + 1] - adjptr[n];
    const long long slot = (long long)d * d * adjptr[n] + (long long)ii * d * deg + (long long)pos[(e * nPe + a) * nPe + b] * d + jj;
    inv[i] = (int)slot;
}

// ---- A2: replay ------------------------------------------------------------------------------------------------
// one warp per node; `acc` = d*d*deg doubles of warp-private shared memory laid out exactly like the node's CSR block
EFB_D void replay_node(const GroupTable& T, int d, long long n, const long long* EFB_RESTRICT rowptr,
                       const long long* EFB_RESTRICT qlist, const long long* EFB_RESTRICT adjptr,
                       const int* EFB_RESTRICT pos, double* acc, double* EFB_RESTRICT out) {
    const long long a0 = adjptr[n];
    const int deg = (int)(adjptr[n + 1] - a0);
    const int rowlen = d * deg, blk = d * rowlen;
    EFB_LANES(lane) {
        for (int i = lane; i < blk; i += 32) acc[i] = 0.0;
    }
    for (long long s = rowptr[n]; s < rowptr[n + 1]; ++s) {
        const long long q = qlist[s];
        const int g = group_of_q(T, q);
        const int nPe = T.nPe[g], ndof = nPe * d;
        const long long ql = q - T.qoff[g];  // e*nPe + a
        const double* src = T.data[g] + ql * (long long)(d * ndof);  // rows a*d .. a*d+d-1 of K_e are contiguous
        const int* prow = pos + (T.poff[g] + ql * nPe);
        EFB_LANES(lane) {
            for (int i = lane; i < d * ndof; i += 32) {
                const int ii = i / ndof, j = i - ii * ndof;
                const int b = j / d, jj = j - b * d;
                acc[ii * rowlen + prow[b] * d + jj] += src[i];
            }
        }
    }
    EFB_LANES(lane) {
        double* dst = out + (long long)d * d * a0;
        for (int i = lane; i < blk; i += 32) dst[i] = acc[i];
    }
}

// single-group fast path: D and NPE are compile-time, so the (row, node, component) decomposition of a lane's values is
// hoisted out of every loop, and the values of SB sources are in flight before the ordered accumulation starts
template <int D, int NPE, int SB>
EFB_D void replay_node_fast(const double* EFB_RESTRICT data, long long n, const long long* EFB_RESTRICT rowptr,
                            const long long* EFB_RESTRICT qlist, const long long* EFB_RESTRICT adjptr,
                            const int* EFB_RESTRICT pos, double* acc, double* EFB_RESTRICT out) {
    constexpr int NDOF = D * NPE, NV = D * NDOF, VPL = (NV + 31) / 32;
    const long long a0 = adjptr[n];
    const int deg = (int)(adjptr[n + 1] - a0);
    const int rowlen = D * deg, blk = D * rowlen;
    const long long s_begin = rowptr[n], s_end = rowptr[n + 1];
    EFB_LANES(lane) {
        for (int i = lane; i < blk; i += 32) acc[i] = 0.0;
    }
    // per-lane registers on the device; the host emulation keeps one copy per emulated lane
    double v[EFB_LANE_COPIES][SB][VPL];
    int pb[EFB_LANE_COPIES][SB][VPL];
    for (long long s0 = s_begin; s0 < s_end; s0 += SB) {
        EFB_LANES(lane) {  // issue the loads of up to SB sources
            EFB_UNROLL
            for (int u = 0; u < SB; ++u) {
                if (s0 + u < s_end) {
                    const long long q = qlist[s0 + u];  // = e*NPE + a
                    const double* src = data + q * (long long)NV;
                    const int* prow = pos + q * NPE;
                    EFB_UNROLL
                    for (int k = 0; k < VPL; ++k) {
                        const int i = lane + 32 * k;
                        if (i < NV) {
                            v[EFB_LANE_SLOT(lane)][u][k] = src[i];
                            pb[EFB_LANE_SLOT(lane)][u][k] = prow[(i % NDOF) / D];
                        }
                    }
                }
            }
        }
        EFB_UNROLL
        for (int u = 0; u < SB; ++u) {  // ordered accumulation: source by source (ascending k of every slot)
            EFB_LANES(lane) {
                if (s0 + u < s_end) {
                    EFB_UNROLL
                    for (int k = 0; k < VPL; ++k) {
                        const int i = lane + 32 * k;
                        if (i < NV) acc[(i / NDOF) * rowlen + pb[EFB_LANE_SLOT(lane)][u][k] * D + (i % NDOF) % D] += v[EFB_LANE_SLOT(lane)][u][k];
                    }
                }
            }
        }
    }
    EFB_LANES(lane) {
        double* dst = out + (long long)D * D * a0;
        for (int i = lane; i < blk; i += 32) dst[i] = acc[i];
    }
}

// dense vector: F[n*d+ii] = ordered sum over the node's element rows, _simu.py:1075-1078 + bincount
EFB_HD void replay_vector_item(const GroupTable& T, int d, long long r, const long long* rowptr, const long long* qlist,
                               double* out) {
    const long long n = r / d;
    const int ii = (int)(r % d);
    double s = 0.0;
    for (long long k = rowptr[n]; k < rowptr[n + 1]; ++k) {
        const long long q = qlist[k];
        const int g = group_of_q(T, q);
        s += T.data[g][(q - T.qoff[g]) * d + ii];
    }
    out[r] = s;
}

}  // namespace efb
