// Fused element integration + CSR assembly for a homogeneous elasticity tensor (A3 `Assembly` for O1 + A2 in one kernel,
// EasyFEA/Simulations/_simu.py:1104-1144 with Operators/Bilinear.py:62-79): K_e never exists, neither in HBM nor in
// shared memory.  Body shared by the device kernel (fused_kernels.cu) and the host emulation of tests/hostcheck.
//
// Algebra.  With B = S G (elem_kernels.cuh, O1) and a homogeneous C2 = S C S, the block of nodes (a, b) of an element is
//     K_e[a,b][i][j] = sum_p w_p sum_{s,r} G[s,(a,i)] C2[s,r] G[r,(b,j)] = sum_{k,l} Dt[i][j][k][l] T_ab[k][l],
//     T_ab[k][l] = sum_p w_p (dN_a/dx_k)(dN_b/dx_l),   Dt[i][j][k][l] = sum_{s,r} E[s][i][k] C2[s][r] E[r][j][l],
// E the 0/1 pattern of G (G[s,(a,i)] = sum_k E[s][i][k] dN_a/dx_k).  Dt does not depend on the Gauss point, so the
// per-point work is the DIM x DIM outer product T_ab (9 FMAs in 3D instead of the 21-27 of the B^T C B form) and, being
// linear, Dt is applied ONCE per CSR block after the contributions of all elements have been summed:
//     K[n,m] = Dt : sum_{e, a, b : conn[e][a] = n, conn[e][b] = m} T_ab(e).
// Gradients are stored pre-multiplied by sqrt(w_p |det F|) (weights must be positive), so T_ab = sum_p ga (x) gb.
//
// Work decomposition.  Nodes are processed in CLUSTERS of S nodes that are close in space (a one-time schedule,
// easyfea_b200/assembly.py FusedSchedule: Morton order of the node lattice).  One CTA per cluster:
//   1. gather the coordinates of the cluster's elements (every element touching one of its nodes) -> shared memory,
//   2. G2-G6 for every (element, Gauss point) of the cluster: one task per thread -> scaled gradients in shared memory
//      (each element is integrated once per cluster that touches it: the geometry redundancy of the node-owned scheme),
//   3. node rows: a warp owns G = 32/LPN nodes; lane (i, b, h) walks the elements of node i in ascending order, computes
//      half h of the Gauss-point sum of T[a_j, b] of element j (CHUNK elements at a time: independent FMA chains), the two
//      halves meet in a shuffle, and the block is added to the node's accumulator at the slot of column node b
//      (read-modify-write in warp-private shared memory, a warp barrier between elements): every CSR slot is summed in
//      ascending element order by one warp — deterministic, no atomics,
//   4. Dt is applied and the node's d x d*deg block is written once, coalesced.
// HBM traffic per launch: connectivity + coordinates of the clusters' elements, the schedule, and the CSR data once.
#pragma once
#include "csr_kernels.cuh"
#include "elem_kernels.cuh"

namespace efb {

// K[i][j] = sum_{k,l} Dt[i*DIM + j][k*DIM + l] * T[k][l] (scale folded in).  `ortho`: Dt[(i,j)][(k,l)] is non-zero only for
// (k,l) = (i,j), for (k,l) = (j,i), and on the diagonal i == j for k == l — what an isotropic or axis-aligned orthotropic C
// gives (no normal/shear coupling, diagonal shear block); the kernel instantiated for it issues 21 FMAs per block, not 81.
struct FusedTerms {
    double Dt[9][9];
    int ortho;
};

template <int DIM>
EFB_HD bool fused_ortho_term(int i, int j, int k, int l) {
    return (i == k && j == l) || (i == l && j == k) || (i == j && k == l);
}

struct FusedView {
    int n_clusters, S, cap_e, max_deg;
    const long long* cl_nodes;  // (n_clusters*S, 4): node id (-1 = padding) | DD*adjptr[n] | deg + (cnt << 32) | first task
    const int* cl_ne;           // (n_clusters) elements of each cluster
    const int* cl_conn;         // (n_clusters, cap_e, NPE) coordinate rows of the clusters' elements, -1 = empty
    const int* desc;            // (n_tasks) local element index | local node a << 16, tasks of a node consecutive
    const int* tpos;            // (n_tasks, NPE) slot of column node b inside adj(node)
    double* out;                // CSR data
};

// NPG > 0: number of Gauss points known at compile time (the per-point loops unroll, addresses become immediates);
// NPG == 0: runtime count (generic instantiation).
template <int DIM, int NPE, int NPG>
struct Fused {
    static constexpr int LB = NPE <= 4 ? 4 : (NPE <= 8 ? 8 : 16);   // lanes per node along the column-node axis
    static constexpr int H = (NPG == 0 || NPG >= 4) ? 2 : 1;        // lanes sharing the Gauss points of one unit (p split)
    static constexpr int LPN = LB * H;                              // lanes per node
    static constexpr int G = 32 / LPN;                              // nodes per warp
    static constexpr int DD = DIM * DIM;
    static constexpr int SLOT = DD | 1;                             // odd stride of a block accumulator: slots spread over all banks
    static constexpr int TS = (DIM * NPE) | 1;                      // odd stride of one Gauss point in the dN table
    // stride of one Gauss point inside an element record [p][k][a]: the geometry tasks of a warp (lanes = Gauss points)
    // store without bank conflicts, and for HEXA8 the two p-halves read by the lanes (b, h=0) / (b, h=1) are 8 banks apart
    static constexpr int GPS = (DIM * NPE) % 8 == 0 ? DIM * NPE + 2 : ((DIM * NPE) | 1);
    static constexpr int CHUNK = 4;                                 // element steps whose descriptors travel in registers
    EFB_HD static int npg(int nPg) { return NPG ? NPG : nPg; }
    EFB_HD static int rec(int nPg) { return (npg(nPg) * GPS) | 1; }
    EFB_HD static int tables(int nPg) { return (npg(nPg) * TS + npg(nPg) + 1) & ~1; }
    EFB_HD static int acc_per_node(int max_deg) { return max_deg * SLOT; }
    EFB_HD static int region(int cap_e, int nwarps, int max_deg) {
        const int x = cap_e * NPE * DIM, a = nwarps * G * acc_per_node(max_deg);
        return ((x > a ? x : a) + 1) & ~1;
    }
    EFB_HD static size_t total(int nPg, int cap_e, int nwarps, int max_deg) {
        return (size_t)tables(nPg) + (size_t)cap_e * rec(nPg) + region(cap_e, nwarps, max_deg);
    }
};

// G2-G6 of one (element, Gauss point): X (NPE x DIM) -> out[k*NPE + a] = sqrt(w |det F|) dN_a/dx_k
template <int DIM, int NPE>
EFB_D void fused_geometry_task(const double* EFB_RESTRICT X, const double* EFB_RESTRICT dNp, double w, double* EFB_RESTRICT out) {
    double dn[DIM][NPE];
    EFB_UNROLL
    for (int r = 0; r < DIM; ++r)
        EFB_UNROLL
        for (int n = 0; n < NPE; ++n) dn[r][n] = dNp[r * NPE + n];
    double F[DIM * DIM], Fi[DIM * DIM];
    EFB_UNROLL
    for (int i = 0; i < DIM * DIM; ++i) F[i] = 0.0;
    EFB_UNROLL
    for (int n = 0; n < NPE; ++n) {
        double x[DIM];
        EFB_UNROLL
        for (int c = 0; c < DIM; ++c) x[c] = X[n * DIM + c];
        EFB_UNROLL
        for (int r = 0; r < DIM; ++r)
            EFB_UNROLL
            for (int c = 0; c < DIM; ++c) F[r * DIM + c] += dn[r][n] * x[c];
    }
    const double det = det_inv<DIM>(F, Fi);
    const double sw = sqrt(w * fabs(det));
    EFB_UNROLL
    for (int i = 0; i < DIM * DIM; ++i) Fi[i] *= sw;
    double gn[DIM][NPE];
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a) {
        EFB_UNROLL
        for (int d = 0; d < DIM; ++d) {
            double s = 0.0;
            EFB_UNROLL
            for (int k = 0; k < DIM; ++k) s += Fi[d * DIM + k] * dn[k][a];
            gn[d][a] = s;
        }
    }
    EFB_UNROLL
    for (int d = 0; d < DIM; ++d)
        EFB_UNROLL
        for (int a = 0; a < NPE; ++a) out[d * NPE + a] = gn[d][a];
}

// partial sum over the Gauss points [p0, p1) of T[a, b] of one element
template <int DIM, int NPE, int NPG>
EFB_D void fused_unit(const double* EFB_RESTRICT rec, int p0, int p1, int a, int b, double (&T)[DIM * DIM]) {
    constexpr int GPS = Fused<DIM, NPE, NPG>::GPS;
    constexpr int PMAX = NPG ? (NPG + Fused<DIM, NPE, NPG>::H - 1) / Fused<DIM, NPE, NPG>::H : 1;
    const double* ra = rec + a;
    const double* rb = rec + b;
    if constexpr (NPG != 0) {
        EFB_UNROLL
        for (int q = 0; q < PMAX; ++q) {  // fixed trip count: the loads of all points are issued before the first FMA
            const int p = p0 + q;
            if (NPG % Fused<DIM, NPE, NPG>::H == 0 || p < p1) {  // equal halves: no range check
                double ga[DIM], gb[DIM];
                EFB_UNROLL
                for (int k = 0; k < DIM; ++k) {
                    ga[k] = ra[p * GPS + k * NPE];
                    gb[k] = rb[p * GPS + k * NPE];
                }
                EFB_UNROLL
                for (int k = 0; k < DIM; ++k)
                    EFB_UNROLL
                    for (int l = 0; l < DIM; ++l) T[k * DIM + l] += ga[k] * gb[l];
            }
        }
    } else {
        for (int p = p0; p < p1; ++p) {
            double ga[DIM], gb[DIM];
            EFB_UNROLL
            for (int k = 0; k < DIM; ++k) {
                ga[k] = ra[p * GPS + k * NPE];
                gb[k] = rb[p * GPS + k * NPE];
            }
            EFB_UNROLL
            for (int k = 0; k < DIM; ++k)
                EFB_UNROLL
                for (int l = 0; l < DIM; ++l) T[k * DIM + l] += ga[k] * gb[l];
        }
    }
}

// per-lane registers of a node group: the node record and the descriptors of the next CHUNK element steps
template <int CHUNK>
struct FusedLane {
    long long out_off, task0;
    int node, deg, cnt;
    int dreg[CHUNK], sreg[CHUNK];
};

// node record of lane's node + descriptors of steps [j0, j0+CHUNK): independent loads, issued together
template <int DIM, int NPE, int NPG>
EFB_D void fused_load_lane(const FusedView& f, long long k0, int lane, int j0, bool with_record,
                           FusedLane<Fused<DIM, NPE, NPG>::CHUNK>& L) {
    using FC = Fused<DIM, NPE, NPG>;
    const int i = lane / FC::LPN, b = (lane % FC::LPN) / FC::H;
    if (with_record) {
        const long long* nr = f.cl_nodes + (k0 + i) * 4;
        L.node = (int)nr[0];
        L.out_off = nr[1];
        L.deg = (int)(nr[2] & 0xffffffffLL);
        L.cnt = (int)(nr[2] >> 32);
        L.task0 = nr[3];
    }
    EFB_UNROLL
    for (int u = 0; u < FC::CHUNK; ++u) {
        const bool on = j0 + u < L.cnt && b < NPE;
        L.dreg[u] = on ? f.desc[L.task0 + j0 + u] : -1;
        L.sreg[u] = on ? f.tpos[(L.task0 + j0 + u) * NPE + b] : 0;
    }
}

// the rows of the G nodes at schedule positions k0 .. k0+G-1, executed by one warp; `accw` = its G accumulators.
// `L`: lane registers already loaded with the node records and the descriptors of the first CHUNK steps.
template <int DIM, int NPE, int NPG, bool ORTHO>
EFB_D void fused_group(const FusedView& f, const FusedTerms& terms, int nPg, const double* recs, double* accw, long long k0,
                       FusedLane<Fused<DIM, NPE, NPG>::CHUNK> (&L)[EFB_LANE_COPIES]) {
    using FC = Fused<DIM, NPE, NPG>;
    constexpr int LPN = FC::LPN, G = FC::G, H = FC::H, DD = FC::DD, SLOT = FC::SLOT, CHUNK = FC::CHUNK;
    const int npg = FC::npg(nPg), RS = FC::rec(nPg), acc_node = FC::acc_per_node(f.max_deg);
    const int PH = (npg + H - 1) / H;  // Gauss points per lane of a unit
    int maxcnt = 0;
    for (int i = 0; i < G; ++i) {  // warp-uniform
        const int cnt = (int)(f.cl_nodes[(k0 + i) * 4 + 2] >> 32);
        maxcnt = cnt > maxcnt ? cnt : maxcnt;
    }
    EFB_LANES(lane) {
        const int i = lane / LPN;
        double* acc = accw + i * acc_node;
        for (int t = lane % LPN; t < L[EFB_LANE_SLOT(lane)].deg * SLOT; t += LPN) acc[t] = 0.0;
    }
    for (int j0 = 0; j0 < maxcnt; j0 += CHUNK) {
        if (j0 > 0) {
            EFB_LANES(lane) { fused_load_lane<DIM, NPE, NPG>(f, k0, lane, j0, false, L[EFB_LANE_SLOT(lane)]); }
        }
        double T[EFB_LANE_COPIES][CHUNK][DD];
        EFB_LANES(lane) {  // the partial blocks of up to CHUNK elements: independent work, no barrier in between
            const int b = (lane % LPN) / H, h = lane % H;
            const int p0 = h * PH, p1 = (p0 + PH < npg) ? p0 + PH : npg;
            EFB_UNROLL
            for (int u = 0; u < CHUNK; ++u) {
                EFB_UNROLL
                for (int c = 0; c < DD; ++c) T[EFB_LANE_SLOT(lane)][u][c] = 0.0;
                const int d = L[EFB_LANE_SLOT(lane)].dreg[u];
                if (d >= 0) fused_unit<DIM, NPE, NPG>(recs + (size_t)(d & 0xffff) * RS, p0, p1, d >> 16, b, T[EFB_LANE_SLOT(lane)][u]);
            }
        }
        EFB_UNROLL
        for (int u = 0; u < CHUNK; ++u) {
            if (j0 + u < maxcnt) {  // warp-uniform
                // ordered accumulation of element step j0+u (ascending element id for every slot).  The two lanes of a unit
                // split the DD components: lane h=0 owns [0, C0), lane h=1 owns [C0, DD); each sends the partner the halves of
                // the components the partner owns and adds what it receives to its own (x + y == y + x: one result per slot)
                constexpr int C0 = (H == 2) ? (DD + 1) / 2 : DD;
                EFB_LANES(lane) {
                    const int i = lane / LPN, h = lane % H;
                    const double* Tm = T[EFB_LANE_SLOT(lane)][u];
                    double own[C0];
                    EFB_UNROLL
                    for (int q = 0; q < C0; ++q) {
                        if (H == 2) {
                            const int cu = C0 + q < DD ? C0 + q : DD - 1;  // component of the upper half (clamped: unused when q is past it)
                            const double mine = h == 0 ? Tm[q] : Tm[cu];
#ifdef __CUDACC__
                            const double send = h == 0 ? Tm[cu] : Tm[q];
                            own[q] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
#else
                            const double* Tp = T[lane ^ 1][u];
                            own[q] = mine + (h == 0 ? Tp[q] : Tp[cu]);
#endif
                        } else {
                            own[q] = Tm[q];
                        }
                    }
                    if (L[EFB_LANE_SLOT(lane)].dreg[u] >= 0) {
                        double* dst = accw + i * acc_node + L[EFB_LANE_SLOT(lane)].sreg[u] * SLOT + (h == 0 ? 0 : C0);
                        const int nq = h == 0 ? C0 : DD - C0;
                        EFB_UNROLL
                        for (int q = 0; q < C0; ++q)
                            if (q < nq) dst[q] += own[q];
                    }
                }
            }
        }
    }
    for (int i = 0; i < G; ++i) {  // Dt : T per block (one lane per column node) and the write of the node's d x d*deg block
        const long long* nr = f.cl_nodes + (k0 + i) * 4;
        if (nr[0] < 0) continue;
        const int deg = (int)(nr[2] & 0xffffffffLL);
        const double* acc = accw + i * acc_node;
        double* dst = f.out + nr[1];
        const int rowlen = deg * DIM;
        EFB_LANES(lane) {
            for (int slot = lane; slot < deg; slot += 32) {
                double Tn[DD], K[DD];
                EFB_UNROLL
                for (int c = 0; c < DD; ++c) Tn[c] = acc[slot * SLOT + c];
                EFB_UNROLL
                for (int ii = 0; ii < DIM; ++ii)
                    EFB_UNROLL
                    for (int jj = 0; jj < DIM; ++jj) {
                        double v = 0.0;
                        EFB_UNROLL
                        for (int k = 0; k < DIM; ++k)
                            EFB_UNROLL
                            for (int l = 0; l < DIM; ++l)
                                if (!ORTHO || fused_ortho_term<DIM>(ii, jj, k, l)) v += terms.Dt[ii * DIM + jj][k * DIM + l] * Tn[k * DIM + l];
                        K[ii * DIM + jj] = v;
                    }
                EFB_UNROLL
                for (int ii = 0; ii < DIM; ++ii)
                    EFB_UNROLL
                    for (int jj = 0; jj < DIM; ++jj) dst[ii * rowlen + slot * DIM + jj] = K[ii * DIM + jj];
            }
        }
    }
}

// one cluster on one CTA; one warp per node group (launch: nthreads = 32 * S / G).  `load_tables`: first cluster of this CTA.
// The caller puts a CTA barrier between two clusters of the same CTA (the accumulators share memory with the next gather).
template <int DIM, int NPE, int NPG, bool ORTHO>
EFB_D void fused_cluster_block(const GroupView& g, const FusedView& f, const FusedTerms& terms, long long cluster, int nthreads,
                               double* smem, bool load_tables = true, long long next_cluster = -1) {
    using FC = Fused<DIM, NPE, NPG>;
    const int nPg = FC::npg(g.nPg), RS = FC::rec(nPg), nwarps = nthreads / 32;
    double* dNt = smem;
    double* wt = smem + nPg * FC::TS;
    double* recs = smem + FC::tables(nPg);
    double* region = recs + (size_t)f.cap_e * RS;  // element coordinates during the geometry stage, accumulators afterwards
    const long long e_lo = cluster * f.cap_e;  // the cluster's slab of cl_conn: no pointer load in front of the gather
    const int nE = f.cl_ne[cluster];
    const int acc_warp = FC::G * FC::acc_per_node(f.max_deg);
#ifdef __CUDACC__
    // the node records and first descriptors of this warp's group travel while the geometry is computed
    FusedLane<FC::CHUNK> L[1];
    const long long k0 = cluster * f.S + (long long)(threadIdx.x >> 5) * FC::G;
    fused_load_lane<DIM, NPE, NPG>(f, k0, threadIdx.x & 31, 0, true, L[0]);
    // the gather of the NEXT cluster of this CTA is a two-level chain (connectivity -> coordinates): its connectivity entries
    // are read now and its coordinate lines are pulled into L2 once the geometry stage below is done, so that the chain costs
    // two L2 hits instead of two DRAM round trips at the head of the next cluster
    int pf_node[2] = {-1, -1};
    if (next_cluster >= 0) {
        EFB_UNROLL
        for (int q = 0; q < 2; ++q) {
            const int idx = (int)threadIdx.x + q * nthreads;
            if (idx < f.cap_e * NPE) pf_node[q] = f.cl_conn[next_cluster * f.cap_e * NPE + idx];
        }
    }
#endif
    EFB_PHASE(tid, nthreads) {
        if (load_tables) {
            for (int i = tid; i < nPg * DIM * NPE; i += nthreads) dNt[(i / (DIM * NPE)) * FC::TS + i % (DIM * NPE)] = g.dN_pg[i];
            for (int i = tid; i < nPg; i += nthreads) wt[i] = g.w_pg[i];
        }
        for (int idx = tid; idx < f.cap_e * NPE; idx += nthreads) {
            const int node = f.cl_conn[e_lo * NPE + idx];
            if (node >= 0) {
                const double* src = g.coord + (long long)node * g.coord_stride;
                EFB_UNROLL
                for (int d = 0; d < DIM; ++d) region[idx * DIM + d] = src[d];
            }
        }
    }
    EFB_PHASE(tid, nthreads) {
        for (int task = tid; task < nE * nPg; task += nthreads) {
            const int le = task / nPg, p = task - le * nPg;
            fused_geometry_task<DIM, NPE>(region + le * (NPE * DIM), dNt + p * FC::TS, wt[p], recs + (size_t)le * RS + p * FC::GPS);
        }
    }
#ifdef __CUDACC__
    EFB_UNROLL
    for (int q = 0; q < 2; ++q)
        if (pf_node[q] >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.coord + (long long)pf_node[q] * g.coord_stride));
    fused_group<DIM, NPE, NPG, ORTHO>(f, terms, nPg, recs, region + (threadIdx.x >> 5) * acc_warp, k0, L);
#else
    for (int wi = 0; wi < nwarps; ++wi) {
        FusedLane<FC::CHUNK> L[EFB_LANE_COPIES];
        const long long k0 = cluster * f.S + (long long)wi * FC::G;
        for (int lane = 0; lane < 32; ++lane) fused_load_lane<DIM, NPE, NPG>(f, k0, lane, 0, true, L[lane]);
        fused_group<DIM, NPE, NPG, ORTHO>(f, terms, nPg, recs, region + wi * acc_warp, k0, L);
    }
#endif
}

// Dt terms of a homogeneous C2 = S C S (row-major ns x ns), scaled; host side
template <int DIM>
inline void fused_terms_from_C2(const double* C2, double scale, FusedTerms& t) {
    constexpr int NS = StrainSize<DIM>::value;
    // E[s][i][k] = 1 where strain row s of G takes dN/dx_k for displacement component i (elem_kernels.cuh, O1)
    int E[NS][DIM][DIM] = {};
    if constexpr (DIM == 2) {  // G[:, (a,0)] = (gx, 0, gy); G[:, (a,1)] = (0, gy, gx)
        E[0][0][0] = 1; E[2][0][1] = 1;
        E[1][1][1] = 1; E[2][1][0] = 1;
    } else {  // G[:, (a,0)] = (gx,0,0,0,gz,gy); G[:, (a,1)] = (0,gy,0,gz,0,gx); G[:, (a,2)] = (0,0,gz,gy,gx,0)
        E[0][0][0] = 1; E[4][0][2] = 1; E[5][0][1] = 1;
        E[1][1][1] = 1; E[3][1][2] = 1; E[5][1][0] = 1;
        E[2][2][2] = 1; E[3][2][1] = 1; E[4][2][0] = 1;
    }
    t.ortho = 1;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j)
            for (int k = 0; k < DIM; ++k)
                for (int l = 0; l < DIM; ++l) {
                    double c = 0.0;
                    for (int s = 0; s < NS; ++s)
                        for (int r = 0; r < NS; ++r)
                            if (E[s][i][k] && E[r][j][l]) c += C2[s * NS + r];
                    t.Dt[i * DIM + j][k * DIM + l] = scale * c;
                    if (c != 0.0 && !fused_ortho_term<DIM>(i, j, k, l)) t.ortho = 0;
                }
}

}  // namespace efb
