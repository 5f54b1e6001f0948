// Block bodies of the element kernels (rows G2-G11, O1-O4, P1 of SURVEY.md §8a).
//
// Layout of one CTA: EPB elements x TPE threads per element (TPE = ndof for vector problems, nPe for scalar).
// Shared memory: the reference-element tables of the group once per CTA, then per element the nodal
// coordinates, F / F^-1 / det / wJ at every Gauss point and the physical shape-function gradients
// gN[p][a][d] = dN_a/dx_d.  B (Kelvin-Mandel, _group_elem.py:1241-1312) is never materialised: every B column has
// `dim` non-zeros, which the contractions below spell out.
#pragma once
#include "frame.cuh"

namespace efb {

struct GroupView {
    int nPg;
    int coord_stride;
    long long Ne;
    const int* connect;
    const double* coord;
    const double* dN_pg;
    const double* N_pg;
    const double* w_pg;
};

// ---------------------------------------------------------------------------------------------------------
// shared-memory map (offsets in doubles)
// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
struct SmemMap {
    // doubles per node in the gradient table gN[p][a][GS]: 3D rows are padded to 4 so (gx,gy) / (gz,-) are two
    // 16-byte loads
    static constexpr int GS = (DIM == 3) ? 4 : 2;
    int nPg, EPB, extra;  // extra = op-specific doubles per element
    bool grad;            // false: no gN table (operators that need no physical gradients, e.g. the mass matrix)
    EFB_HD SmemMap(int nPg_, int EPB_, int extra_, bool grad_ = true) : nPg(nPg_), EPB(EPB_), extra(extra_), grad(grad_) {}
    // tables
    EFB_HD int off_dN() const { return 0; }
    EFB_HD int off_N() const { return nPg * DIM * NPE; }
    EFB_HD int off_w() const { return off_N() + nPg * NPE; }
    EFB_HD int tables() const { return off_w() + nPg; }
    // per element
    EFB_HD int o_X() const { return 0; }
    EFB_HD int o_F() const { return NPE * DIM; }
    EFB_HD int o_Fi() const { return o_F() + nPg * DIM * DIM; }
    EFB_HD int o_det() const { return o_Fi() + nPg * DIM * DIM; }
    EFB_HD int o_wJ() const { return o_det() + nPg; }
    EFB_HD int o_gN() const { return (o_wJ() + nPg + 1) & ~1; }  // 16-byte aligned
    EFB_HD int o_extra() const { return o_gN() + (grad ? nPg * NPE * GS : 0); }
    // even stride (keeps 16-byte alignment) that is not a multiple of 16 doubles, so the same field of the elements
    // sharing a warp falls in different banks
    EFB_HD int per_elem() const {
        int n = (o_extra() + extra + 1) & ~1;
        return (n % 16 == 0) ? n + 2 : n;
    }
    EFB_HD int tables_padded() const { return (tables() + 1) & ~1; }
    EFB_HD int total() const { return tables_padded() + EPB * per_elem(); }
    EFB_HD double* elem(double* smem, int el) const { return smem + tables_padded() + el * per_elem(); }
};

// closed-form det / inverse with the reference's association of products, EasyFEA/FEM/_linalg.py:533-656
template <int DIM>
EFB_HD double det_inv(const double* F, double* Fi);

template <>
EFB_HD double det_inv<2>(const double* F, double* Fi) {
    const double a = F[0], b = F[1], c = F[2], d = F[3];
    const double det = (a * d) - (c * b);
    const double r = 1.0 / det;
    Fi[0] = r * d;
    Fi[1] = r * (-b);
    Fi[2] = r * (-c);
    Fi[3] = r * a;
    return det;
}

template <>
EFB_HD double det_inv<3>(const double* F, double* Fi) {
    const double a11 = F[0], a12 = F[1], a13 = F[2];
    const double a21 = F[3], a22 = F[4], a23 = F[5];
    const double a31 = F[6], a32 = F[7], a33 = F[8];
    const double det = a11 * ((a22 * a33) - (a32 * a23)) - a12 * ((a21 * a33) - (a31 * a23)) +
                       a13 * ((a21 * a32) - (a31 * a22));
    const double r = 1.0 / det;
    // adjugate (transposed cofactors)
    Fi[0] = r * ((a22 * a33) - (a23 * a32));
    Fi[1] = r * (-((a12 * a33) - (a13 * a32)));
    Fi[2] = r * ((a12 * a23) - (a13 * a22));
    Fi[3] = r * (-((a21 * a33) - (a23 * a31)));
    Fi[4] = r * ((a11 * a33) - (a13 * a31));
    Fi[5] = r * (-((a11 * a23) - (a13 * a21)));
    Fi[6] = r * ((a21 * a32) - (a22 * a31));
    Fi[7] = r * (-((a11 * a32) - (a12 * a31)));
    Fi[8] = r * ((a11 * a22) - (a12 * a21));
    return det;
}

// ---------------------------------------------------------------------------------------------------------
// G2-G6: tables + coordinates -> F, det, wJ, F^-1, gN for the EPB elements of this CTA (4 phases)
// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
EFB_D void geometry_phases(const GroupView& g, const SmemMap<DIM, NPE>& sm, long long e0, int TPE, int nthreads,
                           double* smem, bool need_grad, bool warp_local = false) {
    const int nPg = g.nPg;
    double* dNt = smem + sm.off_dN();
    double* Nt = smem + sm.off_N();
    double* wt = smem + sm.off_w();

    EFB_PHASE(tid, nthreads) {
        for (int i = tid; i < nPg * DIM * NPE; i += nthreads) dNt[i] = g.dN_pg[i];
        for (int i = tid; i < nPg * NPE; i += nthreads) Nt[i] = g.N_pg[i];
        for (int i = tid; i < nPg; i += nthreads) wt[i] = g.w_pg[i];
        const int el = tid / TPE, t = tid % TPE;
        const long long e = e0 + el;
        if (el < sm.EPB && e < g.Ne) {
            double* X = sm.elem(smem, el) + sm.o_X();
            for (int i = t; i < NPE * DIM; i += TPE) {
                const int a = i / DIM, d = i % DIM;
                X[i] = g.coord[(long long)g.connect[e * NPE + a] * g.coord_stride + d];
            }
        }
    }
    EFB_PHASE_E(tid, nthreads, warp_local) {  // F[p][r][c] = sum_n dN[p][r][n] x[n][c]   _group_elem.py:864-867
        const int el = tid / TPE, t = tid % TPE;
        if (el < sm.EPB && e0 + el < g.Ne) {
            double* E = sm.elem(smem, el);
            const double* X = E + sm.o_X();
            double* F = E + sm.o_F();
            for (int i = t; i < nPg * DIM * DIM; i += TPE) {
                const int pr = i / DIM, c = i % DIM;
                const double* row = dNt + pr * NPE;
                double s = 0.0;
                EFB_UNROLL
                for (int n = 0; n < NPE; ++n) s += row[n] * X[n * DIM + c];
                F[i] = s;
            }
        }
    }
    EFB_PHASE_E(tid, nthreads, warp_local) {  // det, |det| w, inverse                      :871-915
        const int el = tid / TPE, t = tid % TPE;
        if (el < sm.EPB && e0 + el < g.Ne) {
            double* E = sm.elem(smem, el);
            for (int p = t; p < nPg; p += TPE) {
                const double det = det_inv<DIM>(E + sm.o_F() + p * DIM * DIM, E + sm.o_Fi() + p * DIM * DIM);
                E[sm.o_det() + p] = det;
                E[sm.o_wJ() + p] = fabs(det) * wt[p];
            }
        }
    }
    if (need_grad) {
        EFB_PHASE_E(tid, nthreads, warp_local) {  // gN[p][a][d] = sum_k Fi[p][d][k] dN[p][k][a]   :1083-1105
            const int el = tid / TPE, t = tid % TPE;
            if (el < sm.EPB && e0 + el < g.Ne) {
                double* E = sm.elem(smem, el);
                const double* Fi = E + sm.o_Fi();
                double* gN = E + sm.o_gN();
                constexpr int GS = SmemMap<DIM, NPE>::GS;
                for (int i = t; i < nPg * NPE * DIM; i += TPE) {
                    const int p = i / (NPE * DIM), a = (i / DIM) % NPE, d = i % DIM;
                    double s = 0.0;
                    EFB_UNROLL
                    for (int k = 0; k < DIM; ++k) s += Fi[(p * DIM + d) * DIM + k] * dNt[(p * DIM + k) * NPE + a];
                    gN[(p * NPE + a) * GS + d] = s;
                }
            }
        }
    }
}

// entry B[s][(a,d)] of the Kelvin-Mandel strain operator from the gradient (gx,gy,gz) of node a
template <int DIM>
EFB_HD double B_entry(int s, int d, const double* g) {
    if constexpr (DIM == 2) {
        if (s == 0) return d == 0 ? g[0] : 0.0;
        if (s == 1) return d == 1 ? g[1] : 0.0;
        return kInvSqrt2 * g[1 - d];
    } else {
        if (s < 3) return d == s ? g[s] : 0.0;
        if (s == 3) return d == 0 ? 0.0 : kInvSqrt2 * g[d == 1 ? 2 : 1];  // yz
        if (s == 4) return d == 1 ? 0.0 : kInvSqrt2 * g[d == 0 ? 2 : 0];  // xz
        return d == 2 ? 0.0 : kInvSqrt2 * g[d == 0 ? 1 : 0];              // xy
    }
}

// coefficient with the broadcast modes of FeArray.broadcast
EFB_HD double coef_at(const double* c, int mode, double scalar, long long e, int p, int nPg) {
    if (c == nullptr || mode == 0) return c ? c[0] : scalar;
    if (mode == 1) return c[e];
    if (mode == 2) return c[p];
    return c[e * nPg + p];
}

// ---------------------------------------------------------------------------------------------------------
// G2-G8 export                                                        _group_elem.py:832-1312
// ---------------------------------------------------------------------------------------------------------
struct GeomOut {
    double *F, *detF, *jac, *wJ, *invF, *dN, *B;
};

template <int DIM, int NPE>
EFB_D void geometry_block(const GroupView& g, const GeomOut& o, int EPB, long long blockId, int nthreads, double* smem) {
    constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    const int TPE = nthreads / EPB;
    const SmemMap<DIM, NPE> sm(g.nPg, EPB, 0);
    const long long e0 = blockId * EPB;
    const int nPg = g.nPg;
    geometry_phases<DIM, NPE>(g, sm, e0, TPE, nthreads, smem, true);
    EFB_PHASE(tid, nthreads) {
        const int el = tid / TPE, t = tid % TPE;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* E = sm.elem(smem, el);
            for (int i = t; i < nPg * DIM * DIM; i += TPE) {
                if (o.F) o.F[e * nPg * DIM * DIM + i] = E[sm.o_F() + i];
                if (o.invF) o.invF[e * nPg * DIM * DIM + i] = E[sm.o_Fi() + i];
            }
            for (int p = t; p < nPg; p += TPE) {
                if (o.detF) o.detF[e * nPg + p] = E[sm.o_det() + p];
                if (o.jac) o.jac[e * nPg + p] = fabs(E[sm.o_det() + p]);
                if (o.wJ) o.wJ[e * nPg + p] = E[sm.o_wJ() + p];
            }
            const double* gN = E + sm.o_gN();
            if (o.dN) {
                for (int i = t; i < nPg * DIM * NPE; i += TPE) {  // output layout (p, d, a)
                    const int p = i / (DIM * NPE), d = (i / NPE) % DIM, a = i % NPE;
                    o.dN[e * nPg * DIM * NPE + i] = gN[(p * NPE + a) * SmemMap<DIM, NPE>::GS + d];
                }
            }
            if (o.B) {
                for (int i = t; i < nPg * NS * NDOF; i += TPE) {  // (p, s, a*DIM+d)
                    const int p = i / (NS * NDOF), s = (i / NDOF) % NS, col = i % NDOF;
                    o.B[e * (long long)(nPg * NS * NDOF) + i] = B_entry<DIM>(s, col % DIM, gN + (p * NPE + col / DIM) * SmemMap<DIM, NPE>::GS);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// G8-G10 export: the cached per-Gauss-point factors of the reference        _group_elem.py:1314-1407
//   leftDisp (Ne,nPg,nPe*DIM,ns)        = wJ B^T                      Get_leftDispPart_e_pg :1315
//   reaction (Ne,nPg,nPe*dof_n,nPe*dof_n) = (wJ N^T) N, N block-diagonal Get_ReactionPart_e_pg :1338
//   diffuse  (Ne,nPg,nPe,DIM)           = wJ dN^T                     Get_DiffusePart_e_pg  :1363
//   source   (Ne,nPg,nPe*dof_n,dof_n)   = wJ N^T                      Get_SourcePart_e_pg   :1383
// ---------------------------------------------------------------------------------------------------------
struct GeomPartsOut {
    double *leftDisp, *reaction, *diffuse, *source;
    int dof_n;
};

template <int DIM, int NPE>
EFB_D void geometry_parts_block(const GroupView& g, const GeomPartsOut& o, int EPB, long long blockId, int nthreads, double* smem) {
    constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    const int TPE = nthreads / EPB;
    const bool grad = o.leftDisp || o.diffuse;
    const SmemMap<DIM, NPE> sm(g.nPg, EPB, 0, grad);
    const long long e0 = blockId * EPB;
    const int nPg = g.nPg;
    const double* Nt = smem + sm.off_N();
    geometry_phases<DIM, NPE>(g, sm, e0, TPE, nthreads, smem, grad);
    EFB_PHASE(tid, nthreads) {
        const int el = tid / TPE, t = tid % TPE;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* E = sm.elem(smem, el);
            const double* wJ = E + sm.o_wJ();
            const double* gN = E + sm.o_gN();
            constexpr int GS = SmemMap<DIM, NPE>::GS;
            if (o.leftDisp) {
                for (int i = t; i < nPg * NDOF * NS; i += TPE) {  // (p, col = a*DIM+d, s)
                    const int p = i / (NDOF * NS), col = (i / NS) % NDOF, s = i % NS;
                    o.leftDisp[e * (long long)(nPg * NDOF * NS) + i] = wJ[p] * B_entry<DIM>(s, col % DIM, gN + (p * NPE + col / DIM) * GS);
                }
            }
            if (o.diffuse) {
                for (int i = t; i < nPg * NPE * DIM; i += TPE) {  // (p, a, d)
                    const int p = i / (NPE * DIM);
                    o.diffuse[e * (long long)(nPg * NPE * DIM) + i] = wJ[p] * gN[(i / DIM) * GS + i % DIM];
                }
            }
            const int dn = o.dof_n, nd = NPE * dn;
            if (o.reaction) {
                for (int i = t; i < nPg * nd * nd; i += TPE) {  // (p, r, c)
                    const int p = i / (nd * nd), r = (i / nd) % nd, c = i % nd;
                    o.reaction[e * (long long)(nPg * nd * nd) + i] =
                        (r % dn == c % dn) ? (wJ[p] * Nt[p * NPE + r / dn]) * Nt[p * NPE + c / dn] : 0.0;
                }
            }
            if (o.source) {
                for (int i = t; i < nPg * nd * dn; i += TPE) {  // (p, r, comp)
                    const int p = i / (nd * dn), r = (i / dn) % nd, comp = i % dn;
                    o.source[e * (long long)(nPg * nd * dn) + i] = (r % dn == comp) ? wJ[p] * Nt[p * NPE + r / dn] : 0.0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// O1: K_e = scale * sum_p wJ B^T C B                                   Operators/Bilinear.py:62-79
//
// With S = diag(1,..,1, 1/sqrt2,..) the Kelvin-Mandel operator is B = S G, G holding the plain gradients (DIM non-zeros
// per column), so K_e = sum_p w G^T (S C S) G: a homogeneous C reaches the kernel as C2 = S C S in the constant bank;
// per-element / per-Gauss-point C is staged raw in shared memory and the S factors are applied to the (few) row vectors.
//
// Work decomposition: thread (el, a, h) owns the DIM rows of node a of element el, restricted to the NB = NPE/CS column
// nodes of chunk h: DIM*NB*DIM accumulators in registers (72 for HEXA8, <= 81 always).  Per Gauss point the thread forms
// bc[i][:] = w G[:, (a,i)]^T C2 once (3 DIM NS flops) and then spends exactly DIM FMAs per accumulator; the only
// shared-memory traffic of the inner loop is one broadcast read of the NB column-node gradients, i.e. ~10 FP64 FMAs per
// shared-memory wavefront, which keeps the kernel on the FP64 pipe instead of the shared-memory pipe (measured: with one
// row per thread the same contraction is bound by shared-memory wavefronts, profiles/r1_ke_variants.md).
// A warp holds EPW = 32/NPE elements of one chunk; a CTA has NCW consumer warps + one producer warp (elem_kernels.cu).
// Geometry (G2-G6) is one producer task per (element, Gauss point): F, det, F^-1 stay in registers, only w|det F| and
// the physical gradients go to shared memory.
// ---------------------------------------------------------------------------------------------------------
struct CMat {
    double v[36];
};

struct alignas(16) Pair {
    double x, y;
};

// C -> S C S in place (row-major ns x ns)
template <int DIM>
EFB_HD double kelvin_factor(int s, int r) {
    const bool a = s >= DIM, b = r >= DIM;
    return (a && b) ? 0.5 : ((a || b) ? kInvSqrt2 : 1.0);
}

template <int DIM>
inline void prescale_C(CMat& C) {
    constexpr int NS = StrainSize<DIM>::value;
    for (int s = 0; s < NS; ++s)
        for (int r = 0; r < NS; ++r) C.v[s * NS + r] *= kelvin_factor<DIM>(s, r);
}

template <int DIM, int NPE>
struct ElasticTile {
    static constexpr int NDOF = DIM * NPE;
    // column chunks: a thread keeps DIM rows x NB column nodes, DIM*NB*DIM <= 81 accumulators
    static constexpr int CS = (DIM * NPE * DIM <= 81) ? 1 : ((NPE % 4 == 0 && DIM * (NPE / 2) * DIM > 81) ? 4 : (NPE % 3 == 0 ? 3 : 2));
    static constexpr int NB = NPE / CS;               // column nodes per thread
    static constexpr int NACC = DIM * NB * DIM;       // accumulators per thread
    static constexpr int EPW = 32 / NPE;              // elements per warp
    static constexpr int G = (CS == 1) ? 3 : 1;       // element groups per CTA
    static constexpr int NCW = G * CS;                // consumer warps
    static constexpr int EPB = G * EPW;               // elements per CTA batch
    static constexpr int THREADS = NCW * 32;          // consumer threads
    // a lane's rows leave through bulk asynchronous copies when every piece is a 16-byte multiple at a 16-byte address
    static constexpr bool kBulk = (NB * DIM) % 2 == 0 && NDOF % 2 == 0;
    // per-lane staging tile: NACC doubles padded to an ODD number of 16-byte units -> conflict-free 128-bit stores
    static constexpr int LANE_STAGE = 2 * ((((NACC + 1) / 2) | 1));
    static_assert(NB * CS == NPE, "CS must divide NPE");
    static_assert(NACC <= 81, "too many accumulators per thread");
    static_assert(EPW >= 1, "an element needs at most one warp of lanes");
};

// shared-memory map of the stiffness kernel (offsets in doubles)
template <int DIM, int NPE>
struct ElasticSmem {
    static constexpr int GS = DIM;                    // doubles per node in gN (packed)
    static constexpr int TS = (DIM * NPE) | 1;        // odd stride of one Gauss point in the dN table: the geometry tasks of
                                                      // a warp read different Gauss points without bank conflicts
    static constexpr int GPS = (NPE * GS) | 1;        // odd stride of one Gauss point in gN (same reason)
    static constexpr int KE = DIM * NPE * DIM * NPE;  // one element matrix
    int nPg, EPB, extra;
    EFB_HD ElasticSmem(int nPg_, int EPB_, int extra_) : nPg(nPg_), EPB(EPB_), extra(extra_) {}
    EFB_HD int off_stage() const { return 0; }  // first: 16-byte aligned for the bulk copies
    EFB_HD int stage_doubles() const { return ElasticTile<DIM, NPE>::NCW * 32 * ElasticTile<DIM, NPE>::LANE_STAGE; }
    EFB_HD int off_dN() const { return (stage_doubles() + 1) & ~1; }
    EFB_HD int off_w() const { return off_dN() + nPg * TS; }
    EFB_HD int tables_padded() const { return (off_w() + nPg + 1) & ~1; }
    EFB_HD int o_X() const { return 0; }
    EFB_HD int o_wJ() const { return NPE * DIM; }
    EFB_HD int o_gN() const { return o_wJ() + nPg; }
    EFB_HD int o_extra() const { return (o_gN() + nPg * GPS + 1) & ~1; }  // 16-byte aligned (filled by cp.async)
    // even stride (alignment of `extra`), not a multiple of 16 doubles: the same field of the elements sharing a warp
    // starts in different banks
    EFB_HD int per_elem() const {
        int n = (o_extra() + extra + 1) & ~1;
        return (n % 16 == 0) ? n + 2 : n;
    }
    EFB_HD int total() const { return tables_padded() + EPB * per_elem(); }
    EFB_HD int total2() const { return tables_padded() + 2 * EPB * per_elem(); }  // two geometry buffers (pipelined kernel)
    EFB_HD double* elem(double* smem, int el) const { return smem + tables_padded() + el * per_elem(); }
};

// DIM rows of node a x columns of nodes [b0, b0+NB): acc[i][b*DIM + j], all Gauss points.  `Cs` = raw C of the element
// (CMODE 1) or of its Gauss points (CMODE 2) in shared memory.
template <int DIM, int NPE, int CMODE>
EFB_D void elastic_rows(const CMat& C2const, const double* EFB_RESTRICT Cs, const double* EFB_RESTRICT wJ,
                        const double* EFB_RESTRICT gN, int nPg, int a, int b0,
                        double (&acc)[DIM][ElasticTile<DIM, NPE>::NB * DIM]) {
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS;
    using SM = ElasticSmem<DIM, NPE>;
    constexpr int GS = SM::GS, GPS = SM::GPS, NB = ElasticTile<DIM, NPE>::NB;
    EFB_UNROLL
    for (int i = 0; i < DIM; ++i)
        EFB_UNROLL
        for (int j = 0; j < NB * DIM; ++j) acc[i][j] = 0.0;
    for (int p = 0; p < nPg; ++p) {
        const double* gp = gN + p * GPS;
        const double w = wJ[p];
        // C2(s, r) of this Gauss point: constant bank (mode 0), or raw C from shared memory with the Kelvin-Mandel factor
        // of row s folded into the weights (fs) and that of column r applied to bc afterwards (fr)
#define EFB_C(s_, r_) (CMODE == 0 ? C2const.v[(s_) * NS + (r_)] : Cs[(CMODE == 2 ? p * NC : 0) + (s_) * NS + (r_)])
        constexpr double fs = (CMODE == 0) ? 1.0 : kInvSqrt2;
        double bc[DIM][NS];
        if constexpr (DIM == 2) {
            // G[:, (a,0)] = (gx, 0, gy); G[:, (a,1)] = (0, gy, gx)
            const double wx = w * gp[a * GS], wy = w * gp[a * GS + 1];
            const double sx = fs * wx, sy = fs * wy;
            EFB_UNROLL
            for (int r = 0; r < NS; ++r) {
                bc[0][r] = wx * EFB_C(0, r) + sy * EFB_C(2, r);
                bc[1][r] = wy * EFB_C(1, r) + sx * EFB_C(2, r);
            }
            if (CMODE != 0) {
                bc[0][2] *= kInvSqrt2;
                bc[1][2] *= kInvSqrt2;
            }
            EFB_UNROLL
            for (int b = 0; b < NB; ++b) {
                const double gx = gp[(b0 + b) * GS], gy = gp[(b0 + b) * GS + 1];
                EFB_UNROLL
                for (int i = 0; i < 2; ++i) {
                    double s0 = acc[i][b * 2 + 0], s1 = acc[i][b * 2 + 1];
                    s0 += bc[i][0] * gx;
                    s0 += bc[i][2] * gy;
                    s1 += bc[i][1] * gy;
                    s1 += bc[i][2] * gx;
                    acc[i][b * 2 + 0] = s0;
                    acc[i][b * 2 + 1] = s1;
                }
            }
        } else {
            // G[:, (a,0)] = (gx,0,0,0,gz,gy); G[:, (a,1)] = (0,gy,0,gz,0,gx); G[:, (a,2)] = (0,0,gz,gy,gx,0)
            const double wx = w * gp[a * GS], wy = w * gp[a * GS + 1], wz = w * gp[a * GS + 2];
            const double sx = fs * wx, sy = fs * wy, sz = fs * wz;
            EFB_UNROLL
            for (int r = 0; r < NS; ++r) {
                bc[0][r] = wx * EFB_C(0, r) + sz * EFB_C(4, r) + sy * EFB_C(5, r);
                bc[1][r] = wy * EFB_C(1, r) + sz * EFB_C(3, r) + sx * EFB_C(5, r);
                bc[2][r] = wz * EFB_C(2, r) + sy * EFB_C(3, r) + sx * EFB_C(4, r);
            }
            if (CMODE != 0) {
                EFB_UNROLL
                for (int i = 0; i < 3; ++i)
                    EFB_UNROLL
                    for (int r = 3; r < 6; ++r) bc[i][r] *= kInvSqrt2;
            }
            EFB_UNROLL
            for (int b = 0; b < NB; ++b) {
                const double gx = gp[(b0 + b) * GS], gy = gp[(b0 + b) * GS + 1], gz = gp[(b0 + b) * GS + 2];
                EFB_UNROLL
                for (int i = 0; i < 3; ++i) {
                    double s0 = acc[i][b * 3 + 0], s1 = acc[i][b * 3 + 1], s2 = acc[i][b * 3 + 2];
                    s0 += bc[i][0] * gx;
                    s0 += bc[i][4] * gz;
                    s0 += bc[i][5] * gy;
                    s1 += bc[i][1] * gy;
                    s1 += bc[i][3] * gz;
                    s1 += bc[i][5] * gx;
                    s2 += bc[i][2] * gz;
                    s2 += bc[i][3] * gy;
                    s2 += bc[i][4] * gx;
                    acc[i][b * 3 + 0] = s0;
                    acc[i][b * 3 + 1] = s1;
                    acc[i][b * 3 + 2] = s2;
                }
            }
        }
#undef EFB_C
    }
}

// ---- the stages of one batch of EPB elements, shared by the phase-structured body (host emulation) and the
// ---- warp-specialised device kernel.  `E0` = first per-element record of the geometry buffer in use.

// gather: nodal coordinates of the batch -> shared memory; executed by `nth` cooperating threads
template <int DIM, int NPE>
EFB_D void elastic_gather(const GroupView& g, const ElasticSmem<DIM, NPE>& sm, long long e0, int nvalid, double* E0, int t, int nth) {
    const int pe = sm.per_elem();
    for (int idx = t; idx < nvalid * NPE; idx += nth) {
        const int el = idx / NPE, a = idx - el * NPE;
        const double* src = g.coord + (long long)g.connect[(e0 + el) * NPE + a] * g.coord_stride;
        double* X = E0 + el * pe + sm.o_X();
        EFB_UNROLL
        for (int d = 0; d < DIM; ++d) X[a * DIM + d] = src[d];
    }
}

// raw copy of the batch's C (per element: `extra` contiguous doubles) -> shared memory (host emulation / fallback)
template <int DIM, int NPE>
EFB_D void elastic_gather_C(const ElasticSmem<DIM, NPE>& sm, const double* EFB_RESTRICT C, long long e0, int nvalid, double* E0,
                            int t, int nth) {
    const int extra = sm.extra, pe = sm.per_elem();
    for (int idx = t; idx < nvalid * extra; idx += nth) {
        const int el = idx / extra, i = idx - el * extra;
        (E0 + el * pe + sm.o_extra())[i] = C[(e0 + el) * (long long)extra + i];
    }
}

// G2-G6 for one (element, Gauss point), in registers                    _group_elem.py:832-1105
// (reference gradients cached in registers, all gN values formed before the first store: see warp_geometry_task)
template <int DIM, int NPE>
EFB_D void elastic_geometry_task(const ElasticSmem<DIM, NPE>& sm, const double* EFB_RESTRICT dNt, const double* EFB_RESTRICT wt,
                                 double scale, double* EFB_RESTRICT E, int p) {
    using SM = ElasticSmem<DIM, NPE>;
    constexpr int GS = SM::GS, TS = SM::TS, GPS = SM::GPS;
    const double* X = E + sm.o_X();
    const double* dNp = dNt + p * TS;
    double F[DIM * DIM], Fi[DIM * DIM];
    EFB_UNROLL
    for (int i = 0; i < DIM * DIM; ++i) F[i] = 0.0;
    if constexpr (DIM * NPE <= 30) {
        double dn[DIM][NPE];
        EFB_UNROLL
        for (int r = 0; r < DIM; ++r)
            EFB_UNROLL
            for (int n = 0; n < NPE; ++n) dn[r][n] = dNp[r * NPE + n];
        EFB_UNROLL
        for (int n = 0; n < NPE; ++n) {
            EFB_UNROLL
            for (int r = 0; r < DIM; ++r)
                EFB_UNROLL
                for (int c = 0; c < DIM; ++c) F[r * DIM + c] += dn[r][n] * X[n * DIM + c];
        }
        const double det = det_inv<DIM>(F, Fi);
        double gn[NPE][DIM];
        EFB_UNROLL
        for (int a = 0; a < NPE; ++a) {
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) {
                double s = 0.0;
                EFB_UNROLL
                for (int k = 0; k < DIM; ++k) s += Fi[d * DIM + k] * dn[k][a];
                gn[a][d] = s;
            }
        }
        E[sm.o_wJ() + p] = scale * (fabs(det) * wt[p]);
        double* gp = E + sm.o_gN() + p * GPS;
        EFB_UNROLL
        for (int a = 0; a < NPE; ++a)
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) gp[a * GS + d] = gn[a][d];
    } else {  // high-order elements: too many reference gradients for the register file, stream them
        EFB_UNROLL
        for (int n = 0; n < NPE; ++n) {  // F[r][c] = sum_n dN[p][r][n] x[n][c]
            EFB_UNROLL
            for (int r = 0; r < DIM; ++r)
                EFB_UNROLL
                for (int c = 0; c < DIM; ++c) F[r * DIM + c] += dNp[r * NPE + n] * X[n * DIM + c];
        }
        const double det = det_inv<DIM>(F, Fi);
        E[sm.o_wJ() + p] = scale * (fabs(det) * wt[p]);
        double* gp = E + sm.o_gN() + p * GPS;
        EFB_UNROLL
        for (int a = 0; a < NPE; ++a) {  // gN[a][d] = sum_k Fi[d][k] dN[p][k][a]
            double s[DIM];
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) {
                s[d] = 0.0;
                EFB_UNROLL
                for (int k = 0; k < DIM; ++k) s[d] += Fi[d * DIM + k] * dNp[k * NPE + a];
            }
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) gp[a * GS + d] = s[d];
        }
    }
}

// which (element of the batch, node, column chunk) consumer thread `tid` owns; false if it owns nothing
template <int DIM, int NPE>
EFB_D bool elastic_owner(int tid, int nvalid, int& el, int& a, int& b0) {
    using Tile = ElasticTile<DIM, NPE>;
    const int warp = tid >> 5, lane = tid & 31;
    const int grp = warp / Tile::CS, h = warp - grp * Tile::CS;
    const int elw = lane / NPE;
    a = lane - elw * NPE;
    el = grp * Tile::EPW + elw;
    b0 = h * Tile::NB;
    return elw < Tile::EPW && el < nvalid;
}

// phase-structured body: every stage ends with a CTA barrier (this is what tests/hostcheck emulates)
template <int DIM, int NPE, int CMODE>
EFB_D void elastic_block(const GroupView& g, const CMat& C2const, const double* EFB_RESTRICT C, double scale,
                         double* EFB_RESTRICT out, long long blockId, int nthreads, double* smem, bool load_tables = true) {
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS, NDOF = DIM * NPE;
    using SM = ElasticSmem<DIM, NPE>;
    using Tile = ElasticTile<DIM, NPE>;
    constexpr int TS = SM::TS, KE = SM::KE, EPB = Tile::EPB, NB = Tile::NB;
    const int nPg = g.nPg;
    const int extra = CMODE == 2 ? nPg * NC : (CMODE == 1 ? NC : 0);
    const SM sm(nPg, EPB, extra);
    const long long e0 = blockId * EPB;
    const int nvalid = (g.Ne - e0 < EPB) ? (int)(g.Ne - e0) : EPB;
    double* dNt = smem + sm.off_dN();
    double* wt = smem + sm.off_w();
    double* E0 = sm.elem(smem, 0);

    if (load_tables) {
        EFB_PHASE(tid, nthreads) {
            for (int i = tid; i < nPg * DIM * NPE; i += nthreads) dNt[(i / (DIM * NPE)) * TS + i % (DIM * NPE)] = g.dN_pg[i];
            for (int i = tid; i < nPg; i += nthreads) wt[i] = g.w_pg[i];
        }
    }
    EFB_PHASE(tid, nthreads) {
        elastic_gather<DIM, NPE>(g, sm, e0, nvalid, E0, tid, nthreads);
        if (CMODE != 0) elastic_gather_C<DIM, NPE>(sm, C, e0, nvalid, E0, tid, nthreads);
    }
    EFB_PHASE(tid, nthreads) {
        for (int task = tid; task < nvalid * nPg; task += nthreads)
            elastic_geometry_task<DIM, NPE>(sm, dNt, wt, scale, E0 + (task / nPg) * sm.per_elem(), task % nPg);
    }
    EFB_PHASE(tid, nthreads) {
        int el, a, b0;
        if (elastic_owner<DIM, NPE>(tid, nvalid, el, a, b0)) {
            const double* E = E0 + el * sm.per_elem();
            double acc[DIM][NB * DIM];
            elastic_rows<DIM, NPE, CMODE>(C2const, E + sm.o_extra(), E + sm.o_wJ(), E + sm.o_gN(), nPg, a, b0, acc);
            double* dst = out + (e0 + el) * (long long)KE + (long long)(a * DIM) * NDOF + b0 * DIM;
            EFB_UNROLL
            for (int i = 0; i < DIM; ++i)
                EFB_UNROLL
                for (int j = 0; j < NB * DIM; ++j) dst[i * NDOF + j] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// O1, homogeneous C, warp-autonomous form (elements whose DIM rows x all columns fit one thread, CS == 1).
//
// Every warp is its own pipeline: it gathers the EPW = 32/NPE elements of its batch (node ids / coordinates of the next
// batches travel in registers while this one is computed), runs the EPW*nPg geometry tasks on its own lanes, contracts,
// stages its element matrices in warp-private shared memory and ships them with bulk asynchronous copies.  Only warp
// barriers: with 8 such warps per SM every sub-partition carries the same FP64 load and the phases of different warps
// drift apart, so gather/geometry/store latency of one warp hides under the FMA stream of its neighbour.
//
// SYM (C == C^T, checked on the host): K_e is symmetric, K[b,a] = K[a,b]^T as DIM x DIM node blocks.  Lane (el, a)
// computes the NBS = NPE/2 + 1 blocks (a, (a+t) mod NPE), t = 0..NPE/2 (a cyclic cover of the block upper triangle,
// the same number of blocks on every lane), and writes both the block and its transpose into the element tile.
// ORTHO (no normal/shear coupling and a diagonal shear block in C, e.g. isotropic or axis-aligned orthotropic): the
// structurally zero products of bc and of the block update are not issued (7 instead of 9 FMAs per block row in 3D);
// the skipped terms are exact zeros, so the result is bit-identical to the full update.
// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
struct ElasticWarp {
    static constexpr int NDOF = DIM * NPE, KE = NDOF * NDOF;
    static constexpr int EPW = 32 / NPE;                 // elements per warp batch
    static constexpr int NBS = NPE / 2 + 1;              // node blocks per lane, symmetric scheme
    static constexpr int NMIR = (NPE - 1) / 2;           // blocks t = 1..NMIR are mirrored (t = 0 and t = NPE/2 (even NPE) are not)
    static constexpr int TS = (DIM * NPE) | 1;           // odd stride of one Gauss point in the dN table
    static constexpr int GPS = (NPE * DIM) | 1;          // odd stride of one Gauss point in gN (== ElasticSmem::GPS)
    static constexpr int XW = (EPW * NPE * DIM + 1) & ~1;  // coordinates of the batch
    // strides = 8 (mod 16) doubles: the per-lane 8-byte accesses of two elements sharing a half-warp fall in
    // complementary banks (HEXA8: 3b and 3b+8 (mod 16), b = 0..7, are 16 distinct bank pairs)
    EFB_HD static constexpr int pad8(int n) { return n + ((8 - n % 16) + 16) % 16; }
    static constexpr int ES = pad8(KE);                  // element tile of the staging area (symmetric scheme)
    EFB_HD static int rec(int nPg) { return pad8(nPg + nPg * GPS); }  // per element: wJ[nPg] | gN[nPg][GPS]
    static constexpr int STAGE_SYM = EPW * ES;
    static constexpr int STAGE_LANE = 32 * ElasticTile<DIM, NPE>::LANE_STAGE;  // per-lane tiles (general scheme)
    EFB_HD static int tables(int nPg) { return (nPg * TS + nPg + 1) & ~1; }
    EFB_HD static int per_warp(int nPg, bool sym) { return (sym ? STAGE_SYM : STAGE_LANE) + XW + EPW * rec(nPg); }
    static_assert(ElasticTile<DIM, NPE>::CS == 1, "warp-autonomous kernel: one thread must hold DIM full rows");
};

// G2-G6 of one (element, Gauss point): X (NPE x DIM) -> wJ, gN[a][d]       _group_elem.py:832-1105
// The reference gradients of the Gauss point are read into registers once and serve both F and gN, and every gN value is
// formed before the first store: the stores alias the loads as far as the compiler knows, and interleaving them
// serialises the task on shared-memory latency (measured: 27 % of the kernel's samples, profiles/README.md).
template <int DIM, int NPE>
EFB_D void warp_geometry_task(const double* EFB_RESTRICT X, const double* EFB_RESTRICT dNp, double wscale,
                              double* EFB_RESTRICT wJp, double* EFB_RESTRICT gp) {
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
    double gn[NPE][DIM];
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a) {
        EFB_UNROLL
        for (int d = 0; d < DIM; ++d) {
            double s = 0.0;
            EFB_UNROLL
            for (int k = 0; k < DIM; ++k) s += Fi[d * DIM + k] * dn[k][a];
            gn[a][d] = s;
        }
    }
    *wJp = wscale * fabs(det);
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a)
        EFB_UNROLL
        for (int d = 0; d < DIM; ++d) gp[a * DIM + d] = gn[a][d];
}

// bc[i][:] = w G[:, (a,i)]^T C2 for the DIM dof rows of node a (C2 = S C S, homogeneous)
template <int DIM, bool ORTHO>
EFB_D void warp_bc(const CMat& C2, double w, const double* ga, double (&bc)[DIM][StrainSize<DIM>::value]) {
    constexpr int NS = StrainSize<DIM>::value;
#define EFB_C2(s_, r_) C2.v[(s_) * NS + (r_)]
    if constexpr (DIM == 2) {
        const double wx = w * ga[0], wy = w * ga[1];
        if constexpr (ORTHO) {
            bc[0][0] = wx * EFB_C2(0, 0); bc[0][1] = wx * EFB_C2(0, 1); bc[0][2] = wy * EFB_C2(2, 2);
            bc[1][0] = wy * EFB_C2(1, 0); bc[1][1] = wy * EFB_C2(1, 1); bc[1][2] = wx * EFB_C2(2, 2);
        } else {
            EFB_UNROLL
            for (int r = 0; r < NS; ++r) {
                bc[0][r] = wx * EFB_C2(0, r) + wy * EFB_C2(2, r);
                bc[1][r] = wy * EFB_C2(1, r) + wx * EFB_C2(2, r);
            }
        }
    } else {
        const double wx = w * ga[0], wy = w * ga[1], wz = w * ga[2];
        if constexpr (ORTHO) {
            bc[0][0] = wx * EFB_C2(0, 0); bc[0][1] = wx * EFB_C2(0, 1); bc[0][2] = wx * EFB_C2(0, 2);
            bc[0][3] = 0.0;               bc[0][4] = wz * EFB_C2(4, 4); bc[0][5] = wy * EFB_C2(5, 5);
            bc[1][0] = wy * EFB_C2(1, 0); bc[1][1] = wy * EFB_C2(1, 1); bc[1][2] = wy * EFB_C2(1, 2);
            bc[1][3] = wz * EFB_C2(3, 3); bc[1][4] = 0.0;               bc[1][5] = wx * EFB_C2(5, 5);
            bc[2][0] = wz * EFB_C2(2, 0); bc[2][1] = wz * EFB_C2(2, 1); bc[2][2] = wz * EFB_C2(2, 2);
            bc[2][3] = wy * EFB_C2(3, 3); bc[2][4] = wx * EFB_C2(4, 4); bc[2][5] = 0.0;
        } else {
            EFB_UNROLL
            for (int r = 0; r < NS; ++r) {
                bc[0][r] = wx * EFB_C2(0, r) + wz * EFB_C2(4, r) + wy * EFB_C2(5, r);
                bc[1][r] = wy * EFB_C2(1, r) + wz * EFB_C2(3, r) + wx * EFB_C2(5, r);
                bc[2][r] = wz * EFB_C2(2, r) + wy * EFB_C2(3, r) + wx * EFB_C2(4, r);
            }
        }
    }
#undef EFB_C2
}

// blk[i][j] += bc[i][:] . G[:, (b,j)] for one column node b with gradient gb
template <int DIM, bool ORTHO>
EFB_D void warp_block_update(const double (&bc)[DIM][StrainSize<DIM>::value], const double* gb, double (&blk)[DIM][DIM]) {
    if constexpr (DIM == 2) {
        const double gx = gb[0], gy = gb[1];
        EFB_UNROLL
        for (int i = 0; i < 2; ++i) {
            double s0 = blk[i][0], s1 = blk[i][1];
            s0 += bc[i][0] * gx;
            s0 += bc[i][2] * gy;
            s1 += bc[i][1] * gy;
            s1 += bc[i][2] * gx;
            blk[i][0] = s0;
            blk[i][1] = s1;
        }
    } else {
        const double gx = gb[0], gy = gb[1], gz = gb[2];
        EFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            double s0 = blk[i][0], s1 = blk[i][1], s2 = blk[i][2];
            s0 += bc[i][0] * gx;
            if (!(ORTHO && i == 1)) s0 += bc[i][4] * gz;
            if (!(ORTHO && i == 2)) s0 += bc[i][5] * gy;
            s1 += bc[i][1] * gy;
            if (!(ORTHO && i == 0)) s1 += bc[i][3] * gz;
            if (!(ORTHO && i == 2)) s1 += bc[i][5] * gx;
            s2 += bc[i][2] * gz;
            if (!(ORTHO && i == 0)) s2 += bc[i][3] * gy;
            if (!(ORTHO && i == 1)) s2 += bc[i][4] * gx;
            blk[i][0] = s0;
            blk[i][1] = s1;
            blk[i][2] = s2;
        }
    }
}

// general scheme: the DIM rows of node a against every column node (same association as elastic_rows, mode 0)
template <int DIM, int NPE, bool ORTHO>
EFB_D void warp_rows_full(const CMat& C2, const double* EFB_RESTRICT geoE, int nPg, int a, double (&acc)[NPE][DIM][DIM]) {
    constexpr int NS = StrainSize<DIM>::value, GPS = ElasticWarp<DIM, NPE>::GPS;
    EFB_UNROLL
    for (int b = 0; b < NPE; ++b)
        EFB_UNROLL
        for (int i = 0; i < DIM; ++i)
            EFB_UNROLL
            for (int j = 0; j < DIM; ++j) acc[b][i][j] = 0.0;
    for (int p = 0; p < nPg; ++p) {
        const double* gp = geoE + nPg + p * GPS;
        double bc[DIM][NS];
        warp_bc<DIM, ORTHO>(C2, geoE[p], gp + a * DIM, bc);
        EFB_UNROLL
        for (int b = 0; b < NPE; ++b) warp_block_update<DIM, ORTHO>(bc, gp + b * DIM, acc[b]);
    }
}

// symmetric scheme: blocks (a, (a+t) mod NPE), t = 0..NPE/2.  The operands are software-pipelined by hand: the own
// gradient and weight of Gauss point p+1 and the column gradient of block t+1 are loaded while block t is accumulated
// (block 0 is the diagonal block: its column gradient is the own gradient, no load).
template <int DIM, int NPE, bool ORTHO>
EFB_D void warp_rows_sym(const CMat& C2, const double* EFB_RESTRICT geoE, int nPg, int a,
                         double (&acc)[ElasticWarp<DIM, NPE>::NBS][DIM][DIM]) {
    constexpr int NS = StrainSize<DIM>::value, GPS = ElasticWarp<DIM, NPE>::GPS, NBS = ElasticWarp<DIM, NPE>::NBS;
    EFB_UNROLL
    for (int t = 0; t < NBS; ++t)
        EFB_UNROLL
        for (int i = 0; i < DIM; ++i)
            EFB_UNROLL
            for (int j = 0; j < DIM; ++j) acc[t][i][j] = 0.0;
    int boff[NBS];  // offsets of the column nodes of this lane
    EFB_UNROLL
    for (int t = 0; t < NBS; ++t) {
        int b = a + t;
        if (b >= NPE) b -= NPE;
        boff[t] = b * DIM;
    }
    const double* gp = geoE + nPg;
    double w_n = geoE[0], ga_n[DIM];
    EFB_UNROLL
    for (int d = 0; d < DIM; ++d) ga_n[d] = gp[boff[0] + d];
    for (int p = 0; p < nPg; ++p, gp += GPS) {
        const double w = w_n;
        double gb[DIM], gb_n[DIM];
        EFB_UNROLL
        for (int d = 0; d < DIM; ++d) gb[d] = ga_n[d];
        if (NBS > 1) {
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) gb_n[d] = gp[boff[NBS > 1 ? 1 : 0] + d];
        }
        if (p + 1 < nPg) {
            w_n = geoE[p + 1];
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) ga_n[d] = gp[GPS + boff[0] + d];
        }
        double bc[DIM][NS];
        warp_bc<DIM, ORTHO>(C2, w, gb, bc);
        EFB_UNROLL
        for (int t = 0; t < NBS; ++t) {
            warp_block_update<DIM, ORTHO>(bc, gb, acc[t]);
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) gb[d] = gb_n[d];
            if (t + 2 < NBS) {
                EFB_UNROLL
                for (int d = 0; d < DIM; ++d) gb_n[d] = gp[boff[t + 2] + d];
            }
        }
    }
}

// block (a, b) and its transpose (b, a) -> element tile (row-major NDOF x NDOF)
template <int DIM, int NPE>
EFB_D void warp_store_sym(const double (&acc)[ElasticWarp<DIM, NPE>::NBS][DIM][DIM], int a, double* tile) {
    using W = ElasticWarp<DIM, NPE>;
    EFB_UNROLL
    for (int t = 0; t < W::NBS; ++t) {
        int b = a + t;
        if (b >= NPE) b -= NPE;
        EFB_UNROLL
        for (int i = 0; i < DIM; ++i)
            EFB_UNROLL
            for (int j = 0; j < DIM; ++j) tile[(a * DIM + i) * W::NDOF + b * DIM + j] = acc[t][i][j];
        if (t >= 1 && t <= W::NMIR) {
            EFB_UNROLL
            for (int i = 0; i < DIM; ++i)
                EFB_UNROLL
                for (int j = 0; j < DIM; ++j) tile[(b * DIM + j) * W::NDOF + a * DIM + i] = acc[t][i][j];
        }
    }
}

// host-checkable composition of one warp batch (the device kernel runs the same stages with warp barriers and bulk copies
// in between, elem_kernels.cu): `wmem` = the warp's shared memory, layout [stage | X | records]
template <int DIM, int NPE, bool SYM, bool ORTHO>
EFB_D void elastic_warp_batch_ref(const GroupView& g, const CMat& C2, double scale, double* EFB_RESTRICT out, long long blk,
                                  const double* dNt, const double* wt, double* wmem) {
    using W = ElasticWarp<DIM, NPE>;
    const int nPg = g.nPg, rec = W::rec(nPg);
    double* stage = wmem;
    double* X = wmem + (SYM ? W::STAGE_SYM : W::STAGE_LANE);
    double* geo = X + W::XW;
    const long long e0 = blk * W::EPW;
    const int nvalid = (g.Ne - e0 < W::EPW) ? (int)(g.Ne - e0) : W::EPW;
    for (int lane = 0; lane < nvalid * NPE; ++lane) {
        const double* src = g.coord + (long long)g.connect[e0 * NPE + lane] * g.coord_stride;
        for (int d = 0; d < DIM; ++d) X[lane * DIM + d] = src[d];
    }
    for (int task = 0; task < nvalid * nPg; ++task) {
        const int el = task / nPg, p = task - el * nPg;
        double* E = geo + el * rec;
        warp_geometry_task<DIM, NPE>(X + el * NPE * DIM, dNt + p * W::TS, scale * wt[p], E + p, E + nPg + p * W::GPS);
    }
    for (int lane = 0; lane < nvalid * NPE; ++lane) {
        const int el = lane / NPE, a = lane - el * NPE;
        double* dst = out + (e0 + el) * (long long)W::KE;
        if constexpr (SYM) {
            double acc[W::NBS][DIM][DIM];
            warp_rows_sym<DIM, NPE, ORTHO>(C2, geo + el * rec, nPg, a, acc);
            warp_store_sym<DIM, NPE>(acc, a, stage + el * W::ES);
        } else {
            double acc[NPE][DIM][DIM];
            warp_rows_full<DIM, NPE, ORTHO>(C2, geo + el * rec, nPg, a, acc);
            for (int i = 0; i < DIM; ++i)
                for (int b = 0; b < NPE; ++b)
                    for (int j = 0; j < DIM; ++j) dst[(a * DIM + i) * W::NDOF + b * DIM + j] = acc[b][i][j];
        }
    }
    if constexpr (SYM)
        for (int el = 0; el < nvalid; ++el)
            for (int i = 0; i < W::KE; ++i) out[(e0 + el) * (long long)W::KE + i] = stage[el * W::ES + i];
}

// ---------------------------------------------------------------------------------------------------------
// O2: M_e = scale * sum_p coef wJ N^T N (block-diagonal N for dof_n > 1)   Bilinear.py:42-59
// O3: K_e = scale * sum_p coef wJ dN^T A dN                                 Bilinear.py:25-39, 229-249
// S4: K_e = scale * sum_p wJ (r N^T N + k dN^T A dN), F_e = scale * sum_p wJ f N   Simulations/_phasefield.py:540-573
// one body: thread (element, node b) accumulates column b of the scalar (nPe x nPe) matrix
// ---------------------------------------------------------------------------------------------------------
struct ScalarOp {
    // reaction part
    const double* r;  // coefficient array or nullptr
    int r_mode;
    double r_scalar;
    bool has_r;
    // diffusion part
    const double* A;  // (dim,dim) with A_mode leading axes, nullptr = identity
    int A_mode;
    const double* k;
    int k_mode;
    double k_scalar;
    bool has_k;
    // source part (vector output)
    const double* f;
    int f_mode;
    double f_scalar;
    bool has_f;
    int dof_n;  // block expansion of the outputs
    double scale;
    double* Ke;  // (Ne, nPe*dof_n, nPe*dof_n) or nullptr
    double* Fe;  // (Ne, nPe*dof_n, dof_n) if f_keep_axis else (Ne, nPe*dof_n)
    bool f_keep_axis;
};

template <int DIM, int NPE>
EFB_D void scalar_block(const GroupView& g, const ScalarOp& op, int EPB, long long blockId, int nthreads, double* smem) {
    const int nPg = g.nPg;
    const SmemMap<DIM, NPE> sm(nPg, EPB, NPE * NPE + NPE, op.has_k);
    const long long e0 = blockId * EPB;
    const double* Nt = smem + sm.off_N();
    geometry_phases<DIM, NPE>(g, sm, e0, NPE, nthreads, smem, op.has_k);
    EFB_PHASE(tid, nthreads) {
        const int el = tid / NPE, b = tid % NPE;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            double* E = sm.elem(smem, el);
            const double* wJ = E + sm.o_wJ();
            const double* gN = E + sm.o_gN();
            double* Ms = E + sm.o_extra();  // [a][b]
            double* Fs = Ms + NPE * NPE;
            double acc[NPE];
            EFB_UNROLL
            for (int a = 0; a < NPE; ++a) acc[a] = 0.0;
            double fb = 0.0;
            for (int p = 0; p < nPg; ++p) {
                const double w = wJ[p];
                const double* Np = Nt + p * NPE;
                if (op.has_r) {
                    const double c = coef_at(op.r, op.r_mode, op.r_scalar, e, p, nPg) * w * Np[b];
                    EFB_UNROLL
                    for (int a = 0; a < NPE; ++a) acc[a] += Np[a] * c;
                }
                if (op.has_k) {
                    const double c = coef_at(op.k, op.k_mode, op.k_scalar, e, p, nPg) * w;
                    const double* gb = gN + (p * NPE + b) * SmemMap<DIM, NPE>::GS;
                    double Ag[DIM];
                    if (op.A) {
                        const double* Ap = op.A + (op.A_mode == 0 ? 0 : (op.A_mode == 1 ? e * DIM * DIM : (e * nPg + p) * (long long)(DIM * DIM)));
                        EFB_UNROLL
                        for (int i = 0; i < DIM; ++i) {
                            double s = 0.0;
                            EFB_UNROLL
                            for (int kk = 0; kk < DIM; ++kk) s += Ap[i * DIM + kk] * gb[kk];
                            Ag[i] = c * s;
                        }
                    } else {
                        EFB_UNROLL
                        for (int i = 0; i < DIM; ++i) Ag[i] = c * gb[i];
                    }
                    EFB_UNROLL
                    for (int a = 0; a < NPE; ++a) {
                        const double* ga = gN + (p * NPE + a) * SmemMap<DIM, NPE>::GS;
                        double s = 0.0;
                        EFB_UNROLL
                        for (int i = 0; i < DIM; ++i) s += ga[i] * Ag[i];
                        acc[a] += s;
                    }
                }
                if (op.has_f) fb += coef_at(op.f, op.f_mode, op.f_scalar, e, p, nPg) * w * Np[b];
            }
            EFB_UNROLL
            for (int a = 0; a < NPE; ++a) Ms[a * NPE + b] = op.scale * acc[a];
            Fs[b] = op.scale * fb;
        }
    }
    EFB_PHASE(tid, nthreads) {  // block-expanded, coalesced write-out
        const int el = tid / NPE, t = tid % NPE;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* Ms = sm.elem(smem, el) + sm.o_extra();
            const double* Fs = Ms + NPE * NPE;
            const int dn = op.dof_n, ndof = NPE * dn;
            if (op.Ke) {
                double* dst = op.Ke + e * (long long)(ndof * ndof);
                for (int i = t; i < ndof * ndof; i += NPE) {
                    const int row = i / ndof, col = i % ndof;
                    dst[i] = (row % dn == col % dn) ? Ms[(row / dn) * NPE + col / dn] : 0.0;
                }
            }
            if (op.Fe) {
                if (op.f_keep_axis) {
                    double* dst = op.Fe + e * (long long)(ndof * dn);
                    for (int i = t; i < ndof * dn; i += NPE) {
                        const int row = i / dn, comp = i % dn;
                        dst[i] = (row % dn == comp) ? Fs[row / dn] : 0.0;
                    }
                } else {
                    double* dst = op.Fe + e * (long long)ndof;
                    for (int i = t; i < ndof; i += NPE) dst[i] = Fs[i / dn];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// P1: eps = B u_e                 Models/Elastic/_laws.py:127-157;  O4b: F_e = sum_p wJ B^T sigma   Linear.py:38-52
// P7: g = (1 - N d_e)^2 + k_res   Models/_phasefield.py:295-317
// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
EFB_D void strain_block(const GroupView& g, const int* EFB_RESTRICT connect_dof, const double* EFB_RESTRICT u,
                        double* EFB_RESTRICT eps, int EPB, long long blockId, int nthreads, double* smem) {
    constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    const int nPg = g.nPg;
    const SmemMap<DIM, NPE> sm(nPg, EPB, NDOF);
    const long long e0 = blockId * EPB;
    geometry_phases<DIM, NPE>(g, sm, e0, NDOF, nthreads, smem, true);
    EFB_PHASE(tid, nthreads) {
        const int el = tid / NDOF, t = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            double* ue = sm.elem(smem, el) + sm.o_extra();
            ue[t] = u[(long long)connect_dof[e * NPE + t / DIM] * DIM + t % DIM];
        }
    }
    EFB_PHASE(tid, nthreads) {
        const int el = tid / NDOF, t = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* E = sm.elem(smem, el);
            const double* gN = E + sm.o_gN();
            const double* ue = E + sm.o_extra();
            for (int i = t; i < nPg * NS; i += NDOF) {
                const int p = i / NS, s = i % NS;
                double acc = 0.0;
                for (int a = 0; a < NPE; ++a) {
                    const double* ga = gN + (p * NPE + a) * SmemMap<DIM, NPE>::GS;
                    EFB_UNROLL
                    for (int d = 0; d < DIM; ++d) acc += B_entry<DIM>(s, d, ga) * ue[a * DIM + d];
                }
                eps[e * (long long)(nPg * NS) + i] = acc;
            }
        }
    }
}

template <int DIM, int NPE>
EFB_D void internal_force_block(const GroupView& g, const double* EFB_RESTRICT sigma, double* EFB_RESTRICT out, int EPB,
                                long long blockId, int nthreads, double* smem) {
    constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    const int nPg = g.nPg;
    const SmemMap<DIM, NPE> sm(nPg, EPB, nPg * NS);
    const long long e0 = blockId * EPB;
    geometry_phases<DIM, NPE>(g, sm, e0, NDOF, nthreads, smem, true);
    EFB_PHASE(tid, nthreads) {
        const int el = tid / NDOF, t = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            double* sg = sm.elem(smem, el) + sm.o_extra();
            for (int i = t; i < nPg * NS; i += NDOF) sg[i] = sigma[e * (long long)(nPg * NS) + i];
        }
    }
    EFB_PHASE(tid, nthreads) {
        const int el = tid / NDOF, t = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* E = sm.elem(smem, el);
            const double* gN = E + sm.o_gN();
            const double* wJ = E + sm.o_wJ();
            const double* sg = E + sm.o_extra();
            const int a = t / DIM, d = t % DIM;
            double acc = 0.0;
            for (int p = 0; p < nPg; ++p) {
                const double* ga = gN + (p * NPE + a) * SmemMap<DIM, NPE>::GS;
                double s = 0.0;
                EFB_UNROLL
                for (int k = 0; k < NS; ++k) s += B_entry<DIM>(k, d, ga) * sg[p * NS + k];
                acc += wJ[p] * s;
            }
            out[e * NDOF + t] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Hyperelastic tangent and residual: `Operators.NonLinear.SecondPiolaKirchhoffStressTensor`
// and its shared core `__second_piola_block`, EasyFEA/FEM/Operators/NonLinear.py:37-201, with the kinematic operator
// `HyperElasticState.Compute_De` / `__Build_De`, EasyFEA/Models/HyperElastic/_state.py:320-392:
//     F = I + grad u,   B[s,(a,i)] = sum_m F[i][m] Blin[s,(a,m)]        (Green-Lagrange strain rate, Kelvin-Mandel rows)
//     K_e = sum_p wJ ( B^T d2W B  +  delta_ij  dN_a . S . dN_b ),  S = matrix of the Kelvin-Mandel vector dW (2nd Piola-Kirchhoff)
//     R_e = sum_p wJ B^T dW
// dW (Ne,nPg,ns) and d2W (Ne,nPg,ns,ns) come from the material law at the same state; outputs are in the interleaved dof
// order (x1,y1,z1,x2,...) the reference returns after its final reorder.  Thread (element, dof column).
// ---------------------------------------------------------------------------------------------------------
template <int DIM>
EFB_HD void km_to_matrix(const double* v, double (&S)[DIM][DIM]) {  // Project_vector_to_matrix, Models/_utils.py
    if constexpr (DIM == 2) {
        S[0][0] = v[0]; S[1][1] = v[1];
        S[0][1] = S[1][0] = v[2] * kInvSqrt2;
    } else {
        S[0][0] = v[0]; S[1][1] = v[1]; S[2][2] = v[2];
        S[1][2] = S[2][1] = v[3] * kInvSqrt2;
        S[0][2] = S[2][0] = v[4] * kInvSqrt2;
        S[0][1] = S[1][0] = v[5] * kInvSqrt2;
    }
}

template <int DIM, int NPE>
struct HyperSmem {  // per-element extras behind SmemMap: u_e | F[p] | Bm[p][s][col]
    static constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    EFB_HD static int extra(int nPg) { return NDOF + nPg * DIM * DIM + nPg * NS * NDOF; }
};

template <int DIM, int NPE>
EFB_D void hyper_block(const GroupView& g, const int* EFB_RESTRICT connect_dof, const double* EFB_RESTRICT u,
                       const double* EFB_RESTRICT dW, const double* EFB_RESTRICT d2W, double scale, double* EFB_RESTRICT Ke,
                       double* EFB_RESTRICT Re, int EPB, long long blockId, int nthreads, double* smem) {
    constexpr int NS = StrainSize<DIM>::value, NDOF = DIM * NPE, GS = SmemMap<DIM, NPE>::GS;
    const int nPg = g.nPg;
    const SmemMap<DIM, NPE> sm(nPg, EPB, HyperSmem<DIM, NPE>::extra(nPg));
    const long long e0 = blockId * EPB;
    geometry_phases<DIM, NPE>(g, sm, e0, NDOF, nthreads, smem, true);
    EFB_PHASE(tid, nthreads) {  // nodal displacements of the element
        const int el = tid / NDOF, t = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) (sm.elem(smem, el) + sm.o_extra())[t] = u[(long long)connect_dof[e * NPE + t / DIM] * DIM + t % DIM];
    }
    EFB_PHASE(tid, nthreads) {  // F[p][i][m] = delta_im + sum_a u_a,i dN_a/dx_m        _state.py:88-143
        const int el = tid / NDOF, t = tid % NDOF;
        if (el < EPB && e0 + el < g.Ne) {
            double* E = sm.elem(smem, el);
            const double* ue = E + sm.o_extra();
            const double* gN = E + sm.o_gN();
            double* Fm = E + sm.o_extra() + NDOF;
            for (int idx = t; idx < nPg * DIM * DIM; idx += NDOF) {
                const int p = idx / (DIM * DIM), i = (idx / DIM) % DIM, m = idx % DIM;
                double s = (i == m) ? 1.0 : 0.0;
                for (int a = 0; a < NPE; ++a) s += ue[a * DIM + i] * gN[(p * NPE + a) * GS + m];
                Fm[idx] = s;
            }
        }
    }
    EFB_PHASE(tid, nthreads) {  // Bm[p][s][(a,i)] = sum_m F[i][m] Blin[s,(a,m)]          NonLinear.py:37-72, _state.py:320-366
        const int el = tid / NDOF, col = tid % NDOF;
        if (el < EPB && e0 + el < g.Ne) {
            double* E = sm.elem(smem, el);
            const double* gN = E + sm.o_gN();
            const double* Fm = E + sm.o_extra() + NDOF;
            double* Bm = E + sm.o_extra() + NDOF + nPg * DIM * DIM;
            const int a = col / DIM, i = col % DIM;
            for (int p = 0; p < nPg; ++p) {
                const double* ga = gN + (p * NPE + a) * GS;
                for (int s = 0; s < NS; ++s) {
                    double v = 0.0;
                    EFB_UNROLL
                    for (int m = 0; m < DIM; ++m) v += Fm[(p * DIM + i) * DIM + m] * B_entry<DIM>(s, m, ga);
                    Bm[(p * NS + s) * NDOF + col] = v;
                }
            }
        }
    }
    EFB_PHASE(tid, nthreads) {  // column (b,j) of K_e and entry (b,j) of R_e
        const int el = tid / NDOF, col = tid % NDOF;
        const long long e = e0 + el;
        if (el < EPB && e < g.Ne) {
            const double* E = sm.elem(smem, el);
            const double* wJ = E + sm.o_wJ();
            const double* gN = E + sm.o_gN();
            const double* Bm = E + sm.o_extra() + NDOF + nPg * DIM * DIM;
            const int b = col / DIM, j = col % DIM;
            double acc[NDOF];
            EFB_UNROLL
            for (int r = 0; r < NDOF; ++r) acc[r] = 0.0;
            double rcol = 0.0;
            for (int p = 0; p < nPg; ++p) {
                const double w = wJ[p];
                const double* dWp = dW + (e * nPg + p) * (long long)NS;
                const double* d2Wp = d2W + (e * nPg + p) * (long long)(NS * NS);
                const double* Bp = Bm + p * NS * NDOF;
                double tmp[NS];  // w * d2W . B[:, col]
                double rs = 0.0;
                EFB_UNROLL
                for (int s = 0; s < NS; ++s) {
                    double v = 0.0;
                    EFB_UNROLL
                    for (int r = 0; r < NS; ++r) v += d2Wp[s * NS + r] * Bp[r * NDOF + col];
                    tmp[s] = w * v;
                    rs += dWp[s] * Bp[s * NDOF + col];
                }
                rcol += w * rs;
                double S[DIM][DIM], Sg[DIM];  // w * S . dN_b
                km_to_matrix<DIM>(dWp, S);
                const double* gb = gN + (p * NPE + b) * GS;
                EFB_UNROLL
                for (int k = 0; k < DIM; ++k) {
                    double v = 0.0;
                    EFB_UNROLL
                    for (int l = 0; l < DIM; ++l) v += S[k][l] * gb[l];
                    Sg[k] = w * v;
                }
                for (int row = 0; row < NDOF; ++row) {
                    double v = 0.0;
                    EFB_UNROLL
                    for (int s = 0; s < NS; ++s) v += Bp[s * NDOF + row] * tmp[s];
                    if (row % DIM == j) {  // geometric tangent: same displacement component only
                        const double* ga = gN + (p * NPE + row / DIM) * GS;
                        EFB_UNROLL
                        for (int k = 0; k < DIM; ++k) v += ga[k] * Sg[k];
                    }
                    acc[row] += v;
                }
            }
            if (Ke)
                for (int row = 0; row < NDOF; ++row) Ke[e * (long long)(NDOF * NDOF) + row * NDOF + col] = scale * acc[row];
            if (Re) Re[e * NDOF + col] = scale * rcol;
        }
    }
}

// no geometry needed: thread per (element, Gauss point)
EFB_HD double degradation_at(const int* connect_dof, const double* d, const double* N_pg, long long e, int p, int nPe,
                             double k_res) {
    double s = 0.0;
    for (int a = 0; a < nPe; ++a) s += N_pg[p * nPe + a] * d[connect_dof[e * nPe + a]];
    const double om = 1.0 - s;
    return om * om + k_res;
}

}  // namespace efb
