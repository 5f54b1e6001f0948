// Warp-autonomous form of the scalar operator body (O2 `UV`, O3 `GradUGradV / GradU_A_GradV`, O4 `V`, S4 damage system;
// Bilinear.py:25-59, 229-249, Linear.py:18-36, Simulations/_phasefield.py:540-573) for elements of at most 8 nodes.
// Same arithmetic as scalar_block (elem_kernels.cuh) — same sums in the same order — but no CTA barrier and no F / F^-1 / det
// arrays in shared memory: a warp owns a batch of 32/LPE elements (LPE = 4 or 8 lanes per element), gathers their coordinates,
// integrates the geometry of a Gauss point per lane in registers (only wJ and the physical gradients go to its private shared
// memory), contracts one column per lane, and writes the block-expanded outputs of the whole batch as one contiguous, coalesced
// range.  Only warp barriers: the phases of the 8 warps of a CTA drift apart and hide each other's latencies; per-element shared
// memory drops from 512 to 112-312 doubles (HEXA8), i.e. 16-24 resident warps per SM instead of 12 behind 5 CTA barriers.
// Written with the lane macros of csr_kernels.cuh so that tests/hostcheck runs the same body on the host.
#pragma once
#include "csr_kernels.cuh"
#include "elem_kernels.cuh"

namespace efb {

template <int DIM, int NPE>
struct ScalarWarp {
    static_assert(NPE <= 8, "warp form: at most 8 nodes per element");
    static constexpr int LPE = NPE <= 4 ? 4 : 8;  // lanes per element
    static constexpr int EPW = 32 / LPE;          // elements per warp batch
    static constexpr int GS = DIM;                // doubles per (Gauss point, node) in the gradient table
    static constexpr int GPS = (NPE * DIM) | 1;   // doubles per Gauss point in it: odd, so the lanes (one Gauss point each) hit different banks
    // doubles per Gauss point in the dN table: = 2 mod 4, so that the lanes of an element (one Gauss point each) read their rows —
    // 128-bit loads — from different 16-byte bank groups (a stride of DIM * NPE = 24 doubles was a 4-way conflict, ncu)
    static constexpr int DPS = ((DIM * NPE + 1) & ~3) + 2;
    // doubles per element in Ms: the elements of a half warp store their columns into different banks
    static constexpr int MSS = ((NPE * NPE + 15) & ~15) + LPE;
    static_assert(DPS >= DIM * NPE && DPS % 4 == 2 && MSS >= NPE * NPE, "table strides");
    EFB_HD static int tables(int nPg) { return (nPg * DPS + nPg * NPE + nPg + 1) & ~1; }
    // per-warp scratch: X | wJ | gradients (has_k) | Ms | Fs
    EFB_HD static int o_wJ() { return EPW * NPE * DIM; }
    EFB_HD static int o_G(int nPg) { return o_wJ() + EPW * nPg; }
    EFB_HD static int o_Ms(int nPg, bool grad) { return o_G(nPg) + (grad ? EPW * nPg * GPS : 0); }
    EFB_HD static int o_Fs(int nPg, bool grad) { return o_Ms(nPg, grad) + EPW * MSS; }
    EFB_HD static int per_warp(int nPg, bool grad) { return (o_Fs(nPg, grad) + EPW * NPE + 1) | 1; }  // odd: warps start in different banks
    EFB_HD static size_t total(int nPg, bool grad, int nwarps) { return (size_t)tables(nPg) + (size_t)nwarps * per_warp(nPg, grad); }
};

// reference-element tables of the group -> shared memory (once per CTA; the caller puts a CTA barrier behind it)
template <int DIM, int NPE>
EFB_D void scalar_warp_tables(const GroupView& g, double* smem, int tid, int nthreads) {
    const int nPg = g.nPg;
    using SW = ScalarWarp<DIM, NPE>;
    double* dNt = smem;
    double* Nt = dNt + nPg * SW::DPS;
    double* wt = Nt + nPg * NPE;
    for (int i = tid; i < nPg * DIM * NPE; i += nthreads) dNt[(i / (DIM * NPE)) * SW::DPS + i % (DIM * NPE)] = g.dN_pg[i];
    for (int i = tid; i < nPg * NPE; i += nthreads) Nt[i] = g.N_pg[i];
    for (int i = tid; i < nPg; i += nthreads) wt[i] = g.w_pg[i];
}

// block-expanded outputs of one batch, lane `lane` of 32; DN = dof_n known at compile time (0: read it from op)
template <int NPE, int DN, int MSS>
EFB_D void scalar_warp_writeout(const ScalarOp& op, long long e0, int nvalid, const double* EFB_RESTRICT Ms, const double* EFB_RESTRICT Fs,
                                int lane) {
    const int dn = DN ? DN : op.dof_n, ndof = NPE * dn;
    if (op.Ke) {
        double* dst = op.Ke + e0 * (long long)(ndof * ndof);
        const int per = ndof * ndof, total = nvalid * per;
        for (int i = lane; i < total; i += 32) {
            const int el = i / per, rem = i - el * per;
            const int row = rem / ndof, col = rem - row * ndof;
            dst[i] = (row % dn == col % dn) ? Ms[el * MSS + (row / dn) * NPE + col / dn] : 0.0;
        }
    }
    if (op.Fe) {
        if (op.f_keep_axis) {
            double* dst = op.Fe + e0 * (long long)(ndof * dn);
            const int per = ndof * dn, total = nvalid * per;
            for (int i = lane; i < total; i += 32) {
                const int el = i / per, rem = i - el * per;
                const int row = rem / dn, comp = rem - row * dn;
                dst[i] = (row % dn == comp) ? Fs[el * NPE + row / dn] : 0.0;
            }
        } else {
            double* dst = op.Fe + e0 * (long long)ndof;
            const int total = nvalid * ndof;
            for (int i = lane; i < total; i += 32) {
                const int el = i / ndof, rem = i - el * ndof;
                dst[i] = Fs[el * NPE + rem / dn];
            }
        }
    }
}

// one batch (elements batch*EPW ...) on one warp; `tab` = the tables, `ws` = this warp's scratch
template <int DIM, int NPE>
EFB_D void scalar_warp_batch(const GroupView& g, const ScalarOp& op, long long batch, const double* tab, double* ws) {
    using SW = ScalarWarp<DIM, NPE>;
    constexpr int LPE = SW::LPE, EPW = SW::EPW, GS = SW::GS, GPS = SW::GPS, DPS = SW::DPS, MSS = SW::MSS;
    const int nPg = g.nPg;
    const double* dNt = tab;
    const double* Nt = dNt + nPg * DPS;
    const double* wt = Nt + nPg * NPE;
    double* X = ws;
    double* wJ = ws + SW::o_wJ();
    double* G = ws + SW::o_G(nPg);
    double* Ms = ws + SW::o_Ms(nPg, op.has_k);
    double* Fs = ws + SW::o_Fs(nPg, op.has_k);
    const long long e0 = batch * EPW;
    const int nvalid = (g.Ne - e0 < EPW) ? (int)(g.Ne - e0) : EPW;

    EFB_LANES(lane) {  // gather: lane (el, t) fetches node t of element el
        const int el = lane / LPE, t = lane % LPE;
        if (el < nvalid && t < NPE) {
            const double* src = g.coord + (long long)g.connect[(e0 + el) * NPE + t] * g.coord_stride;
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) X[(el * NPE + t) * DIM + d] = src[d];
        }
    }
    EFB_LANES(lane) {  // geometry of the Gauss points t, t + LPE, ... of element el, in registers   _group_elem.py:864-915, 1083-1105
        const int el = lane / LPE, t = lane % LPE;
        if (el < nvalid) {
            const double* Xe = X + el * NPE * DIM;
            for (int p = t; p < nPg; p += LPE) {
                double F[DIM * DIM], Fi[DIM * DIM];
                EFB_UNROLL
                for (int r = 0; r < DIM; ++r)
                    EFB_UNROLL
                    for (int c = 0; c < DIM; ++c) {
                        const double* row = dNt + p * DPS + r * NPE;
                        double s = 0.0;
                        EFB_UNROLL
                        for (int n = 0; n < NPE; ++n) s += row[n] * Xe[n * DIM + c];
                        F[r * DIM + c] = s;
                    }
                const double det = det_inv<DIM>(F, Fi);
                wJ[el * nPg + p] = fabs(det) * wt[p];
                if (op.has_k) {
                    double* gp = G + (el * nPg + p) * GPS;
                    EFB_UNROLL
                    for (int a = 0; a < NPE; ++a)
                        EFB_UNROLL
                        for (int d = 0; d < DIM; ++d) {
                            double s = 0.0;
                            EFB_UNROLL
                            for (int k = 0; k < DIM; ++k) s += Fi[d * DIM + k] * dNt[p * DPS + k * NPE + a];
                            gp[a * GS + d] = s;
                        }
                }
            }
        }
    }
    EFB_LANES(lane) {  // contraction: lane (el, b) accumulates column b (the sums of scalar_block, in its order)
        const int el = lane / LPE, b = lane % LPE;
        if (el < nvalid && b < NPE) {
            const long long e = e0 + el;
            double acc[NPE];
            EFB_UNROLL
            for (int a = 0; a < NPE; ++a) acc[a] = 0.0;
            double fb = 0.0;
            for (int p = 0; p < nPg; ++p) {
                const double w = wJ[el * nPg + p];
                const double* Np = Nt + p * NPE;
                if (op.has_r) {
                    const double c = coef_at(op.r, op.r_mode, op.r_scalar, e, p, nPg) * w * Np[b];
                    EFB_UNROLL
                    for (int a = 0; a < NPE; ++a) acc[a] += Np[a] * c;
                }
                if (op.has_k) {
                    const double c = coef_at(op.k, op.k_mode, op.k_scalar, e, p, nPg) * w;
                    const double* gp = G + (el * nPg + p) * GPS;
                    const double* gb = gp + b * GS;
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
                        const double* ga = gp + a * GS;
                        double s = 0.0;
                        EFB_UNROLL
                        for (int i = 0; i < DIM; ++i) s += ga[i] * Ag[i];
                        acc[a] += s;
                    }
                }
                if (op.has_f) fb += coef_at(op.f, op.f_mode, op.f_scalar, e, p, nPg) * w * Np[b];
            }
            EFB_UNROLL
            for (int a = 0; a < NPE; ++a) Ms[el * MSS + a * NPE + b] = op.scale * acc[a];
            Fs[el * NPE + b] = op.scale * fb;
        }
    }
    // block-expanded write-out: the batch's outputs are one contiguous range (dof_n is 1, 2 or 3 in the reference: compile-time
    // divisors; any other value takes the generic loop)
    EFB_LANES(lane) {
        if (op.dof_n == 1) scalar_warp_writeout<NPE, 1, MSS>(op, e0, nvalid, Ms, Fs, lane);
        else if (op.dof_n == 2) scalar_warp_writeout<NPE, 2, MSS>(op, e0, nvalid, Ms, Fs, lane);
        else if (op.dof_n == 3) scalar_warp_writeout<NPE, 3, MSS>(op, e0, nvalid, Ms, Fs, lane);
        else scalar_warp_writeout<NPE, 0, MSS>(op, e0, nvalid, Ms, Fs, lane);
    }
}

}  // namespace efb
