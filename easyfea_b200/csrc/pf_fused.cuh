// Fused S3 of the phase-field displacement sub-problem for one-point elements (TRI3, TETRA4: BASELINE configs 3 and 4):
// `PhaseField.__Construct_Elastic_Matrix`, EasyFEA/Simulations/_phasefield.py:444-482, in ONE pass per element —
// strain (P1, Models/Elastic/_laws.py:127-157) -> split (P2-P5, Models/_phasefield.py:396-749) -> C = g(d) cP + cM
// (:462-469, g = Get_g_e_pg :295-317) -> K_e = thickness * wJ B^T C B (Operators/Bilinear.py:62-79).  The strain, g and
// C(d) arrays of the four-kernel composition never exist.  Body shared with the host emulation of tests/hostcheck.
#pragma once
#include "elem_kernels.cuh"
#include "pf_math.cuh"

namespace efb {

template <int DIM>
EFB_HD void pf_elastic_simplex_item(const PfMat& m, const GroupView& g, const int* EFB_RESTRICT connect_dof,
                                    const double* EFB_RESTRICT u, const double* EFB_RESTRICT d, double k_res, double scale, long long e,
                                    double* EFB_RESTRICT Ke) {
    constexpr int NPE = DIM + 1, NS = StrainSize<DIM>::value, NDOF = DIM * NPE;
    double X[NPE][DIM], ue[NPE][DIM], de[NPE];
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a) {
        const double* src = g.coord + (long long)g.connect[e * NPE + a] * g.coord_stride;
        const long long n = connect_dof[e * NPE + a];
        EFB_UNROLL
        for (int c = 0; c < DIM; ++c) {
            X[a][c] = src[c];
            ue[a][c] = u[n * DIM + c];
        }
        de[a] = d[n];
    }
    // G2-G6 at the single Gauss point: F = dN X, det, inverse, gN_a = F^-1 dN_a          _group_elem.py:832-1105
    double F[DIM * DIM], Fi[DIM * DIM];
    EFB_UNROLL
    for (int r = 0; r < DIM; ++r)
        EFB_UNROLL
        for (int c = 0; c < DIM; ++c) {
            double s = 0.0;
            EFB_UNROLL
            for (int a = 0; a < NPE; ++a) s += g.dN_pg[r * NPE + a] * X[a][c];
            F[r * DIM + c] = s;
        }
    const double det = det_inv<DIM>(F, Fi);
    const double wJ = fabs(det) * g.w_pg[0];
    double gN[NPE][DIM];
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a)
        EFB_UNROLL
        for (int c = 0; c < DIM; ++c) {
            double s = 0.0;
            EFB_UNROLL
            for (int k = 0; k < DIM; ++k) s += Fi[c * DIM + k] * g.dN_pg[k * NPE + a];
            gN[a][c] = s;
        }
    // B (ns x ndof), eps = B u_e, g = (1 - N d_e)^2 + k_res
    double B[NS][NDOF];
    EFB_UNROLL
    for (int s = 0; s < NS; ++s)
        EFB_UNROLL
        for (int col = 0; col < NDOF; ++col) B[s][col] = B_entry<DIM>(s, col % DIM, gN[col / DIM]);
    double eps[NS];
    EFB_UNROLL
    for (int s = 0; s < NS; ++s) {
        double acc = 0.0;
        EFB_UNROLL
        for (int a = 0; a < NPE; ++a)
            EFB_UNROLL
            for (int c = 0; c < DIM; ++c) acc += B_entry<DIM>(s, c, gN[a]) * ue[a][c];
        eps[s] = acc;
    }
    double nd = 0.0;
    EFB_UNROLL
    for (int a = 0; a < NPE; ++a) nd += g.N_pg[a] * de[a];
    const double om = 1.0 - nd, gd = om * om + k_res;
    // split: with one Gauss point the per-ELEMENT case flags of the 3D eigen-solver are the point's own
    int bits = 0;
    if (DIM == 3 && split_is_spectral(m.split)) {
        double v[NS];
        decomposed_vector<NS>(m, eps, v);
        Inv3 q;
        invariants_3d(v, q);
        bits = case_bits_3d(q);
    }
    double cP[NS * NS], cM[NS * NS];
    split_point<DIM>(m, eps, bits, cP, cM);
    // K_e = scale * wJ B^T (g cP + cM) B
    double* dst = Ke + e * (long long)(NDOF * NDOF);
    const double w = scale * wJ;
    double CB[NS][NDOF];
    EFB_UNROLL
    for (int s = 0; s < NS; ++s)
        EFB_UNROLL
        for (int col = 0; col < NDOF; ++col) {
            double acc = 0.0;
            EFB_UNROLL
            for (int r = 0; r < NS; ++r) acc += (gd * cP[s * NS + r] + cM[s * NS + r]) * B[r][col];
            CB[s][col] = w * acc;
        }
    for (int row = 0; row < NDOF; ++row)
        EFB_UNROLL
        for (int col = 0; col < NDOF; ++col) {
            double acc = 0.0;
            EFB_UNROLL
            for (int s = 0; s < NS; ++s) acc += B[s][row] * CB[s][col];
            dst[row * NDOF + col] = acc;
        }
}

}  // namespace efb
