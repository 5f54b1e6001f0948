// Per-Gauss-point phase-field law (rows P2-P6 of SURVEY.md §8a): closed-form 2x2 / 3x3 eigen-decompositions in
// registers, spectral projectors, and the Amor / Miehe / Stress / He / Bourdin splits.
// Mirrors EasyFEA/Models/_phasefield.py:396-1243 expression by expression (same definitions: projM = I - projP,
// M2 = I - M1 - M3, Heaviside(0) = 1/2, per-ELEMENT case flags in 3D) so that values agree to rounding.
#pragma once
#include "frame.cuh"

namespace efb {

struct PfMat {
    int dim, split, planeStress, pad;
    double E, v, lambda, mu, bulk;
    double C[36], sqrtC[36], inv_sqrtC[36];
};

enum { SPLIT_BOURDIN = 0, SPLIT_AMOR = 1, SPLIT_MIEHE = 2, SPLIT_STRESS = 3, SPLIT_HE = 4 };
// element-level case bits of the 3D eigen-solver (Models/_phasefield.py:861-906)
enum { CASE_T1 = 1, CASE_T2 = 2, CASE_T3 = 4 };

EFB_HD double heaviside_half(double x) { return x < 0.0 ? 0.0 : (x == 0.0 ? 0.5 : (x > 0.0 ? 1.0 : x)); }
EFB_HD double sign_np(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }  // np.sign (0 -> 0, NaN -> NaN)
EFB_HD bool finite_d(double x) { return (x - x) == 0.0; }

template <int N>
EFB_HD void matvec(const double* A, const double* x, double* y) {
    EFB_UNROLL
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        EFB_UNROLL
        for (int j = 0; j < N; ++j) s += A[i * N + j] * x[j];
        y[i] = s;
    }
}

template <int N>
EFB_HD void matmul(const double* A, const double* B, double* Cm) {  // Cm = A B (no aliasing)
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            EFB_UNROLL
            for (int k = 0; k < N; ++k) s += A[i * N + k] * B[k * N + j];
            Cm[i * N + j] = s;
        }
}

template <int N>
EFB_HD void matmul_tn(const double* A, const double* B, double* Cm) {  // Cm = A^T B
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            EFB_UNROLL
            for (int k = 0; k < N; ++k) s += A[k * N + i] * B[k * N + j];
            Cm[i * N + j] = s;
        }
}

// ---------------------------------------------------------------------------------------------------------
// 2D: `_Eigen_values_vectors_projectors` :777-807 and `__Spectral_Decomposition` :1084-1120
// ---------------------------------------------------------------------------------------------------------
EFB_HD void spectral_projector_2d(const double* v, double* P /* 3x3 */) {
    const double m00 = v[0], m11 = v[1], m01 = v[2] / kSqrt2;
    const double det = (m00 * m11) - (m01 * m01);
    const double tr = m00 + m11;
    double delta = tr * tr - (4.0 * det);
    if (delta < 0.0) delta = 0.0;  // repair: the reference takes sqrt of the rounded-negative value (NaN)
    const double root = sqrt(delta);
    const double e0 = (tr - root) / 2.0, e1 = (tr + root) / 2.0;
    double M1[3] = {1.0, 0.0, 0.0};  // (00, 11, 01)
    if (e0 != e1) {
        const double den = e0 - e1;
        M1[0] = (m00 - e1) / den;
        M1[1] = (m11 - e1) / den;
        M1[2] = m01 / den;
    }
    const double M2[3] = {1.0 - M1[0], 1.0 - M1[1], 0.0 - M1[2]};
    const double m1[3] = {M1[0], M1[1], M1[2] * kSqrt2};
    const double m2[3] = {M2[0], M2[1], M2[2] * kSqrt2};
    const double p0 = (e0 + fabs(e0)) / 2.0, p1 = (e1 + fabs(e1)) / 2.0;
    double den = e0 - e1;
    if (den == 0.0) den = 1.0;
    const double beta = (p0 - p1) / den;
    const double g0 = heaviside_half(e0) - beta, g1 = heaviside_half(e1) - beta;
    EFB_UNROLL
    for (int i = 0; i < 3; ++i)
        EFB_UNROLL
        for (int j = 0; j < 3; ++j)
            P[i * 3 + j] = ((i == j ? beta : 0.0) + g0 * (m1[i] * m1[j])) + g1 * (m2[i] * m2[j]);
}

// ---------------------------------------------------------------------------------------------------------
// 3D: invariants + Lode angle (:809-833), the four cases (:836-934), normalisation (:944-948)
// ---------------------------------------------------------------------------------------------------------
struct Inv3 {
    double M[9];
    double I1, g, sg, arg;
    bool gnz;
};

EFB_HD void invariants_3d(const double* v, Inv3& q) {
    double* M = q.M;
    M[0] = v[0];
    M[4] = v[1];
    M[8] = v[2];
    M[5] = M[7] = v[3] / kSqrt2;
    M[2] = M[6] = v[4] / kSqrt2;
    M[1] = M[3] = v[5] / kSqrt2;
    const double I1 = M[0] + M[4] + M[8];
    // trace(M M)
    const double trMM = (M[0] * M[0] + M[1] * M[3] + M[2] * M[6]) + (M[3] * M[1] + M[4] * M[4] + M[5] * M[7]) +
                        (M[6] * M[2] + M[7] * M[5] + M[8] * M[8]);
    const double I2 = 0.5 * (I1 * I1 - trMM);
    const double I3 = M[0] * ((M[4] * M[8]) - (M[7] * M[5])) - M[1] * ((M[3] * M[8]) - (M[6] * M[5])) +
                      M[2] * ((M[3] * M[7]) - (M[6] * M[4]));
    const double g = I1 * I1 - 3.0 * I2;
    q.I1 = I1;
    q.g = g;
    q.sg = sqrt(g);
    q.gnz = (g != 0.0);
    double arg = 0.5 * (2.0 * (I1 * I1 * I1) - 9.0 * I1 * I2 + 27.0 * I3);
    if (q.gnz) arg = arg / (g * q.sg);
    q.arg = arg;
}

// case bits this Gauss point raises, exactly as the reference tests them (NaN theta -> "distinct")
EFB_HD int case_bits_3d(const Inv3& q) {
    if (!q.gnz) return 0;
    const double theta = (1.0 / 3.0) * acos(q.arg);
    const double pi3 = 1.0471975511965976;  // np.pi / 3
    if (theta == pi3) return CASE_T2;
    if (theta == 0.0) return CASE_T3;
    return CASE_T1;
}

// element-level selection from the OR of the points' bits: case2 / case3 apply (cumulatively) to every point of
// an element in which ANY point meets them; case1 = some point distinct, and neither case2 nor case3
EFB_HD void eigen_3d_cases(const Inv3& q, double arg, bool c2, bool c3, bool c1, double* vals, double* M1, double* M3) {
    const double* M = q.M;
    const double sg = q.sg;
    double v1 = q.I1 / 3.0, v2 = v1, v3 = v1;
    EFB_UNROLL
    for (int i = 0; i < 9; ++i) M1[i] = M3[i] = 0.0;
    M1[0] = 1.0;
    M3[8] = 1.0;
    const double irg = (1.0 / 3.0) * (q.I1 - sg);
    const double gm12 = 1.0 / sg;
    if (c2) {  // two maximum eigenvalues :861-874
        v1 += -2.0 / 3.0 * sg;
        v2 += 1.0 / 3.0 * sg;
        v3 += 1.0 / 3.0 * sg;
        EFB_UNROLL
        for (int i = 0; i < 9; ++i) {
            const double id = (i % 4 == 0) ? 1.0 : 0.0;
            M1[i] = gm12 * (irg * id - M[i]);
            M3[i] = 0.5 * (id - M1[i]);
        }
    }
    if (c3) {  // two minimum eigenvalues :882-895
        v1 += -1.0 / 3.0 * sg;
        v2 += -1.0 / 3.0 * sg;
        v3 += 2.0 / 3.0 * sg;
        EFB_UNROLL
        for (int i = 0; i < 9; ++i) {
            const double id = (i % 4 == 0) ? 1.0 : 0.0;
            M3[i] = gm12 * (M[i] - irg * id);
            M1[i] = 0.5 * (id - M3[i]);
        }
    }
    if (c1) {  // three distinct eigenvalues :902-934
        const double theta = (1.0 / 3.0) * acos(arg);
        const double tpi3 = 2.0943951023931953;  // 2*np.pi/3
        v1 += 2.0 / 3.0 * (sg * cos(tpi3 + theta));
        v2 += 2.0 / 3.0 * (sg * cos(tpi3 - theta));
        v3 += 2.0 / 3.0 * (sg * cos(theta));
        double A[9], B[9], Cm[9];
        EFB_UNROLL
        for (int i = 0; i < 9; ++i) {
            const double id = (i % 4 == 0) ? 1.0 : 0.0;
            A[i] = M[i] - v2 * id;
            B[i] = M[i] - v3 * id;
            Cm[i] = M[i] - v1 * id;
        }
        const double d1 = (v1 - v2) * (v1 - v3), d3 = (v3 - v1) * (v3 - v2);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s1 = 0.0, s3 = 0.0;
                EFB_UNROLL
                for (int k = 0; k < 3; ++k) {
                    s1 += A[r * 3 + k] * B[k * 3 + c];   // (M - v2 I)(M - v3 I)
                    s3 += Cm[r * 3 + k] * A[k * 3 + c];  // (M - v1 I)(M - v2 I)
                }
                M1[r * 3 + c] = s1 / d1;
                M3[r * 3 + c] = s3 / d3;
            }
    }
    double n1 = 0.0, n3 = 0.0;
    EFB_UNROLL
    for (int i = 0; i < 9; ++i) {
        n1 += M1[i] * M1[i];
        n3 += M3[i] * M3[i];
    }
    n1 = sqrt(n1);
    n3 = sqrt(n3);
    EFB_UNROLL
    for (int i = 0; i < 9; ++i) {
        M1[i] = M1[i] / n1;
        M3[i] = M3[i] / n3;
    }
    vals[0] = v1;
    vals[1] = v2;
    vals[2] = v3;
}

EFB_HD void spectral_projector_3d(const double* v, int elem_bits, double* P /* 6x6 */) {
    Inv3 q;
    invariants_3d(v, q);
    const bool c2 = (elem_bits & CASE_T2) != 0, c3 = (elem_bits & CASE_T3) != 0;
    const bool c1 = (elem_bits & CASE_T1) != 0 && !(c2 || c3);
    double vals[3], M1[9], M3[9];
    eigen_3d_cases(q, q.arg, c2, c3, c1, vals, M1, M3);
    bool ok = finite_d(vals[0]) && finite_d(vals[1]) && finite_d(vals[2]);
    EFB_UNROLL
    for (int i = 0; i < 9; ++i) ok = ok && finite_d(M1[i]) && finite_d(M3[i]);
    if (!ok) {
        // repair policy (DESIGN.md "degenerate states"): where the reference formulas are non-finite at this point,
        // recompute it alone with the Lode argument clamped to [-1,1] and the case chosen per point
        // (g <= 0 or NaN counts as the triple-eigenvalue case)
        const bool gpos = q.g > 0.0;
        double a = gpos ? q.arg : 0.0;
        if (!(a == a)) a = 0.0;
        a = a > 1.0 ? 1.0 : (a < -1.0 ? -1.0 : a);
        Inv3 qq = q;
        qq.arg = a;
        qq.gnz = gpos;
        const int bits = case_bits_3d(qq);
        eigen_3d_cases(q, a, (bits & CASE_T2) != 0, (bits & CASE_T3) != 0, (bits & CASE_T1) != 0, vals, M1, M3);
    }
    double M2[9];
    EFB_UNROLL
    for (int i = 0; i < 9; ++i) M2[i] = ((i % 4 == 0) ? 1.0 : 0.0) - (M1[i] + M3[i]);
    const double* Ms[3] = {M1, M2, M3};
    // Kelvin-Mandel vectors m_a (:950-963 via Project_matrix_to_vector, Models/_utils.py:191-222)
    double m[3][6];
    EFB_UNROLL
    for (int a = 0; a < 3; ++a) {
        m[a][0] = Ms[a][0];
        m[a][1] = Ms[a][4];
        m[a][2] = Ms[a][8];
        m[a][3] = Ms[a][5] * kSqrt2;
        m[a][4] = Ms[a][2] * kSqrt2;
        m[a][5] = Ms[a][1] * kSqrt2;
    }
    double valp[3], H[3];
    EFB_UNROLL
    for (int a = 0; a < 3; ++a) {
        valp[a] = (vals[a] + fabs(vals[a])) / 2.0;
        H[a] = heaviside_half(vals[a]);
    }
    const int pa[3] = {0, 0, 1}, pb[3] = {1, 2, 2};
    double th[3];
    EFB_UNROLL
    for (int k = 0; k < 3; ++k) {
        double den = vals[pa[k]] - vals[pb[k]];
        if (den == 0.0) den = 1.0;
        th[k] = (valp[pa[k]] - valp[pb[k]]) / (2.0 * den);
    }
    const int rI[6] = {0, 1, 2, 1, 0, 0}, rJ[6] = {0, 1, 2, 2, 2, 1};
    for (int I = 0; I < 6; ++I)
        for (int J = 0; J < 6; ++J) {
            double s = 0.0;
            EFB_UNROLL
            for (int a = 0; a < 3; ++a) s += (m[a][I] * m[a][J]) * H[a];
            const int i = rI[I], j = rJ[I], k = rI[J], l = rJ[J];
            const double scale = (I < 3 && J < 3) ? 1.0 : ((I >= 3 && J >= 3) ? 2.0 : kSqrt2);  // _km_scale :1165-1168
            double gs = 0.0;
            EFB_UNROLL
            for (int pr = 0; pr < 3; ++pr) {
                const double* A = Ms[pa[pr]];
                const double* B = Ms[pb[pr]];
                const double G = (A[i * 3 + k] * B[j * 3 + l] + A[i * 3 + l] * B[j * 3 + k] + B[i * 3 + k] * A[j * 3 + l] +
                                  B[i * 3 + l] * A[j * 3 + k]) * scale;
                gs += G * th[pr];
            }
            P[I * 6 + J] = s + gs;
        }
}

// ---------------------------------------------------------------------------------------------------------
// the vector whose spectrum a split decomposes: eps (Miehe), sigma = C eps (Stress), sqrtC eps (He)
// ---------------------------------------------------------------------------------------------------------
template <int NS>
EFB_HD void decomposed_vector(const PfMat& m, const double* eps, double* out) {
    if (m.split == SPLIT_STRESS)
        matvec<NS>(m.C, eps, out);
    else if (m.split == SPLIT_HE)
        matvec<NS>(m.sqrtC, eps, out);
    else {
        EFB_UNROLL
        for (int i = 0; i < NS; ++i) out[i] = eps[i];
    }
}

EFB_HD bool split_is_spectral(int split) { return split == SPLIT_MIEHE || split == SPLIT_STRESS || split == SPLIT_HE; }

// cP, cM (NS x NS) at one Gauss point; `Calc_C` Models/_phasefield.py:396-431
template <int DIM>
EFB_HD void split_point(const PfMat& m, const double* eps, int elem_bits, double* cP, double* cM) {
    constexpr int NS = StrainSize<DIM>::value;
    constexpr int NC = NS * NS;
    if (m.split == SPLIT_BOURDIN) {  // :433-449
        EFB_UNROLL
        for (int i = 0; i < NC; ++i) {
            cP[i] = m.C[i];
            cM[i] = 0.0;
        }
        return;
    }
    double vec[NS];
    decomposed_vector<NS>(m, eps, vec);
    double tr = vec[0] + vec[1];
    if constexpr (DIM == 3) tr += vec[2];
    const double Rp = (1.0 + sign_np(tr)) / 2.0, Rm = (1.0 + sign_np(-tr)) / 2.0;  // __Rp_Rm :485-501
#define EFB_IXI(i, j) (((i) < DIM && (j) < DIM) ? 1.0 : 0.0)
    if (m.split == SPLIT_AMOR) {  // :451-483
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j < NS; ++j) {
                const double ixi = EFB_IXI(i, j);
                cP[i * NS + j] = m.bulk * (Rp * ixi) + (2.0 * m.mu) * ((i == j ? 1.0 : 0.0) - (1.0 / DIM) * ixi);
                cM[i * NS + j] = m.bulk * (Rm * ixi);
            }
        return;
    }
    double P[NC];
    if constexpr (DIM == 2)
        spectral_projector_2d(vec, P);
    else
        spectral_projector_3d(vec, elem_bits, P);
    if (m.split == SPLIT_MIEHE) {  // :515-537
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j < NS; ++j) {
                const double ixi = EFB_IXI(i, j), id = (i == j ? 1.0 : 0.0);
                cP[i * NS + j] = m.lambda * (Rp * ixi) + (2.0 * m.mu) * P[i * NS + j];
                cM[i * NS + j] = m.lambda * (Rm * ixi) + (2.0 * m.mu) * (id - P[i * NS + j]);
            }
        return;
    }
    if (m.split == SPLIT_STRESS) {  // :573-632
        double a, b;
        if (DIM == 2) {
            a = (1.0 + m.v) / m.E;
            b = m.planeStress ? m.v / m.E : m.v * (1.0 + m.v) / m.E;
        } else {
            a = 1.0 / (2.0 * m.mu);
            b = m.v / m.E;
        }
        double sP[NC], sM[NC], T[NC];
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j < NS; ++j) {
                const double ixi = EFB_IXI(i, j), id = (i == j ? 1.0 : 0.0);
                sP[i * NS + j] = (a * P[i * NS + j]) - (b * Rp * ixi);
                sM[i * NS + j] = (a * (id - P[i * NS + j])) - (b * Rm * ixi);
            }
        matmul_tn<NS>(m.C, sP, T);  // C^T sP
        matmul<NS>(T, m.C, cP);
        matmul_tn<NS>(m.C, sM, T);
        matmul<NS>(T, m.C, cM);
        return;
    }
    // He :680-749 — proj = sqrtS projTilde sqrtC, c = C proj
    double T[NC], Q[NC];
    matmul<NS>(m.inv_sqrtC, P, T);
    matmul<NS>(T, m.sqrtC, Q);
    matmul<NS>(m.C, Q, cP);
    EFB_UNROLL
    for (int i = 0; i < NC; ++i) P[i] = ((i % (NS + 1) == 0) ? 1.0 : 0.0) - P[i];
    matmul<NS>(m.inv_sqrtC, P, T);
    matmul<NS>(T, m.sqrtC, Q);
    matmul<NS>(m.C, Q, cM);
#undef EFB_IXI
}

// psi = 1/2 eps . (c eps)        Models/_phasefield.py:335-394
template <int NS>
EFB_HD double energy_density(const double* c, const double* eps) {
    double s = 0.0;
    for (int i = 0; i < NS; ++i) {
        double si = 0.0;
        EFB_UNROLL
        for (int j = 0; j < NS; ++j) si += c[i * NS + j] * eps[j];
        s += 0.5 * eps[i] * si;
    }
    return s;
}

}  // namespace efb
