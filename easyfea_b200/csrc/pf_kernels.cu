// C-ABI entry points of the phase-field law kernels (P2-P7): one thread per Gauss point, everything in registers.
#include <string.h>

#include "common.cuh"
#include "pf_fused.cuh"

namespace efb {

static_assert(sizeof(PfMat) == sizeof(efb_pf_material), "efb_pf_material layout");

template <int DIM>
__global__ void k_pf_case_bits(PfMat m, const double* __restrict__ eps, long long n, int nPg, int* __restrict__ elem_bits) {
    constexpr int NS = StrainSize<DIM>::value;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[NS], v[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) e[k] = eps[i * NS + k];
    decomposed_vector<NS>(m, e, v);
    Inv3 q;
    invariants_3d(v, q);
    const int bits = case_bits_3d(q);
    if (bits) atomicOr(elem_bits + i / nPg, bits);  // OR is order-independent: deterministic
}

template <int DIM>
__global__ void __launch_bounds__(128)
    k_pf_split(PfMat m, const double* __restrict__ eps, long long n, int nPg, const int* __restrict__ elem_bits,
               double* __restrict__ cP_out, double* __restrict__ cM_out, double* __restrict__ psiP, double* __restrict__ psiM,
               const double* __restrict__ g, double* __restrict__ Cdeg) {
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[NS], cP[NC], cM[NC];
#pragma unroll
    for (int k = 0; k < NS; ++k) e[k] = eps[i * NS + k];
    const int bits = elem_bits ? elem_bits[i / nPg] : 0;
    split_point<DIM>(m, e, bits, cP, cM);
    if (cP_out)
        for (int k = 0; k < NC; ++k) cP_out[i * NC + k] = cP[k];
    if (cM_out)
        for (int k = 0; k < NC; ++k) cM_out[i * NC + k] = cM[k];
    if (psiP) psiP[i] = energy_density<NS>(cP, e);
    if (psiM) psiM[i] = energy_density<NS>(cM, e);
    if (Cdeg) {
        const double gi = g[i];
        for (int k = 0; k < NC; ++k) Cdeg[i * NC + k] = gi * cP[k] + cM[k];  // Simulations/_phasefield.py:462-469
    }
}

template <int DIM>
__global__ void __launch_bounds__(128) k_pf_elastic_simplex(PfMat m, GroupView g, const int* __restrict__ connect_dof,
                                                            const double* __restrict__ u, const double* __restrict__ d, double k_res,
                                                            double scale, double* __restrict__ Ke) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < g.Ne) pf_elastic_simplex_item<DIM>(m, g, connect_dof, u, d, k_res, scale, e, Ke);
}

__global__ void k_pf_history_rf(double* __restrict__ psiP, const double* __restrict__ psiP_old, long long n, int regu, double Gc,
                                double l0, double* __restrict__ r, double* __restrict__ f) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p = psiP[i];
    if (psiP_old) {
        const double old = psiP_old[i];
        if (p - old < 0.0) p = old;  // inc_H < 0 -> keep the history value, Simulations/_phasefield.py:526-530
        psiP[i] = p;
    }
    if (regu == EFB_REGU_AT1) {  // Models/_phasefield.py:253-293
        if (r) r[i] = 2.0 * p;
        if (f) {
            const double fv = 2.0 * p - ((3.0 * Gc) / (8.0 * l0));
            f[i] = (fv + fabs(fv)) / 2.0;
        }
    } else {
        if (r) r[i] = 2.0 * p + (Gc / l0);
        if (f) f[i] = 2.0 * p;
    }
}

}  // namespace efb

using namespace efb;

extern "C" int efb_pf_split(const efb_pf_material* m, const double* eps, int64_t Ne, int32_t nPg, int32_t* elem_bits,
                            double* cP, double* cM, double* psiP, double* psiM, const double* g_e_pg, double* Cdeg,
                            void* stream) {
    if (!m || !eps || (m->dim != 2 && m->dim != 3) || m->split < 0 || m->split > EFB_SPLIT_HE || (Cdeg && !g_e_pg)) {
        set_error("efb_pf_split: bad arguments");
        return 1;
    }
    const long long n = (long long)Ne * nPg;
    if (n == 0) return 0;
    PfMat pm;
    memcpy(&pm, m, sizeof(pm));
    cudaStream_t st = as_stream(stream);
    const unsigned nblk = (unsigned)((n + 127) / 128);
    if (m->dim == 2) {
        k_pf_split<2><<<nblk, 128, 0, st>>>(pm, eps, n, nPg, nullptr, cP, cM, psiP, psiM, g_e_pg, Cdeg);
    } else {
        const int* bits = nullptr;
        if (split_is_spectral(m->split)) {
            if (!elem_bits) {
                set_error("efb_pf_split: 3D spectral splits need the elem_bits workspace (Ne int32)");
                return 1;
            }
            cudaMemsetAsync(elem_bits, 0, sizeof(int) * (size_t)Ne, st);
            k_pf_case_bits<3><<<nblk, 128, 0, st>>>(pm, eps, n, nPg, elem_bits);
            bits = elem_bits;
        }
        k_pf_split<3><<<nblk, 128, 0, st>>>(pm, eps, n, nPg, bits, cP, cM, psiP, psiM, g_e_pg, Cdeg);
    }
    return check_launch("efb_pf_split");
}

extern "C" int efb_pf_elastic_Ke(const efb_pf_material* m, const efb_group* g, const int32_t* connect_dof, const double* u,
                                 const double* d, double k_res, double scale, double* Ke, void* stream) {
    if (!m || !g || !connect_dof || !u || !d || !Ke || !g->connect || !g->coord || !g->dN_pg || !g->N_pg || !g->w_pg) {
        set_error("efb_pf_elastic_Ke: bad arguments");
        return 1;
    }
    if (g->nPg != 1 || g->nPe != g->dim + 1 || m->dim != g->dim || m->split < 0 || m->split > EFB_SPLIT_HE) {
        set_error("efb_pf_elastic_Ke: one-point simplex elements only (dim %d, nPe %d, nPg %d)", (int)g->dim, (int)g->nPe, (int)g->nPg);
        return 3;  // the caller composes strain -> degradation -> split -> stiffness instead
    }
    if (g->Ne == 0) return 0;
    PfMat pm;
    memcpy(&pm, m, sizeof(pm));
    GroupView v;
    v.nPg = g->nPg; v.coord_stride = g->coord_stride; v.Ne = g->Ne; v.connect = g->connect; v.coord = g->coord;
    v.dN_pg = g->dN_pg; v.N_pg = g->N_pg; v.w_pg = g->w_pg;
    const unsigned nblk = (unsigned)((g->Ne + 127) / 128);
    if (g->dim == 2)
        k_pf_elastic_simplex<2><<<nblk, 128, 0, as_stream(stream)>>>(pm, v, connect_dof, u, d, k_res, scale, Ke);
    else
        k_pf_elastic_simplex<3><<<nblk, 128, 0, as_stream(stream)>>>(pm, v, connect_dof, u, d, k_res, scale, Ke);
    return check_launch("efb_pf_elastic_Ke");
}

extern "C" int efb_pf_history_rf(double* psiP, const double* psiP_old, int64_t n, int regu, double Gc, double l0, double* r,
                                 double* f, void* stream) {
    if (!psiP || (regu != EFB_REGU_AT1 && regu != EFB_REGU_AT2)) {
        set_error("efb_pf_history_rf: bad arguments");
        return 1;
    }
    if (n == 0) return 0;
    k_pf_history_rf<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(psiP, psiP_old, n, regu, Gc, l0, r, f);
    return check_launch("efb_pf_history_rf");
}
