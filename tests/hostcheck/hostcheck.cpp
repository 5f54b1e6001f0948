// TEST-ONLY host emulation of the kernel bodies (compiled by g++, never part of the product library).
// The CUDA kernels are written as phase-structured block bodies (csrc/frame.cuh); here the phase macro expands to
// a loop over thread ids, so the exact indexing / shared-memory logic of every kernel can be checked against the
// oracle in the GPU-less authoring container.  Pointers are HOST pointers.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../easyfea_b200/csrc/csr_kernels.cuh"
#include "../../easyfea_b200/csrc/elem_kernels.cuh"
#include "../../easyfea_b200/csrc/pf_math.cuh"
#include "../../easyfea_b200/csrc/scalar_warp.cuh"
#include "../../include/easyfea_b200.h"

using namespace efb;

static GroupView view_of(const efb_group* g) {
    GroupView v;
    v.nPg = g->nPg;
    v.coord_stride = g->coord_stride;
    v.Ne = g->Ne;
    v.connect = g->connect;
    v.coord = g->coord;
    v.dN_pg = g->dN_pg;
    v.N_pg = g->N_pg;
    v.w_pg = g->w_pg;
    return v;
}

static int epb_for(int TPE) {
    int epb = 192 / TPE;
    return epb < 1 ? 1 : (epb > 16 ? 16 : epb);
}

#define FOR_EACH(X) X(2, 3) X(2, 4) X(2, 6) X(2, 9) X(3, 4) X(3, 8) X(3, 10) X(3, 27)

extern "C" int hc_geometry(const efb_group* g, double* F, double* detF, double* jac, double* wJ, double* invF, double* dN, double* B) {
    GeomOut o{F, detF, jac, wJ, invF, dN, B};
#define X(D, N)                                                                                   \
    if (g->dim == D && g->nPe == N) {                                                             \
        const int TPE = D * N, EPB = epb_for(TPE);                                                \
        SmemMap<D, N> sm(g->nPg, EPB, 0);                                                         \
        std::vector<double> smem(sm.total());                                                     \
        for (long long b = 0; b * EPB < g->Ne; ++b) geometry_block<D, N>(view_of(g), o, EPB, b, EPB * TPE, smem.data()); \
        return 0;                                                                                 \
    }
    FOR_EACH(X)
#undef X
    return 2;
}

extern "C" int hc_geometry_parts(const efb_group* g, int dof_n, double* leftDisp, double* reaction, double* diffuse, double* source) {
    GeomPartsOut o{leftDisp, reaction, diffuse, source, dof_n};
    const bool grad = leftDisp || diffuse;
#define X(D, N)                                                                                   \
    if (g->dim == D && g->nPe == N) {                                                             \
        const int TPE = D * N, EPB = epb_for(TPE);                                                \
        SmemMap<D, N> sm(g->nPg, EPB, 0, grad);                                                   \
        std::vector<double> smem(sm.total());                                                     \
        for (long long b = 0; b * EPB < g->Ne; ++b) geometry_parts_block<D, N>(view_of(g), o, EPB, b, EPB * TPE, smem.data()); \
        return 0;                                                                                 \
    }
    FOR_EACH(X)
#undef X
    return 2;
}

template <int D, int N, int CMODE>
static void run_elastic(const efb_group* g, const double* C, double scale, double* out) {
    constexpr int NS = StrainSize<D>::value;
    using Tile = ElasticTile<D, N>;
    const int EPB = Tile::EPB;
    ElasticSmem<D, N> sm(g->nPg, EPB, CMODE == 2 ? g->nPg * NS * NS : (CMODE == 1 ? NS * NS : 0));
    std::vector<double> smem(sm.total());
    CMat Cc;
    memset(&Cc, 0, sizeof(Cc));
    if (CMODE == 0) {
        memcpy(Cc.v, C, sizeof(double) * NS * NS);
        prescale_C<D>(Cc);
    }
    for (long long b = 0; b * EPB < g->Ne; ++b)  // tables are staged by the first batch only, like the persistent kernel
        elastic_block<D, N, CMODE>(view_of(g), Cc, C, scale, out, b, Tile::THREADS, smem.data(), b == 0);
}

extern "C" int hc_elastic_Ke(const efb_group* g, const double* C, int C_mode, double scale, double* out) {
#define X(D, N)                                                       \
    if (g->dim == D && g->nPe == N) {                                 \
        if (C_mode == 0) run_elastic<D, N, 0>(g, C, scale, out);      \
        else if (C_mode == 1) run_elastic<D, N, 1>(g, C, scale, out); \
        else run_elastic<D, N, 2>(g, C, scale, out);                  \
        return 0;                                                     \
    }
    FOR_EACH(X)
#undef X
    return 2;
}

// warp-autonomous homogeneous-C forms (elem_kernels.cuh, ElasticWarp): form 1 = general, 2 = symmetric, 3 = symmetric + ortho
template <int D, int N, bool SYM, bool ORTHO>
static void run_elastic_warp(const efb_group* g, const double* C, double scale, double* out) {
    constexpr int NS = StrainSize<D>::value;
    using W = ElasticWarp<D, N>;
    CMat Cc;
    memset(&Cc, 0, sizeof(Cc));
    memcpy(Cc.v, C, sizeof(double) * NS * NS);
    prescale_C<D>(Cc);
    const GroupView v = view_of(g);
    std::vector<double> tab(W::tables(g->nPg)), wmem(W::per_warp(g->nPg, SYM));
    for (int i = 0; i < g->nPg * D * N; ++i) tab[(i / (D * N)) * W::TS + i % (D * N)] = g->dN_pg[i];
    for (int i = 0; i < g->nPg; ++i) tab[g->nPg * W::TS + i] = g->w_pg[i];
    for (long long b = 0; b * W::EPW < g->Ne; ++b)
        elastic_warp_batch_ref<D, N, SYM, ORTHO>(v, Cc, scale, out, b, tab.data(), tab.data() + g->nPg * W::TS, wmem.data());
}

extern "C" int hc_elastic_Ke_warp(const efb_group* g, const double* C, int form, double scale, double* out) {
#define X(D, N)                                                                  \
    if (g->dim == D && g->nPe == N) {                                            \
        if (form == 1) run_elastic_warp<D, N, false, false>(g, C, scale, out);   \
        else if (form == 2) run_elastic_warp<D, N, true, false>(g, C, scale, out); \
        else run_elastic_warp<D, N, true, true>(g, C, scale, out);               \
        return 0;                                                                \
    }
    X(2, 3) X(2, 4) X(2, 6) X(2, 9) X(3, 4) X(3, 8)
#undef X
    return 2;
}

template <int D, int N>
static int run_scalar(const efb_group* g, const ScalarOp& op) {
    if constexpr (N <= 8) {  // the warp-autonomous form the device takes for these element types (HC_SCALAR_BLOCK=1: the block form)
        if (!getenv("HC_SCALAR_BLOCK")) {
            using SW = ScalarWarp<D, N>;
            std::vector<double> smem(SW::total(g->nPg, op.has_k, 1));
            const GroupView v = view_of(g);
            scalar_warp_tables<D, N>(v, smem.data(), 0, 1);
            for (long long b = 0; b * SW::EPW < g->Ne; ++b) scalar_warp_batch<D, N>(v, op, b, smem.data(), smem.data() + SW::tables(g->nPg));
            return 0;
        }
    }
    const int TPE = N, EPB = epb_for(TPE);
    SmemMap<D, N> sm(g->nPg, EPB, N * N + N);
    std::vector<double> smem(sm.total());
    for (long long b = 0; b * EPB < g->Ne; ++b) scalar_block<D, N>(view_of(g), op, EPB, b, EPB * TPE, smem.data());
    return 0;
}

extern "C" int hc_scalar(const efb_group* g, const double* r, int r_mode, double r_scalar, int has_r, const double* A, int A_mode,
                         const double* k, int k_mode, double k_scalar, int has_k, const double* f, int f_mode, double f_scalar,
                         int has_f, int dof_n, double scale, double* Ke, double* Fe, int f_keep_axis) {
    ScalarOp op;
    op.r = r; op.r_mode = r_mode; op.r_scalar = r_scalar; op.has_r = has_r;
    op.A = A; op.A_mode = A_mode; op.k = k; op.k_mode = k_mode; op.k_scalar = k_scalar; op.has_k = has_k;
    op.f = f; op.f_mode = f_mode; op.f_scalar = f_scalar; op.has_f = has_f;
    op.dof_n = dof_n; op.scale = scale; op.Ke = Ke; op.Fe = Fe; op.f_keep_axis = f_keep_axis;
#define X(D, N) \
    if (g->dim == D && g->nPe == N) return run_scalar<D, N>(g, op);
    FOR_EACH(X)
#undef X
    return 2;
}

extern "C" int hc_strain(const efb_group* g, const int32_t* connect_dof, const double* u, double* eps) {
#define X(D, N)                                                                                             \
    if (g->dim == D && g->nPe == N) {                                                                       \
        const int TPE = D * N, EPB = epb_for(TPE);                                                          \
        SmemMap<D, N> sm(g->nPg, EPB, TPE);                                                                 \
        std::vector<double> smem(sm.total());                                                               \
        for (long long b = 0; b * EPB < g->Ne; ++b)                                                         \
            strain_block<D, N>(view_of(g), connect_dof, u, eps, EPB, b, EPB * TPE, smem.data());            \
        return 0;                                                                                           \
    }
    FOR_EACH(X)
#undef X
    return 2;
}

extern "C" int hc_internal_force(const efb_group* g, const double* sigma, double* out) {
#define X(D, N)                                                                                             \
    if (g->dim == D && g->nPe == N) {                                                                       \
        constexpr int NS = StrainSize<D>::value;                                                            \
        const int TPE = D * N, EPB = epb_for(TPE);                                                          \
        SmemMap<D, N> sm(g->nPg, EPB, g->nPg * NS);                                                         \
        std::vector<double> smem(sm.total());                                                               \
        for (long long b = 0; b * EPB < g->Ne; ++b)                                                         \
            internal_force_block<D, N>(view_of(g), sigma, out, EPB, b, EPB * TPE, smem.data());             \
        return 0;                                                                                           \
    }
    FOR_EACH(X)
#undef X
    return 2;
}

extern "C" int hc_hyper(const efb_group* g, const int32_t* connect_dof, const double* u, const double* dW, const double* d2W, double scale,
                        double* Ke, double* Re) {
#define X(D, N)                                                                                    \
    if (g->dim == D && g->nPe == N) {                                                              \
        const int TPE = D * N, EPB = epb_for(TPE);                                                 \
        SmemMap<D, N> sm(g->nPg, EPB, HyperSmem<D, N>::extra(g->nPg));                             \
        std::vector<double> smem(sm.total());                                                      \
        for (long long b = 0; b * EPB < g->Ne; ++b)                                                \
            hyper_block<D, N>(view_of(g), connect_dof, u, dW, d2W, scale, Ke, Re, EPB, b, EPB * TPE, smem.data()); \
        return 0;                                                                                  \
    }
    X(2, 3) X(2, 4) X(2, 6) X(2, 9) X(3, 4) X(3, 8) X(3, 10)
#undef X
    return 2;
}

extern "C" int hc_degradation(const efb_group* g, const int32_t* connect_dof, const double* d, double k_res, double* out) {
    for (long long i = 0; i < g->Ne * g->nPg; ++i)
        out[i] = degradation_at(connect_dof, d, g->N_pg, i / g->nPg, (int)(i % g->nPg), g->nPe, k_res);
    return 0;
}

template <int DIM>
static void pf_split_host(const PfMat& pm, const double* eps, long long Ne, int nPg, double* cP, double* cM, double* psiP,
                          double* psiM, const double* g, double* Cdeg) {
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS;
    std::vector<int> bits(Ne, 0);
    if (DIM == 3 && split_is_spectral(pm.split)) {
        for (long long i = 0; i < Ne * nPg; ++i) {
            double v[NS];
            decomposed_vector<NS>(pm, eps + i * NS, v);
            Inv3 q;
            invariants_3d(v, q);
            bits[i / nPg] |= case_bits_3d(q);
        }
    }
    for (long long i = 0; i < Ne * nPg; ++i) {
        double p[NC], m[NC];
        split_point<DIM>(pm, eps + i * NS, bits[i / nPg], p, m);
        if (cP) memcpy(cP + i * NC, p, sizeof(p));
        if (cM) memcpy(cM + i * NC, m, sizeof(m));
        if (psiP) psiP[i] = energy_density<NS>(p, eps + i * NS);
        if (psiM) psiM[i] = energy_density<NS>(m, eps + i * NS);
        if (Cdeg)
            for (int k = 0; k < NC; ++k) Cdeg[i * NC + k] = g[i] * p[k] + m[k];
    }
}

extern "C" int hc_pf_split(const efb_pf_material* m, const double* eps, int64_t Ne, int32_t nPg, double* cP, double* cM,
                           double* psiP, double* psiM, const double* g, double* Cdeg) {
    PfMat pm;
    memcpy(&pm, m, sizeof(pm));
    if (m->dim == 2)
        pf_split_host<2>(pm, eps, Ne, nPg, cP, cM, psiP, psiM, g, Cdeg);
    else
        pf_split_host<3>(pm, eps, Ne, nPg, cP, cM, psiP, psiM, g, Cdeg);
    return 0;
}

// ---- CSR pipeline, all stages on the host -----------------------------------------------------------------------
struct HcPattern {
    std::vector<long long> rowptr, qlist, adjptr;
    std::vector<int> adj, pos;
    int max_deg;
};

static void make_table(GroupTable& T, int n_groups, const int32_t* const* connect, const double* const* data, const int64_t* Ne,
                       const int32_t* nPe, int dof_n) {
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
}

extern "C" void* hc_pattern_build(int n_groups, const int32_t* const* connect, const int64_t* Ne, const int32_t* nPe, int64_t Nn) {
    GroupTable T;
    make_table(T, n_groups, connect, nullptr, Ne, nPe, 1);
    HcPattern* P = new HcPattern;
    std::vector<int> cnt(Nn, 0), cursor(Nn, 0);
    for (int g = 0; g < n_groups; ++g)
        for (long long i = 0; i < Ne[g] * nPe[g]; ++i) count_node_rows_item(connect[g], i, cnt.data());
    P->rowptr.assign(Nn + 1, 0);
    for (long long n = 0; n < Nn; ++n) P->rowptr[n + 1] = P->rowptr[n] + cnt[n];
    P->qlist.assign(P->rowptr[Nn], -1);
    // fill in REVERSE order to prove the sort restores determinism
    for (int g = n_groups - 1; g >= 0; --g)
        for (long long i = Ne[g] * nPe[g] - 1; i >= 0; --i)
            fill_node_rows_item(connect[g], i, T.qoff[g], P->rowptr.data(), cursor.data(), P->qlist.data());
    for (long long n = 0; n < Nn; ++n) sort_node_rows_item(n, P->rowptr.data(), P->qlist.data());
    std::vector<int> deg(Nn), buf(kAdjCap);
    P->max_deg = 0;
    for (long long n = 0; n < Nn; ++n) {
        deg[n] = gather_neighbours(T, n, P->rowptr.data(), P->qlist.data(), buf.data());
        if (deg[n] > P->max_deg) P->max_deg = deg[n];
    }
    P->adjptr.assign(Nn + 1, 0);
    for (long long n = 0; n < Nn; ++n) P->adjptr[n + 1] = P->adjptr[n] + deg[n];
    P->adj.resize(P->adjptr[Nn]);
    for (long long n = 0; n < Nn; ++n) {
        const int len = gather_neighbours(T, n, P->rowptr.data(), P->qlist.data(), buf.data());
        memcpy(P->adj.data() + P->adjptr[n], buf.data(), sizeof(int) * len);
    }
    P->pos.resize(T.poff[n_groups]);
    for (int g = 0; g < n_groups; ++g)
        for (long long i = 0; i < Ne[g] * nPe[g] * nPe[g]; ++i)
            slot_map_item(connect[g], nPe[g], i, P->adjptr.data(), P->adj.data(), P->pos.data() + T.poff[g]);
    return P;
}

extern "C" void hc_pattern_free(void* p) { delete (HcPattern*)p; }
extern "C" int64_t hc_pattern_nnz_node(void* p) { return ((HcPattern*)p)->adj.size(); }

extern "C" void hc_pattern_expand(void* p, int64_t Nn, int d, int64_t Ndof, int32_t* indptr, int32_t* indices) {
    HcPattern* P = (HcPattern*)p;
    for (long long r = 0; r <= Ndof; ++r) expand_indptr_item<int>(r, Nn, d, P->adjptr.data(), indptr);
    for (long long n = 0; n < Nn; ++n)
        for (long long t = P->adjptr[n]; t < P->adjptr[n + 1]; ++t) expand_indices_item<int>(t, n, d, P->adjptr.data(), P->adj.data(), indices);
}

extern "C" void hc_pattern_inv(void* p, int n_groups, const int32_t* const* connect, const int64_t* Ne, const int32_t* nPe, int d,
                               int32_t* inv) {
    HcPattern* P = (HcPattern*)p;
    GroupTable T;
    make_table(T, n_groups, connect, nullptr, Ne, nPe, d);
    for (int g = 0; g < n_groups; ++g)
        for (long long i = 0; i < T.koff[g + 1] - T.koff[g]; ++i)
            inv_map_item(connect[g], nPe[g], d, i, P->adjptr.data(), P->pos.data() + T.poff[g], inv + T.koff[g]);
}

extern "C" void hc_replay_matrix(void* p, int n_groups, const double* const* data, const int64_t* Ne, const int32_t* nPe, int d,
                                 int64_t Nn, double* out) {
    HcPattern* P = (HcPattern*)p;
    GroupTable T;
    make_table(T, n_groups, nullptr, data, Ne, nPe, d);
    std::vector<double> acc((size_t)d * d * P->max_deg);
    for (long long n = 0; n < Nn; ++n) {
#define FAST(D, N)                                                                                                          \
    if (n_groups == 1 && d == D && nPe[0] == N) {                                                                           \
        replay_node_fast<D, N, 4>(data[0], n, P->rowptr.data(), P->qlist.data(), P->adjptr.data(), P->pos.data(), acc.data(), out); \
        continue;                                                                                                           \
    }
        FAST(3, 8) FAST(3, 4) FAST(3, 27) FAST(2, 3) FAST(2, 9) FAST(1, 3) FAST(1, 8) FAST(1, 27)
#undef FAST
        replay_node(T, d, n, P->rowptr.data(), P->qlist.data(), P->adjptr.data(), P->pos.data(), acc.data(), out);
    }
}

extern "C" void hc_replay_vector(void* p, int n_groups, const double* const* data, const int64_t* Ne, const int32_t* nPe, int d,
                                 int64_t Nn, double* out) {
    HcPattern* P = (HcPattern*)p;
    GroupTable T;
    make_table(T, n_groups, nullptr, data, Ne, nPe, d);
    for (long long r = 0; r < Nn * d; ++r) replay_vector_item(T, d, r, P->rowptr.data(), P->qlist.data(), out);
}

#include "../../easyfea_b200/csrc/pf_fused.cuh"

extern "C" int hc_pf_elastic_Ke(const efb_pf_material* m, const efb_group* g, const int32_t* connect_dof, const double* u, const double* d,
                                double k_res, double scale, double* Ke) {
    if (g->nPg != 1 || g->nPe != g->dim + 1) return 3;
    PfMat pm;
    memcpy(&pm, m, sizeof(pm));
    for (long long e = 0; e < g->Ne; ++e) {
        if (g->dim == 2) pf_elastic_simplex_item<2>(pm, view_of(g), connect_dof, u, d, k_res, scale, e, Ke);
        else pf_elastic_simplex_item<3>(pm, view_of(g), connect_dof, u, d, k_res, scale, e, Ke);
    }
    return 0;
}

// ---- fused element integration + assembly (csrc/fused_kernels.cuh): one emulated CTA per cluster ----
#include "../../easyfea_b200/csrc/fused_kernels.cuh"

extern "C" int hc_assemble_elastic(const efb_group* g, const double* C, double scale, int n_clusters, int S, int cap_e, int max_deg,
                                   const int64_t* cl_nodes, const int32_t* cl_ne, const int32_t* cl_conn, const int32_t* desc,
                                   const int32_t* tpos, double* out) {
    FusedView f;
    f.n_clusters = n_clusters; f.S = S; f.cap_e = cap_e; f.max_deg = max_deg;
    f.cl_nodes = (const long long*)cl_nodes; f.cl_ne = cl_ne; f.cl_conn = cl_conn; f.desc = desc; f.tpos = tpos;
    f.out = out;
    FusedTerms terms;
    memset(&terms, 0, sizeof(terms));
    CMat C2;
    memset(&C2, 0, sizeof(C2));
#define X(D, N, P)                                                                                 \
    if (g->dim == D && g->nPe == N && (P == g->nPg || P == 0)) {                                   \
        using FC = Fused<D, N, P>;                                                                 \
        if (S % FC::G) return 1;                                                                   \
        const int nwarps = S / FC::G;                                                              \
        memcpy(C2.v, C, sizeof(double) * StrainSize<D>::value * StrainSize<D>::value);             \
        prescale_C<D>(C2);                                                                         \
        fused_terms_from_C2<D>(C2.v, scale, terms);                                                \
        std::vector<double> smem(FC::total(g->nPg, cap_e, nwarps, max_deg));                       \
        for (long long c = 0; c < n_clusters; ++c) {                                               \
            if (terms.ortho) fused_cluster_block<D, N, P, true>(view_of(g), f, terms, c, nwarps * 32, smem.data()); \
            else fused_cluster_block<D, N, P, false>(view_of(g), f, terms, c, nwarps * 32, smem.data()); \
        }                                                                                          \
        return 0;                                                                                  \
    }
    X(2, 3, 1) X(2, 4, 4) X(2, 6, 3) X(2, 9, 9) X(3, 4, 1) X(3, 8, 8) X(3, 10, 4) X(2, 3, 0) X(2, 4, 0) X(2, 6, 0) X(2, 9, 0) X(3, 4, 0) X(3, 8, 0) X(3, 10, 0)
#undef X
    return 2;
}
