// C-ABI entry point of the fused element-integration + CSR-assembly kernel (bodies and rationale: fused_kernels.cuh).
#include "common.cuh"
#include "fused_kernels.cuh"

#include <string.h>

namespace efb {

template <int DIM, int NPE, int NPG, bool ORTHO>
__global__ void __launch_bounds__(512) k_assemble_elastic(GroupView g, FusedView f, FusedTerms terms) {
    extern __shared__ __align__(16) double smem[];
    // persistent CTAs: the reference-element tables are staged once, clusters blockIdx.x, blockIdx.x + gridDim.x, ...
    bool first = true;
    for (long long cluster = blockIdx.x; cluster < f.n_clusters; cluster += gridDim.x) {
        if (!first) __syncthreads();  // the accumulators of the previous cluster share memory with this cluster's gather
        const long long nxt = cluster + gridDim.x;
        fused_cluster_block<DIM, NPE, NPG, ORTHO>(g, f, terms, cluster, blockDim.x, smem, first, nxt < f.n_clusters ? nxt : -1);
        first = false;
    }
}

// element types with an instantiation of the fused kernel: the record of one element (nPg*DIM*NPE doubles) times the
// elements of a cluster must fit shared memory, which rules out the high-order 3D elements
// X(DIM, NPE, NPG): the quadratures of the stiffness term (FEM/_gauss.py:423-488) are compiled in; NPG = 0 takes any other
#define EFB_FOR_EACH_FUSED(X)                                                                                    \
    X(2, 3, 1) X(2, 4, 4) X(2, 6, 3) X(2, 9, 9) X(3, 4, 1) X(3, 8, 8) X(3, 10, 4)                                \
    X(2, 3, 0) X(2, 4, 0) X(2, 6, 0) X(2, 8, 0) X(2, 9, 0) X(3, 4, 0) X(3, 8, 0) X(3, 10, 0)

template <int DIM, int NPE, int NPG>
static int launch_assemble(const efb_group* g, const FusedView& f, const FusedTerms& terms, cudaStream_t st) {
    using FC = Fused<DIM, NPE, NPG>;
    if (f.S % FC::G != 0 || f.S / FC::G > 16) {
        set_error("efb_assemble_elastic: S=%d must be a multiple of %d and at most %d", f.S, FC::G, 16 * FC::G);
        return 1;
    }
    const int nwarps = f.S / FC::G;
    const size_t bytes = sizeof(double) * FC::total(g->nPg, f.cap_e, nwarps, f.max_deg);
    if (bytes > 227 * 1024) {
        set_error("efb_assemble_elastic: a cluster of %d elements / max_deg %d needs %zu bytes of shared memory", f.cap_e, f.max_deg, bytes);
        return 3;
    }
    auto kern = terms.ortho ? k_assemble_elastic<DIM, NPE, NPG, true> : k_assemble_elastic<DIM, NPE, NPG, false>;
    if (ensure_smem(kern, bytes)) return 1;
    GroupView v;
    v.nPg = g->nPg; v.coord_stride = g->coord_stride; v.Ne = g->Ne; v.connect = g->connect; v.coord = g->coord;
    v.dN_pg = g->dN_pg; v.N_pg = g->N_pg; v.w_pg = g->w_pg;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nwarps * 32, bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm * 4;  // a few clusters per resident CTA slot: boundary clusters are lighter
    if (grid > f.n_clusters) grid = f.n_clusters;
    kern<<<(unsigned)grid, nwarps * 32, bytes, st>>>(v, f, terms);
    return check_launch("efb_assemble_elastic");
}

}  // namespace efb

using namespace efb;

extern "C" int efb_assemble_elastic_smem(int dim, int nPe, int nPg, int S, int cap_e, int max_deg) {
#define X(D, N, P)                                                                \
    if (dim == D && nPe == N && (P == nPg || P == 0)) {                           \
        if (S % Fused<D, N, P>::G != 0 || S / Fused<D, N, P>::G > 16) return -1;  \
        return (int)(sizeof(double) * Fused<D, N, P>::total(nPg, cap_e, S / Fused<D, N, P>::G, max_deg)); \
    }
    EFB_FOR_EACH_FUSED(X)
#undef X
    return -1;
}

extern "C" int efb_assemble_elastic_group(int dim, int nPe, int nPg) {
#define X(D, N, P) \
    if (dim == D && nPe == N && (P == nPg || P == 0)) return Fused<D, N, P>::G;
    EFB_FOR_EACH_FUSED(X)
#undef X
    return -1;
}

extern "C" int efb_assemble_elastic(const efb_group* g, const double* C_host, const double* w_pg_host, double scale, int n_clusters,
                                    int S, int cap_e, int max_deg, const int64_t* cl_nodes, const int32_t* cl_ne,
                                    const int32_t* cl_conn, const int32_t* desc, const int32_t* tpos, double* out, void* stream) {
    if (!g || !C_host || !w_pg_host || !cl_nodes || !cl_ne || !cl_conn || !desc || !tpos || !out || n_clusters < 0) {
        set_error("efb_assemble_elastic: bad arguments");
        return 1;
    }
    for (int p = 0; p < g->nPg; ++p)
        if (!(w_pg_host[p] > 0.0)) {
            set_error("efb_assemble_elastic: the quadrature has a non-positive weight (gradients travel scaled by sqrt(w))");
            return 3;
        }
    if (n_clusters == 0) return 0;
    FusedView f;
    f.n_clusters = n_clusters; f.S = S; f.cap_e = cap_e; f.max_deg = max_deg;
    f.cl_nodes = (const long long*)cl_nodes; f.cl_ne = cl_ne; f.cl_conn = cl_conn; f.desc = desc; f.tpos = tpos;
    f.out = out;
    FusedTerms terms;
    memset(&terms, 0, sizeof(terms));
    CMat C2;
    memset(&C2, 0, sizeof(C2));
#define X(D, N, P)                                                               \
    if (g->dim == D && g->nPe == N && (P == g->nPg || P == 0)) {                 \
        memcpy(C2.v, C_host, sizeof(double) * StrainSize<D>::value * StrainSize<D>::value); \
        prescale_C<D>(C2);                                                       \
        fused_terms_from_C2<D>(C2.v, scale, terms);                              \
        return launch_assemble<D, N, P>(g, f, terms, as_stream(stream));         \
    }
    EFB_FOR_EACH_FUSED(X)
#undef X
    set_error("efb_assemble_elastic: no fused instantiation for dim=%d nPe=%d", (int)g->dim, (int)g->nPe);
    return 3;
}
