// C-ABI entry points of the element kernels: launch configuration + (dim, nPe) dispatch.
#include "common.cuh"
#include "elem_kernels.cuh"

#include <string.h>

namespace efb {

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

// elements per CTA: ~256 threads, and at most ~72 KB of shared memory so that 3 CTAs fit an SM
template <int DIM, int NPE>
static int elems_per_block(int TPE, int nPg, int extra) {
    int epb = 256 / TPE;
    if (epb < 1) epb = 1;
    if (epb > 16) epb = 16;
    while (epb > 1 && SmemMap<DIM, NPE>(nPg, epb, extra).total() * sizeof(double) > 72 * 1024) --epb;
    return epb;
}

static int validate(const efb_group* g) {
    if (!g || !g->connect || !g->coord || !g->dN_pg || !g->N_pg || !g->w_pg) {
        set_error("efb_group: null pointer");
        return 1;
    }
    if (g->nPg <= 0 || g->Ne < 0 || g->coord_stride < g->dim) {
        set_error("efb_group: bad sizes (nPg=%d, Ne=%lld, coord_stride=%d)", g->nPg, (long long)g->Ne, g->coord_stride);
        return 1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_geometry(GroupView g, GeomOut o, int EPB) {
    extern __shared__ double smem[];
    geometry_block<DIM, NPE>(g, o, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_geometry(const efb_group* g, const GeomOut& o, cudaStream_t st) {
    const int TPE = DIM * NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, 0);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, 0);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_geometry<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_geometry<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), o, EPB);
    return check_launch("efb_geometry");
}

// ---------------------------------------------------------------------------------------------------------
// persistent: each CTA walks element batches blockIdx.x, blockIdx.x + gridDim.x, ...; the reference-element tables are
// staged once per CTA; several CTAs share an SM, so one CTA's gather/geometry phases overlap another's FP64 main loop
template <int DIM, int NPE>
constexpr int elastic_min_blocks() {
    // aim at <= 112 registers per thread where the CTA is small enough for that to matter
#ifndef EFB_ELASTIC_MINB96
#define EFB_ELASTIC_MINB96 6
#endif
    return ElasticTile<DIM, NPE>::THREADS <= 96 ? EFB_ELASTIC_MINB96 : (ElasticTile<DIM, NPE>::THREADS <= 128 ? 4 : 1);
}

template <int DIM, int NPE, int CMODE>
__global__ void __launch_bounds__(ElasticTile<DIM, NPE>::THREADS, elastic_min_blocks<DIM, NPE>())
    k_elastic(GroupView g, CMat C2, const double* C, double scale, double* out, long long nblk) {
    extern __shared__ __align__(16) double smem[];
    bool first = true;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        elastic_block<DIM, NPE, CMODE>(g, C2, C, scale, out, blk, blockDim.x, smem, first);
        first = false;
    }
    if (ElasticTile<DIM, NPE>::kBulk && threadIdx.x == 0) bulk_store_wait_read();  // shared memory outlives the last copy
}

// CTAs that keep every SM full: resident CTAs per SM (occupancy query) x number of SMs
template <class K>
static long long persistent_grid(K kernel, int threads, size_t smem_bytes, long long nblk) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long full = (long long)sms * per_sm;
    return nblk < full ? nblk : full;
}

template <int DIM, int NPE, int CMODE>
static int launch_elastic_mode(const efb_group* g, const CMat& C2, const double* C, double scale, double* out, cudaStream_t st) {
    constexpr int NS = StrainSize<DIM>::value;
    using SM = ElasticSmem<DIM, NPE>;
    using Tile = ElasticTile<DIM, NPE>;
    const int extra = CMODE == 2 ? g->nPg * NS * NS : (CMODE == 1 ? NS * NS : 0);
    const SM sm(g->nPg, Tile::EPB, extra);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_elastic<DIM, NPE, CMODE>, bytes)) return 1;
    const long long nblk = (g->Ne + Tile::EPB - 1) / Tile::EPB;
    if (nblk == 0) return 0;
    const long long grid = persistent_grid(k_elastic<DIM, NPE, CMODE>, Tile::THREADS, bytes, nblk);
    k_elastic<DIM, NPE, CMODE><<<(unsigned)grid, Tile::THREADS, bytes, st>>>(view_of(g), C2, C, scale, out, nblk);
    return check_launch("efb_elastic_Ke");
}

template <int DIM, int NPE>
static int launch_elastic(const efb_group* g, const double* C, const double* C_host, int C_mode, double scale, double* out,
                          cudaStream_t st) {
    constexpr int NS = StrainSize<DIM>::value;
    CMat Cconst;
    memset(&Cconst, 0, sizeof(Cconst));
    if (C_mode == 0) {
        // a homogeneous C travels as a kernel argument (constant bank)
        if (C_host) {
            memcpy(Cconst.v, C_host, sizeof(double) * NS * NS);
        } else {  // only a device copy was given: fetch it (synchronises the stream)
            cudaError_t err = cudaMemcpyAsync(Cconst.v, C, sizeof(double) * NS * NS, cudaMemcpyDeviceToHost, st);
            if (err == cudaSuccess) err = cudaStreamSynchronize(st);
            if (err != cudaSuccess) {
                set_error("efb_elastic_Ke: reading C: %s", cudaGetErrorString(err));
                return 1;
            }
        }
        prescale_C<DIM>(Cconst);  // Kelvin-Mandel C -> S C S (elem_kernels.cuh, O1)
        return launch_elastic_mode<DIM, NPE, 0>(g, Cconst, C, scale, out, st);
    }
    if (C_mode == 1) return launch_elastic_mode<DIM, NPE, 1>(g, Cconst, C, scale, out, st);
    return launch_elastic_mode<DIM, NPE, 2>(g, Cconst, C, scale, out, st);
}

// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_scalar(GroupView g, ScalarOp op, int EPB) {
    extern __shared__ double smem[];
    scalar_block<DIM, NPE>(g, op, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_scalar(const efb_group* g, const ScalarOp& op, cudaStream_t st, const char* what) {
    const int TPE = NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, NPE * NPE + NPE);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, NPE * NPE + NPE);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_scalar<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_scalar<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), op, EPB);
    return check_launch(what);
}

// ---------------------------------------------------------------------------------------------------------
template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_strain(GroupView g, const int* connect_dof, const double* u, double* eps, int EPB) {
    extern __shared__ double smem[];
    strain_block<DIM, NPE>(g, connect_dof, u, eps, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_strain(const efb_group* g, const int* connect_dof, const double* u, double* eps, cudaStream_t st) {
    const int TPE = DIM * NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, TPE);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, TPE);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_strain<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_strain<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), connect_dof, u, eps, EPB);
    return check_launch("efb_strain");
}

template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_internal_force(GroupView g, const double* sigma, double* out, int EPB) {
    extern __shared__ double smem[];
    internal_force_block<DIM, NPE>(g, sigma, out, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_internal_force(const efb_group* g, const double* sigma, double* out, cudaStream_t st) {
    constexpr int NS = StrainSize<DIM>::value;
    const int TPE = DIM * NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, g->nPg * NS);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, g->nPg * NS);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_internal_force<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_internal_force<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), sigma, out, EPB);
    return check_launch("efb_internal_force");
}

__global__ void k_degradation(const int* connect_dof, const double* d, const double* N_pg, long long Ne, int nPg, int nPe,
                              double k_res, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ne * nPg) return;
    out[i] = degradation_at(connect_dof, d, N_pg, i / nPg, (int)(i % nPg), nPe, k_res);
}

}  // namespace efb

using namespace efb;

#define EFB_DISPATCH(D, N, CALL) \
    if (g->dim == D && g->nPe == N) return CALL;
#define EFB_NO_INSTANCE(g)                                                                  \
    set_error("no kernel instantiation for dim=%d nPe=%d", (int)(g)->dim, (int)(g)->nPe); \
    return 2;

extern "C" int efb_geometry(const efb_group* g, double* F, double* detF, double* jac, double* wJ, double* invF, double* dN,
                            double* B, void* stream) {
    if (validate(g)) return 1;
    GeomOut o{F, detF, jac, wJ, invF, dN, B};
#define X(D, N) EFB_DISPATCH(D, N, (launch_geometry<D, N>(g, o, as_stream(stream))))
    EFB_FOR_EACH_ELEM(X)
#undef X
    EFB_NO_INSTANCE(g)
}

extern "C" int efb_elastic_Ke(const efb_group* g, const double* C, const double* C_host, int C_mode, double scale, double* out,
                              void* stream) {
    if (validate(g)) return 1;
    if ((!C && !(C_mode == 0 && C_host)) || !out || C_mode < 0 || C_mode > 2) {
        set_error("efb_elastic_Ke: bad arguments");
        return 1;
    }
#define X(D, N) EFB_DISPATCH(D, N, (launch_elastic<D, N>(g, C, C_host, C_mode, scale, out, as_stream(stream))))
    EFB_FOR_EACH_ELEM(X)
#undef X
    EFB_NO_INSTANCE(g)
}

static int run_scalar(const efb_group* g, const ScalarOp& op, void* stream, const char* what) {
    if (validate(g)) return 1;
#define X(D, N) EFB_DISPATCH(D, N, (launch_scalar<D, N>(g, op, as_stream(stream), what)))
    EFB_FOR_EACH_ELEM(X)
#undef X
    EFB_NO_INSTANCE(g)
}

static ScalarOp empty_op() {
    ScalarOp op;
    op.r = nullptr; op.r_mode = 0; op.r_scalar = 0.0; op.has_r = false;
    op.A = nullptr; op.A_mode = 0; op.k = nullptr; op.k_mode = 0; op.k_scalar = 0.0; op.has_k = false;
    op.f = nullptr; op.f_mode = 0; op.f_scalar = 0.0; op.has_f = false;
    op.dof_n = 1; op.scale = 1.0; op.Ke = nullptr; op.Fe = nullptr; op.f_keep_axis = false;
    return op;
}

extern "C" int efb_mass_Me(const efb_group* g, const double* coef, int coef_mode, double coef_scalar, int dof_n, double scale,
                           double* out, void* stream) {
    ScalarOp op = empty_op();
    op.has_r = true; op.r = coef; op.r_mode = coef_mode; op.r_scalar = coef_scalar;
    op.dof_n = dof_n; op.scale = scale; op.Ke = out;
    return run_scalar(g, op, stream, "efb_mass_Me");
}

extern "C" int efb_diffusion_Ke(const efb_group* g, const double* A, int A_mode, const double* coef, int coef_mode,
                                double coef_scalar, double scale, double* out, void* stream) {
    ScalarOp op = empty_op();
    op.has_k = true; op.A = A; op.A_mode = A_mode; op.k = coef; op.k_mode = coef_mode; op.k_scalar = coef_scalar;
    op.scale = scale; op.Ke = out;
    return run_scalar(g, op, stream, "efb_diffusion_Ke");
}

extern "C" int efb_source_Fe(const efb_group* g, const double* f, int f_mode, double f_scalar, int dof_n, double scale,
                             double* out, void* stream) {
    ScalarOp op = empty_op();
    op.has_f = true; op.f = f; op.f_mode = f_mode; op.f_scalar = f_scalar;
    op.dof_n = dof_n; op.scale = scale; op.Fe = out; op.f_keep_axis = true;
    return run_scalar(g, op, stream, "efb_source_Fe");
}

extern "C" int efb_pf_damage_Ke_Fe(const efb_group* g, const double* r, const double* f, const double* A, double k,
                                   double scale, double* Ke, double* Fe, void* stream) {
    ScalarOp op = empty_op();
    op.has_r = true; op.r = r; op.r_mode = EFB_COEF_E_PG;
    op.has_k = true; op.A = A; op.A_mode = 0; op.k_scalar = k;
    op.has_f = true; op.f = f; op.f_mode = EFB_COEF_E_PG;
    op.scale = scale; op.Ke = Ke; op.Fe = Fe; op.f_keep_axis = false;
    return run_scalar(g, op, stream, "efb_pf_damage_Ke_Fe");
}

extern "C" int efb_internal_force(const efb_group* g, const double* sigma, double* out, void* stream) {
    if (validate(g)) return 1;
#define X(D, N) EFB_DISPATCH(D, N, (launch_internal_force<D, N>(g, sigma, out, as_stream(stream))))
    EFB_FOR_EACH_ELEM(X)
#undef X
    EFB_NO_INSTANCE(g)
}

extern "C" int efb_strain(const efb_group* g, const int32_t* connect_dof, const double* u, double* eps, void* stream) {
    if (validate(g)) return 1;
#define X(D, N) EFB_DISPATCH(D, N, (launch_strain<D, N>(g, connect_dof, u, eps, as_stream(stream))))
    EFB_FOR_EACH_ELEM(X)
#undef X
    EFB_NO_INSTANCE(g)
}

extern "C" int efb_pf_degradation(const efb_group* g, const int32_t* connect_dof, const double* d, double k_res, double* out,
                                  void* stream) {
    if (validate(g)) return 1;
    const long long n = g->Ne * g->nPg;
    if (n == 0) return 0;
    k_degradation<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(connect_dof, d, g->N_pg, g->Ne, g->nPg, g->nPe,
                                                                                k_res, out);
    return check_launch("efb_pf_degradation");
}
