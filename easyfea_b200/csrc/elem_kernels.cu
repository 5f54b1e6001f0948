// C-ABI entry points of the element kernels: launch configuration + (dim, nPe) dispatch.
#include <cuda.h>

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
// Warp-specialised persistent kernel.  CTA = NCW consumer warps (the contraction, pure FP64 + broadcast LDS) + one
// producer warp (gather of connectivity/coordinates/C and the per-Gauss-point geometry), decoupled by two geometry
// buffers in shared memory and four named barriers (full[b] / empty[b]): the producer runs up to two batches ahead, so
// the global-load latency of the gather and the serial inverse-Jacobian chain never stall the FP64 warps.
// Which warp of the CTA plays the producer ROTATES with a per-SM launch counter: warps map to the four SM sub-partitions
// by warp id % 4, and each sub-partition has its own FP64 pipe, so co-resident CTAs must not all park their producer
// (or, with 3 consumer warps, their idle slot) on the same sub-partition.
// A consumer warp stages its finished rows in its own shared-memory tile and writes them with ONE TMA tensor store
// (box = NB*DIM columns x NPE nodes x EPW elements of the view K_e[e][a][i][col]); the only synchronisation it needs for
// that is its own bulk-group wait, so consumer warps never wait for each other.
// Each CTA walks batches blockIdx.x, blockIdx.x + gridDim.x, ...
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

enum { kBarFull = 1, kBarEmpty = 3, kBarConsumers = 5 };  // ids 1,2 / 3,4 / 5 (0 is __syncthreads)

__device__ unsigned int g_sm_launch_counter[1024];  // per-SM CTA arrival counter (only its value mod #warps matters)

__device__ __forceinline__ void tma_store_tile_4d(const CUtensorMap* map, const double* ssrc, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int DIM, int NPE>
constexpr int elastic_threads() {
    return ElasticTile<DIM, NPE>::THREADS + 32;
}

template <int DIM, int NPE>
constexpr int elastic_min_blocks() {
#ifndef EFB_ELASTIC_MINB
#define EFB_ELASTIC_MINB 4
#endif
    return elastic_threads<DIM, NPE>() <= 160 ? EFB_ELASTIC_MINB : 1;
}

template <int DIM, int NPE, int CMODE>
__global__ void __launch_bounds__(elastic_threads<DIM, NPE>(), elastic_min_blocks<DIM, NPE>())
    k_elastic(GroupView g, CMat C2, const double* __restrict__ C, double scale, double* __restrict__ out, long long nblk,
              const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS;
    using SM = ElasticSmem<DIM, NPE>;
    using Tile = ElasticTile<DIM, NPE>;
    constexpr int TS = SM::TS, KE = SM::KE, EPB = Tile::EPB, EPW = Tile::EPW, NB = Tile::NB, CS = Tile::CS, WPG = Tile::WPG;
    constexpr int NCT = Tile::THREADS, NT = NCT + 32, NW = NT / 32;  // consumer threads, all threads, warps
    const int nPg = g.nPg;
    const int extra = CMODE == 2 ? nPg * NC : (CMODE == 1 ? NC : 0);
    const SM sm(nPg, EPB, extra);
    double* dNt = smem + sm.off_dN();
    double* wt = smem + sm.off_w();
    double* stage = smem + sm.off_stage();
    const int buf_stride = EPB * sm.per_elem();
    double* bufs = sm.elem(smem, 0);
    __shared__ int s_rot;

    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_rot = (int)(atomicAdd(&g_sm_launch_counter[smid & 1023], 1u) % NW);
    }
    for (int i = threadIdx.x; i < nPg * DIM * NPE; i += NT) dNt[(i / (DIM * NPE)) * TS + i % (DIM * NPE)] = g.dN_pg[i];
    for (int i = threadIdx.x; i < nPg; i += NT) wt[i] = g.w_pg[i];
    __syncthreads();
    // role of this warp: roles 0..NW-2 are the consumer warps (i, h) of elastic_contract, role NW-1 is the producer
    const int lane = threadIdx.x & 31;
    const int role = ((threadIdx.x >> 5) + NW - s_rot) % NW;
    const int tid = role * 32 + lane;  // virtual thread id

    if (role == NW - 1) {
        // ---------------- producer warp ----------------
        int k = 0;
        for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++k) {
            const int b = k & 1;
            if (k >= 2) named_sync(kBarEmpty + b, NT);  // consumers have finished reading buffer b (batch k-2)
            double* E0 = bufs + b * buf_stride;
            const long long e0 = blk * EPB;
            const int nvalid = (g.Ne - e0 < EPB) ? (int)(g.Ne - e0) : EPB;
            elastic_gather<DIM, NPE, CMODE>(g, sm, C, e0, nvalid, E0, lane, 32);
            __syncwarp();
            for (int task = lane; task < nvalid * nPg; task += 32)
                elastic_geometry_task<DIM, NPE>(sm, dNt, wt, scale, E0 + (task / nPg) * sm.per_elem(), task % nPg);
            __threadfence_block();
            named_arrive(kBarFull + b, NT);
        }
    } else {
        // ---------------- consumer warps ----------------
        const int grp = role / WPG, wg = role - grp * WPG;
        const int ci = wg / CS, ch = wg - ci * CS;  // row component and column chunk of this warp
        int k = 0;
        for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++k) {
            const int b = k & 1;
            const double* E0 = bufs + b * buf_stride;
            const long long e0 = blk * EPB;
            const int nvalid = (g.Ne - e0 < EPB) ? (int)(g.Ne - e0) : EPB;
            if constexpr (Tile::kTensor) {
                if (lane == 0) bulk_store_wait_read();  // my previous tile has left shared memory
                __syncwarp();
                named_sync(kBarFull + b, NT);  // geometry of batch k is in buffer b
                elastic_contract<DIM, NPE, CMODE, true>(sm, C2, nPg, nvalid, E0, stage, tid);
                named_arrive(kBarEmpty + b, NT);  // buffer b may be refilled
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic-proxy stores -> async proxy
                __syncwarp();
                if (lane == 0 && grp * EPW < nvalid)  // elements past Ne are clipped by the tensor map bounds
                    tma_store_tile_4d(&out_map, stage + role * Tile::WARP_TILE, ch * NB * DIM, ci, 0, (int)(e0 + grp * EPW));
            } else {
                double* gdst = out + e0 * (long long)KE;
                named_sync(kBarFull + b, NT);
                elastic_contract<DIM, NPE, CMODE, false>(sm, C2, nPg, nvalid, E0, stage, tid);
                named_arrive(kBarEmpty + b, NT);
                named_sync(kBarConsumers, NCT);
                for (int idx = tid; idx < nvalid * KE; idx += NCT) gdst[idx] = stage[idx];
                named_sync(kBarConsumers, NCT);
            }
        }
        if (Tile::kTensor && lane == 0) bulk_store_wait_read();  // shared memory outlives the last copy
    }
}

// 4-D tensor map of the output viewed as K_e[e][a][i][col] (innermost first: col, i, a, e) with the store box of one warp
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <int DIM, int NPE>
static int make_out_map(CUtensorMap* map, double* out, long long Ne) {
    using Tile = ElasticTile<DIM, NPE>;
    memset(map, 0, sizeof(*map));
    if (!Tile::kTensor) return 0;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    constexpr cuuint64_t NDOF = DIM * NPE;
    const cuuint64_t gdim[4] = {NDOF, (cuuint64_t)DIM, (cuuint64_t)NPE, (cuuint64_t)Ne};
    const cuuint64_t gstride[3] = {NDOF * 8, DIM * NDOF * 8, NDOF * NDOF * 8};  // bytes, dims 1..3
    const cuuint32_t box[4] = {(cuuint32_t)(Tile::NB * DIM), 1, (cuuint32_t)NPE, (cuuint32_t)Tile::EPW};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return 1;
    }
    return 0;
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
    const size_t bytes = sizeof(double) * sm.total2();
    if (ensure_smem(k_elastic<DIM, NPE, CMODE>, bytes)) return 1;
    const long long nblk = (g->Ne + Tile::EPB - 1) / Tile::EPB;
    if (nblk == 0) return 0;
    constexpr int NT = elastic_threads<DIM, NPE>();
    CUtensorMap map;
    if (make_out_map<DIM, NPE>(&map, out, g->Ne)) return 1;
    const long long grid = persistent_grid(k_elastic<DIM, NPE, CMODE>, NT, bytes, nblk);
    k_elastic<DIM, NPE, CMODE><<<(unsigned)grid, NT, bytes, st>>>(view_of(g), C2, C, scale, out, nblk, map);
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
