// C-ABI entry points of the element kernels: launch configuration + (dim, nPe) dispatch.
#include "common.cuh"
#include "elem_kernels.cuh"
#include "scalar_warp.cuh"

#include <stdlib.h>
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

// elements per CTA: ~256 threads, and at most ~72 KB of shared memory so that 3 CTAs fit an SM; high-order elements
// (HEXA27: 34 KB of per-element geometry) would be left with one or two elements = a warp or two per CTA, so the budget
// grows to 2 and then 1 CTA per SM until the CTA has at least 4 warps
template <int DIM, int NPE>
static int elems_per_block(int TPE, int nPg, int extra, bool grad = true) {
    int best = 1;
    for (const size_t budget : {72 * 1024, 110 * 1024, 220 * 1024}) {
        int epb = 256 / TPE;
        if (epb < 1) epb = 1;
        if (epb > 16) epb = 16;
        while (epb > 1 && SmemMap<DIM, NPE>(nPg, epb, extra, grad).total() * sizeof(double) > budget) --epb;
        best = epb;
        if (epb * TPE >= 128) break;
    }
    return best;
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

template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_geometry_parts(GroupView g, GeomPartsOut o, int EPB) {
    extern __shared__ double smem[];
    geometry_parts_block<DIM, NPE>(g, o, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_geometry_parts(const efb_group* g, const GeomPartsOut& o, cudaStream_t st) {
    const bool grad = o.leftDisp || o.diffuse;
    const int TPE = DIM * NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, 0, grad);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, 0, grad);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_geometry_parts<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_geometry_parts<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), o, EPB);
    return check_launch("efb_geometry_parts");
}

// ---------------------------------------------------------------------------------------------------------
// Warp-specialised persistent kernel.  CTA = NCW consumer warps (the contraction: FP64 FMAs fed by broadcast LDS) + one
// producer warp (gather of connectivity/coordinates/C and the per-Gauss-point geometry), decoupled by two geometry
// buffers in shared memory and four named barriers (full[b] / empty[b]): the producer runs up to two batches ahead, so
// the global-load latency of the gather and the serial inverse-Jacobian chain never stall the FP64 warps.
// Which warp of the CTA plays the producer ROTATES with a per-SM launch counter: warps map to the four SM sub-partitions
// by warp id % 4 and each sub-partition has its own FP64 pipe, so co-resident CTAs must not all park their producer on
// the same sub-partition.
// A consumer lane stages its DIM finished rows in its own padded (bank-conflict-free) shared-memory tile and sends them
// to HBM with bulk asynchronous copies (TMA engine, cp.async.bulk): the rows of one node are contiguous in K_e, so one
// copy per lane (CS == 1) or per row (CS > 1) is enough, and the only synchronisation is the lane's own bulk-group wait —
// consumer warps never wait for each other and never sit on store queues.
// Each CTA walks batches blockIdx.x, blockIdx.x + gridDim.x, ...
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

enum { kBarFull = 1, kBarEmpty = 3 };  // ids 1,2 / 3,4 (0 is __syncthreads)

__device__ unsigned int g_sm_launch_counter[1024];  // per-SM CTA arrival counter (only its value mod #warps matters)

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(double* sdst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int DIM, int NPE>
constexpr int elastic_threads() {
    return ElasticTile<DIM, NPE>::THREADS + 32;
}

template <int DIM, int NPE>
constexpr int elastic_min_blocks() {
#ifndef EFB_ELASTIC_MINB
#define EFB_ELASTIC_MINB 2
#endif
    return EFB_ELASTIC_MINB;
}

template <int DIM, int NPE, int CMODE>
__global__ void __launch_bounds__(elastic_threads<DIM, NPE>(), elastic_min_blocks<DIM, NPE>())
    k_elastic(GroupView g, CMat C2, const double* __restrict__ C, double scale, double* __restrict__ out, long long nblk) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NS = StrainSize<DIM>::value, NC = NS * NS, NDOF = DIM * NPE;
    using SM = ElasticSmem<DIM, NPE>;
    using Tile = ElasticTile<DIM, NPE>;
    constexpr int TS = SM::TS, KE = SM::KE, EPB = Tile::EPB, NB = Tile::NB, CS = Tile::CS;
    constexpr int NCT = Tile::THREADS, NT = NCT + 32, NW = NT / 32;  // consumer threads, all threads, warps
    constexpr int ROW = NB * DIM;                                    // doubles of one row piece
    const int nPg = g.nPg;
    const int extra = CMODE == 2 ? nPg * NC : (CMODE == 1 ? NC : 0);
    const SM sm(nPg, EPB, extra);
    double* dNt = smem + sm.off_dN();
    double* wt = smem + sm.off_w();
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
    // role of this warp: roles 0..NW-2 are the consumer warps, role NW-1 is the producer
    const int lane = threadIdx.x & 31;
    const int role = ((threadIdx.x >> 5) + NW - s_rot) % NW;
    const int tid = role * 32 + lane;  // virtual thread id

    if (role == NW - 1) {
        // ---------------- producer warp ----------------
        // two-level software pipeline of the gather: node ids of batch k+2 and coordinates of batch k+1 are in flight
        // (in registers) while the geometry of batch k is computed
        constexpr int NLOAD = (EPB * NPE + 31) / 32;  // nodes per lane and batch
        int nid[NLOAD];
        double xc[NLOAD][DIM];
        const long long stride = gridDim.x;
        auto load_ids = [&](long long blk) {
            const long long e0 = blk * EPB;
            EFB_UNROLL
            for (int q = 0; q < NLOAD; ++q) {
                const long long i = e0 * NPE + lane + 32 * q;
                nid[q] = (blk < nblk && lane + 32 * q < EPB * NPE && i < g.Ne * NPE) ? g.connect[i] : -1;
            }
        };
        auto load_coords = [&]() {
            EFB_UNROLL
            for (int q = 0; q < NLOAD; ++q) {
                if (nid[q] >= 0) {
                    const double* src = g.coord + (long long)nid[q] * g.coord_stride;
                    EFB_UNROLL
                    for (int d = 0; d < DIM; ++d) xc[q][d] = src[d];
                }
            }
        };
        load_ids(blockIdx.x);
        load_coords();                      // coordinates of batch 0
        load_ids(blockIdx.x + stride);      // ids of batch 1
        int k = 0;
        for (long long blk = blockIdx.x; blk < nblk; blk += stride, ++k) {
            const int b = k & 1;
            if (k >= 2) named_sync(kBarEmpty + b, NT);  // consumers have finished reading buffer b (batch k-2)
            double* E0 = bufs + b * buf_stride;
            const long long e0 = blk * EPB;
            const int nvalid = (g.Ne - e0 < EPB) ? (int)(g.Ne - e0) : EPB;
            if (CMODE != 0) {  // raw C of the batch: 16 bytes per asynchronous copy, no register staging
                if ((extra & 1) == 0) {
                    const int per = extra / 2;  // 16-byte pieces per element
                    for (int idx = lane; idx < nvalid * per; idx += 32) {
                        const int el = idx / per, i = idx - el * per;
                        cp_async16(E0 + el * sm.per_elem() + sm.o_extra() + 2 * i, C + (e0 + el) * (long long)extra + 2 * i);
                    }
                } else {  // an odd number of doubles per element (2D, ns*ns = 9): pieces are not 16-byte aligned
                    elastic_gather_C<DIM, NPE>(sm, C, e0, nvalid, E0, lane, 32);
                }
            }
            EFB_UNROLL
            for (int q = 0; q < NLOAD; ++q) {  // coordinates of batch k: registers -> shared memory
                const int idx = lane + 32 * q;
                if (idx < nvalid * NPE) {
                    const int el = idx / NPE, a = idx - el * NPE;
                    double* X = E0 + el * sm.per_elem() + sm.o_X();
                    EFB_UNROLL
                    for (int d = 0; d < DIM; ++d) X[a * DIM + d] = xc[q][d];
                }
            }
            load_coords();                  // batch k+1 (its ids arrived during the previous iteration)
            load_ids(blk + 2 * stride);     // batch k+2
            __syncwarp();
            for (int task = lane; task < nvalid * nPg; task += 32)
                elastic_geometry_task<DIM, NPE>(sm, dNt, wt, scale, E0 + (task / nPg) * sm.per_elem(), task % nPg);
            if (CMODE != 0) cp_async_wait_all();
            __threadfence_block();
            named_arrive(kBarFull + b, NT);
        }
    } else {
        // ---------------- consumer warps ----------------
        double* my_stage = smem + sm.off_stage() + tid * Tile::LANE_STAGE;
        int k = 0;
        for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++k) {
            const int b = k & 1;
            const double* E0 = bufs + b * buf_stride;
            const long long e0 = blk * EPB;
            const int nvalid = (g.Ne - e0 < EPB) ? (int)(g.Ne - e0) : EPB;
            int el, a, b0;
            const bool mine = elastic_owner<DIM, NPE>(tid, nvalid, el, a, b0);
            named_sync(kBarFull + b, NT);  // geometry of batch k is in buffer b
            double acc[DIM][ROW];
            if (mine) {
                const double* E = E0 + el * sm.per_elem();
                elastic_rows<DIM, NPE, CMODE>(C2, E + sm.o_extra(), E + sm.o_wJ(), E + sm.o_gN(), nPg, a, b0, acc);
            }
            named_arrive(kBarEmpty + b, NT);  // buffer b may be refilled
            if (mine) {
                double* dst = out + (e0 + el) * (long long)KE + (long long)(a * DIM) * NDOF + b0 * DIM;
                if constexpr (Tile::kBulk) {
                    bulk_wait_read();  // my previous rows have left my staging tile
                    EFB_UNROLL
                    for (int i = 0; i < DIM; ++i)
                        EFB_UNROLL
                        for (int j = 0; j < ROW; j += 2) {
                            Pair v;
                            v.x = acc[i][j];
                            v.y = acc[i][j + 1];
                            *reinterpret_cast<Pair*>(my_stage + i * ROW + j) = v;
                        }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic-proxy stores -> async proxy
                    if constexpr (CS == 1) {
                        bulk_store(dst, my_stage, DIM * ROW * sizeof(double));  // DIM full rows are contiguous in K_e
                    } else {
                        EFB_UNROLL
                        for (int i = 0; i < DIM; ++i) bulk_store(dst + i * NDOF, my_stage + i * ROW, ROW * sizeof(double));
                    }
                    bulk_commit();
                } else {
                    EFB_UNROLL
                    for (int i = 0; i < DIM; ++i)
                        EFB_UNROLL
                        for (int j = 0; j < ROW; ++j) dst[i * NDOF + j] = acc[i][j];
                }
            }
        }
        if (Tile::kBulk) bulk_wait_read();  // shared memory outlives the last copy
    }
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
    const long long grid = persistent_grid(k_elastic<DIM, NPE, CMODE>, NT, bytes, nblk);
    k_elastic<DIM, NPE, CMODE><<<(unsigned)grid, NT, bytes, st>>>(view_of(g), C2, C, scale, out, nblk);
    return check_launch("efb_elastic_Ke");
}

// ---------------------------------------------------------------------------------------------------------
// Warp-autonomous stiffness kernel for a homogeneous C (bodies and rationale: elem_kernels.cuh, ElasticWarp).
// NW warps per CTA, no CTA barrier after the table load; warp `gw` walks batches gw, gw + GW, ...
// ---------------------------------------------------------------------------------------------------------
constexpr int kWarpKernelWarps = 4;

template <int DIM, int NPE, bool SYM, bool ORTHO>
__global__ void __launch_bounds__(kWarpKernelWarps * 32, 2)
    k_elastic_w(GroupView g, CMat C2, double scale, double* __restrict__ out, long long nbatch) {
    extern __shared__ __align__(128) double smem[];
    using W = ElasticWarp<DIM, NPE>;
    using Tile = ElasticTile<DIM, NPE>;
    constexpr int EPW = W::EPW, NDOF = W::NDOF, KE = W::KE, NW = kWarpKernelWarps;
    const int nPg = g.nPg, rec = W::rec(nPg);
    double* dNt = smem;
    double* wt = smem + nPg * W::TS;
    for (int i = threadIdx.x; i < nPg * DIM * NPE; i += NW * 32) dNt[(i / (DIM * NPE)) * W::TS + i % (DIM * NPE)] = g.dN_pg[i];
    for (int i = threadIdx.x; i < nPg; i += NW * 32) wt[i] = g.w_pg[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* stage = smem + W::tables(nPg) + warp * W::per_warp(nPg, SYM);
    double* X = stage + (SYM ? W::STAGE_SYM : W::STAGE_LANE);
    double* geo = X + W::XW;
    const int el = lane / NPE, a = lane - el * NPE;
    const long long gw = (long long)blockIdx.x * NW + warp, GW = (long long)gridDim.x * NW;

    // gather pipeline: node id of batch k+2 and coordinates of batch k+1 are in flight while batch k is computed
    int nid;
    double xc[DIM];
    auto load_id = [&](long long blk) {
        const long long i = blk * (EPW * NPE) + lane;
        nid = (lane < EPW * NPE && blk < nbatch && i < g.Ne * NPE) ? g.connect[i] : -1;
    };
    auto load_xc = [&]() {
        if (nid >= 0) {
            const double* src = g.coord + (long long)nid * g.coord_stride;
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) xc[d] = src[d];
        }
    };
    load_id(gw);
    load_xc();
    load_id(gw + GW);

    for (long long blk = gw; blk < nbatch; blk += GW) {
        const long long e0 = blk * EPW;
        const int nvalid = (g.Ne - e0 < EPW) ? (int)(g.Ne - e0) : EPW;
        if (lane < nvalid * NPE) {
            EFB_UNROLL
            for (int d = 0; d < DIM; ++d) X[lane * DIM + d] = xc[d];
        }
        load_xc();               // batch k+1 (its ids arrived during the previous iteration)
        load_id(blk + 2 * GW);   // batch k+2
        __syncwarp();
        for (int task = lane; task < nvalid * nPg; task += 32) {
            const int te = task / nPg, p = task - te * nPg;
            double* E = geo + te * rec;
            warp_geometry_task<DIM, NPE>(X + te * NPE * DIM, dNt + p * W::TS, scale * wt[p], E + p, E + nPg + p * W::GPS);
        }
        __syncwarp();
        const bool mine = el < nvalid;  // el < EPW holds for every lane below EPW*NPE
        if constexpr (SYM) {
            double acc[W::NBS][DIM][DIM];
            if (mine) warp_rows_sym<DIM, NPE, ORTHO>(C2, geo + el * rec, nPg, a, acc);
            if (mine && a == 0) bulk_wait_read();  // the previous copy of this element tile has left shared memory
            __syncwarp();
            if (mine) warp_store_sym<DIM, NPE>(acc, a, stage + el * W::ES);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (mine && a == 0) {
                bulk_store(out + (e0 + el) * (long long)KE, stage + el * W::ES, KE * sizeof(double));
                bulk_commit();
            }
        } else {
            double acc[NPE][DIM][DIM];
            if (mine) {
                warp_rows_full<DIM, NPE, ORTHO>(C2, geo + el * rec, nPg, a, acc);
                double* dst = out + (e0 + el) * (long long)KE + (long long)(a * DIM) * NDOF;
                if constexpr (Tile::kBulk) {
                    double* my_stage = stage + lane * Tile::LANE_STAGE;
                    bulk_wait_read();
                    EFB_UNROLL
                    for (int i = 0; i < DIM; ++i)
                        EFB_UNROLL
                        for (int c = 0; c < NDOF; c += 2) {
                            Pair v;
                            v.x = acc[c / DIM][i][c % DIM];
                            v.y = acc[(c + 1) / DIM][i][(c + 1) % DIM];
                            *reinterpret_cast<Pair*>(my_stage + i * NDOF + c) = v;
                        }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_store(dst, my_stage, DIM * NDOF * sizeof(double));
                    bulk_commit();
                } else {
                    EFB_UNROLL
                    for (int i = 0; i < DIM; ++i)
                        EFB_UNROLL
                        for (int c = 0; c < NDOF; ++c) dst[i * NDOF + c] = acc[c / DIM][i][c % DIM];
                }
            }
        }
    }
    bulk_wait_read();  // shared memory outlives the last copy
}

// which form of the homogeneous-C kernel runs: EFB_ELASTIC_KERNEL = v6 | w | wsym | wortho overrides the default
// (the fastest form the structure of C allows); dev/tuning knob, read at every call
enum ElasticForm { kFormV6 = 0, kFormW = 1, kFormWSym = 2, kFormWSymOrtho = 3 };

template <int DIM>
static int elastic_form_for(const CMat& C2) {
    constexpr int NS = StrainSize<DIM>::value;
    bool sym = true, ortho = true;
    for (int s = 0; s < NS; ++s)
        for (int r = 0; r < NS; ++r) {
            if (C2.v[s * NS + r] != C2.v[r * NS + s]) sym = false;
            const bool structural_zero = (s != r) && (s >= DIM || r >= DIM);
            if (structural_zero && C2.v[s * NS + r] != 0.0) ortho = false;
        }
    int form = sym ? (ortho ? kFormWSymOrtho : kFormWSym) : kFormW;
    if (const char* env = getenv("EFB_ELASTIC_KERNEL")) {
        if (!strcmp(env, "v6")) form = kFormV6;
        else if (!strcmp(env, "w")) form = kFormW;
        else if (!strcmp(env, "wsym") && sym) form = kFormWSym;
        else if (!strcmp(env, "wortho") && sym && ortho) form = kFormWSymOrtho;
    }
    return form;
}

template <int DIM, int NPE, bool SYM, bool ORTHO>
static int launch_elastic_w(const efb_group* g, const CMat& C2, double scale, double* out, cudaStream_t st) {
    using W = ElasticWarp<DIM, NPE>;
    const size_t bytes = sizeof(double) * (W::tables(g->nPg) + kWarpKernelWarps * W::per_warp(g->nPg, SYM));
    if (ensure_smem(k_elastic_w<DIM, NPE, SYM, ORTHO>, bytes)) return 1;
    const long long nbatch = (g->Ne + W::EPW - 1) / W::EPW;
    if (nbatch == 0) return 0;
    const long long nblk = (nbatch + kWarpKernelWarps - 1) / kWarpKernelWarps;
    const long long grid = persistent_grid(k_elastic_w<DIM, NPE, SYM, ORTHO>, kWarpKernelWarps * 32, bytes, nblk);
    k_elastic_w<DIM, NPE, SYM, ORTHO><<<(unsigned)grid, kWarpKernelWarps * 32, bytes, st>>>(view_of(g), C2, scale, out, nbatch);
    return check_launch("efb_elastic_Ke");
}

// element types that have the warp-autonomous instantiations (FP64-bound, DIM rows x all columns in one thread)
template <int DIM, int NPE>
struct HasWarpKernel {
    static constexpr bool value = (DIM == 3 && NPE == 8);
};

template <int DIM, int NPE>
static int launch_elastic_const(const efb_group* g, const CMat& C2, const double* C, double scale, double* out, cudaStream_t st) {
    if constexpr (HasWarpKernel<DIM, NPE>::value) {
        if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            switch (elastic_form_for<DIM>(C2)) {
                case kFormW: return launch_elastic_w<DIM, NPE, false, false>(g, C2, scale, out, st);
                case kFormWSym: return launch_elastic_w<DIM, NPE, true, false>(g, C2, scale, out, st);
                case kFormWSymOrtho: return launch_elastic_w<DIM, NPE, true, true>(g, C2, scale, out, st);
                default: break;
            }
        }
    }
    return launch_elastic_mode<DIM, NPE, 0>(g, C2, C, scale, out, st);
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
        return launch_elastic_const<DIM, NPE>(g, Cconst, C, scale, out, st);
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

// warp-autonomous form (scalar_warp.cuh) for elements of at most 8 nodes: persistent CTAs of 8 warps, one batch of 32/LPE elements
// per warp and pass, only warp barriers
#ifndef EFB_SCALAR_W_MINB
#define EFB_SCALAR_W_MINB 3  // 80 registers (HEXA8: 94 without the bound, no spills): 3 CTAs per SM for the operators without a gradient table
#endif
template <int DIM, int NPE>
__global__ void __launch_bounds__(256, EFB_SCALAR_W_MINB) k_scalar_w(GroupView g, ScalarOp op, long long nbatches) {
    extern __shared__ double smem[];
    using SW = ScalarWarp<DIM, NPE>;
    scalar_warp_tables<DIM, NPE>(g, smem, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* ws = smem + SW::tables(g.nPg) + warp * SW::per_warp(g.nPg, op.has_k);
    for (long long b = (long long)blockIdx.x * nw + warp; b < nbatches; b += (long long)gridDim.x * nw)
        scalar_warp_batch<DIM, NPE>(g, op, b, smem, ws);
}

template <int DIM, int NPE>
static int launch_scalar(const efb_group* g, const ScalarOp& op, cudaStream_t st, const char* what) {
    if constexpr (NPE <= 8) {
        // dev: EFB_SCALAR_BLOCK=1 forces the block form, =-1 the warp form
        static const int force = [] { const char* e = getenv("EFB_SCALAR_BLOCK"); return e ? atoi(e) : 0; }();
        // measured on B200 (scripts/scalar_probe.py, profiles/README.md): the warp form wins for every operator and element type
        // of at most 8 nodes (TRI3 2.2x, TETRA4 1.5-2.5x, QUAD4 1.3-1.5x, HEXA8 1.3-1.8x)
        const bool warp_wins = true;
        if (force < 0 || (force == 0 && warp_wins)) {
            using SW = ScalarWarp<DIM, NPE>;
            constexpr int NW = 8;
            const size_t bytes = sizeof(double) * SW::total(g->nPg, op.has_k, NW);
            if (bytes <= 100 * 1024) {
                if (ensure_smem(k_scalar_w<DIM, NPE>, bytes)) return 1;
                const long long nbatches = (g->Ne + SW::EPW - 1) / SW::EPW;
                if (nbatches == 0) return 0;
                int dev = 0, sms = 148, per_sm = 1;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scalar_w<DIM, NPE>, NW * 32, bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
                long long grid = (nbatches + NW - 1) / NW;
                if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
                k_scalar_w<DIM, NPE><<<(unsigned)grid, NW * 32, bytes, st>>>(view_of(g), op, nbatches);
                return check_launch(what);
            }
        }
    }
    const int TPE = NPE, EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, NPE * NPE + NPE, op.has_k);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, NPE * NPE + NPE, op.has_k);
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

template <int DIM, int NPE>
__global__ void __launch_bounds__(256) k_hyper(GroupView g, const int* connect_dof, const double* u, const double* dW, const double* d2W,
                                               double scale, double* Ke, double* Re, int EPB) {
    extern __shared__ double smem[];
    hyper_block<DIM, NPE>(g, connect_dof, u, dW, d2W, scale, Ke, Re, EPB, blockIdx.x, blockDim.x, smem);
}

template <int DIM, int NPE>
static int launch_hyper(const efb_group* g, const int* connect_dof, const double* u, const double* dW, const double* d2W, double scale,
                        double* Ke, double* Re, cudaStream_t st) {
    const int TPE = DIM * NPE, extra = HyperSmem<DIM, NPE>::extra(g->nPg), EPB = elems_per_block<DIM, NPE>(TPE, g->nPg, extra);
    const SmemMap<DIM, NPE> sm(g->nPg, EPB, extra);
    const size_t bytes = sizeof(double) * sm.total();
    if (ensure_smem(k_hyper<DIM, NPE>, bytes)) return 1;
    const long long nblk = (g->Ne + EPB - 1) / EPB;
    if (nblk == 0) return 0;
    k_hyper<DIM, NPE><<<(unsigned)nblk, EPB * TPE, bytes, st>>>(view_of(g), connect_dof, u, dW, d2W, scale, Ke, Re, EPB);
    return check_launch("efb_hyperelastic_Ke_Re");
}

__global__ void k_degradation(const int* connect_dof, const double* d, const double* N_pg, long long Ne, int nPg, int nPe,
                              double k_res, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ne * nPg) return;
    out[i] = degradation_at(connect_dof, d, N_pg, i / nPg, (int)(i % nPg), nPe, k_res);
}

}  // namespace efb

using namespace efb;

// element types of the hyperelastic operator: one thread keeps a whole K_e column (ndof accumulators) in registers
#define EFB_FOR_EACH_HYPER(X) X(2, 3) X(2, 4) X(2, 6) X(2, 8) X(2, 9) X(3, 4) X(3, 8) X(3, 10)
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

extern "C" int efb_geometry_parts(const efb_group* g, int dof_n, double* leftDisp, double* reaction, double* diffuse,
                                  double* source, void* stream) {
    if (validate(g)) return 1;
    if (dof_n < 1 || dof_n > 3) {
        set_error("efb_geometry_parts: dof_n must be 1..3");
        return 1;
    }
    GeomPartsOut o{leftDisp, reaction, diffuse, source, dof_n};
#define X(D, N) EFB_DISPATCH(D, N, (launch_geometry_parts<D, N>(g, o, as_stream(stream))))
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

extern "C" int efb_hyperelastic_Ke_Re(const efb_group* g, const int32_t* connect_dof, const double* u, const double* dWde,
                                      const double* d2Wde, double scale, double* Ke, double* Re, void* stream) {
    if (validate(g)) return 1;
    if (!connect_dof || !u || !dWde || !d2Wde || (!Ke && !Re)) {
        set_error("efb_hyperelastic_Ke_Re: bad arguments");
        return 1;
    }
#define X(D, N) EFB_DISPATCH(D, N, (launch_hyper<D, N>(g, connect_dof, u, dWde, d2Wde, scale, Ke, Re, as_stream(stream))))
    EFB_FOR_EACH_HYPER(X)
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
