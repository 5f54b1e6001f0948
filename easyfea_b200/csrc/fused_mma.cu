// Fused element integration + CSR assembly of HEXA8 / 8 Gauss points on the FP64 tensor-core instruction (DMMA.8x8x4):
// `efb_assemble_elastic_mma`, the fast instantiation of A3 `Assembly` for O1 with a homogeneous C
// (EasyFEA/Simulations/_simu.py:1104-1144 with Operators/Bilinear.py:62-79).  Algebra and ownership are those of
// fused_kernels.cuh: K[n,m] = Dt : sum_e T_ab(e), T_ab = sum_p g_a(p) (x) g_b(p), g = sqrt(w |det F|) grad N, every CSR slot
// summed in a fixed order by one lane, no atomics.  What changes is WHERE the contraction runs.
//
// The batched element contraction is a dense FP64 GEMM: with G (8 Gauss points x 24) the scaled gradients of one element,
// T_e = G^T G (24 x 8 by 8 x 24).  With the gradient index ordered (component, node) the m8n8k4 fragments fall out so that
//   * the B fragment of column tile l (columns = the 8 nodes b, component l) held by lane (b, p mod 4) is G[p][l][b] for
//     p = p mod 4 and p mod 4 + 4,
//   * a ROW tile = one pair of nodes (2j, 2j+1) x 3 components (6 of 8 rows used), so that row tiles whose nodes are not
//     owned by this cluster are skipped (the cluster needs T_ab only for OWNED a: 5 of 11.25 pairs per element on 4x2x2 bricks),
//   * the accumulator fragment of (row tile j, column tile l) gives lane ((a', k), b/2) the entries T_ab[k][l] of the two
//     blocks (a = 2j + a', b = 2 (lane mod 4) + {0, 1}): complete rows of 3x3 blocks, stored with one 128-bit store each.
// The Jacobians are a GEMM too: F[(p, r)][(e, c)] = dN[(p, r)][n] X[n][(e, c)] for four elements at a time, whose accumulator
// fragments hand lane (p, e) the complete 3x3 Jacobian of (element e, Gauss point p).
//
// One persistent CTA per SM, warp-specialised (no CTA-wide barrier in the steady state):
//   * a LOADER warp streams the per-cluster records (connectivity, staging slots, node records, round headers) and gather
//     programs into a 4-slot shared-memory ring with bulk asynchronous copies (cp.async.bulk on mbarriers), three clusters ahead;
//   * 11 INTEGRATION warps (16 warps per CTA is what 128 registers per thread allow; 20 warps at 96 registers measured 20-30 %
//     slower) take passes of 4 elements over the pass sequence, the assignment rotating from cluster to cluster: Jacobians (12 DMMA); lane (p, e) inverts its Jacobian and writes the 24 scaled gradients of
//     (e, p) to a per-warp table (bank-swizzled: conflict-free stores and fragment loads); per element the B fragments (6
//     loads), the needed row tiles (2 loads + 6 DMMA each) and the T rows of the owned nodes into one of TWO staging buffers
//     stage[slot][k*3+l][b ^ x] (slot = (owned node, element), x = node swizzle); the coordinates of a pass are requested one
//     pass ahead and carried in registers;
//   * 4 GATHER warps run the gather PROGRAM of the previous cluster out of the other staging buffer (built once,
//     assembly.MmaSchedule): every CSR block (node, slot) of the cluster is owned by one lane, which sums the staged
//     contributions of that block in ascending element order (blocks are sorted by their number of contributions so that the
//     32 lanes of a round run the same trip count), applies Dt and writes the 3x3 block.
// Ring slots and staging buffers are handed over with full/empty mbarriers (arrive = release, wait = acquire).
// FP64 DMMA runs on the FP64 pipe at the same flop rate as DFMA on B200 (measured 37.1 TFLOP/s, scripts/micro/dmma_peak.cu): the
// gain is 8x fewer issue slots and no operand traffic through shared memory, which is what bounded the DFMA form.
#include "common.cuh"
#include "fused_kernels.cuh"
#include "tma.cuh"

#include <string.h>

namespace efb {

struct MmaView {
    int n_clusters, cap4, t_cap, rmax, rec_words, pw_max;
    const int* recs;            // (n_clusters, rec_words): conn (cap4, 8) | rowslot (cap4, 8) int16 | nodes (16, 2) int64 | hdr (1 + rmax)
    const long long* prog_off;  // (n_clusters + 1) offset (int32 words) of every cluster's gather program
    const int* prog;            // per round: [dest of the 32 lanes][32 x cpad uint16 sources (staging offsets in doubles)]
    double* out;
};

constexpr int kMmaS = 16;                   // nodes per cluster
constexpr int kMmaTS = 72;                  // doubles per staged task: [k*3+l][b]
constexpr int kMmaZero = 72;                // zero tail of a staging buffer
constexpr int kMmaGE = 192;                 // doubles of the gradient table of one element: [p][l][b ^ swizzle(p)]
constexpr int kMmaWarpBuf = 2 * kMmaGE;     // per-warp scratch: two tables (element e + 1 is written while e is read)
#ifndef EFB_MMA_P2
#define EFB_MMA_P2 11
#endif
#ifndef EFB_MMA_P3
#define EFB_MMA_P3 4
#endif
#ifndef EFB_MMA_STAGES
#define EFB_MMA_STAGES 4
#endif
constexpr int kMmaP2 = EFB_MMA_P2;          // integration warps (16 warps per CTA: 128 registers per thread); -D overrides are for A/B builds
constexpr int kMmaP3 = EFB_MMA_P3;          // gather warps
constexpr int kMmaStages = EFB_MMA_STAGES;  // ring slots
constexpr int kMmaThreads = (kMmaP2 + kMmaP3 + 1) * 32;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// a global load that stays where it is written (the prefetch distance is the point: the compiler must not sink it to its use)
__device__ __forceinline__ double ldg_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int pad_trips(int c) { return c <= 1 ? 1 : (c <= 2 ? 2 : (c <= 4 ? 4 : (c + 7) & ~7)); }

__host__ __device__ inline size_t mma_slot_words(int rec_words, int pw_max) { return (size_t)rec_words + pw_max; }

template <bool ORTHO>
__global__ void __launch_bounds__(kMmaThreads, 1) k_assemble_hexa8_mma(GroupView g, MmaView f, FusedTerms terms) {
    extern __shared__ __align__(16) double smem[];
    const int stage_doubles = f.t_cap * kMmaTS + kMmaZero;  // + a block of zeros: the source of lanes without a contribution
    double* stage0 = smem;                                            // [2][t_cap * 72]
    double* gbuf0 = stage0 + 2 * (size_t)stage_doubles;               // [P2 warps][kMmaWarpBuf]
    int* ring0 = reinterpret_cast<int*>(gbuf0 + kMmaP2 * kMmaWarpBuf); // [stages][rec_words + pw_max]
    const int slot_words = (int)mma_slot_words(f.rec_words, f.pw_max);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring0 + (size_t)kMmaStages * slot_words);
    unsigned long long* full = bars;                  // [stages] record + program have landed
    unsigned long long* empty = bars + kMmaStages;    // [stages] every consumer warp is done with the slot
    unsigned long long* sfull = bars + 2 * kMmaStages;      // [2] the integration warps have staged the cluster
    unsigned long long* sempty = bars + 2 * kMmaStages + 2; // [2] the gather warps have consumed the buffer

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2 * kMmaZero) stage0[(size_t)(threadIdx.x / kMmaZero) * stage_doubles + f.t_cap * kMmaTS + threadIdx.x % kMmaZero] = 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kMmaP2 + kMmaP3);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sfull[b], kMmaP2);
            mbar_init(&sempty[b], kMmaP3);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const long long nk = (long long)blockIdx.x < f.n_clusters ? (f.n_clusters - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // offsets inside a ring slot (int32 words)
    const int o_rs = f.cap4 * 8, o_nodes = o_rs + f.cap4 * 4, o_hdr = o_nodes + 4 * kMmaS, o_prog = f.rec_words;

    if (warp == kMmaP2 + kMmaP3) {
        // ------------------------------------------------ loader ------------------------------------------------
        for (long long k = 0; k < nk; ++k) {
            const int s = (int)(k % kMmaStages);
            const long long c = blockIdx.x + k * gridDim.x;
            const long long po = f.prog_off[c], pn = f.prog_off[c + 1];
            mbar_wait(&empty[s], (unsigned)(((k / kMmaStages) & 1) ^ 1));
            if (elect_one()) {
                int* dst = ring0 + (size_t)s * slot_words;
                const unsigned pbytes = (unsigned)((pn - po) * 4);
                mbar_expect_tx(&full[s], (unsigned)f.rec_words * 4u + pbytes);
                bulk_load(dst, f.recs + c * f.rec_words, (unsigned)f.rec_words * 4u, &full[s]);
                if (pbytes) bulk_load(dst + o_prog, f.prog + po, pbytes, &full[s]);
            }
            __syncwarp();
        }
    } else if (warp >= kMmaP2) {
        // ------------------------------------------------ gather ------------------------------------------------
        const int j = warp - kMmaP2;
        for (long long k = 0; k < nk; ++k) {
            const int s = (int)(k % kMmaStages), b = (int)(k & 1);
            const int* slot = ring0 + (size_t)s * slot_words;
            const double* stage = stage0 + (size_t)b * stage_doubles;
            mbar_wait(&full[s], (unsigned)((k / kMmaStages) & 1));
            mbar_wait(&sfull[b], (unsigned)((k >> 1) & 1));
            const int R = slot[o_hdr];
            const long long* nrec = reinterpret_cast<const long long*>(slot + o_nodes);
#pragma unroll 1
            for (int r = j; r < R; r += kMmaP3) {
                const int h = slot[o_hdr + 1 + r];
                const int c = h & 0xff, cpad = pad_trips(c);
                const int* rd = slot + o_prog + (h >> 8);
                const int dest = rd[lane];
                const unsigned short* sp = reinterpret_cast<const unsigned short*>(rd + 32) + lane * cpad;
                double acc[9];
                EFB_UNROLL
                for (int q = 0; q < 9; ++q) acc[q] = 0.0;
#pragma unroll 2
                for (int it = 0; it < c; ++it) {  // lanes with fewer contributions read the zero block
                    const double* q0 = stage + sp[it];
                    EFB_UNROLL
                    for (int q = 0; q < 9; ++q) acc[q] += q0[q * 8];
                }
                if (dest >= 0) {
                    const int i = dest >> 16, sl = dest & 0xffff;
                    const long long off = nrec[i * 2];
                    const int rowlen = (int)nrec[i * 2 + 1] * 3;
                    double* dst = f.out + off + sl * 3;
                    EFB_UNROLL
                    for (int ii = 0; ii < 3; ++ii)
                        EFB_UNROLL
                        for (int jj = 0; jj < 3; ++jj) {
                            double v = 0.0;
                            EFB_UNROLL
                            for (int kk = 0; kk < 3; ++kk)
                                EFB_UNROLL
                                for (int l = 0; l < 3; ++l)
                                    if (!ORTHO || fused_ortho_term<3>(ii, jj, kk, l)) v += terms.Dt[ii * 3 + jj][kk * 3 + l] * acc[kk * 3 + l];
                            dst[(long long)ii * rowlen + jj] = v;
                        }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sempty[b]);
                mbar_arrive(&empty[s]);
            }
        }
    } else {
        // ---------------------------------------------- integration ----------------------------------------------
        double* gbuf = gbuf0 + warp * kMmaWarpBuf;
        const int lq = lane >> 2, lr = lane & 3;
        // reference-element constants of this lane
        double AF[3][2];   // Jacobian GEMM, A fragment: dN[p = lq][r][n = lr + 4 s]
        double dNq[3][2];  // gradient table: this lane writes G[p = lq][l][b = 2 lr + {0, 1}] and needs dN[p][m][b]
        EFB_UNROLL
        for (int r = 0; r < 3; ++r)
            EFB_UNROLL
            for (int s = 0; s < 2; ++s) {
                AF[r][s] = g.dN_pg[(lq * 3 + r) * 8 + lr + 4 * s];
                dNq[r][s] = g.dN_pg[(lq * 3 + r) * 8 + 2 * lr + s];
            }
        const double wq = g.w_pg[lq];
        const int arow = lq < 6 ? lq / 3 : 0, krow = lq < 6 ? lq % 3 : 0;  // row (a', k) of a row tile held by this lane
        const bool row_on = lq < 6;
        const int ef = lq >> 1, cpar = lq & 1;  // Jacobian GEMM, B fragment: column lq = (element lq / 2, coordinate lq % 2)
        const int npass = f.cap4 >> 2;
        // gradient table [p][l][b ^ (p & 2 ? 4 : 0)] (p-stride 24 doubles): conflict-free 128-bit stores and fragment loads
        const int gw_off = lq * 24 + ((2 * lr) ^ ((lq & 2) ? 4 : 0));        // this lane's pair (b = 2 lr, 2 lr + 1) of Gauss point lq
        const int gb_off = lr * 24 + (lq ^ ((lr & 2) ? 4 : 0));              // B fragment: G[p = lr (+4)][l][b = lq]
        const int ga_off = lr * 24 + krow * 8;                               // A fragment: G[p = lr (+4)][k][2 j + a'] (+ swizzled column)
        const int ga_swz = (lr & 2) ? 4 : 0;
        const int st_row = krow * 24;                                        // staging: row k of the lane's node, columns 2 lr, 2 lr + 1

        // coordinates of the lane's two nodes (lr, lr + 4) of element 4 i + lq / 2 of the cluster in ring slot `slot`
        auto load_coords = [&](const int* slot, int i, double (&x)[4]) {
            const int id0 = slot[(i * 4 + ef) * 8 + lr], id1 = slot[(i * 4 + ef) * 8 + lr + 4];
            const double* c0 = g.coord + (long long)(id0 < 0 ? 0 : id0) * g.coord_stride;
            const double* c1 = g.coord + (long long)(id1 < 0 ? 0 : id1) * g.coord_stride;
            x[0] = ldg_f64(c0 + cpar);
            x[1] = ldg_f64(c1 + cpar);
            x[2] = ldg_f64(c0 + 2);
            x[3] = ldg_f64(c1 + 2);
        };
                // lane (p = lq, e = lr): H = sqrt(w |det F|) F^-1 of (element e, Gauss point p) of the pass with coordinates x
        auto pass_jacobians = [&](const double (&x)[4], double (&H)[9]) {
            {
                const double z0 = cpar ? 0.0 : x[2], z1 = cpar ? 0.0 : x[3];
                double F[9];
                EFB_UNROLL
                for (int r = 0; r < 3; ++r) {
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                    dmma(a0, a1, AF[r][0], x[0]);
                    dmma(a0, a1, AF[r][1], x[1]);
                    dmma(b0, b1, AF[r][0], z0);
                    dmma(b0, b1, AF[r][1], z1);
                    F[r * 3 + 0] = a0;  // F[r][c] = sum_n dN[p][r][n] x_e[n][c]
                    F[r * 3 + 1] = a1;
                    F[r * 3 + 2] = b0;
                }
                const double det = det_inv<3>(F, H);
                const double sw = sqrt(wq * fabs(det));
                EFB_UNROLL
                for (int i = 0; i < 9; ++i) H[i] *= sw;
            }
        };
        // gradient tables, row tiles and staging of the 4 elements of a pass; rsp = their staging slots (4 x 8 int16)
        auto pass_tiles = [&](const double (&H)[9], const int* rsp, double* stage) {
            // gradient table of element e: the four lanes of Gauss point p fetch F^-1 of (e, p) from lane (p, e) and write two
            // columns each
            auto write_table = [&](int e, double* gt) {
                double He[9];
                EFB_UNROLL
                for (int i = 0; i < 9; ++i) He[i] = __shfl_sync(0xffffffffu, H[i], (lane & ~3) | e);
                EFB_UNROLL
                for (int l = 0; l < 3; ++l) {
                    const double v0 = He[l * 3] * dNq[0][0] + He[l * 3 + 1] * dNq[1][0] + He[l * 3 + 2] * dNq[2][0];
                    const double v1 = He[l * 3] * dNq[0][1] + He[l * 3 + 1] * dNq[1][1] + He[l * 3 + 2] * dNq[2][1];
                    *reinterpret_cast<double2*>(gt + gw_off + l * 8) = make_double2(v0, v1);
                }
            };
            write_table(0, gbuf);
            __syncwarp();
            EFB_UNROLL
            for (int e = 0; e < 4; ++e) {
                const double* ge = gbuf + (e & 1) * kMmaGE;
                if (e < 3) write_table(e + 1, gbuf + ((e + 1) & 1) * kMmaGE);
                const int4 rs = *reinterpret_cast<const int4*>(rsp + e * 4);
                const int pairs[4] = {rs.x, rs.y, rs.z, rs.w};
                if ((pairs[0] & pairs[1] & pairs[2] & pairs[3]) != -1) {  // an owned node (else: empty element slot), warp-uniform
                    double gB[2][3];
                    EFB_UNROLL
                    for (int s = 0; s < 2; ++s)
                        EFB_UNROLL
                        for (int l = 0; l < 3; ++l) gB[s][l] = ge[gb_off + s * 96 + l * 8];
                    EFB_UNROLL
                    for (int j = 0; j < 4; ++j) {
                        if (pairs[j] == -1) continue;  // neither node of the pair is owned here (warp-uniform)
                        double gA0 = 0.0, gA1 = 0.0;
                        if (row_on) {
                            gA0 = ge[ga_off + ((2 * j + arow) ^ ga_swz)];
                            gA1 = ge[ga_off + 96 + ((2 * j + arow) ^ ga_swz)];
                        }
                        const int v = arow ? (pairs[j] >> 16) : (int)(short)(pairs[j] & 0xffff);
                        const bool on = row_on && v >= 0;
                        const int xs = (v >> 8) & 7;
                        double* d0 = stage + (on ? (v & 0xff) : 0) * kMmaTS + st_row;
                        double* d1 = d0 + ((2 * lr + 1) ^ xs);
                        d0 += (2 * lr) ^ xs;
                        double c[3][2];
                        EFB_UNROLL
                        for (int l = 0; l < 3; ++l) {
                            c[l][0] = c[l][1] = 0.0;
                            dmma(c[l][0], c[l][1], gA0, gB[0][l]);
                        }
                        EFB_UNROLL
                        for (int l = 0; l < 3; ++l) dmma(c[l][0], c[l][1], gA1, gB[1][l]);
                        if (on) {
                            EFB_UNROLL
                            for (int l = 0; l < 3; ++l) {
                                d0[l * 8] = c[l][0];
                                d1[l * 8] = c[l][1];
                            }
                        }
                    }
                }
                __syncwarp();  // table e + 1 is complete, table e may be overwritten
            }
        };

        // this warp's passes: pass i of cluster k belongs to warp (k (npass + 1) + i) mod P2 — the assignment rotates from cluster
        // to cluster, so that light passes (boundary elements, the last partly filled pass) and heavy ones visit every warp;
        // (kn, in) = the next pass of this warp, xn its coordinates
        auto first_in = [&](long long k) { return (int)((warp + kMmaP2 - (int)((k * (npass + 1)) % kMmaP2)) % kMmaP2); };
        long long kn = 0;
        int in = first_in(0);
        while (kn < nk && in >= npass) {
            ++kn;
            in = first_in(kn);
        }
        double xn[4] = {0.0, 0.0, 0.0, 0.0};
        bool have = false;  // xn holds the coordinates of pass (kn, in)
#pragma unroll 1
        for (long long k = 0; k < nk; ++k) {
            const int s = (int)(k % kMmaStages), b = (int)(k & 1);
            const int* slot = ring0 + (size_t)s * slot_words;
            double* stage = stage0 + (size_t)b * stage_doubles;
            mbar_wait(&full[s], (unsigned)((k / kMmaStages) & 1));
            mbar_wait(&sempty[b], (unsigned)(((k >> 1) & 1) ^ 1));
#pragma unroll 1
            while (kn == k) {
                if (!have) load_coords(slot, in, xn);
                double H[9];
                pass_jacobians(xn, H);
                const int i = in;
                in += kMmaP2;
                while (kn < nk && in >= npass) {
                    ++kn;
                    in = first_in(kn);
                }
                have = false;
                if (kn < nk && kn - k < kMmaStages) {  // the next pass's coordinates travel while this pass is computed
                    if (kn != k) mbar_wait(&full[kn % kMmaStages], (unsigned)((kn / kMmaStages) & 1));
                    load_coords(ring0 + (size_t)(kn % kMmaStages) * slot_words, in, xn);
                    have = true;
                }
                pass_tiles(H, slot + o_rs + i * 16, stage);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sfull[b]);
                mbar_arrive(&empty[s]);
            }
        }
    }
}

static size_t mma_smem_bytes(int t_cap, int rec_words, int pw_max) {
    return sizeof(double) * (2 * ((size_t)t_cap * kMmaTS + kMmaZero) + (size_t)kMmaP2 * kMmaWarpBuf) +
           sizeof(int) * kMmaStages * mma_slot_words(rec_words, pw_max) + sizeof(unsigned long long) * (2 * kMmaStages + 4);
}

}  // namespace efb

using namespace efb;

extern "C" int efb_assemble_elastic_mma_smem(int t_cap, int rec_words, int pw_max) { return (int)mma_smem_bytes(t_cap, rec_words, pw_max); }

extern "C" int efb_assemble_elastic_mma(const efb_group* g, const double* C_host, const double* w_pg_host, double scale, int n_clusters,
                                        int cap4, int t_cap, int rmax, int rec_words, int pw_max, const int32_t* recs,
                                        const int64_t* prog_off, const int32_t* prog, double* out, void* stream) {
    if (!g || !C_host || !w_pg_host || !recs || !prog_off || !prog || !out || n_clusters < 0) {
        set_error("efb_assemble_elastic_mma: bad arguments");
        return 1;
    }
    if (g->dim != 3 || g->nPe != 8 || g->nPg != 8) {
        set_error("efb_assemble_elastic_mma: HEXA8 with 8 Gauss points only (dim=%d nPe=%d nPg=%d)", (int)g->dim, (int)g->nPe, (int)g->nPg);
        return 3;
    }
    for (int p = 0; p < g->nPg; ++p)
        if (!(w_pg_host[p] > 0.0)) {
            set_error("efb_assemble_elastic_mma: the quadrature has a non-positive weight (gradients travel scaled by sqrt(w))");
            return 3;
        }
    const size_t bytes = mma_smem_bytes(t_cap, rec_words, pw_max);
    if (t_cap < 1 || t_cap > 256 || (cap4 & 3) || rmax < 1 || (rec_words & 3) || (pw_max & 3) ||
        rec_words < cap4 * 12 + 4 * kMmaS + 1 + rmax || bytes > 227 * 1024) {
        set_error("efb_assemble_elastic_mma: clusters of %d tasks / %d element slots / %d rounds / %d program words are outside the kernel",
                  t_cap, cap4, rmax, pw_max);
        return 3;
    }
    if (n_clusters == 0) return 0;
    MmaView f;
    f.n_clusters = n_clusters; f.cap4 = cap4; f.t_cap = t_cap; f.rmax = rmax; f.rec_words = rec_words; f.pw_max = pw_max;
    f.recs = recs; f.prog_off = (const long long*)prog_off; f.prog = prog; f.out = out;
    FusedTerms terms;
    memset(&terms, 0, sizeof(terms));
    CMat C2;
    memset(&C2, 0, sizeof(C2));
    memcpy(C2.v, C_host, sizeof(double) * 36);
    prescale_C<3>(C2);
    fused_terms_from_C2<3>(C2.v, scale, terms);
    auto kern = terms.ortho ? k_assemble_hexa8_mma<true> : k_assemble_hexa8_mma<false>;
    if (ensure_smem(kern, bytes)) return 1;
    GroupView v;
    v.nPg = g->nPg; v.coord_stride = g->coord_stride; v.Ne = g->Ne; v.connect = g->connect; v.coord = g->coord;
    v.dN_pg = g->dN_pg; v.N_pg = g->N_pg; v.w_pg = g->w_pg;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = sms;  // persistent: one CTA per SM, clusters round-robin
    if (grid > n_clusters) grid = n_clusters;
    kern<<<(unsigned)grid, kMmaThreads, bytes, as_stream(stream)>>>(v, f, terms);
    return check_launch("efb_assemble_elastic_mma");
}
