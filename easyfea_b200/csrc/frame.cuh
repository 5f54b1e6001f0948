// Kernel frame: one source for the device kernels and for a host EMULATION used only by tests/hostcheck.
//
// Block-level kernel bodies are written as sequences of PHASES; a phase is executed by every thread of the block
// and ends with a block barrier.  No per-thread state lives across phases (it lives in shared memory), so the
// same body can be compiled by g++ with the phase macro expanding to a loop over thread ids.  The host build is
// a CHECKER for indexing/math in the GPU-less authoring container (tests/hostcheck) — the product never runs it.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define EFB_HD __host__ __device__ __forceinline__
#define EFB_D __device__ __forceinline__
// run the following statement once for this thread, then barrier
#define EFB_PHASE(tid, nthreads) for (int tid = threadIdx.x, _efb_once = 1; _efb_once; _efb_once = 0, __syncthreads())
// same, but when `warp_local` (every element's threads sit inside one warp) a warp barrier is enough, so the warps of
// a CTA drift apart and one warp's geometry phases overlap another warp's FP64 main loop
#define EFB_PHASE_E(tid, nthreads, warp_local) \
    for (int tid = threadIdx.x, _efb_once = 1; _efb_once; _efb_once = 0, ((warp_local) ? __syncwarp() : __syncthreads()))
#define EFB_RESTRICT __restrict__
#define EFB_UNROLL _Pragma("unroll")
#else
#define EFB_HD inline
#define EFB_D inline
#define EFB_PHASE(tid, nthreads) for (int tid = 0; tid < (nthreads); ++tid)
#define EFB_PHASE_E(tid, nthreads, warp_local) for (int tid = 0; tid < (nthreads); ++tid)
#define EFB_RESTRICT
#define EFB_UNROLL
#endif

namespace efb {

constexpr double kInvSqrt2 = 0.70710678118654746;  // 1/np.sqrt(2) rounded to nearest (0x3FE6A09E667F3BCC)
constexpr double kSqrt2 = 1.4142135623730951;

template <int DIM>
struct StrainSize {
    static constexpr int value = (DIM == 2) ? 3 : 6;
};

}  // namespace efb
