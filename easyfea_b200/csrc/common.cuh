// Shared host-side helpers of the C-ABI translation units: error reporting, launch checks, dispatch.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/easyfea_b200.h"

namespace efb {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// opt in to large dynamic shared memory once per kernel
template <class K>
inline int ensure_smem(K kernel, size_t bytes) {
    // these kernels live in shared memory, not L1: ask for the largest carve-out so several CTAs fit an SM
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (bytes > 48 * 1024) {
        if (bytes > 227 * 1024) {
            set_error("kernel needs %zu bytes of shared memory (> 227 KB)", bytes);
            return 1;
        }
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (err != cudaSuccess) {
            set_error("cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(err));
            return 1;
        }
    }
    return 0;
}

}  // namespace efb

// (dim, nPe) pairs with a compiled instantiation; X(DIM, NPE)
#define EFB_FOR_EACH_ELEM(X) \
    X(2, 3) X(2, 4) X(2, 6) X(2, 8) X(2, 9) X(2, 10) X(3, 4) X(3, 6) X(3, 8) X(3, 10) X(3, 15) X(3, 18) X(3, 20) X(3, 27)
