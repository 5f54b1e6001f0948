// dev microbenchmark: sustained DFMA throughput of one B200 for different numbers of warps per SM and independent chains per
// thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a dfma_peak.cu -o dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(int warps_per_sm, double* out) {
    const int iters = 4000;
    const int threads = 128;
    const int blocks = 148 * warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dfma<NACC><<<blocks, threads>>>(out, 1.0000001, 1e-9, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_dfma<NACC><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * (double)blocks * threads * NACC * 8.0 * iters;
    printf("warps/SM %2d  chains/thread %2d : %.3f ms  %.2f TFLOP/s\n", warps_per_sm, NACC, ms, flops / ms / 1e9);
}

int main() {
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 64 * 32);
    for (int w : {4, 8, 16, 32}) {
        run<1>(w, out);
        run<4>(w, out);
        run<8>(w, out);
        run<24>(w, out);
    }
    return 0;
}
