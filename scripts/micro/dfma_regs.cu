// dev microbenchmark: DFMA throughput with register operands shaped like the K_e inner loop:
// acc[i][b*3+j] += bc[i][s] * g[b][d]  (3 register operands per DFMA), 72 accumulators per thread.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128, 2) k(double* out, const double* in, int iters) {
    double acc[3][24];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 24; ++j) acc[i][j] = 0.0;
    double bc[3][6];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 6; ++r) bc[i][r] = in[(threadIdx.x + i * 6 + r) & 63];
    __shared__ double sg[8 * 3 * 8];
    for (int i = threadIdx.x; i < 192; i += blockDim.x) sg[i] = in[i & 63] * 1e-3;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const double* gp = sg + (it & 7) * 24;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double gx, gy, gz;
            if (MODE == 0) {  // gradients from shared memory (broadcast), like the kernel
                gx = gp[b * 3]; gy = gp[b * 3 + 1]; gz = gp[b * 3 + 2];
            } else {  // gradients derived from registers only
                gx = bc[0][b % 6]; gy = bc[1][(b + 1) % 6]; gz = bc[2][(b + 2) % 6];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s0 = acc[i][b * 3 + 0], s1 = acc[i][b * 3 + 1], s2 = acc[i][b * 3 + 2];
                s0 += bc[i][0] * gx; s0 += bc[i][4] * gz; s0 += bc[i][5] * gy;
                s1 += bc[i][1] * gy; s1 += bc[i][3] * gz; s1 += bc[i][5] * gx;
                s2 += bc[i][2] * gz; s2 += bc[i][3] * gy; s2 += bc[i][4] * gx;
                acc[i][b * 3 + 0] = s0; acc[i][b * 3 + 1] = s1; acc[i][b * 3 + 2] = s2;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 24; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(int blocks_per_sm, double* out, double* in) {
    const int iters = 4000, threads = 128, blocks = 148 * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, in, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d  CTAs/SM %d (%d warps/SM): %.3f ms  %.2f TFLOP/s\n", MODE, blocks_per_sm, blocks_per_sm * 4, ms,
           2.0 * blocks * threads * 216.0 * iters / ms / 1e9);
}

int main() {
    double *out, *in;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 128);
    cudaMalloc(&in, sizeof(double) * 64);
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + i * 1e-3;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int b : {1, 2}) { run<0>(b, out, in); run<1>(b, out, in); }
    return 0;
}
