// dev microbenchmark: FP64 tensor-core (mma.sync m8n8k4 / m16n8k8 / m16n8k16 .f64) throughput of one B200, alone and mixed with DFMA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_peak.cu -o dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC, int NFMA>
__global__ void k_884(double* out, double a, double b, int iters) {
    double c[NACC][2];
    double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < NFMA; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NFMA; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int K>
__global__ void k_16(double* out, double a, double b, int iters) {
    double c[NACC][4];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (K == 8) dmma1688(c[i], av, bv); else dmma16816(c[i], av, bv);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
void timeit(const char* name, int warps_per_sm, double fma_per_thread_iter, double dfma_per_thread_iter, F launch) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double threads = 148.0 * warps_per_sm * 32;
    printf("%-28s warps/SM %2d : %.3f ms  tensor %.2f TFLOP/s  + dfma %.2f TFLOP/s\n", name, warps_per_sm, ms,
           2.0 * threads * fma_per_thread_iter * iters / ms / 1e9, 2.0 * threads * dfma_per_thread_iter * iters / ms / 1e9);
}

int main() {
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 64 * 32);
    for (int w : {4, 8, 16, 32}) {
        const int threads = 128, blocks = 148 * w * 32 / threads;
        // m8n8k4: 256 FMA per warp instruction = 8 per thread
        timeit("m8n8k4 x1 chain", w, 4 * 1 * 8.0, 0, [&](int it) { k_884<1, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m8n8k4 x4 chains", w, 4 * 4 * 8.0, 0, [&](int it) { k_884<4, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m8n8k4 x9 chains", w, 4 * 9 * 8.0, 0, [&](int it) { k_884<9, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m8n8k4 x9 + 9 dfma", w, 4 * 9 * 8.0, 4 * 9.0, [&](int it) { k_884<9, 9><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m8n8k4 x4 + 16 dfma", w, 4 * 4 * 8.0, 4 * 16.0, [&](int it) { k_884<4, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m16n8k8 x4 chains", w, 4 * 4 * 32.0, 0, [&](int it) { k_16<4, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
        timeit("m16n8k16 x4 chains", w, 4 * 4 * 64.0, 0, [&](int it) { k_16<4, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9, it); });
    }
    return 0;
}
