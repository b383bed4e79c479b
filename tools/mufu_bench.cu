// Micro-benchmark: MUFU.EX2 throughput per SM as a function of resident warps (build: nvcc -arch=sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 20000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<148, threads>>>(out, 100, 0.5f);
        cudaEventRecord(e0);
        k<<<148, threads>>>(out, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        double ops = double(threads) * iters * 8;   // per SM
        printf("threads/SM %4d: %.3f ms  %.2f ex2/ns/SM  (~%.1f per clk at %d MHz nominal)\n", threads, ms, ops / (ms * 1e6),
               ops / (ms * 1e6) / (clk / 1e6), clk / 1000);
    }
    return 0;
}
