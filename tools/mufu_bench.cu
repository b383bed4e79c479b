// Micro-benchmark: exp2 throughput per SM for the softmax inner loop (build: nvcc -arch=sm_100a).
//   f32     ex2.approx.ftz.f32            (MUFU, one result per lane-op)
//   f16x2   ex2.approx.f16x2              (MUFU, two results per lane-op if the pipe is packed)
//   bf16x2  ex2.approx.ftz.bf16x2
//   poly    Cody-Waite + degree-3 polynomial on the FMA / ALU pipes (no MUFU)
//   mix     3 of 4 elements on MUFU f32, 1 of 4 by polynomial
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_poly(float x) {
    // 2^x = 2^n * p(r), n = round(x), r = x - n in [-0.5, 0.5]
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: integer part lands in the low mantissa bits
    const float r = x - (t - 12582912.0f);
    float p = fmaf(r, 0.0555041f, 0.2402265f);
    p = fmaf(p, r, 0.6931472f);
    p = fmaf(p, r, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            } else if (MODE == 1) {
                uint32_t u = __float_as_uint(a[i]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
                a[i] = __uint_as_float(u);
            } else if (MODE == 2) {
                uint32_t u = __float_as_uint(a[i]);
                asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u));
                a[i] = __uint_as_float(u);
            } else if (MODE == 3) {
                a[i] = ex2_poly(a[i]) - 1.0f;
            } else if (MODE == 4) {
                if ((i & 3) == 3) a[i] = ex2_poly(a[i]) - 1.0f;
                else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            } else if (MODE == 5) {  // F2FP only: pack pairs to bf16x2 (round to nearest)
                if ((i & 1) == 0) {
                    uint32_t u;
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i + 1]), "f"(a[i]));
                    a[i] = __uint_as_float(u & 0x3f803f80u) + 1.0f;
                }
            } else if (MODE == 6) {  // the softmax mix: one MUFU per element + one F2FP per pair
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if ((i & 1) == 1) {
                    uint32_t u;
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(a[i - 1]));
                    a[i - 1] = __uint_as_float(u & 0x3f803f80u);
                }
            } else {                 // MUFU + truncating pack on the ALU (PRMT) + masked copy for the row sum (LOP3)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if ((i & 1) == 1) {
                    const uint32_t u = __byte_perm(__float_as_uint(a[i - 1]), __float_as_uint(a[i]), 0x7632);
                    a[i - 1] = __uint_as_float(u & 0x3f803f80u);
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int per_op) {
    for (int threads : {256, 512}) {
        const int iters = 20000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, threads>>>(out, 100, 0.5f);
        cudaEventRecord(e0);
        k<MODE><<<148, threads>>>(out, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = double(threads) * iters * 8 * per_op;   // exp2 results per SM
        printf("%-7s threads/SM %4d: %.3f ms  %.2f exp2/ns/SM\n", name, threads, ms, ops / (ms * 1e6));
    }
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    run<0>("f32", out, 1);
    run<1>("f16x2", out, 2);
    run<2>("bf16x2", out, 2);
    run<3>("poly", out, 1);
    run<4>("mix3:1", out, 1);
    run<5>("f2fp", out, 1);       // per-element rate (one cvt packs two)
    run<6>("ex2+f2fp", out, 1);
    run<7>("ex2+prmt", out, 1);
    return 0;
}
