// Multi-head self-attention, head_dim 64, non-causal, no mask (reference
// dinov2_layers/attention.py:49-62; q is pre-scaled: 64^-0.5 is folded into the packed qkv weights).
//
//   bf16 mode : attention_tc5.cu (tcgen05.mma, S / P / O / row sums in TMEM, TMA-fed)
//   fp32 mode : attention_f32_kernel below - verification mode, one query per thread, fp32 FFMA, expf.
//
// qkv layout: [B*N, 3*D] rows = tokens, columns = (3, heads, 64) as produced by the qkv GEMM.
#include <cstdlib>

#include "common.h"
#include "elementwise.h"

namespace dad {

namespace {

constexpr int HD = 64;

// ------------------------------------------------------------------ fp32 verification kernel
constexpr int FQ = 128, FK = 32;

__global__ void __launch_bounds__(FQ) attention_f32_kernel(const float* qkv, float* out, int N, int D) {
    __shared__ float sK[FK][HD];
    __shared__ float sV[FK][HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * FQ + threadIdx.x;
    const long long ld = 3LL * D;
    const float* base = qkv + static_cast<long long>(b) * N * ld + h * HD;
    float q[HD], acc[HD];
    const bool qok = qi < N;
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = qok ? base[static_cast<long long>(qi) * ld + d] : 0.f; acc[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < N; k0 += FK) {
        __syncthreads();
        for (int i = threadIdx.x; i < FK * HD; i += FQ) {
            const int r = i / HD, d = i - r * HD;
            const bool ok = (k0 + r) < N;
            sK[r][d] = ok ? base[static_cast<long long>(k0 + r) * ld + D + d] : 0.f;
            sV[r][d] = ok ? base[static_cast<long long>(k0 + r) * ld + 2 * D + d] : 0.f;
        }
        __syncthreads();
        const int kn = min(FK, N - k0);
        float sc[FK];
        float cmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < FK; ++r) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], sK[r][d], s);
            sc[r] = (r < kn) ? s : -INFINITY;
            cmax = fmaxf(cmax, sc[r]);
        }
        const float mn = fmaxf(m, cmax);
        const float scale = expf(m - mn);
        m = mn;
        l *= scale;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= scale;
#pragma unroll
        for (int r = 0; r < FK; ++r) {
            const float p = expf(sc[r] - m);
            l += p;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, sV[r][d], acc[d]);
        }
    }
    if (qok) {
        float* o = out + (static_cast<long long>(b) * N + qi) * D + h * HD;
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = acc[d] * inv;
    }
}

}  // namespace

int attention(const void* qkv, void* out, int is_bf16, int B, int N, int heads, cudaStream_t st) {
    DAD_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0, "attention: bad arguments");
    const int D = heads * HD;
    ProfScope prof(PROF_ATTN, 4.0 * B * static_cast<double>(N) * N * D, st);
    if (is_bf16) {
        // tcgen05 / TMEM kernels.  Default: attention_tc5.cu (round 2: packed FFMA2 + MUFU / FMA-pipe exponentials, row sums
        // on the tensor core, in-kernel exact fallback).  DAD_ATT_VARIANT=2 / 3 select the round-1 kernels (attention_tc.cu:
        // lazy rescale, 2 CTAs / SM; attention_tc3.cu: 4 serial CTAs / SM, exact per-tile maximum) for A/B measurements.
        const char* ev = getenv("DAD_ATT_VARIANT");   // read per call: tests switch it at run time
        const char* ep = getenv("DAD_ATT_POLY5");      // pairs of every 8 exponentiated on the FMA pipe (0, 2, 3, 4, 5)
        const int variant = ev ? atoi(ev) : 5;
        const int poly = ep ? atoi(ep) : 2;
        if (variant == 7 || variant == 8)   // 8: two single-thread MMA issuers
            DAD_TRY(attention_tc7(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, N, heads, poly, variant == 8, st));
        else if (variant == 6)
            DAD_TRY(attention_tc6(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, N, heads, poly, st));
        else if (variant == 3)
            DAD_TRY(attention_tc3(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, N, heads, st));
        else if (variant == 2)
            DAD_TRY(attention_tc(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, N, heads, st));
        else
            DAD_TRY(attention_tc5(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B, N, heads, poly, st));
        return DAD_OK;
    }
    {
        const dim3 grid(cdiv(N, FQ), heads, B);
        attention_f32_kernel<<<grid, FQ, 0, st>>>(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(out), N, D);
    }
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
