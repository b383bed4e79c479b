// Shared GEMM epilogue: bias -> activation -> LayerScale -> +row table -> +residuals
// -> store (fp32 / bf16 / ReLU copy / ConvTranspose scatter / fused output head).
// Processes 4 consecutive columns of one output row at a time (all channel
// counts on the path are multiples of 8, so 16-byte / 8-byte vector accesses
// are always aligned).
#pragma once
#include "gemm.h"

namespace dad {

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void load4(const void* base, long long off, int is_bf16, float (&r)[4]) {
    if (is_bf16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + off);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        r[0] = __low2float(a); r[1] = __high2float(a); r[2] = __low2float(b); r[3] = __high2float(b);
    } else {
        const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
        r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = f.w;
    }
}

__device__ __forceinline__ void store4_bf16(bf16* base, long long off, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(base + off) = u;
}

// grow: logical A-row index (token / input pixel); orow: output row index.
__device__ __forceinline__ void epilogue_store4(const Epilogue& e, int N, long long grow, long long orow,
                                                int col, float (&v)[4]) {
    int bcol = col;
    long long off;
    if (e.scat_k) {
        const int kk = col / e.scat_CoP;
        const int co = col - kk * e.scat_CoP;
        if (co >= e.scat_Co) return;
        const int ky = kk / e.scat_k, kx = kk - ky * e.scat_k;
        const int hw = e.scat_H * e.scat_W;
        const int b = static_cast<int>(grow / hw);
        const int rem = static_cast<int>(grow - static_cast<long long>(b) * hw);
        const int y = rem / e.scat_W, x = rem - y * e.scat_W;
        const long long opix =
            (static_cast<long long>(b) * (e.scat_k * e.scat_H) + (e.scat_k * y + ky)) * (e.scat_k * e.scat_W) +
            (e.scat_k * x + kx);
        off = opix * e.ldc + co;
        bcol = co;
    } else {
        off = orow * e.ldc + col;
    }
    if (e.bias) {
        const float4 b4 = *reinterpret_cast<const float4*>(e.bias + bcol);
        v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
    }
    if (e.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
    } else if (e.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (e.gamma) {
        const float4 g4 = *reinterpret_cast<const float4*>(e.gamma + bcol);
        v[0] *= g4.x; v[1] *= g4.y; v[2] *= g4.z; v[3] *= g4.w;
    }
    if (e.rowtab) {
        const long long tr = grow % e.rowtab_period;
        const float4 t4 = *reinterpret_cast<const float4*>(e.rowtab + tr * N + col);
        v[0] += t4.x; v[1] += t4.y; v[2] += t4.z; v[3] += t4.w;
    }
    if (e.res1) {
        float r[4];
        load4(e.res1, off, e.res1_bf16, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += r[i];
    }
    if (e.res2) {
        float r[4];
        load4(e.res2, off, e.res2_bf16, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += r[i];
    }
    if (e.out) {
        if (e.out_bf16) store4_bf16(reinterpret_cast<bf16*>(e.out), off, v);
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (e.out_relu) {
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = fmaxf(v[i], 0.f);
        if (e.out_bf16) store4_bf16(reinterpret_cast<bf16*>(e.out_relu), off, r);
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out_relu) + off) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

}  // namespace dad
