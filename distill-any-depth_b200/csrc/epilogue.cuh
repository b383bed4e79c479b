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

// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) with MUFU reciprocal / exp2 in flush-to-zero mode (no range
// fix-up code): used when the result is rounded to bf16 anyway (bf16 ulp 4e-3 relative); fp32 outputs use erff.
// 15 instructions per element: 5 FMUL, 7 FFMA, 2 MUFU, 1 LOP3.
__device__ __forceinline__ float gelu_fast(float x) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752440f;                 // |x| / sqrt(2)
    const float zl = ax * 0.84932180028801904272f;                // |x| * sqrt(log2(e) / 2): zl^2 = z^2 * log2(e)
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(zl * zl)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erf_abs = fmaf(-(poly * t), e, 1.0f);
    const float hx = 0.5f * x;
    return fmaf(hx, copysignf(erf_abs, x), hx);
}

// The epilogue of one (row, 4 columns) group is split in two so that callers can issue the global
// loads of several groups (residuals / row table) before any dependent math or store:
//   epilogue_prefetch : output offset + residual / table loads
//   epilogue_finish   : bias -> activation -> LayerScale -> + table -> + residuals -> stores
// KIND specialises the hot encoder epilogues at compile time (small code, no dead branches);
// EK_GENERIC evaluates every option at run time.
enum { EK_GENERIC = 0, EK_BIAS_BF16 = 1, EK_GELU_BF16 = 2, EK_RES_F32 = 3,
       EK_GENERIC_NOGELU = 4 /* run-time generic without the GELU code (tensor-core engine: fc1 is EK_GELU_BF16) */ };

// which specialised kind (if any) computes exactly what `e` asks for
inline int epilogue_kind(const Epilogue& e) {
    if (e.scat_k || e.rowtab || e.res2 || e.out_relu || e.head_out || !e.out || !e.bias) return EK_GENERIC;
    if (e.out_bf16 && !e.gamma && !e.res1 && e.act == ACT_NONE) return EK_BIAS_BF16;
    if (e.out_bf16 && !e.gamma && !e.res1 && e.act == ACT_GELU) return EK_GELU_BF16;
    if (!e.out_bf16 && e.gamma && e.res1 && !e.res1_bf16 && e.act == ACT_NONE && e.res1 == e.out) return EK_RES_F32;  // x += g*(acc+b)
    return EK_GENERIC;
}

struct EpiPre {
    long long off;
    int bcol;
    bool skip;
    float r1[4], r2[4], tab[4];
};

// grow: logical A-row index (token / input pixel); orow: output row index.
// ConvTranspose scatter: output pixel index of input row `grow` for the (ky, kx) = (0, 0) sub-pixel
__device__ __forceinline__ long long epilogue_scatter_base(const Epilogue& e, long long grow) {
    const int hw = e.scat_H * e.scat_W;
    const int b = static_cast<int>(grow / hw);
    const int rem = static_cast<int>(grow - static_cast<long long>(b) * hw);
    const int y = rem / e.scat_W, x = rem - y * e.scat_W;
    return (static_cast<long long>(b) * (e.scat_k * e.scat_H) + e.scat_k * y) * (e.scat_k * e.scat_W) + e.scat_k * x;
}

// `sbase`: epilogue_scatter_base(grow) when e.scat_k != 0 (callers hoist it: it only depends on the row)
template <int KIND = EK_GENERIC>
__device__ __forceinline__ void epilogue_prefetch(const Epilogue& e, int N, long long grow, long long orow, int col,
                                                  EpiPre& p, long long sbase = 0) {
    constexpr bool G = KIND == EK_GENERIC || KIND == EK_GENERIC_NOGELU;
    p.bcol = col;
    p.skip = false;
    if (G && e.scat_k) {
        const int kk = col / e.scat_CoP;
        const int co = col - kk * e.scat_CoP;
        if (co >= e.scat_Co) { p.skip = true; return; }
        const int ky = kk / e.scat_k, kx = kk - ky * e.scat_k;
        p.off = (sbase + static_cast<long long>(ky) * (e.scat_k * e.scat_W) + kx) * e.ldc + co;
        p.bcol = co;
    } else {
        p.off = orow * e.ldc + col;
    }
    if (G && e.rowtab) {
        const long long tr = grow % e.rowtab_period;
        const float4 t4 = *reinterpret_cast<const float4*>(e.rowtab + tr * N + col);
        p.tab[0] = t4.x; p.tab[1] = t4.y; p.tab[2] = t4.z; p.tab[3] = t4.w;
    }
    if (G ? (e.res1 != nullptr) : (KIND == EK_RES_F32)) load4(e.res1, p.off, G ? e.res1_bf16 : 0, p.r1);
    if (G && e.res2) load4(e.res2, p.off, e.res2_bf16, p.r2);
}

// per-column vectors of a 4-column group (bias, LayerScale gamma): loaded once per group of rows
struct EpiCols {
    float4 b4, g4;
};
template <int KIND = EK_GENERIC>
__device__ __forceinline__ void epilogue_load_cols(const Epilogue& e, int col, EpiCols& c) {
    constexpr bool G = KIND == EK_GENERIC || KIND == EK_GENERIC_NOGELU;
    int bcol = col;
    c.b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    c.g4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (G && e.scat_k) {
        bcol = col % e.scat_CoP;
        if (bcol >= e.scat_Co) return;  // padded output channel: never stored
    }
    if (G ? (e.bias != nullptr) : true) c.b4 = *reinterpret_cast<const float4*>(e.bias + bcol);
    if (G ? (e.gamma != nullptr) : (KIND == EK_RES_F32)) c.g4 = *reinterpret_cast<const float4*>(e.gamma + bcol);
}

template <int KIND = EK_GENERIC>
__device__ __forceinline__ void epilogue_finish(const Epilogue& e, const EpiPre& p, const EpiCols& c, float (&v)[4]) {
    constexpr bool G = KIND == EK_GENERIC || KIND == EK_GENERIC_NOGELU;
    if (G && p.skip) return;
    if (G ? (e.bias != nullptr) : true) {
        v[0] += c.b4.x; v[1] += c.b4.y; v[2] += c.b4.z; v[3] += c.b4.w;
    }
    const int act = G ? e.act : (KIND == EK_GELU_BF16 ? ACT_GELU : ACT_NONE);
    const bool obf = G ? (e.out_bf16 != 0) : (KIND != EK_RES_F32);
    if (KIND != EK_GENERIC_NOGELU && act == ACT_GELU) {
        if (obf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_fast(v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
        }
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (G ? (e.gamma != nullptr) : (KIND == EK_RES_F32)) {
        v[0] *= c.g4.x; v[1] *= c.g4.y; v[2] *= c.g4.z; v[3] *= c.g4.w;
    }
    if (G && e.rowtab) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += p.tab[i];
    }
    if (G ? (e.res1 != nullptr) : (KIND == EK_RES_F32)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += p.r1[i];
    }
    if (G && e.res2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += p.r2[i];
    }
    if (G ? (e.out != nullptr) : true) {
        if (obf) store4_bf16(reinterpret_cast<bf16*>(e.out), p.off, v);
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + p.off) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (G && e.out_relu) {
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = fmaxf(v[i], 0.f);
        if (obf) store4_bf16(reinterpret_cast<bf16*>(e.out_relu), p.off, r);
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out_relu) + p.off) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

__device__ __forceinline__ void epilogue_store4(const Epilogue& e, int N, long long grow, long long orow, int col,
                                                float (&v)[4]) {
    EpiPre p;
    EpiCols c;
    epilogue_load_cols<EK_GENERIC>(e, col, c);
    epilogue_prefetch<EK_GENERIC>(e, N, grow, orow, col, p, e.scat_k ? epilogue_scatter_base(e, grow) : 0);
    epilogue_finish<EK_GENERIC>(e, p, c, v);
}

}  // namespace dad
