// Training-backward kernels of the fp32 verification engine.  The reference obtains these
// gradients from PyTorch autograd over depth_anything_v2/dpt.py:150-225 / dinov2.py:212-321 / util/blocks.py:29-148
// (tools/train_distillation.py:1556-1575); here each adjoint is written out: one strided FFMA GEMM covers every
// data-gradient / weight-gradient contraction (including the implicit-im2col weight gradient of the convolutions),
// the rest are HBM-bound row / pixel kernels.  Parameter gradients ACCUMULATE (atomicAdd) into caller-owned buffers.
#include <algorithm>

#include "backward.h"

namespace dad {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// activation tensors are fp32 (verification engine) or bf16 (tensor-core engine): run-time element type, uniform branch
__device__ __forceinline__ float ldf(const void* p, long long i, int bf) {
    return bf ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stf(void* p, long long i, int bf, float v) {
    if (bf) reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[i] = v;
}
__device__ __forceinline__ float4 ldf4(const void* p, long long i, int bf) {   // i: element index, multiple of 4
    if (bf) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p) + i);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    }
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
}
__device__ __forceinline__ void stf4(void* p, long long i, int bf, float4 v) {
    if (bf) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p) + i) = u;
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
    }
}

__global__ void __launch_bounds__(256) sgemm_kernel(const SGemm g, int ksplit, int kchunk) {
    __shared__ float sA[TK][TM + 4];
    __shared__ float sB[TK][TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int z = blockIdx.z / ksplit, ks = blockIdx.z - z * ksplit;
    const int z1 = z / g.nb2, z2 = z - z1 * g.nb2;
    const float* A = g.A + z1 * g.a1 + z2 * g.a2;
    const float* B = g.B + z1 * g.b1 + z2 * g.b2;
    float* C = g.C + z1 * g.c1 + z2 * g.c2;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int kbeg = ks * kchunk;
    const int kend = min(g.K, kbeg + kchunk);

    // loader maps: consecutive threads walk the contiguous axis of each operand
    const bool a_kfast = g.sak == 1;
    const int a_m = a_kfast ? (tid >> 2) : (tid & 63);
    const int a_k = a_kfast ? (tid & 3) * 4 : (tid >> 6) * 4;
    const bool b_kfast = !g.conv_taps && g.sbk == 1;
    const int b_n = b_kfast ? (tid >> 2) : (tid & 63);
    const int b_k = b_kfast ? (tid & 3) * 4 : (tid >> 6) * 4;
    const int am = m0 + a_m, bn = n0 + b_n;
    int cc = 0, cdy = 0, cdx = 0;
    if (g.conv_taps) {
        const int tap = bn / g.convC;
        cc = bn - tap * g.convC;
        if (g.conv_taps == 9) { cdy = tap / 3 - 1; cdx = tap - (tap / 3) * 3 - 1; }
    }
    const int hw = g.convHo * g.convWo;

    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + a_k + i;
            sA[a_k + i][a_m] = (am < g.M && k < kend) ? A[am * g.sam + k * g.sak] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + b_k + i;
            float v = 0.f;
            if (bn < g.N && k < kend) {
                if (!g.conv_taps) {
                    v = B[k * g.sbk + bn * g.sbn];
                } else {
                    const int b = k / hw;
                    const int r = k - b * hw;
                    const int oy = r / g.convWo, ox = r - oy * g.convWo;
                    const int y = oy * g.conv_stride + cdy, x = ox * g.conv_stride + cdx;
                    if (y >= 0 && y < g.convH && x >= 0 && x < g.convW)
                        v = B[((static_cast<long long>(b) * g.convH + y) * g.convW + x) * g.convC + cc];
                }
            }
            sB[b_k + i][b_n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            long long idx;
            if (g.cmap == 0) {
                idx = m * g.scm + n * g.scn;
            } else if (g.cmap == 1) {
                const int tap = n / g.convC, c = n - tap * g.convC;
                idx = (static_cast<long long>(m) * g.convC + c) * g.conv_taps + tap;
            } else {
                const int t = m / g.ct_CoP, co = m - t * g.ct_CoP;
                if (co >= g.ct_Co) continue;
                idx = (static_cast<long long>(n) * g.ct_Co + co) * g.ct_kk + t;
            }
            const float v = g.alpha * acc[i][j];
            if (!g.accumulate) C[idx] = v;
            else if (ksplit > 1) atomicAdd(C + idx, v);
            else C[idx] += v;
        }
    }
}

// ---------------------------------------------------------------- column sums (bias / LayerScale gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const void* X, int xbf, long long ldx, const void* Y, int ybf, long long ldy,
                                                     long long rows, int N, float* out, const float* scale, void* scaled_out) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (n < N) {
        const float sc = scale ? scale[n] : 1.f;
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + ry; r < rows; r += static_cast<long long>(gridDim.y) * 8) {
            const float x = ldf(X, r * ldx + n, xbf);
            s += Y ? x * ldf(Y, r * ldy + n, ybf) : x;
            if (scaled_out) stf(scaled_out, r * ldx + n, ybf, x * sc);
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && n < N && out) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        atomicAdd(out + n, t);
    }
}

// bf16 inputs, N % 8 == 0: 16-byte loads, 8 columns per thread (the scalar kernel above moves 64 bytes per warp instruction and
// is load-instruction bound: 2.2 ms per train step over 99 launches).  Same row partition and reduction order per column.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(h[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const bf16* X, long long ldx, const bf16* Y, long long ldy, long long rows, int N,
                                                            float* out, const float* scale, bf16* scaled_out) {
    __shared__ float red[8][32][9];
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 32 + lane) * 8;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (n0 < N) {
        float sc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sc[j] = scale ? scale[n0 + j] : 1.f;
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + ry; r < rows; r += static_cast<long long>(gridDim.y) * 8) {
            float xf[8];
            unpack8(*reinterpret_cast<const uint4*>(X + r * ldx + n0), xf);
            if (Y) {
                float yf[8];
                unpack8(*reinterpret_cast<const uint4*>(Y + r * ldy + n0), yf);
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += xf[j] * yf[j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += xf[j];
            }
            if (scaled_out) {
                uint4 o;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h[j].x = __float2bfloat16_rn(xf[2 * j] * sc[2 * j]);
                    h[j].y = __float2bfloat16_rn(xf[2 * j + 1] * sc[2 * j + 1]);
                }
                *reinterpret_cast<uint4*>(scaled_out + r * ldx + n0) = o;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ry][lane][j] = s[j];
    __syncthreads();
    if (ry == 0 && n0 < N && out) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][lane][j];
            atomicAdd(out + n0 + j, t);
        }
    }
}

// ---------------------------------------------------------------- LayerNorm backward
// One warp per row, persistent over rows.  The per-column dw / db partial sums live in PER-WARP shared-memory accumulators and
// the LayerNorm weight in shared memory (round 1 kept all three in registers: 164 registers at D = 768, one 8-warp block per
// SM, i.e. 24 KB of loads in flight per SM and 5x off the HBM time): ~80 registers, 2 - 3 blocks per SM.
template <int VPT>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const float* x, const float* w, const void* dy, int dybf, float* dx, float* dw,
                                                        float* db, long long rows, int D, int out_period, int in_period,
                                                        int in_offset, float eps) {
    extern __shared__ float4 ln_bwd_smem[];
    const int D4 = D / 4;
    float4* swt = ln_bwd_smem;                                        // [D4] weight
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4* accw = ln_bwd_smem + D4 + static_cast<size_t>(wid) * 2 * D4;   // this warp's [D4] dw and [D4] db partial sums
    float4* accb = accw + D4;
    for (int i = threadIdx.x; i < D4; i += 256) swt[i] = reinterpret_cast<const float4*>(w)[i];
    for (int i = lane; i < D4; i += 32) {
        accw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        accb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const long long gw = static_cast<long long>(blockIdx.x) * 8 + wid;
    const long long nw = static_cast<long long>(gridDim.x) * 8;
    for (long long r = gw; r < rows; r += nw) {
        const long long ir = (r / out_period) * in_period + in_offset + r % out_period;
        const float4* src = reinterpret_cast<const float4*>(x + ir * D);
        float4 v[VPT], gq[VPT];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = lane + j * 32;
            if (idx * 4 < D) {
                v[j] = src[idx];
                gq[j] = ldf4(dy, r * D + idx * 4, dybf);
                sum += v[j].x + v[j].y + v[j].z + v[j].w;
            } else {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                gq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / D;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = lane + j * 32;
            if (idx * 4 < D) {
                const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
                var += a * a + b * b + c * c + d * d;
            }
        }
        for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var / D + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = lane + j * 32;
            if (idx * 4 < D) {
                // v <- xhat; accumulate dw / db; gq <- dy * w
                v[j].x = (v[j].x - mean) * rstd; v[j].y = (v[j].y - mean) * rstd;
                v[j].z = (v[j].z - mean) * rstd; v[j].w = (v[j].w - mean) * rstd;
                float4 aw = accw[idx], ab = accb[idx];
                aw.x += gq[j].x * v[j].x; aw.y += gq[j].y * v[j].y; aw.z += gq[j].z * v[j].z; aw.w += gq[j].w * v[j].w;
                ab.x += gq[j].x; ab.y += gq[j].y; ab.z += gq[j].z; ab.w += gq[j].w;
                accw[idx] = aw;
                accb[idx] = ab;
                const float4 wv = swt[idx];
                gq[j].x *= wv.x; gq[j].y *= wv.y; gq[j].z *= wv.z; gq[j].w *= wv.w;
                s1 += gq[j].x + gq[j].y + gq[j].z + gq[j].w;
                s2 += gq[j].x * v[j].x + gq[j].y * v[j].y + gq[j].z * v[j].z + gq[j].w * v[j].w;
            }
        }
        for (int o = 16; o; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / D, m2 = s2 / D;
        float4* dst = reinterpret_cast<float4*>(dx + ir * D);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int idx = lane + j * 32;
            if (idx * 4 < D) {
                float4 o = dst[idx];
                o.x += rstd * (gq[j].x - m1 - v[j].x * m2);
                o.y += rstd * (gq[j].y - m1 - v[j].y * m2);
                o.z += rstd * (gq[j].z - m1 - v[j].z * m2);
                o.w += rstd * (gq[j].w - m1 - v[j].w * m2);
                dst[idx] = o;
            }
        }
    }
    __syncthreads();
    const float* accf = reinterpret_cast<const float*>(ln_bwd_smem + D4);   // [8 warps][2][D]
    for (int i = threadIdx.x; i < D; i += 256) {
        float tw = 0.f, tb = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            tw += accf[static_cast<size_t>(k) * 2 * D + i];
            tb += accf[static_cast<size_t>(k) * 2 * D + D + i];
        }
        if (dw) atomicAdd(dw + i, tw);
        if (db) atomicAdd(db + i, tb);
    }
}

// ---------------------------------------------------------------- small elementwise kernels
__global__ void __launch_bounds__(256) ls_residual_kernel(const float* xold, const void* y, int bf, const float* gamma, float* xnew,
                                                          long long n, int D) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) xnew[i] = xold[i] + ldf(y, i, bf) * gamma[i % D];   // x + ls(y), ls(y) = y * gamma (layer_scale.py:27-28)
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const void* pre, void* out, int bf, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) { const float x = ldf(pre, i, bf); stf(out, i, bf, 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f))); }
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const void* pre, const void* dout, void* dpre, int bf, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) {
        const float x = ldf(pre, i, bf);
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
        stf(dpre, i, bf, ldf(dout, i, bf) * (cdf + x * pdf));
    }
}

// SwiGLU gate (swiglu_ffn.py:30-34): g = silu(x1) * x2 with [x1 | x2] = x12 row; dx1 = dg * x2 * sig * (1 + x1 * (1 - sig)),
// dx2 = dg * silu(x1)
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const void* x12, const void* dg, void* dx12, int bf, long long rows, int Hd) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= rows * Hd) return;
    const long long r = i / Hd;
    const int c = static_cast<int>(i - r * Hd);
    const long long o = r * 2 * Hd + c;
    const float x1 = ldf(x12, o, bf), x2 = ldf(x12, o + Hd, bf), g = ldf(dg, i, bf);
    const float sig = 1.0f / (1.0f + expf(-x1));
    stf(dx12, o, bf, g * x2 * sig * (1.0f + x1 * (1.0f - sig)));
    stf(dx12, o + Hd, bf, g * x1 * sig);
}

// dst[r * ldd + c] = src[r * lds + c] for c < cols (same element type): column slices of the readout's concat gradient
__global__ void __launch_bounds__(256) copy_cols_kernel(const void* src, long long lds, void* dst, long long ldd, long long rows,
                                                        int cols, int bf) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    stf(dst, r * ldd + c, bf, ldf(src, r * lds + c, bf));
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const void* g, const void* y, const void* add, void* out, int bf, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) stf(out, i, bf, (add ? ldf(add, i, bf) : 0.f) + (ldf(y, i, bf) > 0.f ? ldf(g, i, bf) : 0.f));
}

__global__ void __launch_bounds__(256) add_kernel(void* dst, int bf, const float* src, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) stf(dst, i, bf, ldf(dst, i, bf) + src[i]);
}

__global__ void __launch_bounds__(256) convert_kernel(const void* src, int sbf, void* dst, int dbf, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) stf(dst, i, dbf, ldf(src, i, sbf));
}

// ---- bf16 tile transposes (64 x 64 tiles, 4-byte accesses on both sides) -------------------------------------------
// T64 stages a [64 rows][64 cols] tile; lanes read column PAIRS of one source row and write row PAIRS of one output row.
struct T64 {
    bf16 t[64][66];
};
__device__ __forceinline__ void t64_store_rows(T64& s, int row, int lane, __nv_bfloat162 v) {
    s.t[row][2 * lane] = __low2bfloat16(v);
    s.t[row][2 * lane + 1] = __high2bfloat16(v);
}
// out_row: pointer to output row c at element offset r0; writes elements r0 + 2*lane, +1 when < rlimit (rlimit even)
__device__ __forceinline__ void t64_write_col(const T64& s, int c, int lane, bf16* out_row, int rlimit_local) {
    if (2 * lane < rlimit_local) {
        __nv_bfloat162 v;
        v.x = s.t[2 * lane][c];
        v.y = s.t[2 * lane + 1][c];
        *reinterpret_cast<__nv_bfloat162*>(out_row + 2 * lane) = v;
    }
}

// out[z][c][r] = in[z][r * ld + c] for r < R (0 for R <= r < Rp); z = blockIdx.z with element strides zin / zout.
// R, Rp, ld even; C even.  K-major operands of the weight-gradient GEMMs and the attention-backward transposes.
__global__ void __launch_bounds__(256) transpose_pad_kernel(const bf16* in, long long ld, int R, int C, bf16* out, int Rp,
                                                            long long zin, long long zout) {
    __shared__ T64 s;
    in += blockIdx.z * zin;
    out += blockIdx.z * zout;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = w + i * 8, r = r0 + rr, c = c0 + 2 * lane;
        __nv_bfloat162 v = zero2;
        if (r < R && c < C) v = *reinterpret_cast<const __nv_bfloat162*>(in + r * ld + c);
        t64_store_rows(s, rr, lane, v);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int cc = w + i * 8, c = c0 + cc;
        if (c < C) t64_write_col(s, cc, lane, out + static_cast<long long>(c) * Rp + r0, Rp - r0);
    }
}

// out[(c * taps + tap)][p] = window(X)[p, tap, c], p = output pixel (b, oy, ox); 0 for P <= p < Pp and outside the image
// taps == 1 with (shift_y, shift_x) and a larger (Ho, Wo) frame writes the channel-major copy of X over ZERO-PADDED pixel
// space: frame pixel (oy, ox) holds X(oy - shift_y, ox - shift_x)
__global__ void __launch_bounds__(256) im2colT_kernel(const bf16* X, int H, int W, int Ci, int taps, int stride, int Ho, int Wo,
                                                      long long P, bf16* out, long long Pp, int shift_y, int shift_x) {
    __shared__ T64 s;
    const long long p0 = static_cast<long long>(blockIdx.x) * 64;
    const int c0 = blockIdx.y * 64, tap = blockIdx.z;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int dy = taps == 9 ? tap / 3 - 1 : -shift_y, dx = taps == 9 ? tap % 3 - 1 : -shift_x;
    const long long hw = static_cast<long long>(Ho) * Wo;
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = w + i * 8;
        const long long p = p0 + rr;
        const int c = c0 + 2 * lane;
        __nv_bfloat162 v = zero2;
        if (p < P && c < Ci) {
            const int b = static_cast<int>(p / hw);
            const int r = static_cast<int>(p - b * hw);
            const int oy = r / Wo, ox = r - oy * Wo;
            const int y = oy * stride + dy, x = ox * stride + dx;
            if (y >= 0 && y < H && x >= 0 && x < W)
                v = *reinterpret_cast<const __nv_bfloat162*>(X + ((static_cast<long long>(b) * H + y) * W + x) * Ci + c);
        }
        t64_store_rows(s, rr, lane, v);
    }
    __syncthreads();
    const long long rem = Pp - p0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int cc = w + i * 8, c = c0 + cc;
        if (c < Ci)
            t64_write_col(s, cc, lane, out + (static_cast<long long>(c) * taps + tap) * Pp + p0, rem > 64 ? 64 : static_cast<int>(rem));
    }
}

// w [N][K] fp32 -> out [K][Np] bf16 (zero for n >= N): the transposed weight of a data-gradient GEMM
__global__ void __launch_bounds__(256) pack_linear_T_kernel(const float* w, bf16* out, int N, int K, int Np) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty + i * 8, k = k0 + tx;
        tile[ty + i * 8][tx] = (n < N && k < K) ? w[static_cast<long long>(n) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty + i * 8, n = n0 + tx;
        if (k < K && n < Np) out[static_cast<long long>(k) * Np + n] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
    }
}

// P = softmax(S) row-wise, S fp32 [rows][ld] -> P bf16 [rows][ld] (columns >= T zeroed)
__global__ void __launch_bounds__(256) softmax_rows_bf16_kernel(const float* S, bf16* P, long long rows, int T, int ld) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* row = S + r * ld;
    bf16* out = P + r * ld;
    float m = -INFINITY;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, row[j]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int j = lane; j < T; j += 32) l += expf(row[j] - m);
    for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
    for (int j = lane; j < ld; j += 32) out[j] = __float2bfloat16_rn(j < T ? expf(row[j] - m) * inv : 0.f);
}

// dS = P * (dP - sum_j P dP) row-wise: P bf16, dP fp32 -> dS bf16 (columns >= T zeroed)
__global__ void __launch_bounds__(256) softmax_bwd_bf16_kernel(const bf16* P, const float* dP, bf16* dS, long long rows, int T, int ld) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const bf16* p = P + r * ld;
    const float* d = dP + r * ld;
    bf16* o = dS + r * ld;
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s = fmaf(__bfloat162float(p[j]), d[j], s);
    for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    for (int j = lane; j < ld; j += 32) o[j] = __float2bfloat16_rn(j < T ? __bfloat162float(p[j]) * (d[j] - s) : 0.f);
}

// per problem z: src fp32 [T][ld] -> dstN bf16 [T][ld] (same layout; columns >= T zeroed) and / or dstT bf16 [T][ld] (transposed)
__global__ void __launch_bounds__(256) cvt_tiles_kernel(const float* src, bf16* dstN, bf16* dstT, int T, int ld) {
    __shared__ float tile[32][33];
    const long long zo = static_cast<long long>(blockIdx.z) * T * ld;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8, c = c0 + tx;
        const float v = (r < T && c < T) ? src[zo + static_cast<long long>(r) * ld + c] : 0.f;
        tile[ty + i * 8][tx] = v;
        if (dstN && r < T && c < ld) dstN[zo + static_cast<long long>(r) * ld + c] = __float2bfloat16_rn(v);
    }
    if (!dstT) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, r = r0 + tx;   // dstT[c][r] = src[r][c]
        if (c < T && r < ld) dstT[zo + static_cast<long long>(c) * ld + r] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
    }
}

// per (image b, head h): dst[(b*heads + h)][d][t] = src[(b*T + t) * lds + h*64 + d], t < T (zero for T <= t < ld), d < 64
__global__ void __launch_bounds__(256) head_transpose_kernel(const bf16* src, long long lds, bf16* dst, int T, int heads, int ld) {
    __shared__ bf16 tile[32][33];
    const int z = blockIdx.z, b = z / heads, h = z - b * heads;
    const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + i * 8;
        tile[ty + i * 8][tx] = t < T ? src[(static_cast<long long>(b) * T + t) * lds + h * 64 + d0 + tx] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + ty + i * 8, t = t0 + tx;
        if (t < ld) dst[(static_cast<long long>(z) * 64 + d) * ld + t] = tile[tx][ty + i * 8];
    }
}

__global__ void __launch_bounds__(256) fill_kernel(float* p, float v, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------- attention probabilities
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* S, long long rows, int T, int ld) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float* row = S + r * ld;
    float m = -INFINITY;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, row[j]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = expf(row[j] - m); row[j] = e; l += e; }
    for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
    for (int j = lane; j < T; j += 32) row[j] *= inv;
}

__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* P, float* dP, long long rows, int T, int ld) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* p = P + r * ld;
    float* d = dP + r * ld;
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s = fmaf(p[j], d[j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int j = lane; j < T; j += 32) d[j] = p[j] * (d[j] - s);
}

// ---------------------------------------------------------------- bilinear adjoint (align_corners=True)
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const void* gout, int bf, float* gin, int B, int Hi, int Wi, int Ho, int Wo,
                                                           int C, float sh, float sw) {
    const int cv = C / 4;
    const long long total = static_cast<long long>(B) * Ho * Wo * cv;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % cv);
    long long p = i / cv;
    const int ox = static_cast<int>(p % Wo);
    p /= Wo;
    const int oy = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    const float fy = sh * oy, fx = sw * ox;   // index maths of bilinear_kernel (ATen area_pixel_compute_source_index)
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, hy = 1.f - ly, lx = fx - x0, hx = 1.f - lx;
    const float4 g = ldf4(gout, ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + c * 4, bf);
    float* base = gin + static_cast<long long>(b) * Hi * Wi * C + c * 4;
    const float wq[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
    const long long off[4] = {(static_cast<long long>(y0) * Wi + x0) * C, (static_cast<long long>(y0) * Wi + x1) * C,
                              (static_cast<long long>(y1) * Wi + x0) * C, (static_cast<long long>(y1) * Wi + x1) * C};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float* d = base + off[q];
        atomicAdd(d + 0, wq[q] * g.x); atomicAdd(d + 1, wq[q] * g.y);
        atomicAdd(d + 2, wq[q] * g.z); atomicAdd(d + 3, wq[q] * g.w);
    }
}

// Gather form of the same adjoint: one thread per INPUT pixel (and 4 channels) sums the output pixels whose bilinear footprint
// contains it - the rows oy with y0(oy) in {iy - 1, iy} and the columns likewise, found by re-evaluating the forward's own
// fp32 index expressions (so the weights are bit-identical to the scatter form's) - and writes the gradient once in the
// activation type: no zero fill, no atomics (deterministic), no fp32 staging + convert pass.
__device__ __forceinline__ int bil_first(int i, int n_out, float s) {   // smallest o with (int)(s * o) >= i - 1
    int e = s > 0.f ? static_cast<int>(static_cast<float>(i - 1) / s) : 0;
    e = e < 0 ? 0 : (e > n_out - 1 ? n_out - 1 : e);
    while (e > 0 && static_cast<int>(s * (e - 1)) >= i - 1) --e;
    while (e < n_out && static_cast<int>(s * e) < i - 1) ++e;
    return e;
}
__device__ __forceinline__ float bil_weight(int o, int i, int n_in, float s) {   // weight of input index i in output index o
    const float f = s * o;
    const int i0 = static_cast<int>(f);
    const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    const float l = f - i0, h = 1.f - l;
    return (i0 == i ? h : 0.f) + (i1 == i ? l : 0.f);
}
__global__ void __launch_bounds__(256) bilinear_bwd_gather_kernel(const void* gout, int bf, void* gin, int B, int Hi, int Wi, int Ho,
                                                                  int Wo, int C, float sh, float sw) {
    const int cv = C / 4;
    const long long total = static_cast<long long>(B) * Hi * Wi * cv;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % cv);
    long long p = i / cv;
    const int ix = static_cast<int>(p % Wi);
    p /= Wi;
    const int iy = static_cast<int>(p % Hi);
    const int b = static_cast<int>(p / Hi);
    const int ox_a = bil_first(ix, Wo, sw);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = bil_first(iy, Ho, sh); oy < Ho && static_cast<int>(sh * oy) <= iy; ++oy) {
        const float wy = bil_weight(oy, iy, Hi, sh);
        if (wy == 0.f) continue;
        const long long row = (static_cast<long long>(b) * Ho + oy) * Wo;
        for (int ox = ox_a; ox < Wo && static_cast<int>(sw * ox) <= ix; ++ox) {
            // the scatter form adds (hy * hx) g etc.: the product of the two 1-D weights, formed the same way
            const float w = wy * bil_weight(ox, ix, Wi, sw);
            if (w == 0.f) continue;
            const float4 g = ldf4(gout, (row + ox) * C + c * 4, bf);
            acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
        }
    }
    const long long o = ((static_cast<long long>(b) * Hi + iy) * Wi + ix) * C + c * 4;
    stf4(gin, o, bf, acc);
}

// ---------------------------------------------------------------- output head adjoint
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* gdepth, const float* depth, const void* t32, int bf, const float* w2,
                                                       void* dt32, float* dw2, float* db2, long long P) {
    __shared__ float sacc[33];
    if (threadIdx.x < 33) sacc[threadIdx.x] = 0.f;
    __syncthreads();
    float wv[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) wv[c] = w2[c];
    float aw[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) aw[c] = 0.f;
    float abias = 0.f;
    for (long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; p < P; p += static_cast<long long>(gridDim.x) * 256) {
        const float g = depth[p] > 0.f ? gdepth[p] : 0.f;   // relu(relu(z)): one mask
        abias += g;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = ldf4(t32, p * 32 + 4 * j, bf);
            aw[4 * j] += g * t.x; aw[4 * j + 1] += g * t.y; aw[4 * j + 2] += g * t.z; aw[4 * j + 3] += g * t.w;
            stf4(dt32, p * 32 + 4 * j, bf, make_float4(t.x > 0.f ? g * wv[4 * j] : 0.f, t.y > 0.f ? g * wv[4 * j + 1] : 0.f,
                                                       t.z > 0.f ? g * wv[4 * j + 2] : 0.f, t.w > 0.f ? g * wv[4 * j + 3] : 0.f));
        }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        float v = aw[c];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[c], v);
    }
    for (int o = 16; o; o >>= 1) abias += __shfl_xor_sync(0xffffffffu, abias, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[32], abias);
    __syncthreads();
    if (threadIdx.x < 32 && dw2) atomicAdd(dw2 + threadIdx.x, sacc[threadIdx.x]);
    if (threadIdx.x == 32 && db2) atomicAdd(db2, sacc[32]);
}

// training-forward output head on a saved (fp32 or bf16) ReLU'd 32-channel map: relu(dot(row, w) + b)
__global__ void __launch_bounds__(256) head1x1_any_kernel(const void* in, int bf, const float* w, const float* bias, float* out, long long P) {
    const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (p >= P) return;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = ldf4(in, p * 32 + 4 * j, bf);
        s = fmaf(v.x, w[4 * j], s); s = fmaf(v.y, w[4 * j + 1], s);
        s = fmaf(v.z, w[4 * j + 2], s); s = fmaf(v.w, w[4 * j + 3], s);
    }
    out[p] = fmaxf(s + __ldg(bias), 0.f);
}

// ---------------------------------------------------------------- ConvTranspose / strided-conv helpers
__global__ void __launch_bounds__(256) convT_gather_kernel(const void* dout, void* G, int bf, int B, int H, int W, int k, int Co,
                                                           int CoP) {
    const long long total = static_cast<long long>(B) * H * W * k * k * CoP;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int co = static_cast<int>(i % CoP);
    long long q = i / CoP;
    const int t = static_cast<int>(q % (k * k));
    q /= (k * k);
    const int x = static_cast<int>(q % W);
    q /= W;
    const int y = static_cast<int>(q % H);
    const int b = static_cast<int>(q / H);
    const int ky = t / k, kx = t - ky * k;
    float v = 0.f;
    if (co < Co) v = ldf(dout, ((static_cast<long long>(b) * (k * H) + k * y + ky) * (k * W) + k * x + kx) * Co + co, bf);
    stf(G, i, bf, v);
}

__global__ void __launch_bounds__(256) col2im_s2_kernel(const void* dcol, void* din, int bf, int B, int H, int W, int C, int Cp, int Ho,
                                                        int Wo) {
    const long long total = static_cast<long long>(B) * H * W * C;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    long long q = i / C;
    const int x = static_cast<int>(q % W);
    q /= W;
    const int y = static_cast<int>(q % H);
    const int b = static_cast<int>(q / H);
    float s = 0.f;
    for (int dy = 0; dy < 3; ++dy) {
        const int ty = y + 1 - dy;
        if (ty < 0 || (ty & 1)) continue;
        const int oy = ty >> 1;
        if (oy >= Ho) continue;
        for (int dx = 0; dx < 3; ++dx) {
            const int tx = x + 1 - dx;
            if (tx < 0 || (tx & 1)) continue;
            const int ox = tx >> 1;
            if (ox >= Wo) continue;
            s += ldf(dcol, ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * 9 * Cp + (dy * 3 + dx) * Cp + c, bf);
        }
    }
    stf(din, i, bf, s);
}

__global__ void __launch_bounds__(256) pack_conv_dgrad_kernel(const float* w, void* out, int bf, int Co, int Ci, int taps, int CoP) {
    const long long total = static_cast<long long>(Ci) * taps * CoP;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int co = static_cast<int>(i % CoP);
    const int t = static_cast<int>((i / CoP) % taps);
    const int ci = static_cast<int>(i / (static_cast<long long>(CoP) * taps));
    stf(out, i, bf, co < Co ? w[(static_cast<long long>(co) * Ci + ci) * taps + (taps - 1 - t)] : 0.f);
}

// ConvTranspose weight gradient: tmp [(t * CoP + co)][Ci] (GEMM output) -> dW[ci][co][t] += tmp
__global__ void __launch_bounds__(256) convT_wgrad_permute_kernel(const float* tmp, float* dW, int Ci, int Co, int CoP, int kk) {
    const long long total = static_cast<long long>(Ci) * Co * kk;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total) return;
    const int t = static_cast<int>(i % kk);
    const int co = static_cast<int>((i / kk) % Co);
    const int ci = static_cast<int>(i / (static_cast<long long>(kk) * Co));
    dW[i] += tmp[(static_cast<long long>(t) * CoP + co) * Ci + ci];
}

__global__ void __launch_bounds__(256) batch_sum_rows_kernel(const float* G, float* dtab, int B, int T, int D) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= static_cast<long long>(T) * D) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += G[static_cast<long long>(b) * T * D + i];
    dtab[i] = s;
}

// adjoint of pos_table_kernel (elementwise.cu): same bicubic taps, scattered
__device__ __forceinline__ float bcubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float bcubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void __launch_bounds__(256) pos_table_bwd_kernel(const float* dtab, float* dpos, float* dcls, float* dpbias, int D, int S,
                                                            int oh, int ow, int identity, float inv_sy, float inv_sx) {
    const int t = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += 256) {
        const float g = dtab[static_cast<long long>(t) * D + d];
        if (t == 0) {
            if (dcls) atomicAdd(dcls + d, g);
            if (dpos) atomicAdd(dpos + d, g);
            continue;
        }
        if (dpbias) atomicAdd(dpbias + d, g);
        if (!dpos) continue;
        const int p = t - 1;
        if (identity) {
            atomicAdd(dpos + static_cast<long long>(1 + p) * D + d, g);
        } else {
            const int oy = p / ow, ox = p - oy * ow;
            const float fy = inv_sy * (oy + 0.5f) - 0.5f, fx = inv_sx * (ox + 0.5f) - 0.5f;
            const int iy = static_cast<int>(floorf(fy)), ix = static_cast<int>(floorf(fx));
            const float ty = fy - iy, tx = fx - ix;
            const float wy[4] = {bcubic2(ty + 1.f), bcubic1(ty), bcubic1(1.f - ty), bcubic2(2.f - ty)};
            const float wx[4] = {bcubic2(tx + 1.f), bcubic1(tx), bcubic1(1.f - tx), bcubic2(2.f - tx)};
            for (int i = 0; i < 4; ++i) {
                const int yy = min(max(iy - 1 + i, 0), S - 1);
                for (int j = 0; j < 4; ++j) {
                    const int xx = min(max(ix - 1 + j, 0), S - 1);
                    atomicAdd(dpos + static_cast<long long>(1 + yy * S + xx) * D + d, wy[i] * wx[j] * g);
                }
            }
        }
    }
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>(cdivl(n, 256)); }

}  // namespace

int sgemm(const SGemm& g, cudaStream_t st) {
    DAD_REQUIRE(g.A && g.B && g.C && g.M > 0 && g.N > 0 && g.K > 0 && g.nb1 > 0 && g.nb2 > 0, "sgemm: bad operands");
    if (g.conv_taps) {
        DAD_REQUIRE((g.conv_taps == 1 || g.conv_taps == 9) && g.convC > 0 && g.N == g.conv_taps * g.convC && g.convHo > 0 &&
                        g.convWo > 0 && g.K % (g.convHo * g.convWo) == 0,
                    "sgemm: bad implicit-conv operand");
    }
    const int batch = g.nb1 * g.nb2;
    const long long tiles = static_cast<long long>(cdiv(g.M, TM)) * cdiv(g.N, TN) * batch;
    int ksplit = 1, kchunk = cdiv(g.K, TK) * TK;
    if (g.accumulate && g.K >= 4096 && tiles < 2LL * num_sms()) {
        ksplit = static_cast<int>(std::min<long long>(cdiv(g.K, 2048), cdivl(4LL * num_sms(), tiles)));
        kchunk = cdiv(cdiv(g.K, ksplit), TK) * TK;
        ksplit = cdiv(g.K, kchunk);
    }
    DAD_REQUIRE(static_cast<long long>(batch) * ksplit <= 65535 && cdiv(g.N, TN) <= 65535, "sgemm: grid too large");
    const dim3 grid(cdiv(g.M, TM), cdiv(g.N, TN), batch * ksplit);
    debug_label("sgemm");
    ProfScope prof(PROF_GEMM_SIMT, 2.0 * g.M * g.N * static_cast<double>(g.K) * batch, st);
    sgemm_kernel<<<grid, 256, 0, st>>>(g, ksplit, kchunk);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int colsum(const void* X, int xbf, long long ldx, const void* Y, int ybf, long long ldy, long long rows, int N, float* out,
           const float* scale, void* scaled_out, cudaStream_t st) {
    DAD_REQUIRE(X && rows > 0 && N > 0 && (out || scaled_out), "colsum: bad arguments");
    const dim3 grid(cdiv(N, 32), static_cast<unsigned>(std::min<long long>(cdivl(rows, 64), 1024)));
    debug_label("colsum");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * N * 4 * (1 + (Y ? 1 : 0) + (scaled_out ? 1 : 0)), st);
    if (xbf && (!Y || ybf) && N % 8 == 0 && ldx % 8 == 0 && (!Y || ldy % 8 == 0) && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
        (!Y || (reinterpret_cast<uintptr_t>(Y) & 15) == 0) && (!scaled_out || (reinterpret_cast<uintptr_t>(scaled_out) & 15) == 0)) {
        const dim3 grid8(cdiv(N, 256), grid.y);
        colsum_bf16x8_kernel<<<grid8, 256, 0, st>>>(reinterpret_cast<const bf16*>(X), ldx, reinterpret_cast<const bf16*>(Y), ldy, rows, N, out,
                                                    scale, reinterpret_cast<bf16*>(scaled_out));
        DAD_CHECK_LAUNCH();
        return DAD_OK;
    }
    colsum_kernel<<<grid, 256, 0, st>>>(X, xbf, ldx, Y, ybf, ldy, rows, N, out, scale, scaled_out);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int layernorm_bwd(const float* x, const float* w, const void* dy, int dybf, float* dx, float* dw, float* db, long long rows, int D,
                  int out_period, int in_period, int in_offset, float eps, cudaStream_t st) {
    DAD_REQUIRE(x && w && dy && dx && D % 4 == 0 && D <= 2048, "layernorm_bwd: bad arguments (D=%d)", D);
    const unsigned grid = static_cast<unsigned>(std::min<long long>(cdivl(rows, 8), 4LL * num_sms()));
    debug_label("layernorm_bwd");
    ProfScope prof(PROF_LN, static_cast<double>(rows) * D * 16, st);
    const int vpt = cdiv(D, 128);
    const size_t smem = static_cast<size_t>(17) * D * 4;   // weight + 8 warps x (dw, db) partial sums
#define LNB(V)                                                                                                                 \
    do {                                                                                                                       \
        static bool configured = false;                                                                                        \
        if (!configured) {                                                                                                     \
            DAD_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 17 * 2048 * 4)); \
            configured = true;                                                                                                 \
        }                                                                                                                      \
        ln_bwd_kernel<V><<<grid, 256, smem, st>>>(x, w, dy, dybf, dx, dw, db, rows, D, out_period, in_period, in_offset, eps);  \
    } while (0)
    if (vpt <= 3) LNB(3);
    else if (vpt <= 6) LNB(6);
    else if (vpt <= 8) LNB(8);
    else LNB(16);
#undef LNB
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int ls_residual(const float* xold, const void* y, int bf, const float* gamma, float* xnew, long long rows, int D, cudaStream_t st) {
    const long long n = rows * D;
    debug_label("ls_residual");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 12, st);
    ls_residual_kernel<<<blocks_for(n), 256, 0, st>>>(xold, y, bf, gamma, xnew, n, D);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int gelu_fwd(const void* pre, void* out, int bf, long long n, cudaStream_t st) {
    debug_label("gelu_fwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 8, st);
    gelu_fwd_kernel<<<blocks_for(n), 256, 0, st>>>(pre, out, bf, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int gelu_bwd(const void* pre, const void* dout, void* dpre, int bf, long long n, cudaStream_t st) {
    debug_label("gelu_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 12, st);
    gelu_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(pre, dout, dpre, bf, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int swiglu_bwd(const void* x12, const void* dg, void* dx12, int bf, long long rows, int Hd, cudaStream_t st) {
    debug_label("swiglu_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * Hd * (bf ? 10 : 20), st);
    swiglu_bwd_kernel<<<blocks_for(rows * Hd), 256, 0, st>>>(x12, dg, dx12, bf, rows, Hd);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int copy_cols(const void* src, long long lds, void* dst, long long ldd, long long rows, int cols, int bf, cudaStream_t st) {
    debug_label("copy_cols");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * cols * (bf ? 4 : 8), st);
    copy_cols_kernel<<<blocks_for(rows * cols), 256, 0, st>>>(src, lds, dst, ldd, rows, cols, bf);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int relu_bwd(const void* g, const void* y, const void* add, void* out, int bf, long long n, cudaStream_t st) {
    debug_label("relu_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * (add ? 16 : 12), st);
    relu_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(g, y, add, out, bf, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int add_inplace(void* dst, int bf, const float* src, long long n, cudaStream_t st) {
    debug_label("add_inplace");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 12, st);
    add_kernel<<<blocks_for(n), 256, 0, st>>>(dst, bf, src, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int convert(const void* src, int src_bf16, void* dst, int dst_bf16, long long n, cudaStream_t st) {
    debug_label("convert");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * ((src_bf16 ? 2 : 4) + (dst_bf16 ? 2 : 4)), st);
    convert_kernel<<<blocks_for(n), 256, 0, st>>>(src, src_bf16, dst, dst_bf16, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int fill_f32(float* p, float v, long long n, cudaStream_t st) {
    debug_label("fill_f32");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 4, st);
    fill_kernel<<<blocks_for(n), 256, 0, st>>>(p, v, n);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int softmax_rows_bf16(const float* S, void* P, long long rows, int T, int ld, cudaStream_t st) {
    debug_label("softmax_rows_bf16");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * ld * 6, st);
    softmax_rows_bf16_kernel<<<blocks_for(rows * 32), 256, 0, st>>>(S, reinterpret_cast<bf16*>(P), rows, T, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int softmax_bwd_bf16(const void* P, const float* dP, void* dS, long long rows, int T, int ld, cudaStream_t st) {
    debug_label("softmax_bwd_bf16");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * ld * 8, st);
    softmax_bwd_bf16_kernel<<<blocks_for(rows * 32), 256, 0, st>>>(reinterpret_cast<const bf16*>(P), dP, reinterpret_cast<bf16*>(dS), rows, T, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int cvt_tiles(const float* src, void* dstN, void* dstT, int Z, int T, int ld, cudaStream_t st) {
    DAD_REQUIRE(src && (dstN || dstT) && Z > 0 && Z <= 65535 && ld >= T, "cvt_tiles: bad arguments");
    const dim3 grid(cdiv(ld, 32), cdiv(ld, 32), Z);
    debug_label("cvt_tiles");
    ProfScope prof(PROF_ELEM, static_cast<double>(Z) * T * ld * (4 + (dstN ? 2 : 0) + (dstT ? 2 : 0)), st);
    cvt_tiles_kernel<<<grid, 256, 0, st>>>(src, reinterpret_cast<bf16*>(dstN), reinterpret_cast<bf16*>(dstT), T, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int head_transpose(const void* src, long long lds, void* dst, int B, int T, int heads, int ld, cudaStream_t st) {
    DAD_REQUIRE(src && dst && B * heads <= 65535 && ld >= T, "head_transpose: bad arguments");
    const dim3 grid(cdiv(ld, 32), 2, B * heads);
    debug_label("head_transpose");
    ProfScope prof(PROF_ELEM, static_cast<double>(B) * heads * 64 * (T + ld) * 2, st);
    head_transpose_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(src), lds, reinterpret_cast<bf16*>(dst), T, heads, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int transpose_pad(const void* in, long long ld, int R, int C, void* out, int Rp, cudaStream_t st) {
    return transpose_pad_batched(in, ld, R, C, out, Rp, 1, 0, 0, st);
}

int transpose_pad_batched(const void* in, long long ld, int R, int C, void* out, int Rp, int Z, long long zin, long long zout,
                          cudaStream_t st) {
    DAD_REQUIRE(in && out && R > 0 && C > 0 && Rp >= R && Z > 0 && Z <= 65535, "transpose_pad: bad arguments");
    DAD_REQUIRE(ld % 2 == 0 && C % 2 == 0 && Rp % 2 == 0 && zin % 2 == 0 && zout % 2 == 0, "transpose_pad: odd extents");
    const dim3 grid(cdiv(Rp, 64), cdiv(C, 64), Z);
    DAD_REQUIRE(grid.y <= 65535, "transpose_pad: too many columns");
    debug_label("transpose_pad");
    ProfScope prof(PROF_ELEM, (static_cast<double>(R) + Rp) * C * 2 * Z, st);
    transpose_pad_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(in), ld, R, C, reinterpret_cast<bf16*>(out), Rp, zin, zout);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

// dW[(co * Ci + ci) * taps + t] += S[co * (taps * ld) + t * ld + ci]: the (tap, ci)-ordered result of the shifted-view
// weight-gradient GEMM back into the reference's [Co, Ci, kh, kw] layout
__global__ void __launch_bounds__(256) wgrad_unshift_kernel(const float* S, float* dW, int Co, int Ci, int taps, int ld) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Co * Ci * taps) return;
    const int t = i % taps, ci = (i / taps) % Ci, co = i / (taps * Ci);
    dW[i] += S[static_cast<long long>(co) * taps * ld + t * ld + ci];
}

int wgrad_unshift(const float* S, float* dW, int Co, int Ci, int taps, int ld, cudaStream_t st) {
    DAD_REQUIRE(S && dW && Co > 0 && Ci > 0 && taps > 0 && ld >= Ci, "wgrad_unshift: bad arguments");
    debug_label("wgrad_unshift");
    wgrad_unshift_kernel<<<cdiv(Co * Ci * taps, 256), 256, 0, st>>>(S, dW, Co, Ci, taps, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int im2colT(const void* X, int B, int H, int W, int Ci, int taps, int stride, int Ho, int Wo, void* out, long long Pp,
            cudaStream_t st, int shift_y, int shift_x) {
    const long long P = static_cast<long long>(B) * Ho * Wo;
    DAD_REQUIRE(X && out && Pp >= P && (taps == 1 || taps == 9) && Ci % 2 == 0 && Pp % 2 == 0, "im2colT: bad arguments");
    const dim3 grid(static_cast<unsigned>(cdivl(Pp, 64)), cdiv(Ci, 64), taps);
    debug_label("im2colT");
    ProfScope prof(PROF_ELEM, static_cast<double>(Pp) * Ci * taps * 4, st);
    DAD_REQUIRE((shift_y == 0 && shift_x == 0) || taps == 1, "im2colT: the padded-space copy is a single-tap operation");
    im2colT_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(X), H, W, Ci, taps, stride, Ho, Wo, P,
                                         reinterpret_cast<bf16*>(out), Pp, shift_y, shift_x);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pack_linear_T(const float* w, void* out, int N, int K, int Np, cudaStream_t st) {
    const dim3 grid(cdiv(Np, 32), cdiv(K, 32));
    debug_label("pack_linear_T");
    ProfScope prof(PROF_ELEM, static_cast<double>(N) * K * 6, st);
    pack_linear_T_kernel<<<grid, 256, 0, st>>>(w, reinterpret_cast<bf16*>(out), N, K, Np);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int head1x1_any(const void* in, int bf, const float* w, const float* bias, float* out, long long P, cudaStream_t st) {
    debug_label("head1x1_any");
    ProfScope prof(PROF_ELEM, static_cast<double>(P) * (32 * (bf ? 2 : 4) + 4), st);
    head1x1_any_kernel<<<blocks_for(P), 256, 0, st>>>(in, bf, w, bias, out, P);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int softmax_rows(float* S, long long rows, int T, int ld, cudaStream_t st) {
    debug_label("softmax_rows");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * T * 8, st);
    softmax_rows_kernel<<<blocks_for(rows * 32), 256, 0, st>>>(S, rows, T, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int softmax_bwd_rows(const float* P, float* dP, long long rows, int T, int ld, cudaStream_t st) {
    debug_label("softmax_bwd_rows");
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * T * 12, st);
    softmax_bwd_rows_kernel<<<blocks_for(rows * 32), 256, 0, st>>>(P, dP, rows, T, ld);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int bilinear_bwd(const void* gout, int bf, float* gin, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st) {
    DAD_REQUIRE(C % 4 == 0, "bilinear_bwd: C=%d must be a multiple of 4", C);
    const float sh = Ho > 1 ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.f;
    const long long total = static_cast<long long>(B) * Ho * Wo * (C / 4);
    debug_label("bilinear_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(B) * C * 4 * (static_cast<double>(Hi) * Wi + static_cast<double>(Ho) * Wo), st);
    bilinear_bwd_kernel<<<blocks_for(total), 256, 0, st>>>(gout, bf, gin, B, Hi, Wi, Ho, Wo, C, sh, sw);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int bilinear_bwd_gather(const void* gout, int bf, void* gin, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st) {
    DAD_REQUIRE(C % 4 == 0, "bilinear_bwd: C=%d must be a multiple of 4", C);
    const float sh = Ho > 1 ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.f;
    const long long total = static_cast<long long>(B) * Hi * Wi * (C / 4);
    debug_label("bilinear_bwd_gather");
    ProfScope prof(PROF_ELEM, static_cast<double>(B) * C * (bf ? 2 : 4) * (static_cast<double>(Hi) * Wi + static_cast<double>(Ho) * Wo), st);
    bilinear_bwd_gather_kernel<<<blocks_for(total), 256, 0, st>>>(gout, bf, gin, B, Hi, Wi, Ho, Wo, C, sh, sw);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int head_bwd(const float* gdepth, const float* depth, const void* t32, int bf, const float* w2, void* dt32, float* dw2, float* db2,
             long long P, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(std::min<long long>(cdivl(P, 256), 8LL * num_sms()));
    debug_label("head_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(P) * (8 + 256), st);
    head_bwd_kernel<<<grid, 256, 0, st>>>(gdepth, depth, t32, bf, w2, dt32, dw2, db2, P);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int convT_wgrad_permute(const float* tmp, float* dW, int Ci, int Co, int CoP, int kk, cudaStream_t st) {
    const long long total = static_cast<long long>(Ci) * Co * kk;
    debug_label("convT_wgrad_permute");
    ProfScope prof(PROF_ELEM, static_cast<double>(total) * 12, st);
    convT_wgrad_permute_kernel<<<blocks_for(total), 256, 0, st>>>(tmp, dW, Ci, Co, CoP, kk);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int convT_gather(const void* dout, void* G, int bf, int B, int H, int W, int k, int Co, int CoP, cudaStream_t st) {
    const long long total = static_cast<long long>(B) * H * W * k * k * CoP;
    debug_label("convT_gather");
    ProfScope prof(PROF_ELEM, static_cast<double>(total) * 8, st);
    convT_gather_kernel<<<blocks_for(total), 256, 0, st>>>(dout, G, bf, B, H, W, k, Co, CoP);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int col2im_s2(const void* dcol, void* din, int bf, int B, int H, int W, int C, int Cp, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = static_cast<long long>(B) * H * W * C;
    debug_label("col2im_s2");
    ProfScope prof(PROF_ELEM, static_cast<double>(total) * 4 + static_cast<double>(B) * Ho * Wo * 9 * Cp * 4, st);
    col2im_s2_kernel<<<blocks_for(total), 256, 0, st>>>(dcol, din, bf, B, H, W, C, Cp, Ho, Wo);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pack_conv_dgrad(const float* w, void* out, int bf, int Co, int Ci, int taps, int CoP, cudaStream_t st) {
    const long long total = static_cast<long long>(Ci) * taps * CoP;
    debug_label("pack_conv_dgrad");
    ProfScope prof(PROF_ELEM, static_cast<double>(total) * 8, st);
    pack_conv_dgrad_kernel<<<blocks_for(total), 256, 0, st>>>(w, out, bf, Co, Ci, taps, CoP);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int batch_sum_rows(const float* G, float* dtab, int B, int T, int D, cudaStream_t st) {
    const long long n = static_cast<long long>(T) * D;
    debug_label("batch_sum_rows");
    ProfScope prof(PROF_ELEM, static_cast<double>(n) * 4 * (B + 1), st);
    batch_sum_rows_kernel<<<blocks_for(n), 256, 0, st>>>(G, dtab, B, T, D);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pos_table_bwd(const float* dtab, float* dpos, float* dcls, float* dpbias, int D, int H, int W, cudaStream_t st) {
    const int S = 37;
    const int ph = H / 14, pw = W / 14;
    const int identity = (ph * pw == S * S && H == W) ? 1 : 0;
    const double sfy = (static_cast<double>(ph) + 0.1) / S, sfx = (static_cast<double>(pw) + 0.1) / S;
    debug_label("pos_table_bwd");
    ProfScope prof(PROF_ELEM, static_cast<double>(1 + ph * pw) * D * 8, st);
    pos_table_bwd_kernel<<<1 + ph * pw, 256, 0, st>>>(dtab, dpos, dcls, dpbias, D, S, ph, pw, identity,
                                                      static_cast<float>(1.0 / sfy), static_cast<float>(1.0 / sfx));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
