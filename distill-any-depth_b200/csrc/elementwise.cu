// HBM-bound helper kernels of the forward path: patch im2col, LayerNorm, bilinear resampling
// (align_corners=True, NHWC), strided-conv im2col, 1x1 output head, bicubic positional-embedding
// resize and the weight-packing kernels.  Activations are bf16 (tensor-core mode) or fp32
// (verification mode); every kernel is templated on that type and moves 16 bytes per thread where
// the layout allows.
#include "common.h"
#include "elementwise.h"

namespace dad {

namespace {

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }

__device__ __forceinline__ void store4(float* dst, const float (&o)[4]) {
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void store4(bf16* dst, const float (&o)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst) = u;
}

// ---------------------------------------------------------------- patch im2col (K1)
// x [B,3,H,W] fp32 -> A [B*(1+ph*pw), Kp]; row b*T is the (zero) cls slot, k = c*196 + ky*14 + kx
template <typename T>
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* x, T* A, int B, int H, int W, int Kp) {
    const int ph = H / 14, pw = W / 14, Tn = 1 + ph * pw;
    const long long row = blockIdx.x;
    const int b = static_cast<int>(row / Tn), t = static_cast<int>(row - static_cast<long long>(b) * Tn);
    T* out = A + row * Kp;
    if (t == 0) {
        for (int k = threadIdx.x; k < Kp; k += 256) out[k] = from_f<T>(0.f);
        return;
    }
    const int py = (t - 1) / pw, px = (t - 1) - py * pw;
    for (int k = threadIdx.x; k < Kp; k += 256) {
        float v = 0.f;
        if (k < 588) {
            const int c = k / 196, r = k - c * 196, ky = r / 14, kx = r - ky * 14;
            v = x[((static_cast<long long>(b) * 3 + c) * H + py * 14 + ky) * W + px * 14 + kx];
        }
        out[k] = from_f<T>(v);
    }
}

// ---------------------------------------------------------------- LayerNorm (K3, K9)
// one warp per output row; in row = (r / out_period) * in_period + in_offset + r % out_period
// (Round 2 measured two persistent streaming restructurings of this kernel on the headline step, bit-identical results:
// a producer thread filling an mbarrier ring with 64 KB cp.async.bulk copies - 0.47 of the HBM peak - and per-warp rings
// of 16-byte cp.async copies two rows ahead, 2 CTAs / SM - 0.62; this one-shot kernel: 0.70 - 0.72.  Kept as is.)
template <typename T, int VPT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* in, const float* w, const float* bvec, T* out,
                                                        float* out_f32, long long rows, int D, int out_period,
                                                        int in_period, int in_offset, float eps) {
    pdl_wait();
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const long long ir = (r / out_period) * in_period + in_offset + r % out_period;
    const float4* src = reinterpret_cast<const float4*>(in + ir * D);
    float4 v[VPT];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int idx = lane + j * 32;
        if (idx * 4 < D) {
            v[j] = src[idx];
            sum += v[j].x + v[j].y + v[j].z + v[j].w;
        } else {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    pdl_launch_dependents();  // row is in registers: only the stores remain
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int idx = lane + j * 32;
        if (idx * 4 < D) {
            const float4 g = reinterpret_cast<const float4*>(w)[idx];
            const float4 be = reinterpret_cast<const float4*>(bvec)[idx];
            float o4[4] = {(v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y,
                           (v[j].z - mean) * rstd * g.z + be.z, (v[j].w - mean) * rstd * g.w + be.w};
            if (out) store4(out + r * D + idx * 4, o4);
            if (out_f32) reinterpret_cast<float4*>(out_f32 + r * D)[idx] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

// ---------------------------------------------------------------- bilinear, align_corners=True (K14/K16)
// NHWC, VEC channels per thread.  Index maths as ATen's area_pixel_compute_source_index.
constexpr int BIL_ROWS = 16;  // output rows per thread: a sliding window of two horizontally blended source rows serves them

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int Hi, int Wi, int Ho,
                                                       int Wo, int C, float sh, float sw) {
    // grid: (x-chunks, output row groups, image); thread -> (output column, VEC-channel group) x BIL_ROWS rows.
    // out = hy * (hx * a + lx * b) + ly * (hx * c + lx * d) in ATen's operation order.  The horizontal blends
    // (hx * a + lx * b) of source rows y0 and y1 are kept in registers: when the next output row needs the same source
    // rows (always, when upsampling) nothing is reloaded, when it moves down by one the lower row is reused - 0.5-0.6
    // source rows are fetched and blended per output row instead of 2 (the kernel was issue-bound, not HBM-bound).
    pdl_wait();
    const int cv = C / VEC;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Wo * cv) return;
    const int ox = i / cv, c = i - ox * cv;
    const int b = blockIdx.z;
    const float fx = sw * ox;
    const int x0 = static_cast<int>(fx);
    const int x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float lx = fx - x0, hx = 1.f - lx;
    const T* base = in + static_cast<long long>(b) * Hi * Wi * C + c * VEC;
    T* obase = out + static_cast<long long>(b) * Ho * Wo * C + static_cast<long long>(ox) * C + c * VEC;
    struct alignas(16) Pack { T v[VEC]; };
    const int rowpitch = Wi * C;
    const int o0 = x0 * C, o1 = x1 * C;
    auto hblend = [&](int y, float (&h)[VEC]) {   // h = hx * in[y, x0] + lx * in[y, x1]
        const T* r = base + static_cast<long long>(y) * rowpitch;
        const Pack a = *reinterpret_cast<const Pack*>(r + o0), bq = *reinterpret_cast<const Pack*>(r + o1);
#pragma unroll
        for (int j = 0; j < VEC; ++j) h[j] = hx * to_f(a.v[j]) + lx * to_f(bq.v[j]);
    };
    float h0[VEC], h1[VEC];
    int cy0 = -1, cy1 = -1;   // source rows currently held in h0 / h1
    const int oy0 = blockIdx.y * BIL_ROWS;
    for (int r = 0; r < BIL_ROWS; ++r) {
        const int oy = oy0 + r;
        if (oy >= Ho) break;
        const float fy = sh * oy;
        const int y0 = static_cast<int>(fy);
        const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0);
        const float ly = fy - y0, hy = 1.f - ly;
        if (y0 != cy0) {          // (block-uniform branches: every thread of the block has the same oy)
            if (y0 == cy1) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) h0[j] = h1[j];
            } else {
                hblend(y0, h0);
            }
            cy0 = y0;
        }
        if (y1 != cy1) {
            if (y1 == y0) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) h1[j] = h0[j];
            } else {
                hblend(y1, h1);
            }
            cy1 = y1;
        }
        Pack o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) o.v[j] = from_f<T>(hy * h0[j] + ly * h1[j]);
        *reinterpret_cast<Pack*>(obase + static_cast<long long>(oy) * Wo * C) = o;
    }
}

// ---------------------------------------------------------------- 3x3 stride-2 pad-1 im2col (K12)
// in NHWC [B,H,W,C] -> A [B*Ho*Wo, 9*Cp], k = tap*Cp + c
template <typename T>
__global__ void __launch_bounds__(256) im2col_s2_kernel(const T* in, T* A, int B, int H, int W, int C, int Cp, int Ho,
                                                        int Wo) {
    const long long row = blockIdx.x;
    const int b = static_cast<int>(row / (Ho * Wo));
    const int r = static_cast<int>(row - static_cast<long long>(b) * Ho * Wo);
    const int oy = r / Wo, ox = r - oy * Wo;
    T* out = A + row * 9 * Cp;
    for (int k = threadIdx.x; k < 9 * Cp; k += 256) {
        const int tap = k / Cp, c = k - tap * Cp;
        const int dy = tap / 3, dx = tap - dy * 3;
        const int y = oy * 2 + dy - 1, x = ox * 2 + dx - 1;
        T v = from_f<T>(0.f);
        if (c < C && y >= 0 && y < H && x >= 0 && x < W) v = in[((static_cast<long long>(b) * H + y) * W + x) * C + c];
        out[k] = v;
    }
}

// ---------------------------------------------------------------- fp32-mode output head: relu(dot(row[32], w) + b)
__global__ void __launch_bounds__(256) head1x1_kernel(const float* in, const float* w, const float* bias, float* out, long long P) {
    const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (p >= P) return;
    const float4* src = reinterpret_cast<const float4*>(in + p * 32);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = src[j];
        s = fmaf(v.x, w[4 * j], s); s = fmaf(v.y, w[4 * j + 1], s);
        s = fmaf(v.z, w[4 * j + 2], s); s = fmaf(v.w, w[4 * j + 3], s);
    }
    out[p] = fmaxf(s + __ldg(bias), 0.f);
}

// ---------------------------------------------------------------- positional-embedding table (K2)
// tab[0] = cls + pos[0]; tab[1 + p] = pos_resized[p] + patch_bias.  Bicubic (A = -0.75,
// align_corners=False, coordinate scale = 1 / scale_factor) as F.interpolate(scale_factor=...)
// in dinov2.py:198-205; identity when the grid is the stored 37 x 37 and the image is square.
__device__ __forceinline__ float cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void __launch_bounds__(256) pos_table_kernel(const float* pos, const float* cls, const float* pbias, float* tab,
                                                        int D, int S, int oh, int ow, int identity, float inv_sy, float inv_sx) {
    const int t = blockIdx.x;  // 0 .. oh*ow
    for (int d = threadIdx.x; d < D; d += 256) {
        if (t == 0) { tab[d] = cls[d] + pos[d]; continue; }
        const int p = t - 1;
        float v;
        if (identity) {
            v = pos[static_cast<long long>(1 + p) * D + d];
        } else {
            const int oy = p / ow, ox = p - oy * ow;
            const float fy = inv_sy * (oy + 0.5f) - 0.5f, fx = inv_sx * (ox + 0.5f) - 0.5f;
            const int iy = static_cast<int>(floorf(fy)), ix = static_cast<int>(floorf(fx));
            const float ty = fy - iy, tx = fx - ix;
            const float wy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
            const float wx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
            v = 0.f;
            for (int i = 0; i < 4; ++i) {
                const int yy = min(max(iy - 1 + i, 0), S - 1);
                float rowv = 0.f;
                for (int j = 0; j < 4; ++j) {
                    const int xx = min(max(ix - 1 + j, 0), S - 1);
                    rowv += wx[j] * pos[static_cast<long long>(1 + yy * S + xx) * D + d];
                }
                v += wy[i] * rowv;
            }
        }
        tab[static_cast<long long>(t) * D + d] = v + pbias[d];
    }
}

// ---------------------------------------------------------------- readout concat (use_clstoken, dpt.py:153-156)
// out[(b, i), 0:D] = tok[(b, i), :]; out[(b, i), D:2D] = cls[b, :]   (torch.cat((x, cls.unsqueeze(1).expand_as(x)), -1))
// VEC elements (16 bytes) per thread.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) concat_cls_kernel(const T* __restrict__ tok, const T* __restrict__ cls,
                                                         T* __restrict__ out, long long rows, int np, int D) {
    const int dv = D / VEC;  // vectors per half row
    const long long total = rows * 2 * dv;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * 256) {
        const long long r = i / (2 * dv);
        const int c = static_cast<int>(i - r * 2 * dv);
        const uint4* src = c < dv ? reinterpret_cast<const uint4*>(tok + r * D) + c
                                  : reinterpret_cast<const uint4*>(cls + (r / np) * D) + (c - dv);
        reinterpret_cast<uint4*>(out + r * 2 * D)[c] = *src;
    }
}

// ---------------------------------------------------------------- SwiGLU gate (ViT-g, swiglu_ffn.py:30-34)
// x12 [rows, 2*Hd] -> out [rows, Hd] = silu(x12[:, :Hd]) * x12[:, Hd:]; fp32 maths, ATen's silu = x / (1 + exp(-x))
template <typename T>
__global__ void __launch_bounds__(256) swiglu_kernel(const T* __restrict__ x12, T* __restrict__ out, long long rows, int Hd) {
    const long long total = rows * Hd;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * 256) {
        const long long r = i / Hd;
        const int c = static_cast<int>(i - r * Hd);
        const float a = to_f(x12[r * 2 * Hd + c]), g = to_f(x12[r * 2 * Hd + Hd + c]);
        out[i] = from_f<T>(a / (1.f + expf(-a)) * g);
    }
}

// ---------------------------------------------------------------- weight packing
template <typename T>
__global__ void __launch_bounds__(256) pack_linear_kernel(const float* w, T* out, int N, int K, int Kp, int scale_rows,
                                                          float scale) {
    const long long total = static_cast<long long>(N) * Kp;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * 256) {
        const int n = static_cast<int>(i / Kp), k = static_cast<int>(i - static_cast<long long>(n) * Kp);
        float v = (k < K) ? w[static_cast<long long>(n) * K + k] : 0.f;
        if (n < scale_rows) v *= scale;
        out[i] = from_f<T>(v);
    }
}

// w [Co, Ci, kh, kw] -> out [Co, taps, Cp]
template <typename T>
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* w, T* out, int Co, int Ci, int taps, int Cp) {
    const long long total = static_cast<long long>(Co) * taps * Cp;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * 256) {
        const int c = static_cast<int>(i % Cp);
        const int tap = static_cast<int>((i / Cp) % taps);
        const int co = static_cast<int>(i / (static_cast<long long>(Cp) * taps));
        const float v = (c < Ci) ? w[(static_cast<long long>(co) * Ci + c) * taps + tap] : 0.f;
        out[i] = from_f<T>(v);
    }
}

// ConvTranspose k=s: w [Ci, Co, k, k] -> out [(ky*k + kx) * CoP + co, Kp], zero rows for co >= Co
template <typename T>
__global__ void __launch_bounds__(256) pack_convT_kernel(const float* w, T* out, int Ci, int Co, int k, int CoP, int Kp) {
    const long long total = static_cast<long long>(k) * k * CoP * Kp;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * 256) {
        const int ci = static_cast<int>(i % Kp);
        const long long n = i / Kp;
        const int co = static_cast<int>(n % CoP), kk = static_cast<int>(n / CoP);
        float v = 0.f;
        if (ci < Ci && co < Co) v = w[(static_cast<long long>(ci) * Co + co) * k * k + kk];
        out[i] = from_f<T>(v);
    }
}

__global__ void __launch_bounds__(256) copy_scale_kernel(const float* in, float* out, long long n, long long scale_n, float scale) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) out[i] = in[i] * (i < scale_n ? scale : 1.f);
}

int grid_for(long long total, int per_block = 256) {
    long long g = cdivl(total, per_block);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define DISPATCH_T(is_bf16, ...)                  \
    do {                                          \
        if (is_bf16) { using T = bf16; __VA_ARGS__; } \
        else { using T = float; __VA_ARGS__; }    \
    } while (0)

int patch_im2col(const float* x, void* A, int is_bf16, int B, int H, int W, int Kp, cudaStream_t st) {
    DAD_REQUIRE(H % 14 == 0 && W % 14 == 0 && Kp >= 588, "patch_im2col: bad dims");
    const long long rows = static_cast<long long>(B) * (1 + (H / 14) * (W / 14));
    ProfScope prof(PROF_ELEM, static_cast<double>(B) * 3 * H * W * 4 + static_cast<double>(rows) * Kp * (is_bf16 ? 2 : 4), st);
    DISPATCH_T(is_bf16, (patch_im2col_kernel<T><<<static_cast<unsigned>(rows), 256, 0, st>>>(x, reinterpret_cast<T*>(A), B, H, W, Kp)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int layernorm(const float* in, const float* w, const float* b, void* out, int is_bf16, float* out_f32, long long rows,
              int D, int out_period, int in_period, int in_offset, float eps, cudaStream_t st) {
    DAD_REQUIRE(D % 4 == 0 && D <= 2048, "layernorm: D=%d unsupported", D);
    const unsigned grid = static_cast<unsigned>(cdivl(rows, 8));
    ProfScope prof(PROF_LN, static_cast<double>(rows) * D * (4 + (out ? (is_bf16 ? 2 : 4) : 0) + (out_f32 ? 4 : 0)), st);
    const int vpt = cdiv(D, 128);
    cudaError_t ln_err = cudaSuccess;
#define LN_LAUNCH(V)                                                                                              \
    DISPATCH_T(is_bf16, (ln_err = launch_pdl(layernorm_kernel<T, V>, dim3(grid), dim3(256), 0, st, in, w, b,           \
                                             reinterpret_cast<T*>(out), out_f32, rows, D, out_period, in_period, \
                                             in_offset, eps)))
    if (vpt <= 3) LN_LAUNCH(3);
    else if (vpt <= 6) LN_LAUNCH(6);
    else if (vpt <= 8) LN_LAUNCH(8);
    else LN_LAUNCH(16);
#undef LN_LAUNCH
    DAD_CHECK_CUDA(ln_err);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int bilinear_nhwc(const void* in, void* out, int is_bf16, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st) {
    DAD_REQUIRE(C % 8 == 0, "bilinear: C=%d must be a multiple of 8", C);
    // ATen: scale = (in - 1) / (out - 1) in fp32 (0 when out == 1)
    const float sh = Ho > 1 ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.f;
    ProfScope prof(PROF_ELEM, static_cast<double>(B) * C * (is_bf16 ? 2 : 4) * (static_cast<double>(Hi) * Wi + static_cast<double>(Ho) * Wo), st);
    DAD_REQUIRE(Ho <= 65535 && B <= 65535, "bilinear: output height / batch too large for the launch grid");
    if (is_bf16) {
        const dim3 grid(cdiv(Wo * (C / 8), 256), cdiv(Ho, BIL_ROWS), B);
        DAD_CHECK_CUDA(launch_pdl(bilinear_kernel<bf16, 8>, grid, dim3(256), 0, st, reinterpret_cast<const bf16*>(in),
                                  reinterpret_cast<bf16*>(out), Hi, Wi, Ho, Wo, C, sh, sw));
    } else {
        const dim3 grid(cdiv(Wo * (C / 4), 256), cdiv(Ho, BIL_ROWS), B);
        bilinear_kernel<float, 4><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), Hi, Wi,
                                                        Ho, Wo, C, sh, sw);
    }
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int im2col_s2(const void* in, void* A, int is_bf16, int B, int H, int W, int C, int Cp, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long rows = static_cast<long long>(B) * Ho * Wo;
    ProfScope prof(PROF_ELEM, (static_cast<double>(B) * H * W * C + static_cast<double>(rows) * 9 * Cp) * (is_bf16 ? 2 : 4), st);
    DISPATCH_T(is_bf16, (im2col_s2_kernel<T><<<static_cast<unsigned>(rows), 256, 0, st>>>(
                            reinterpret_cast<const T*>(in), reinterpret_cast<T*>(A), B, H, W, C, Cp, Ho, Wo)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int head1x1(const float* in, const float* w, const float* bias, float* out, long long P, cudaStream_t st) {
    ProfScope prof(PROF_ELEM, static_cast<double>(P) * 33 * 4, st);
    head1x1_kernel<<<static_cast<unsigned>(cdivl(P, 256)), 256, 0, st>>>(in, w, bias, out, P);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int concat_cls(const void* tok, const void* cls, void* out, int is_bf16, int B, int np, int D, cudaStream_t st) {
    DAD_REQUIRE(D % 8 == 0, "concat_cls: D=%d must be a multiple of 8", D);
    const long long rows = static_cast<long long>(B) * np;
    const size_t es = is_bf16 ? 2 : 4;
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * D * es * 3 + static_cast<double>(B) * D * es, st);
    if (is_bf16)
        concat_cls_kernel<bf16, 8><<<grid_for(rows * 2 * (D / 8)), 256, 0, st>>>(
            reinterpret_cast<const bf16*>(tok), reinterpret_cast<const bf16*>(cls), reinterpret_cast<bf16*>(out), rows, np, D);
    else
        concat_cls_kernel<float, 4><<<grid_for(rows * 2 * (D / 4)), 256, 0, st>>>(
            reinterpret_cast<const float*>(tok), reinterpret_cast<const float*>(cls), reinterpret_cast<float*>(out), rows, np, D);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int swiglu(const void* x12, void* out, int is_bf16, long long rows, int Hd, cudaStream_t st) {
    ProfScope prof(PROF_ELEM, static_cast<double>(rows) * Hd * 3 * (is_bf16 ? 2 : 4), st);
    DISPATCH_T(is_bf16, (swiglu_kernel<T><<<grid_for(rows * Hd), 256, 0, st>>>(reinterpret_cast<const T*>(x12),
                                                                               reinterpret_cast<T*>(out), rows, Hd)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pos_table(const float* pos, const float* cls, const float* pbias, float* tab, int D, int H, int W, cudaStream_t st) {
    const int S = 37;
    const int ph = H / 14, pw = W / 14;  // reference: w0 <- image HEIGHT // 14 (dinov2.py:213), h0 <- width // 14
    const int identity = (ph * pw == S * S && H == W) ? 1 : 0;
    // scale_factor = (n + 0.1) / 37 per axis, applied to (rows, cols) of the 37 x 37 grid
    const double sfy = (static_cast<double>(ph) + 0.1) / S, sfx = (static_cast<double>(pw) + 0.1) / S;
    pos_table_kernel<<<1 + ph * pw, 256, 0, st>>>(pos, cls, pbias, tab, D, S, ph, pw, identity,
                                                  static_cast<float>(1.0 / sfy), static_cast<float>(1.0 / sfx));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pack_linear(const float* w, void* out, int is_bf16, int N, int K, int Kp, int scale_rows, float scale, cudaStream_t st) {
    const long long total = static_cast<long long>(N) * Kp;
    DISPATCH_T(is_bf16, (pack_linear_kernel<T><<<grid_for(total), 256, 0, st>>>(w, reinterpret_cast<T*>(out), N, K, Kp, scale_rows, scale)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pack_conv(const float* w, void* out, int is_bf16, int Co, int Ci, int taps, int Cp, cudaStream_t st) {
    const long long total = static_cast<long long>(Co) * taps * Cp;
    DISPATCH_T(is_bf16, (pack_conv_kernel<T><<<grid_for(total), 256, 0, st>>>(w, reinterpret_cast<T*>(out), Co, Ci, taps, Cp)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int pack_convT(const float* w, void* out, int is_bf16, int Ci, int Co, int k, int CoP, int Kp, cudaStream_t st) {
    const long long total = static_cast<long long>(k) * k * CoP * Kp;
    DISPATCH_T(is_bf16, (pack_convT_kernel<T><<<grid_for(total), 256, 0, st>>>(w, reinterpret_cast<T*>(out), Ci, Co, k, CoP, Kp)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int copy_scale(const float* in, float* out, long long n, long long scale_n, float scale, cudaStream_t st) {
    copy_scale_kernel<<<static_cast<unsigned>(cdivl(n, 256)), 256, 0, st>>>(in, out, n, scale_n, scale);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
