// Pre- / post-processing either side of the forward, so that infer_image() stays on the GPU:
//   preprocess   uint8 BGR/RGB HWC image -> /255 -> cv2.resize(INTER_CUBIC) -> (x - mean) / std -> CHW fp32
//                (reference: DepthAnythingV2.image2tensor, depth_anything_v2/dpt.py:237-262 with
//                 util/transform.py:109-148; tools/testers/infer.py:125-128,173-177)
//   resize_depth bilinear align_corners=True on [B,1,H,W] fp32 (dpt.py:233, ATen upsample_bilinear2d index maths)
//   minmax_norm  per-image (d - min) / (max - min)              (tools/testers/infer.py:135)
// The cubic resize follows OpenCV's resizeGeneric_ for CV_64F input (the reference divides by 255.0 in float64
// before resizing): float coefficients from interpolateCubic (A = -0.75) at fx = (float)((dx + 0.5) * scale - 0.5),
// replicate border, double accumulation horizontally then vertically.
#include "common.h"
#include "elementwise.h"

namespace dad {

namespace {

__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
    const float A = -0.75f;
    // ((A*(x + 1) - 5*A)*(x + 1) + 8*A)*(x + 1) - 4*A, evaluated without contraction like the host build
    const float x1 = __fadd_rn(x, 1.0f);
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5 * A), x1), 8 * A), x1), 4 * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2, x), A + 3), x), x), 1.0f);
    const float y = __fsub_rn(1.0f, x);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2, y), A + 3), y), y), 1.0f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}

struct PreArgs {
    const uint8_t* src;   // [h, w, 3] with `pitch` bytes per row
    float* dst;           // [3, nh, nw]
    int h, w, nh, nw;
    long long pitch;
    int swap_rb;          // 1: source is BGR, output channel 0 = R
    double scale_x, scale_y;
    double mean[3], stdv[3];
};

__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= a.nw || dy >= a.nh) return;
    float fx = static_cast<float>((dx + 0.5) * a.scale_x - 0.5);
    float fy = static_cast<float>((dy + 0.5) * a.scale_y - 0.5);
    const int sx = static_cast<int>(floorf(fx)), sy = static_cast<int>(floorf(fy));
    fx -= sx;
    fy -= sy;
    float ca[4], cb[4];
    if (a.nw == a.w) { ca[0] = 0.f; ca[1] = 1.f; ca[2] = 0.f; ca[3] = 0.f; } else cubic_coeffs(fx, ca);
    if (a.nh == a.h) { cb[0] = 0.f; cb[1] = 1.f; cb[2] = 0.f; cb[3] = 0.f; } else cubic_coeffs(fy, cb);
    int xs[4], ys[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xs[k] = min(max(sx - 1 + k, 0), a.w - 1);
        ys[k] = min(max(sy - 1 + k, 0), a.h - 1);
    }
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        const uint8_t* row = a.src + ys[ky] * a.pitch;
        double hrow[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            const uint8_t* px = row + xs[kx] * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double v = static_cast<double>(px[a.swap_rb ? 2 - c : c]) / 255.0;
                hrow[c] = __dadd_rn(hrow[c], __dmul_rn(v, static_cast<double>(ca[kx])));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = __dadd_rn(acc[c], __dmul_rn(hrow[c], static_cast<double>(cb[ky])));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        a.dst[(static_cast<long long>(c) * a.nh + dy) * a.nw + dx] = static_cast<float>((acc[c] - a.mean[c]) / a.stdv[c]);
}

// ATen upsample_bilinear2d, align_corners=True, one channel
__global__ void __launch_bounds__(256) resize_depth_kernel(const float* in, float* out, int Hi, int Wi, int Ho, int Wo,
                                                           float sh, float sw) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
    const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= Wo || oy >= Ho) return;
    const int b = blockIdx.z;
    const float fy = sh * oy, fx = sw * ox;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + static_cast<long long>(b) * Hi * Wi;
    const float v = hy * (hx * p[y0 * Wi + x0] + lx * p[y0 * Wi + x1]) + ly * (hx * p[y1 * Wi + x0] + lx * p[y1 * Wi + x1]);
    out[(static_cast<long long>(b) * Ho + oy) * Wo + ox] = v;
}

__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float keyf(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__global__ void mm_init_kernel(uint32_t* mm, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { mm[2 * i] = 0xFFFFFFFFu; mm[2 * i + 1] = 0u; }
}
__global__ void __launch_bounds__(256) mm_reduce_kernel(const float* x, long long L, uint32_t* mm) {
    const int b = blockIdx.y;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < L; i += static_cast<long long>(gridDim.x) * 256) {
        const uint32_t k = fkey(x[b * L + i]);
        mn = min(mn, k);
        mx = max(mx, k);
    }
    for (int o = 16; o; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * b], mn); atomicMax(&mm[2 * b + 1], mx); }
}
__global__ void __launch_bounds__(256) mm_apply_kernel(const float* x, float* out, long long L, const uint32_t* mm) {
    const int b = blockIdx.y;
    const float mn = keyf(mm[2 * b]), mx = keyf(mm[2 * b + 1]);
    const float range = mx - mn;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < L; i += static_cast<long long>(gridDim.x) * 256)
        out[b * L + i] = (x[b * L + i] - mn) / range;
}

// colorize_depth_maps (distillanydepth/utils/image_util.py:69-118): x = clip((d - min) / (max - min), 0, 1) in fp32 (or
// d * 0 when min == max), matplotlib's Colormap.__call__ index  min(int(x * 256), 255)  into the 256-entry LUT, masked
// pixels set to 0.  Writes the float CHW image and / or the uint8 HWC image of tools/testers/infer.py:139-140
// ((rgb * 255).astype(uint8), chw2hwc) in the same pass.
__global__ void __launch_bounds__(256) colorize_kernel(const float* __restrict__ d, const uint8_t* __restrict__ valid, long long HW,
                                                       float dmin, float dmax, int degenerate, const float* __restrict__ lut,
                                                       const uint8_t* __restrict__ lut_u8, float* __restrict__ out_chw,
                                                       uint8_t* __restrict__ out_hwc) {
    __shared__ float slut[768];
    __shared__ uint8_t slut8[768];   // (float64 table * 255) truncated on the host: numpy's float64 product, bit for bit
    for (int i = threadIdx.x; i < 768; i += 256) { slut[i] = lut[i]; slut8[i] = lut_u8 ? lut_u8[i] : 0; }
    __syncthreads();
    const int b = blockIdx.y;
    const float range = dmax - dmin;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < HW; i += static_cast<long long>(gridDim.x) * 256) {
        const float v = d[b * HW + i];
        float x = degenerate ? v * 0.f : fminf(fmaxf(__fdiv_rn(v - dmin, range), 0.f), 1.f);
        int idx = static_cast<int>(x * 256.f);          // numpy: float32 * 256 -> astype(int) truncates
        idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);    // x == 1 -> N - 1; NaN -> matplotlib's "bad" colour is out of scope
        const bool ok = valid == nullptr || valid[b * HW + i] != 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float col = ok ? slut[idx * 3 + c] : 0.f;
            if (out_chw) out_chw[(b * 3 + c) * HW + i] = col;
            if (out_hwc) out_hwc[(b * HW + i) * 3 + c] = ok ? slut8[idx * 3 + c] : 0;
        }
    }
}

}  // namespace

int colorize_depth(const float* depth, const uint8_t* valid, int B, long long HW, float dmin, float dmax, int degenerate,
                   const float* lut, const uint8_t* lut_u8, float* out_chw, uint8_t* out_hwc, cudaStream_t st) {
    DAD_REQUIRE(depth && lut && (out_chw || out_hwc) && B > 0 && HW > 0 && B <= 65535, "colorize_depth: bad arguments");
    DAD_REQUIRE(!out_hwc || lut_u8, "colorize_depth: the uint8 output needs the uint8 table");
    const long long blocks = cdivl(HW, 256 * 4);
    const int gx = static_cast<int>(blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks));
    ProfScope prof(PROF_ELEM, B * static_cast<double>(HW) * (4.0 + (out_chw ? 12.0 : 0.0) + (out_hwc ? 3.0 : 0.0)), st);
    colorize_kernel<<<dim3(gx, B), 256, 0, st>>>(depth, valid, HW, dmin, dmax, degenerate, lut, lut_u8, out_chw, out_hwc);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int preprocess_image(const uint8_t* src, int h, int w, long long pitch, int swap_rb, int nh, int nw, const double* mean,
                     const double* stdv, float* dst, cudaStream_t st) {
    DAD_REQUIRE(src && dst && mean && stdv, "preprocess_image: null argument");
    DAD_REQUIRE(h > 0 && w > 0 && nh > 0 && nw > 0 && pitch >= 3LL * w, "preprocess_image: bad dims");
    PreArgs a{};
    a.src = src; a.dst = dst; a.h = h; a.w = w; a.nh = nh; a.nw = nw; a.pitch = pitch; a.swap_rb = swap_rb;
    // cv::resize: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
    a.scale_x = 1.0 / (static_cast<double>(nw) / w);
    a.scale_y = 1.0 / (static_cast<double>(nh) / h);
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
    ProfScope prof(PROF_ELEM, 3.0 * h * w + 12.0 * nh * nw, st);
    preprocess_kernel<<<dim3(cdiv(nw, 32), cdiv(nh, 8)), 256, 0, st>>>(a);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int resize_depth(const float* in, int B, int Hi, int Wi, int Ho, int Wo, float* out, cudaStream_t st) {
    DAD_REQUIRE(in && out && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && B <= 65535, "resize_depth: bad arguments");
    const float sh = Ho > 1 ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.f;
    const float sw = Wo > 1 ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.f;
    ProfScope prof(PROF_ELEM, 4.0 * B * (static_cast<double>(Hi) * Wi + static_cast<double>(Ho) * Wo), st);
    resize_depth_kernel<<<dim3(cdiv(Wo, 32), cdiv(Ho, 8), B), 256, 0, st>>>(in, out, Hi, Wi, Ho, Wo, sh, sw);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int minmax_normalize(const float* in, int B, long long L, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(in && out && B > 0 && L > 0 && B <= 65535, "minmax_normalize: bad arguments");
    DAD_REQUIRE(ws && ws_bytes >= static_cast<size_t>(B) * 8, "minmax_normalize: workspace too small (8 bytes per image)");
    uint32_t* mm = reinterpret_cast<uint32_t*>(ws);
    const int gx = static_cast<int>(cdivl(L, 256 * 8) < 1 ? 1 : (cdivl(L, 256 * 8) > 592 ? 592 : cdivl(L, 256 * 8)));
    ProfScope prof(PROF_ELEM, 8.0 * B * L, st, 3);
    mm_init_kernel<<<cdiv(B, 128), 128, 0, st>>>(mm, B);
    mm_reduce_kernel<<<dim3(gx, B), 256, 0, st>>>(in, L, mm);
    mm_apply_kernel<<<dim3(gx, B), 256, 0, st>>>(in, out, L, mm);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
