// Internal C++ interface of the HBM-bound helper kernels and the attention kernels.
#pragma once
#include "common.h"

namespace dad {

int patch_im2col(const float* x, void* A, int is_bf16, int B, int H, int W, int Kp, cudaStream_t st);
// rows = output rows; input row = (r / out_period) * in_period + in_offset + r % out_period
int layernorm(const float* in, const float* w, const float* b, void* out, int is_bf16, float* out_f32, long long rows,
              int D, int out_period, int in_period, int in_offset, float eps, cudaStream_t st);
int bilinear_nhwc(const void* in, void* out, int is_bf16, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st);
int im2col_s2(const void* in, void* A, int is_bf16, int B, int H, int W, int C, int Cp, cudaStream_t st);
int head1x1(const float* in, const float* w, const float* bias, float* out, long long P, cudaStream_t st);
// use_clstoken readout input: out [B*np, 2D] = [tok | cls[b]]  (dpt.py:153-156)
int concat_cls(const void* tok, const void* cls, void* out, int is_bf16, int B, int np, int D, cudaStream_t st);
// ViT-g SwiGLU gate: out [rows, Hd] = silu(x12[:, :Hd]) * x12[:, Hd:]  (swiglu_ffn.py:30-34)
int swiglu(const void* x12, void* out, int is_bf16, long long rows, int Hd, cudaStream_t st);
int pos_table(const float* pos, const float* cls, const float* pbias, float* tab, int D, int H, int W, cudaStream_t st);
int pack_linear(const float* w, void* out, int is_bf16, int N, int K, int Kp, int scale_rows, float scale, cudaStream_t st);
int pack_conv(const float* w, void* out, int is_bf16, int Co, int Ci, int taps, int Cp, cudaStream_t st);
int pack_convT(const float* w, void* out, int is_bf16, int Ci, int Co, int k, int CoP, int Kp, cudaStream_t st);
int copy_scale(const float* in, float* out, long long n, long long scale_n, float scale, cudaStream_t st);
// imageproc.cu: pre- / post-processing next to the forward
int preprocess_image(const uint8_t* src, int h, int w, long long pitch, int swap_rb, int nh, int nw, const double* mean,
                     const double* stdv, float* dst, cudaStream_t st);
int resize_depth(const float* in, int B, int Hi, int Wi, int Ho, int Wo, float* out, cudaStream_t st);
int minmax_normalize(const float* in, int B, long long L, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int colorize_depth(const float* depth, const uint8_t* valid, int B, long long HW, float dmin, float dmax, int degenerate,
                   const float* lut, const uint8_t* lut_u8, float* out_chw, uint8_t* out_hwc, cudaStream_t st);
int attention_tc(const bf16* qkv, bf16* out, int B, int N, int heads, cudaStream_t st);  // tcgen05 / TMEM
int attention_tc3(const bf16* qkv, bf16* out, int B, int N, int heads, cudaStream_t st); // 4 CTAs / SM variant
// round-2 default: FFMA2 + MUFU / FMA-pipe exponentials, tensor-core row sums, in-kernel exact fallback (attention_tc5.cu)
int attention_tc5(const bf16* qkv, bf16* out, int B, int N, int heads, int poly_pairs, cudaStream_t st);
// same with three in-place score buffers and the row sums in the softmax threads (attention_tc6.cu)
int attention_tc6(const bf16* qkv, bf16* out, int B, int N, int heads, int poly_pairs, cudaStream_t st);
// one CTA per SM, two softmax warpgroups alternating the key tiles of one item (attention_tc7.cu)
int attention_tc7(const bf16* qkv, bf16* out, int B, int N, int heads, int poly_pairs, int split_issuers, cudaStream_t st);
int attention(const void* qkv, void* out, int is_bf16, int B, int N, int heads, cudaStream_t st);

}  // namespace dad
