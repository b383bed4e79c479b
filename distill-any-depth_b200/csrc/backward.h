// Internal interface of the training-backward kernels (gradient of the student forward,
// reference tools/train_distillation.py:1556-1575 = autograd over depth_anything_v2/dpt.py:150-225).
// fp32 verification engine: every kernel is plain FFMA fp32 on row-major / NHWC fp32 tensors.
#pragma once
#include "common.h"

namespace dad {

// Batched, fully strided GEMM   C[z][m, n] (+)= alpha * sum_k A[z][m, k] * B[z][k, n]
// element (m, k) of A at A + za + m*sam + k*sak (likewise B, C); batch z = z1 * nb2 + z2 with offsets
// z1 * x1 + z2 * x2 per operand.  Optionally B is an IMPLICIT operand (the im2col view of an NHWC tensor) and the
// C index goes through a weight-layout map, which turns the kernel into the weight-gradient pass of a convolution.
struct SGemm {
    const float* A = nullptr;
    const float* B = nullptr;
    float* C = nullptr;
    int M = 0, N = 0, K = 0;
    long long sam = 0, sak = 0, sbk = 0, sbn = 0, scm = 0, scn = 0;
    int nb1 = 1, nb2 = 1;
    long long a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
    float alpha = 1.f;
    int accumulate = 0;  // 1: C += result (the K loop may then be split over CTAs, atomicAdd)
    // implicit B: k = output pixel (b, oy, ox) of a 3x3 (pad 1) or 1x1 window with `conv_stride`, n = tap * convC + c
    //             -> B = In[b, oy*stride + dy - 1, ox*stride + dx - 1, c] (0 outside); In is NHWC [*, convH, convW, convC]
    int conv_taps = 0, convC = 0, convH = 0, convW = 0, convHo = 0, convWo = 0, conv_stride = 1;
    // C index map: 0 strided; 1 conv weight [Co][Ci][taps]: m = co, n = tap*convC + c -> (m*convC + c)*taps + tap;
    //              2 ConvTranspose weight [Ci][Co][kk]: m = t*ct_CoP + co, n = ci -> (n*ct_Co + co)*ct_kk + t (co < ct_Co)
    int cmap = 0, ct_CoP = 0, ct_Co = 0, ct_kk = 0;
};
int sgemm(const SGemm& g, cudaStream_t st);

// activation tensors carry a run-time element type: *_bf16 / bf = 1 for bf16 (tensor-core engine), 0 for fp32
// out[n] += sum_r X[r*ldx + n] * (Y ? Y[r*ldy + n] : 1);  if `scaled_out` (type of Y): scaled_out[r*ldx + n] = X[r*ldx + n] * scale[n]
int colsum(const void* X, int x_bf16, long long ldx, const void* Y, int y_bf16, long long ldy, long long rows, int N, float* out,
           const float* scale, void* scaled_out, cudaStream_t st);
// LayerNorm backward, row mapping as layernorm(): dy row r <-> x row ir.  dx[ir] += ..., dw += ..., db += ...
int layernorm_bwd(const float* x, const float* w, const void* dy, int dy_bf16, float* dx, float* dw, float* db, long long rows, int D,
                  int out_period, int in_period, int in_offset, float eps, cudaStream_t st);
// xnew = xold + gamma * y   (LayerScale + residual of the training forward)
int ls_residual(const float* xold, const void* y, int bf, const float* gamma, float* xnew, long long rows, int D, cudaStream_t st);
int gelu_fwd(const void* pre, void* out, int bf, long long n, cudaStream_t st);
int gelu_bwd(const void* pre, const void* dout, void* dpre, int bf, long long n, cudaStream_t st);
// SwiGLU gate adjoint: x12 [rows, 2*Hd] = [x1 | x2], dg [rows, Hd] -> dx12 [rows, 2*Hd]  (swiglu_ffn.py:30-34)
int swiglu_bwd(const void* x12, const void* dg, void* dx12, int bf, long long rows, int Hd, cudaStream_t st);
// dst[r*ldd + c] = src[r*lds + c], c < cols (one element type)
int copy_cols(const void* src, long long lds, void* dst, long long ldd, long long rows, int cols, int bf, cudaStream_t st);
// out = (add ? add : 0) + g * (y > 0)
int relu_bwd(const void* g, const void* y, const void* add, void* out, int bf, long long n, cudaStream_t st);
int add_inplace(void* dst, int bf, const float* src, long long n, cudaStream_t st);   // dst += src (fp32)
int convert(const void* src, int src_bf16, void* dst, int dst_bf16, long long n, cudaStream_t st);
int fill_f32(float* p, float v, long long n, cudaStream_t st);
// bf16 only: out[c][r] = in[r*ld + c] (r < R; zero for R <= r < Rp) - K-major operands of the weight-gradient GEMMs
int transpose_pad(const void* in, long long ld, int R, int C, void* out, int Rp, cudaStream_t st);
// the same for Z problems (element strides zin / zout between them)
int transpose_pad_batched(const void* in, long long ld, int R, int C, void* out, int Rp, int Z, long long zin, long long zout,
                          cudaStream_t st);
// bf16 only: out[(c*taps + tap)][p] = window(X)[p, tap, c] over output pixels p (zero padded to Pp columns)
int im2colT(const void* X, int B, int H, int W, int Ci, int taps, int stride, int Ho, int Wo, void* out, long long Pp, cudaStream_t st,
            int shift_y = 0, int shift_x = 0);
int wgrad_unshift(const float* S, float* dW, int Co, int Ci, int taps, int ld, cudaStream_t st);
// w [N][K] fp32 -> out [K][Np] bf16
int pack_linear_T(const float* w, void* out, int N, int K, int Np, cudaStream_t st);
int head1x1_any(const void* in, int bf, const float* w, const float* bias, float* out, long long P, cudaStream_t st);
// row-wise softmax of S [rows, T] in place; and dS = P * (dP - sum_j P*dP) in place of dP
int softmax_rows(float* S, long long rows, int T, int ld, cudaStream_t st);
int softmax_bwd_rows(const float* P, float* dP, long long rows, int T, int ld, cudaStream_t st);
// bf16 engine: P = softmax(S) (fp32 -> bf16), dS = P * (dP - sum P dP) (bf16, fp32 -> bf16); pad columns [T, ld) zeroed
int softmax_rows_bf16(const float* S, void* P, long long rows, int T, int ld, cudaStream_t st);
int softmax_bwd_bf16(const void* P, const float* dP, void* dS, long long rows, int T, int ld, cudaStream_t st);
// Z problems of [T][ld] fp32 -> bf16 copy in the same layout (dstN) and / or transposed (dstT); pad columns zeroed
int cvt_tiles(const float* src, void* dstN, void* dstT, int Z, int T, int ld, cudaStream_t st);
// bf16: dst[(b*heads + h)][d][t] = src[(b*T + t)*lds + h*64 + d] (64 x ld per problem, zero for t >= T)
int head_transpose(const void* src, long long lds, void* dst, int B, int T, int heads, int ld, cudaStream_t st);
// adjoint of bilinear_nhwc (align_corners=True): gin [B,Hi,Wi,C] must be zero-filled by the caller
int bilinear_bwd(const void* gout, int bf, float* gin, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st);
// the same adjoint in gather form: writes gin (activation type) once, deterministic, no zero fill / atomics
int bilinear_bwd_gather(const void* gout, int bf, void* gin, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t st);
// output head: depth = relu(dot(t32, w2) + b2), t32 = relu(conv + b) saved.  dt32 [P,32]; dw2 [32] / db2 [1] accumulate
int head_bwd(const float* gdepth, const float* depth, const void* t32, int bf, const float* w2, void* dt32, float* dw2,
             float* db2, long long P, cudaStream_t st);
// ConvTranspose k=s: G[(b,y,x), t*CoP + co] = dOut[b, k*y+ky, k*x+kx, co] (0 for co >= Co)
int convT_gather(const void* dout, void* G, int bf, int B, int H, int W, int k, int Co, int CoP, cudaStream_t st);
// ConvTranspose weight gradient: tmp [(t*CoP + co)][Ci] (GEMM output) -> dW[ci][co][t] += tmp
int convT_wgrad_permute(const float* tmp, float* dW, int Ci, int Co, int CoP, int kk, cudaStream_t st);
// stride-2 3x3 conv: din[b,y,x,c] = sum over taps of dcol[(b,oy,ox), tap*Cp + c] with y = 2*oy+dy-1, x = 2*ox+dx-1
int col2im_s2(const void* dcol, void* din, int bf, int B, int H, int W, int C, int Cp, cudaStream_t st);
// dgrad weights of a stride-1 conv: w [Co][Ci][taps] -> out [Ci][taps][CoP], out[ci][t][co] = w[co][ci][taps-1-t]
int pack_conv_dgrad(const float* w, void* out, int bf, int Co, int Ci, int taps, int CoP, cudaStream_t st);
// dtab[t, d] = sum_b G[(b*T + t), d]
int batch_sum_rows(const float* G, float* dtab, int B, int T, int D, cudaStream_t st);
// gradients of pos_table(): cls_token, pos_embed (bicubic adjoint), patch-embed bias; all accumulate (any may be null)
int pos_table_bwd(const float* dtab, float* dpos, float* dcls, float* dpbias, int D, int H, int W, cudaStream_t st);

}  // namespace dad
