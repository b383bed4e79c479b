// C ABI wrappers for the loss kernels and the kernel-level test entry points (include/dad_b200.h).
#include "../../include/dad_b200.h"

#include "elementwise.h"
#include "gemm.h"
#include "losses.h"

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int dad_abi_version(void) { return 1; }

size_t dad_loss_workspace_bytes(int rows, int num_contexts) { return dad::loss_workspace_bytes(rows, num_contexts); }

int dad_masked_shift_and_scale(const float* pred, const float* gt, const uint8_t* mask, int rows, int64_t L,
                               float* pred_aligned, float* gt_aligned, void* ws, size_t wsb, void* stream) {
    if (!pred_aligned || !gt_aligned) return dad::set_error(DAD_ERR_INVALID, "masked_shift_and_scale: null output");
    return dad::masked_shift_and_scale(pred, gt, mask, rows, L, pred_aligned, gt_aligned, ws, wsb, ST(stream));
}

int dad_ssi_loss(const float* pred, const float* gt, const uint8_t* mask, int rows, int64_t L, float* dense_out,
                 float* out_scalar, double* partials, void* ws, size_t wsb, void* stream) {
    return dad::ssi_loss(pred, gt, mask, rows, L, dense_out, out_scalar, partials, ws, wsb, ST(stream));
}

int dad_contexts_dr(int level, const float* gt, const uint8_t* mask, int B, int64_t L, uint8_t* ctx_out, void* ws,
                    size_t wsb, void* stream) {
    return dad::contexts_dr(level, gt, mask, B, L, ctx_out, ws, wsb, ST(stream));
}

int dad_ssi_loss_bwd(const float* pred, const float* gt, const uint8_t* mask, int rows, int64_t L, const float* grad_out,
                     float* grad_pred, void* ws, size_t wsb, void* stream) {
    return dad::ssi_loss_bwd(pred, gt, mask, rows, L, grad_out, grad_pred, ws, wsb, ST(stream));
}

int dad_hdn_loss_dr_bwd(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L,
                        const float* grad_out, float* grad_pred, void* ws, size_t wsb, void* stream) {
    return dad::hdn_loss_dr_bwd(level, pred, gt, mask, B, L, grad_out, grad_pred, ws, wsb, ST(stream));
}

int dad_hdn_loss_bwd(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, int64_t L, const float* grad_out,
                     float* grad_pred, void* ws, size_t wsb, void* stream) {
    return dad::hdn_loss_ctx_bwd(pred, gt, ctx, K, B, L, grad_out, grad_pred, ws, wsb, ST(stream));
}

int dad_feat_cos_loss_bwd(const float* student, const float* teacher, int B, int N, int Ds, int Dt, const float* grad_out,
                          float* grad_student, void* stream) {
    return dad::feat_cos_loss_bwd(student, teacher, B, N, Ds, Dt, grad_out, grad_student, ST(stream));
}

int dad_distill_loss_bwd(const float* student, const float* teacher, int strategy, int num_segments, int B, int64_t L,
                         const float* grad_out, float* grad_student, void* ws, size_t wsb, void* stream) {
    return dad::distill_loss_bwd(student, teacher, strategy, num_segments, B, L, grad_out, grad_student, ws, wsb, ST(stream));
}

int dad_grad_loss_bwd(const float* depth, int B, int H, int W, const float* grad_out, float* grad_depth, void* stream) {
    return dad::grad_loss_bwd(depth, B, H, W, grad_out, grad_depth, ST(stream));
}

int dad_contexts_dp(int level, const float* gt, const uint8_t* mask, int B, int64_t L, uint8_t* ctx_out, void* ws,
                    size_t wsb, void* stream) {
    return dad::contexts_dp(level, gt, mask, B, L, ctx_out, ws, wsb, ST(stream));
}

int dad_contexts_ds(int level, const uint8_t* mask, int B, int H, int W, uint8_t* ctx_out, void* stream) {
    return dad::contexts_ds(level, mask, B, H, W, ctx_out, ST(stream));
}

int dad_hdn_loss_dr(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L,
                    float* out_scalar, double* partials, void* ws, size_t wsb, void* stream) {
    return dad::hdn_loss_dr(level, pred, gt, mask, B, L, out_scalar, partials, ws, wsb, ST(stream));
}

int dad_ssi_hdn_dr_loss(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L, float* out_ssi,
                        float* out_hdn, double* partials_ssi, double* partials_hdn, void* ws, size_t wsb, void* stream) {
    return dad::ssi_hdn_dr_fused(level, pred, gt, mask, B, L, out_ssi, out_hdn, partials_ssi, partials_hdn, ws, wsb, ST(stream));
}

int dad_hdn_loss(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, int64_t L, float* out_scalar,
                 double* partials, void* ws, size_t wsb, void* stream) {
    if (!ctx) return dad::set_error(DAD_ERR_INVALID, "hdn_loss: null contexts");
    return dad::hdn_loss_ctx(pred, gt, ctx, K, B, L, out_scalar, partials, ws, wsb, ST(stream));
}

int dad_grad_loss(const float* depth, int B, int H, int W, float* out_scalar, double* partials, void* ws, size_t wsb,
                  void* stream) {
    return dad::grad_loss(depth, B, H, W, out_scalar, partials, ws, wsb, ST(stream));
}

int dad_feat_cos_loss(const float* s, const float* t, int B, int N, int Ds, int Dt, float* out_scalar, double* partials,
                      void* ws, size_t wsb, void* stream) {
    return dad::feat_cos_loss(s, t, B, N, Ds, Dt, out_scalar, partials, ws, wsb, ST(stream));
}

int dad_distill_loss(const float* student, const float* teacher, int strategy, int num_segments, int B, int64_t L,
                     float* out_scalar, double* partials, float* norm_student, float* norm_teacher, void* ws,
                     size_t wsb, void* stream) {
    return dad::distill_loss(student, teacher, strategy, num_segments, B, L, out_scalar, partials, norm_student,
                             norm_teacher, ws, wsb, ST(stream));
}

int dad_gemm(const void* A, const void* W, const float* bias, float* out, int M, int N, int K, int mode, void* stream) {
    dad::GemmProblem p;
    p.A = A; p.M = M; p.K = K; p.lda = K; p.Wt = W; p.N = N; p.Kp = K;
    p.epi.bias = bias; p.epi.out = out; p.epi.ldc = N;
    return mode == 0 ? dad::gemm_tc(p, ST(stream)) : dad::gemm_simt(p, ST(stream));
}

int dad_gemm_ex(const void* A, const void* W, const float* bias, const float* gamma, const void* res, int res_bf16,
                void* out, int out_bf16, int act, int M, int N, int K, int mode, void* stream) {
    dad::GemmProblem p;
    p.A = A; p.M = M; p.K = K; p.lda = K; p.Wt = W; p.N = N; p.Kp = K;
    p.epi.bias = bias; p.epi.gamma = gamma; p.epi.res1 = res; p.epi.res1_bf16 = res_bf16; p.epi.act = act;
    p.epi.out = out; p.epi.out_bf16 = out_bf16; p.epi.ldc = N;
    return mode == 0 ? dad::gemm_tc(p, ST(stream)) : dad::gemm_simt(p, ST(stream));
}

int dad_gemm_splitk(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int N, int K,
                    int lda, int ksplit, void* stream) {
    dad::GemmProblem p;
    p.A = A; p.M = M; p.K = K; p.lda = lda; p.Wt = W; p.N = N; p.Kp = lda;
    p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = N;
    p.ksplit = ksplit;
    return dad::gemm_tc(p, ST(stream));
}

int dad_gemm_splitk_mn(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int N, int K,
                       int lda, int ldw, int ksplit, void* stream) {
    dad::GemmProblem p;
    p.mn = 1; p.A = A; p.M = M; p.K = K; p.lda = lda; p.Wt = W; p.N = N; p.ldw = ldw;
    p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = N;
    p.ksplit = ksplit;
    return dad::gemm_tc(p, ST(stream));
}

int dad_conv_wgrad(const void* dY, const void* X, const float* zeros, const float* ones, float* out, int B, int H, int W, int Co,
                   int Ci, int ksplit, void* stream) {
    dad::GemmProblem p;
    const int CiP = dad::cdiv(Ci, 128) * 128;
    p.mn = 2; p.A = dY; p.M = Co; p.lda = Co; p.Wt = X; p.ldw = Ci; p.B = B; p.H = H; p.W = W;
    p.shift_rows = Ci; p.shift_ld = CiP; p.N = 9 * CiP;
    p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = 9 * CiP;
    p.ksplit = ksplit;
    return dad::gemm_tc(p, ST(stream));
}

int dad_gemm_shifted(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int rows, int K,
                     int lda, int taps, int ld, const int* offsets, int ksplit, void* stream) {
    if (!offsets || taps < 1 || taps > 9) return dad::set_error(DAD_ERR_INVALID, "dad_gemm_shifted: 1..9 taps with their offsets");
    dad::GemmProblem p;
    p.A = A; p.M = M; p.K = K; p.lda = lda; p.Wt = W; p.N = taps * ld; p.Kp = lda;
    p.shift_taps = taps; p.shift_rows = rows; p.shift_ld = ld;
    for (int t = 0; t < taps; ++t) p.shift_off[t] = offsets[t];   // (all views start at row 0 here)
    p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = p.N;
    p.ksplit = ksplit;
    return dad::gemm_tc(p, ST(stream));
}

int dad_conv_nhwc(const void* in, const void* Wpacked, const float* bias, float* out, int B, int H, int W, int C,
                  int Co, int taps, int mode, void* stream) {
    dad::GemmProblem p;
    p.A = in; p.conv = 1; p.B = B; p.H = H; p.W = W; p.C = C; p.taps = taps; p.ldp = C;
    p.Wt = Wpacked; p.N = Co; p.Kp = taps * dad::cdiv(C, 64) * 64;
    p.epi.bias = bias; p.epi.out = out; p.epi.ldc = Co;
    return mode == 0 ? dad::gemm_tc(p, ST(stream)) : dad::gemm_simt(p, ST(stream));
}

int dad_preprocess_image(const uint8_t* image, int h, int w, int64_t pitch_bytes, int swap_rb, int nh, int nw,
                         const double* mean3, const double* std3, float* out_chw, void* stream) {
    return dad::preprocess_image(image, h, w, pitch_bytes, swap_rb, nh, nw, mean3, std3, out_chw, ST(stream));
}

int dad_resize_depth(const float* in, int B, int H, int W, int h, int w, float* out, void* stream) {
    return dad::resize_depth(in, B, H, W, h, w, out, ST(stream));
}

int dad_colorize_depth(const float* depth, const uint8_t* valid, int B, int64_t HW, float dmin, float dmax, int degenerate,
                       const float* lut, const uint8_t* lut_u8, float* out_chw, uint8_t* out_hwc, void* stream) {
    return dad::colorize_depth(depth, valid, B, HW, dmin, dmax, degenerate, lut, lut_u8, out_chw, out_hwc, ST(stream));
}

int dad_minmax_normalize(const float* in, int B, int64_t L, float* out, void* ws, size_t wsb, void* stream) {
    return dad::minmax_normalize(in, B, L, out, ws, wsb, ST(stream));
}

int dad_conv_nhwc_ex(const void* in, const void* Wpacked, const float* bias, float* out, int B, int H, int W, int C,
                     int Co, int taps, int stride, int mode, void* stream) {
    if (mode != 0 && stride != 1)
        return dad::set_error(DAD_ERR_UNSUPPORTED, "dad_conv_nhwc_ex: strided convolution exists in the tensor-core engine only");
    dad::GemmProblem p;
    p.A = in; p.conv = 1; p.B = B; p.H = H; p.W = W; p.C = C; p.taps = taps; p.ldp = C; p.stride = stride;
    p.Wt = Wpacked; p.N = Co; p.Kp = taps * dad::cdiv(C, 64) * 64;
    p.epi.bias = bias; p.epi.out = out; p.epi.ldc = Co;
    return mode == 0 ? dad::gemm_tc(p, ST(stream)) : dad::gemm_simt(p, ST(stream));
}

int dad_attention(const void* qkv, void* out, int B, int N, int heads, int mode, void* stream) {
    return dad::attention(qkv, out, mode == 0, B, N, heads, ST(stream));
}

int dad_layernorm(const float* in, const float* weight, const float* bias, void* out, float* out_f32, long long rows, int D,
                  int out_period, int in_period, int in_offset, float eps, int mode, void* stream) {
    return dad::layernorm(in, weight, bias, out, mode == 0, out_f32, rows, D, out_period, in_period, in_offset, eps, ST(stream));
}

}  // extern "C"
