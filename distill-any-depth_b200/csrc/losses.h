// Internal C++ interface of the loss kernels (see losses.cu).  All pointers are device pointers;
// `ws` is a caller-owned, 256-byte aligned scratch buffer of at least loss_workspace_bytes(B, K).
#pragma once
#include "common.h"

namespace dad {

size_t loss_workspace_bytes(int B, int K);

int masked_shift_and_scale(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L,
                           float* pred_aligned, float* gt_aligned, void* ws, size_t ws_bytes, cudaStream_t st);
int ssi_loss(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L, float* dense_out,
             float* out_scalar, double* partials, void* ws, size_t ws_bytes, cudaStream_t st);
int hdn_loss_dr(int level, const float* pred, const float* gt, const uint8_t* mask, int B, long long L, float* out_scalar,
                double* partials, void* ws, size_t ws_bytes, cudaStream_t st);
// SSILoss()(pred, gt, mask) and compute_hdn_loss(SSILoss(), pred, gt, get_contexts_dr(level, gt, mask)) in one sweep
// sequence (losses_fused.cu); level 1..3; workspace >= ssi_hdn_fused_workspace_bytes(B) (<= loss_workspace_bytes(B, 8))
size_t ssi_hdn_fused_workspace_bytes(int B);
int ssi_hdn_dr_fused(int level, const float* pred, const float* gt, const uint8_t* mask, int B, long long L, float* out_ssi,
                     float* out_hdn, double* partials_ssi, double* partials_hdn, void* ws, size_t ws_bytes, cudaStream_t st);
int hdn_loss_ctx(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, long long L, float* out_scalar,
                 double* partials, void* ws, size_t ws_bytes, cudaStream_t st);
int contexts_dr(int level, const float* gt, const uint8_t* mask, int B, long long L, uint8_t* ctx_out, void* ws,
                size_t ws_bytes, cudaStream_t st);
// backward w.r.t. pred (gt detached); `gout` = device pointer to the upstream gradient of the scalar loss
int ssi_loss_bwd(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L, const float* gout,
                 float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st);
int hdn_loss_dr_bwd(int level, const float* pred, const float* gt, const uint8_t* mask, int B, long long L, const float* gout,
                    float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st);
int hdn_loss_ctx_bwd(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, long long L, const float* gout,
                     float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st);
int feat_cos_loss_bwd(const float* s, const float* t, int B, int N, int Ds, int Dt, const float* gout, float* grad_s,
                      cudaStream_t st);
int distill_loss_bwd(const float* student, const float* teacher, int strategy, int num_segments, int B, long long L,
                     const float* gout, float* grad_student, void* ws, size_t ws_bytes, cudaStream_t st);
int grad_loss_bwd(const float* depth, int B, int H, int W, const float* gout, float* grad_depth, cudaStream_t st);
// quantile-bin (DP) and spatial-grid (DS) HDN contexts; the HDN loss over them goes through hdn_loss_ctx
int contexts_dp(int level, const float* gt, const uint8_t* mask, int B, long long L, uint8_t* ctx_out, void* ws,
                size_t ws_bytes, cudaStream_t st);
int contexts_ds(int level, const uint8_t* mask, int B, int H, int W, uint8_t* ctx_out, cudaStream_t st);
int grad_loss(const float* depth, int B, int H, int W, float* out_scalar, double* partials, void* ws, size_t ws_bytes,
              cudaStream_t st);
int feat_cos_loss(const float* s, const float* t, int B, int N, int Ds, int Dt, float* out_scalar, double* partials,
                  void* ws, size_t ws_bytes, cudaStream_t st);
// strategy: 0 none, 1 global, 2 hybrid/local
int distill_loss(const float* student, const float* teacher, int strategy, int num_segments, int B, long long L,
                 float* out_scalar, double* partials, float* norm_student, float* norm_teacher, void* ws,
                 size_t ws_bytes, cudaStream_t st);

}  // namespace dad
