/*
 * dad_b200.h - C ABI of libdad_b200.so: the B200-native (sm_100a) implementation of the
 * Distill-Any-Depth hot path.
 *
 * The reference has no FFI layer: its "operator boundary" is a set of Python signatures
 * (SURVEY.md 8b).  Each entry point below names the reference symbol it replaces
 * (paths relative to the reference tree).  The Python host side
 * (distill-any-depth_b200/{dpt,dam,losses}.py) binds these with ctypes and keeps the reference's
 * class / function names, argument meaning and error behaviour.
 *
 * Conventions
 *   - every pointer except names / descriptors is a DEVICE pointer owned by the caller;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream); no entry point synchronises the host except dad_model_prepare();
 *   - no entry point allocates device memory except dad_model_set_weight() / dad_model_prepare();
 *     scratch space is the caller's `workspace` (256-byte aligned for losses, 1024-byte for forward);
 *   - returns 0 on success, negative on error (message: dad_last_error(), thread-local);
 *   - masks / contexts are 1 byte per element (torch.bool layout), non-zero = true;
 *   - `partials` (optional, may be NULL) receives {numerator, denominator} as two doubles so that
 *     ranks of a data-parallel job can all-reduce them and form the full-batch loss
 *     (loss = num / (den + eps); for dad_feat_cos_loss loss = 1 - num / den).
 */
#ifndef DAD_B200_H
#define DAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAD_OK 0
#define DAD_ERR_INVALID (-1)     /* invalid argument (maps to ValueError / AssertionError) */
#define DAD_ERR_UNSUPPORTED (-2) /* configuration outside the hot path (NotImplementedError) */
#define DAD_ERR_CUDA (-3)        /* CUDA runtime / driver error (RuntimeError) */
#define DAD_ERR_WORKSPACE (-4)   /* workspace too small */

#define DAD_MODE_BF16 0 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate / residual / LN / softmax */
#define DAD_MODE_FP32 1 /* fp32 FFMA verification mode (no tensor cores, no TF32) */

const char* dad_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int dad_abi_version(void);

/* ------------------------------------------------------------------ model
 * Replaces DepthAnythingV2.__init__/forward (distillanydepth/depth_anything_v2/dpt.py:187-225)
 * and DepthAnything.forward (distillanydepth/modeling/archs/dam/dam.py:396-419): DINOv2 ViT
 * encoder (dinov2.py:212-321, dinov2_layers/{patch_embed,attention,block,mlp,layer_scale}.py)
 * feeding DPTHead.forward (dpt.py:150-184, util/blocks.py:29-148). */
typedef struct dad_model dad_model;

typedef struct dad_model_desc {
    int embed_dim;       /* 384 / 768 / 1024 (= num_heads * 64)        dinov2.py:339-378 */
    int depth;           /* 12 / 12 / 24 transformer blocks                              */
    int num_heads;       /* 6 / 12 / 16                                                  */
    int taps[4];         /* intermediate_layer_idx                     dpt.py:198-203    */
    int features;        /* DPT head width                             dpt.py:190        */
    int out_channels[4]; /* reassemble widths                          dpt.py:191        */
} dad_model_desc;

int dad_model_create(const dad_model_desc* desc, dad_model** out);
void dad_model_destroy(dad_model* m);
/* Copy one fp32 parameter (student state-dict key, e.g. "pretrained.blocks.0.attn.qkv.weight";
 * nn.Module.load_state_dict, tools/train_distillation.py:771) into the model.
 * Two optional parts of the reference model are selected by the parameters PRESENT (no descriptor field):
 *   "depth_head.readout_projects.{0..3}.0.{weight,bias}"   use_clstoken readout (dpt.py:116-122, 153-156)
 *   "pretrained.blocks.N.mlp.{w12,w3}.{weight,bias}"       SwiGLU FFN of ViT-g instead of mlp.fc1 / fc2
 *                                                          (dinov2_layers/swiglu_ffn.py:13-63, dinov2.py:410)
 * Both are honoured by dad_forward and by dad_forward_train / dad_backward. */
int dad_model_set_weight(dad_model* m, const char* name, const float* dev_ptr, int64_t numel, void* stream);
/* Pack weights for `mode` and build the positional table for (H, W) (interpolate_pos_encoding,
 * dinov2.py:179-210).  Idempotent; must precede dad_forward for that (mode, H, W). */
int dad_model_prepare(dad_model* m, int mode, int H, int W, void* stream);
size_t dad_forward_workspace_bytes(dad_model* m, int B, int H, int W, int mode);
/* x [B,3,H,W] fp32 NCHW -> depth_out [B,1,H,W] fp32, feat_out [B,(H/14)(W/14),D] fp32 (may be NULL). */
int dad_forward(dad_model* m, const float* x, int B, int H, int W, int mode, float* depth_out, float* feat_out,
                void* workspace, size_t workspace_bytes, void* stream);
/* Test hook: copy a named intermediate (as fp32) into dst during the next forwards
 * ("tokens", "block0", "block_last", "layer_rn1".."layer_rn4", "path_4", "path_1"); dst = NULL clears. */
int dad_model_debug_capture(dad_model* m, const char* name, float* dst, int64_t numel);

/* ------------------------------------------------------------------ training forward / backward
 * Replaces `loss.backward()` over the student forward (tools/train_distillation.py:1556-1575; autograd through
 * dpt.py:150-225, dinov2.py:212-321, util/blocks.py:29-148).  fp32 engine only (mode = DAD_MODE_FP32; mode 0 returns
 * DAD_ERR_UNSUPPORTED).  dad_forward_train computes the same outputs as dad_forward and keeps the activations it
 * needs on a tape inside `workspace`; dad_backward, given the SAME workspace (untouched in between) and the upstream
 * gradients of both outputs, ACCUMULATES d loss / d parameter into the fp32 buffers registered per parameter with
 * dad_model_set_grad (student state-dict key; numel must match; dev_ptr = NULL unregisters, i.e. freezes it).
 * Nothing is allocated or synchronised; weights must not change between the two calls. */
int dad_model_set_grad(dad_model* m, const char* name, float* dev_ptr, int64_t numel);
size_t dad_train_workspace_bytes(dad_model* m, int B, int H, int W, int mode);
int dad_forward_train(dad_model* m, const float* x, int B, int H, int W, int mode, float* depth_out, float* feat_out,
                      void* workspace, size_t workspace_bytes, void* stream);
/* grad_depth [B,1,H,W] fp32 (required), grad_feat [B,(H/14)(W/14),D] fp32 (may be NULL = no gradient) */
int dad_backward(dad_model* m, int B, int H, int W, int mode, const float* grad_depth, const float* grad_feat,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ losses
 * All maps are fp32 [rows, L] (rows = B*C images, L = H*W pixels). */
size_t dad_loss_workspace_bytes(int rows, int num_contexts);

/* masked_shift_and_scale(depth_preds, depth_gt, mask_valid)  tools/train_distillation.py:449-533 */
int dad_masked_shift_and_scale(const float* pred, const float* gt, const uint8_t* mask /*NULL = all valid*/, int rows,
                               int64_t L, float* pred_aligned, float* gt_aligned, void* workspace,
                               size_t workspace_bytes, void* stream);
/* SSILoss.forward(pred, gt, mask, dense) = masked_l1_loss(masked_shift_and_scale(...))  :535-542, :675-684.
 * dense_out (may be NULL) receives the [rows, L] loss map; out_scalar (may be NULL) the reduced loss. */
int dad_ssi_loss(const float* pred, const float* gt, const uint8_t* mask, int rows, int64_t L, float* dense_out,
                 float* out_scalar, double* partials, void* workspace, size_t workspace_bytes, void* stream);
/* get_contexts_dr(level, depth_gt, mask_valid) -> bool [2^level - 1, B, L]               :544-576 */
int dad_contexts_dr(int level, const float* gt, const uint8_t* mask, int B, int64_t L, uint8_t* ctx_out,
                    void* workspace, size_t workspace_bytes, void* stream);
/* get_contexts_dp(level, depth_gt, mask_valid) -> bool [2^level - 1, B, L]: bins between the
 * torch.nanquantile (linear interpolation) values of the valid pixels                     :578-644
 * workspace: dad_loss_workspace_bytes(B, 2^level + 2) */
int dad_contexts_dp(int level, const float* gt, const uint8_t* mask, int B, int64_t L, uint8_t* ctx_out,
                    void* workspace, size_t workspace_bytes, void* stream);
/* get_contexts_ds(level, mask_valid) -> bool [1 + 4 + .. + 4^(level-1), B, H, W]: mask AND the
 * n x n grid templates of init_temp_masks_ds (square maps, as upstream); level <= 3     :646-673 */
int dad_contexts_ds(int level, const uint8_t* mask /*NULL = all valid*/, int B, int H, int W, uint8_t* ctx_out,
                    void* stream);
/* compute_hdn_loss(SSILoss(), pred, gt, get_contexts_dr(level, gt, mask)) fused: contexts are never
 * materialised.                                                                          :544-576, :686-707 */
int dad_hdn_loss_dr(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L,
                    float* out_scalar, double* partials, void* workspace, size_t workspace_bytes, void* stream);
/* BOTH of  SSILoss()(pred, gt, mask)  (:449-542)  and  compute_hdn_loss(SSILoss(), pred, gt, get_contexts_dr(level,
 * gt, mask))  (:544-576, :686-707)  from one shared sweep over the maps (the loss half of a training / evaluation step
 * that reports both): the 7 depth-range contexts and the full-mask row share every pass.  level 1..3; outputs /
 * (numerator, denominator) partials may each be NULL.  workspace: dad_loss_workspace_bytes(B, 8). */
int dad_ssi_hdn_dr_loss(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L,
                        float* out_ssi, float* out_hdn, double* partials_ssi, double* partials_hdn, void* workspace,
                        size_t workspace_bytes, void* stream);
/* compute_hdn_loss(SSILoss(), pred, gt, mask_valid_list) with explicit contexts [K, B, L], K <= 21  :686-707 */
int dad_hdn_loss(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, int64_t L, float* out_scalar,
                 double* partials, void* workspace, size_t workspace_bytes, void* stream);
/* gradient_preservation_loss(depth [B,1,H,W])                                            :430-446 */
int dad_grad_loss(const float* depth, int B, int H, int W, float* out_scalar, double* partials, void* workspace,
                  size_t workspace_bytes, void* stream);
/* feature_distillation_loss(student [B,N,Ds], teacher [B,N,Dt]) tensor branch, equal N    :284-413 */
int dad_feat_cos_loss(const float* student, const float* teacher, int B, int N, int Ds, int Dt, float* out_scalar,
                      double* partials, void* workspace, size_t workspace_bytes, void* stream);
/* distillation_loss(student, teacher, norm_strategy, num_segments)                        :173-282
 * strategy: 0 'none', 1 'global', 2 'hybrid' / 'local'.  norm_student / norm_teacher (may be NULL)
 * receive the normalised maps (global_normalize / hybrid_normalize). */
int dad_distill_loss(const float* student, const float* teacher, int strategy, int num_segments, int B, int64_t L,
                     float* out_scalar, double* partials, float* norm_student, float* norm_teacher, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ loss gradients (first slice)
 * d loss / d pred of the scalar losses above, with the target map detached (it is the no_grad teacher output at the
 * training call sites, tools/train_distillation.py:1512-1553).  `grad_out` is a DEVICE pointer to the upstream
 * gradient of the scalar; the statistics (medians, scales) are recomputed, so no state is carried from the forward.
 * The semantics are PyTorch autograd's on the reference code: d median / d p lives on the selected element (lowest
 * index on ties), |x| has gradient 0 at 0, masked elements get 0.  Workspace as for the forward call. */
int dad_ssi_loss_bwd(const float* pred, const float* gt, const uint8_t* mask, int rows, int64_t L, const float* grad_out,
                     float* grad_pred, void* workspace, size_t workspace_bytes, void* stream);
int dad_hdn_loss_dr_bwd(int level, const float* pred, const float* gt, const uint8_t* mask, int B, int64_t L,
                        const float* grad_out, float* grad_pred, void* workspace, size_t workspace_bytes, void* stream);
int dad_hdn_loss_bwd(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, int64_t L, const float* grad_out,
                     float* grad_pred, void* workspace, size_t workspace_bytes, void* stream);
int dad_grad_loss_bwd(const float* depth, int B, int H, int W, const float* grad_out, float* grad_depth, void* stream);
/* feature_distillation_loss: gradient w.r.t. the student features [B,N,Ds] (channels dropped by the nearest resize get 0) */
int dad_feat_cos_loss_bwd(const float* student, const float* teacher, int B, int N, int Ds, int Dt, const float* grad_out,
                          float* grad_student, void* stream);
/* distillation_loss: gradient w.r.t. its FIRST map (swap the arguments for the second: the loss is symmetric).
 * Segment / median statistics of both maps are recomputed; each map is normalised by its own statistics. */
int dad_distill_loss_bwd(const float* student, const float* teacher, int strategy, int num_segments, int B, int64_t L,
                         const float* grad_out, float* grad_student, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ pre- / post-processing
 * DepthAnythingV2.image2tensor (depth_anything_v2/dpt.py:237-262; util/transform.py:109-148): uint8 HWC image
 * (device pointer, `pitch_bytes` per row; swap_rb = 1 for a BGR source) -> /255 -> cv2.resize(INTER_CUBIC) to
 * (nh, nw) -> (x - mean) / std -> CHW fp32 [3, nh, nw].  mean3 / std3 are HOST pointers to 3 doubles. */
int dad_preprocess_image(const uint8_t* image, int h, int w, int64_t pitch_bytes, int swap_rb, int nh, int nw,
                         const double* mean3, const double* std3, float* out_chw, void* stream);
/* F.interpolate(depth, (h, w), mode="bilinear", align_corners=True) on [B,1,H,W] fp32 (dpt.py:233). */
int dad_resize_depth(const float* in, int B, int H, int W, int h, int w, float* out, void* stream);
/* per image (d - min) / (max - min) (tools/testers/infer.py:135); workspace: 8 bytes per image. */
int dad_minmax_normalize(const float* in, int B, int64_t L, float* out, void* workspace, size_t workspace_bytes,
                         void* stream);
/* colorize_depth_maps (distillanydepth/utils/image_util.py:69-118) + the uint8 HWC conversion of
 * tools/testers/infer.py:137-140: depth [B, HW] fp32 -> x = clip((d - dmin) / (dmax - dmin), 0, 1) (d * 0 when
 * `degenerate`, i.e. dmin == dmax) -> lut[min(int(x * 256), 255)] (matplotlib's Colormap.__call__; `lut` = device
 * pointer to 256 x 3 fp32) -> pixels with valid == 0 set to 0 (valid may be NULL) -> out_chw [B,3,HW] fp32 and / or
 * out_hwc [B,HW,3] uint8 = lut_u8[index] where lut_u8 (256 x 3 bytes, device) holds (float64 table * 255) truncated, i.e.
 * numpy's `(rgb * 255).astype(np.uint8)` bit for bit; either output may be NULL (lut_u8 may be NULL without out_hwc). */
int dad_colorize_depth(const float* depth, const uint8_t* valid, int B, int64_t HW, float dmin, float dmax, int degenerate,
                       const float* lut, const uint8_t* lut_u8, float* out_chw, uint8_t* out_hwc, void* stream);

/* ------------------------------------------------------------------ kernel-level test entry points
 * out[M,N] (fp32) = A[M,K] (bf16 bits / fp32) * W[N,K]^T through the tcgen05 (mode 0) or FFMA (mode 1)
 * engine with an optional bias[N]; used by the parity tests to bisect the GEMM engine alone. */
int dad_gemm(const void* A, const void* W, const float* bias, float* out, int M, int N, int K, int mode, void* stream);
/* Same with the fused epilogues of the encoder: out = [res +] [gamma *] act(A W^T + bias); out / res are fp32
 * or bf16 (out_bf16 / res_bf16); act: 0 none, 1 GELU(erf), 2 ReLU.  res may alias out (in-place residual). */
int dad_gemm_ex(const void* A, const void* W, const float* bias, const float* gamma, const void* res, int res_bf16,
                void* out, int out_bf16, int act, int M, int N, int K, int mode, void* stream);
/* Weight-gradient GEMM of the bf16 backward: out[M,N] (fp32) += A[M,K] W[N,K]^T with both bf16 operands K-major with row
 * pitch `lda` elements (zero beyond K); the K loop is split over `ksplit` work items per tile whose fp32 partial tiles
 * are reduce-added through TMA.  zeros / ones: device vectors of N floats (the epilogue's bias / scale). */
int dad_gemm_splitk(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int N, int K,
                    int lda, int ksplit, void* stream);
/* The same weight-gradient GEMM with MN-MAJOR operands, i.e. read in the layout the activations already have:
 * out[M,N] (fp32) += sum_k A[k, m] W[k, n], A = [K rows][lda] (M valid columns), W = [K rows][ldw] (N valid columns), bf16:
 * dW = dY^T X straight from dY [tokens, Nout] and X [tokens, Kin], no transposed copies (train.inl wgrad_linear). */
int dad_gemm_splitk_mn(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int N, int K,
                       int lda, int ldw, int ksplit, void* stream);
/* 3x3 / stride-1 / zero-padded convolution weight gradient straight from the NHWC bf16 tensors dY [B,H,W,Co] and X [B,H,W,Ci]
 * (autograd of torch.nn.Conv2d for the decoder convolutions, blocks.py / dpt.py): out[co, t * CiP + ci] (fp32, CiP = Ci
 * rounded up to 128, t = ky * 3 + kx) += sum_pixels dY[p, co] X[p + (ky - 1, kx - 1), ci]; the contraction runs over 8 x 8
 * pixel patches fetched by TMA (zero fill = the padding); no im2col operand, no transposes (train.inl conv_wgrad). */
int dad_conv_wgrad(const void* dY, const void* X, const float* zeros, const float* ones, float* out, int B, int H, int W, int Co,
                   int Ci, int ksplit, void* stream);
/* An earlier form of the 3x3 convolution weight gradient's GEMM (DAD_WGRAD_PATH=1): the B operand's rows are `taps` SHIFTED VIEWS of one
 * [rows, lda] matrix W: out[m, t * ld + r] (fp32) += sum_k A[m, k] * W[r, k + offsets[t]] (zero outside [0, lda) and for
 * r >= rows); ld % 128 == 0, ksplit >= 2, offsets is a HOST array.  Over zero-padded pixel space a convolution tap is such a
 * constant offset, so no im2col operand is materialised (train.inl conv_wgrad; reference: autograd of torch.nn.Conv2d,
 * dpt.py / blocks.py convolutions). */
int dad_gemm_shifted(const void* A, const void* W, const float* zeros, const float* ones, float* out, int M, int rows, int K,
                     int lda, int taps, int ld, const int* offsets, int ksplit, void* stream);
/* conv3x3 / 1x1 (stride 1, zero padding) on NHWC input [B,H,W,C] with packed weights [Co, taps*Cp]. */
int dad_conv_nhwc(const void* in, const void* Wpacked, const float* bias, float* out, int B, int H, int W, int C,
                  int Co, int taps, int mode, void* stream);
/* Same with a stride (1 or 2; 3x3 "same" padding): out is [B, (H-1)/stride+1, (W-1)/stride+1, Co].  Stride 2 is the
 * resize_layers[3] convolution (dpt.py:101-106), an implicit GEMM whose TMA box walks the input with element stride 2. */
int dad_conv_nhwc_ex(const void* in, const void* Wpacked, const float* bias, float* out, int B, int H, int W, int C,
                     int Co, int taps, int stride, int mode, void* stream);
/* attention over qkv [B*N, 3*heads*64] (q pre-scaled) -> out [B*N, heads*64]; bf16 bits (mode 0) or fp32. */
int dad_attention(const void* qkv, void* out, int B, int N, int heads, int mode, void* stream);
/* LayerNorm over the last dimension (dinov2.py:213-214 norm1 / norm2 of every block, :304-305 the shared final norm;
 * torch.nn.LayerNorm, biased variance, eps inside the square root).  in: fp32 rows of D; output row r reads input row
 * (r / out_period) * in_period + in_offset + r % out_period (the final norm drops the class token this way).  out (bf16 bits
 * in mode 0, fp32 in mode 1) and / or out_f32 may be NULL.  Kernel-level test entry (the model calls the same function). */
int dad_layernorm(const float* in, const float* weight, const float* bias, void* out, float* out_f32, long long rows, int D,
                  int out_period, int in_period, int in_offset, float eps, int mode, void* stream);

/* ------------------------------------------------------------------ measurement hooks (bench.py)
 * Launch counter of this library's kernels, and optional per-kernel-class CUDA-event timing
 * (class: 0 gemm_tc, 1 gemm_simt, 2 attention, 3 layernorm, 4 elementwise, 5 loss; `work` = algorithmic FLOPs
 * for classes 0-2, algorithmic bytes for 3-5).  Enabling clears previous records. */
long long dad_launch_count(void);
void dad_profile_enable(int on);
int dad_profile_get(int cls, double* ms, double* work, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* DAD_B200_H */
