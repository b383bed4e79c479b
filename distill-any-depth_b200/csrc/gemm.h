// Internal interface of the GEMM / implicit-GEMM-convolution engine.
//
//   D[row, col] = epilogue( sum_k A[row, k] * Wt[col, k] )
//
// A is either a plain row-major matrix ("linear") or an NHWC activation tensor
// read as a 3x3 / 1x1 stride-1 zero-padded convolution window ("conv").  Wt is
// the K-major weight matrix ([N, Kp]); for conv, Kp = taps * cchunks * 64 with
// each tap's channels zero-padded to a multiple of 64.
//
// Two engines implement it:
//   gemm_tc   - bf16 operands, tcgen05.mma + TMEM accumulators, TMA-fed (bf16 mode)
//   gemm_simt - fp32 operands, FFMA, gather loads (fp32 verification mode)
// Both share the epilogue in epilogue.cuh.
#pragma once
#include "common.h"

namespace dad {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct Epilogue {
    const float* bias = nullptr;     // [N] (indexed by output channel)
    const float* gamma = nullptr;    // [N] LayerScale, applied after bias/act
    const float* rowtab = nullptr;   // [rowtab_period, N] table added per (row % period)
    int rowtab_period = 0;
    int act = ACT_NONE;
    const void* res1 = nullptr;      // residual inputs, same indexing as out
    const void* res2 = nullptr;
    int res1_bf16 = 0, res2_bf16 = 0;
    void* out = nullptr;
    int out_bf16 = 0;
    long long ldc = 0;               // elements between output rows
    void* out_relu = nullptr;        // optional copy with ReLU applied (same dtype / indexing as out)
    // ConvTranspose k=s (non-overlapping): column = (ky*k + kx) * CoP + co,
    // input row = (b, y, x) over scat_H x scat_W; out[b, k*y+ky, k*x+kx, co].
    int scat_k = 0, scat_CoP = 0, scat_Co = 0, scat_H = 0, scat_W = 0;
    // Fused output head (N == 32): out_head[row] = relu(dot(relu(acc+bias), head_w) + head_b)
    const float* head_w = nullptr;
    const float* head_b = nullptr;   // device pointer (a by-value scalar would be baked into captured graphs)
    float* head_out = nullptr;
};

struct GemmProblem {
    // A operand
    const void* A = nullptr;  // bf16 (tc) or fp32 (simt)
    int conv = 0;             // 0 linear, 1 conv
    // linear
    int M = 0, K = 0;
    long long lda = 0;        // elements between rows
    // conv (stride 1, "same" zero padding): NHWC with `ldp` elements between pixels
    int B = 0, H = 0, W = 0, C = 0, taps = 1;
    long long ldp = 0;
    int stride = 1;           // 3x3 only: output is ((H - 1) / stride + 1) x ((W - 1) / stride + 1) (gemm_tc; stride 1 or 2)
    // weights
    const void* Wt = nullptr; // [N, Kp] K-major; bf16 (tc) or fp32 (simt)
    int N = 0;
    int Kp = 0;               // padded K (row length of Wt)
    int ksplit = 1;           // gemm_tc, linear, `out += gamma * (acc + bias)` epilogue only: split the K loop over this many
                              // work items per tile, each reduce-adding its partial product (weight-gradient GEMMs: K = tokens)
    // gemm_tc, linear, generic epilogue only: a BATCH of independent problems z = b * batch_h + h (attention backward: one per
    // image and head).  M / N / K / Kp are per problem; operand element strides per head / image; Wt has `w_rows` valid rows
    // (<= N; rows beyond are zero-filled, so N can be padded).  Output element (row, col) of problem (b, h) goes to
    // out[(b * c_row_b + h * c_row_h + row) * ldc + h * c_col_h + col].
    int batch_h = 0, batch_b = 0;
    long long a_sh = 0, a_sb = 0, w_sh = 0, w_sb = 0;
    int w_rows = 0;
    long long ldw = 0;        // batched only: elements between rows of Wt (0 = Kp)
    // gemm_tc, linear split-K only: the N rows of the B operand are `shift_taps` SHIFTED VIEWS of one [shift_rows, Kp] matrix
    // Wt: output column n = tap * shift_ld + r reads row shift_row[tap] + r of Wt at K coordinate k + shift_off[tap] (rows >=
    // shift_rows and coordinates outside [0, Kp) read as zero).  N = shift_taps * shift_ld, shift_ld % 128 == 0, every
    // shift_off a multiple of 8 (TMA needs a 16-byte aligned box start in the contiguous dimension - measured: an odd offset
    // raises an illegal-instruction fault).  This is the 3x3 convolution weight gradient over zero-padded pixel space, where
    // a vertical tap is a constant offset (train.inl conv_wgrad).
    int shift_taps = 0, shift_rows = 0, shift_ld = 0;
    int shift_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int shift_row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    // gemm_tc, split-K only: MN-MAJOR operands - the contraction index is the OUTER dimension of both operands, i.e. they
    // are read in the layout the activations already have (no transposed copies for the weight gradients):
    //   mn == 1  linear: A is [K rows][lda] with M valid columns, Wt is [K rows][ldw] with N valid columns
    //            (D[m, n] = sum_k A[k, m] Wt[k, n]: dW = dY^T X straight from dY [rows, Nout] and X [rows, Kin]);
    //   mn == 2  3x3 / stride-1 convolution weight gradient: A = dY, Wt = X, both NHWC [B, H, W, C] (pixel pitches lda /
    //            ldw), M = Co, shift_rows = Ci, shift_ld = Ci rounded up to 128, N = 9 * shift_ld; the contraction runs over
    //            8 x 8 pixel patches, tap t = n / shift_ld reads X shifted by (t / 3 - 1, t % 3 - 1) with TMA's zero fill as
    //            the convolution padding: out[co, t * shift_ld + ci] += sum_p dY[p, co] X[p + shift_t, ci].
    //   mn == 3  as 1, but only Wt is MN-major; A stays K-major ([M rows][lda]) - dQ = dS K of the attention backward.
    // mn 1 / 3 also take a BATCH of problems (batch_h > 0, generic epilogue, no split-K): the strides a_sh / a_sb / w_sh / w_sb
    // are then between the heads / images of the [rows][ld] operands and w_rows is the number of valid Wt COLUMNS.
    int mn = 0;
    long long c_row_b = 0, c_row_h = 0;
    int c_col_h = 0;
    // batched + the fp32 `out += gamma * (acc + bias)` epilogue only: c_store = 1 writes each tile with a plain TMA STORE clipped
    // to its problem's M rows (3-D output map) instead of the reduce-add - the destination needs no zero fill and is not
    // read.  Needs contiguous problems (c_row_b == batch_h * c_row_h == batch_h * M) and one work item per output tile.
    int c_store = 0;
    Epilogue epi;
};

int gemm_tc(const GemmProblem& p, cudaStream_t stream);
// 2-CTA (cta_group::2) kernel for wide linear GEMMs with the encoder epilogues; gemm_tc() dispatches to it.
bool gemm_tc2_eligible(const GemmProblem& p);
int gemm_tc2(const GemmProblem& p, cudaStream_t stream);
// 2-CTA halo-mode 3x3 convolution for wide N (conv_tc2.cu); gemm_tc() dispatches to it.
bool conv_tc2_eligible(const GemmProblem& p);
int conv_tc2(const GemmProblem& p, cudaStream_t stream);
int gemm_simt(const GemmProblem& p, cudaStream_t stream);

}  // namespace dad
