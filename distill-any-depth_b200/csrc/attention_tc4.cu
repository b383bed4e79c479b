// EXPERIMENTAL (DAD_ATT_VARIANT=4; NOT the default and NOT yet measured on hardware - written at the end of round 1 from
// the ncu evidence in DESIGN.md section 10, to be validated first thing in round 2).
//
// Fused multi-head attention, head_dim 64, non-causal (reference dinov2_layers/attention.py:49-62), same tiling and TMEM
// layout as attention_tc.cu (CTA = image x head x 128 queries, 2 CTAs / SM, S and P double-buffered in TMEM, P consumed from
// TMEM by the second MMA) with two changes aimed at the idle MUFU time (ncu: XU pipe 68 % busy, two softmax warps per
// scheduler):
//   1. TWO softmax warpgroups per CTA.  Warps 0-3 exponentiate columns [0, 32) of every 128 x 64 score tile, warps 4-7
//      columns [32, 64) (a warp may touch TMEM lanes 32 * (warp % 4) .. + 31, so both groups see the same rows).  Each
//      scheduler then has FOUR independent softmax warps (2 CTAs x 2 groups) instead of two, each holding 32 scores.
//   2. No in-kernel rescale.  P = exp2(s * log2e - m_ref) is taken relative to the row maximum of the FIRST key tile and
//      never rebased: bf16 / fp32 carry an 8-bit exponent, so P up to 2^60 (a later score 41 nats above the first tile's
//      maximum) keeps full relative precision in P, l and O.  A row that exceeds that (or produces a non-finite sum) raises
//      a per-CTA flag; the launcher then runs attention_tc3 (exact per-tile maximum) on the flagged CTAs only, so the result
//      is exact for any input while the two warpgroups of a row never have to agree on a new reference mid-stream.
//      The two groups exchange the first tile's half-row maxima and, at the end, their partial row sums through shared
//      memory (one named barrier each).
//   warp 8 = TMA producer, warp 9 = tcgen05.mma issuer (as warps 4 / 5 of attention_tc.cu).
#include <cstdlib>
#include <type_traits>

#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

int attention_tc3_flagged(const bf16* qkv, bf16* out, int B, int N, int heads, const uint8_t* only_if, cudaStream_t st);

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64, HALF = 32;
constexpr int Q_BYTES = BQ * HD * 2;      // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;    // 8 KB
constexpr int KV_STAGES = 4;
constexpr int SOFTMAX_THREADS = 256;
constexpr int ATT_THREADS = SOFTMAX_THREADS + 64;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, P_COL = 128, O_COL = 192;  // S0 [0,64) S1 [64,128) | P0 [128,160) P1 [160,192) | O [192,256)
constexpr int BAR_BYTES = 1024;
constexpr int XCH_BYTES = 2 * 2 * BQ * 4 + 64;      // half-row maxima [2][128], partial sums [2][128], flag
constexpr int ATT_SMEM = Q_BYTES + 2 * KV_STAGES * KV_BYTES + BAR_BYTES + XCH_BYTES + 1024;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_LIMIT = 1.152921504606846976e18f;  // 2^60

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, uint8_t* __restrict__ flags, int N, int D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + Q_BYTES;
    uint8_t* sV = smem + Q_BYTES + KV_STAGES * KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + 2 * KV_STAGES * KV_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                      // [KV_STAGES]
    uint64_t* v_full = bars + 1 + KV_STAGES;          // [KV_STAGES]
    uint64_t* kv_empty = bars + 1 + 2 * KV_STAGES;    // [KV_STAGES]
    uint64_t* s_full = bars + 1 + 3 * KV_STAGES;      // [2]
    uint64_t* p_full = s_full + 2;                    // [2], 256 arrivals each
    uint64_t* done = s_full + 4;                      // last P V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);
    float* sMax = reinterpret_cast<float*>(smem + Q_BYTES + 2 * KV_STAGES * KV_BYTES + BAR_BYTES);   // [2][BQ]
    float* sSum = sMax + 2 * BQ;                                                                     // [2][BQ]
    int* sBad = reinterpret_cast<int*>(sSum + 2 * BQ);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int T = (N + BKV - 1) / BKV;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp == 9) {
        if (lane == 0) {
            ptx::mbar_init(q_full, 1);
            for (int i = 0; i < KV_STAGES; ++i) {
                ptx::mbar_init(&k_full[i], 1);
                ptx::mbar_init(&v_full[i], 1);
                ptx::mbar_init(&kv_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&s_full[i], 1);
                ptx::mbar_init(&p_full[i], SOFTMAX_THREADS);
            }
            ptx::mbar_init(done, 1);
            *sBad = 0;
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();  // prologue above overlaps the previous kernel's tail; global memory is touched only below

    if (warp == 8) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
            ptx::tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % KV_STAGES;
                const uint32_t ph = (j / KV_STAGES) & 1;
                ptx::mbar_wait(&kv_empty[s], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&k_full[s], KV_BYTES);
                ptx::tma_load_3d(sK + s * KV_BYTES, &tmK, &k_full[s], h * HD, j * BKV, b);
                ptx::mbar_arrive_expect_tx(&v_full[s], KV_BYTES);
                ptx::tma_load_3d(sV + s * KV_BYTES, &tmV, &v_full[s], h * HD, j * BKV, b);
            }
        }
    } else if (warp == 9) {
        // -------------------------------------------------------------------- MMA issuer (as attention_tc.cu)
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
        constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
        const uint32_t v_lo0 = ptx::desc_lo_mn_sw128(ptx::smem_u32(sV));
        auto issue_qk = [&](int j) {  // S[j & 1] = Q K_j^T
            const int s = j % KV_STAGES;
            ptx::mbar_wait(&k_full[s], (j / KV_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t k_lo = k_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_bf16(tmem + S_COL + (j & 1) * BKV, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                                   ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        ptx::mbar_wait(q_full, 0);
        issue_qk(0);
        if (T > 1) issue_qk(1);
        for (int j = 0; j < T; ++j) {
            const int s = j % KV_STAGES;
            ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1);       // both halves of P_j written, S[j & 1] drained
            ptx::mbar_wait(&v_full[s], (j / KV_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)   // 16 keys = 16 rows of 128 B per k-step
                    ptx::umma_bf16_ts(tmem + O_COL, tmem + P_COL + (j & 1) * (BKV / 2) + k * 8,
                                      ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn), idesc_pv, (j | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&kv_empty[s]);                  // K_j / V_j stage free once P V_j retires
                if (j == T - 1) ptx::umma_commit(done);
            }
            __syncwarp();
            if (j + 2 < T) issue_qk(j + 2);
        }
        pdl_launch_dependents();
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-7)
        const int hf = warp >> 2, quarter = warp & 3;           // column half, TMEM lane quarter
        const int rloc = quarter * 32 + lane;                    // row inside the query tile
        const int col0 = hf * HALF;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL + col0;
        const uint32_t tP = tmem + lane_base + P_COL + hf * (HALF / 2);
        const uint32_t tO = tmem + lane_base + O_COL + col0;
        float m_ref = 0.f, l = 0.f;
        bool bad = false;
        uint32_t v[HALF];

        // exp2 of this thread's 32 columns against m_ref -> 16 packed bf16 pairs; returns the partial row sum
        auto exp32 = [&](auto masked, int nvalid, uint32_t (&pk)[16]) -> float {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < HALF; i += 2) {
                float p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[i]), LOG2E, -m_ref));
                float p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[i + 1]), LOG2E, -m_ref));
                if (decltype(masked)::value) {
                    if (col0 + i >= nvalid) p0 = 0.f;
                    if (col0 + i + 1 >= nvalid) p1 = 0.f;
                }
                s0 += p0;
                s1 += p1;
                __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
            }
            return s0 + s1;
        };

        ptx::mbar_wait(&s_full[0], 0);
        ptx::tc_fence_after();
        ptx::tmem_ld_32x32(tS, v);
        for (int j = 0; j < T; ++j) {
            const int buf = j & 1;
            const int nvalid = min(BKV, N - j * BKV);
            ptx::tmem_ld_wait();
            if (j == 0) {
                // reference maximum = row maximum of the first key tile: the two halves exchange theirs once
                float a = -INFINITY, c = -INFINITY;
#pragma unroll
                for (int i = 0; i < HALF; i += 2) {
                    a = fmaxf(a, (col0 + i < nvalid) ? __uint_as_float(v[i]) : -INFINITY);
                    c = fmaxf(c, (col0 + i + 1 < nvalid) ? __uint_as_float(v[i + 1]) : -INFINITY);
                }
                sMax[hf * BQ + rloc] = fmaxf(a, c);
                ptx::named_bar_sync(1, SOFTMAX_THREADS);
                m_ref = fmaxf(sMax[rloc], sMax[BQ + rloc]) * LOG2E;   // column 0 is always valid, so this is finite
            }
            uint32_t pk[16];
            const float lt = (nvalid == BKV) ? exp32(std::false_type{}, nvalid, pk) : exp32(std::true_type{}, nvalid, pk);
            ptx::tmem_st_32x16(tP + buf * (BKV / 2), pk);
            bad |= !(lt <= P_LIMIT);   // a score > 2^60 above the reference (or a non-finite sum): leave this CTA to the exact kernel
            l += lt;
            if (j + 1 < T) {           // request the next tile's columns before draining the stores
                ptx::mbar_wait(&s_full[buf ^ 1], ((j + 1) >> 1) & 1);
                ptx::tc_fence_after();
                ptx::tmem_ld_32x32(tS + (buf ^ 1) * BKV, v);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&p_full[buf]);
        }
        // row sum = sum of the two halves; flag the CTA if any row left the safe range
        sSum[hf * BQ + rloc] = l;
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(sBad, 1);
        ptx::named_bar_sync(1, SOFTMAX_THREADS);
        const float inv = 1.0f / (sSum[rloc] + sSum[BQ + rloc]);
        if (flags && threadIdx.x == 0)
            flags[(static_cast<long long>(b) * gridDim.y + h) * gridDim.x + blockIdx.x] = *sBad ? 1 : 0;
        // final: this thread's 32 columns of O / l -> bf16 -> global (64 contiguous bytes of the row)
        ptx::mbar_wait(done, 0);
        ptx::tc_fence_after();
        const int row = q0 + rloc;
        bf16* dst = out + (static_cast<long long>(b) * N + row) * D + h * HD + col0;
        uint32_t o[32];
        ptx::tmem_ld_32x32(tO, o);
        ptx::tmem_ld_wait();
        if (row < N) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 w;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
                w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + i) = w;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, TMEM_COLS);
    }
}

}  // namespace

// qkv [B*N, 3*D] bf16 (q pre-scaled) -> out [B*N, D] bf16.  `flags`: one byte per CTA (B * heads * ceil(N / 128)),
// device memory owned by the caller; CTAs whose rows left the 2^60 range are recomputed by attention_tc3.
int attention_tc4(const bf16* qkv, bf16* out, int B, int N, int heads, uint8_t* flags, cudaStream_t st) {
    DAD_REQUIRE(flags, "attention_tc4: needs a flag buffer (one byte per CTA)");
    const int D = heads * HD;
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    CUtensorMap tm[3];
    for (int i = 0; i < 3; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)3 * D * 2 * N};
        const cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)(i == 0 ? BQ : BKV), 1};
        DAD_TRY(make_tmap_bf16(&tm[i], qkv + static_cast<long long>(i) * D, 3, dims, strides, box));
    }
    const dim3 grid(cdiv(N, BQ), heads, B);
    DAD_CHECK_CUDA(launch_pdl(attention_tc4_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, flags, N, D));
    DAD_CHECK_LAUNCH();
    // exactness net: the flagged CTAs (none in practice) are recomputed with the per-tile-maximum kernel
    return attention_tc3_flagged(qkv, out, B, N, heads, flags, st);
}

}  // namespace dad
