// Fused multi-head attention on the 5th-gen tensor cores (sm_100a), head_dim 64, non-causal, no mask
// (reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights).
//
// Round-2 kernel.  Same tiling as attention_tc.cu (CTA = image x head x 128 queries, two CTAs per SM, 64-key tiles,
// S double-buffered in TMEM, P consumed from TMEM as the A operand of the second MMA), rebuilt around what
// tools/softmax_bench.cu measured on a B200 (exp2 results per clock per SM, 2 softmax warps per scheduler):
//     FFMA + MUFU + FADD + CVT per element (round 1)            13.3
//     FFMA2 + MUFU + CVT, half of the pairs on the FMA pipe      19.4     (MUFU alone: 16 by construction)
// so the softmax threads now execute NOTHING but  a = s * log2e - ref  (one packed FFMA2 per pair), the exponential
// (MUFU for some pairs, a packed cubic on the FMA pipe for the others) and the bf16 pack:
//   * the row sum  l = sum_j P  is computed by the TENSOR CORE: a third MMA per key tile multiplies P (TMEM) by a constant
//     all-ones [64 x 16] operand into 16 extra accumulator columns (+12 % tensor-pipe time, which has slack);
//   * there is no running maximum and no rescale in the loop: P is taken relative to the row maximum of the FIRST key
//     tile (floating point keeps full relative precision for P up to 2^127).  If a later score exceeds that reference
//     by more than 127 log2 units the row sum comes out non-finite; the CTA then repeats its work INSIDE the same launch:
//     one pass that only tracks the exact row maximum, one pass that exponentiates against it.  No flag buffer, no second
//     kernel, nothing allocated: exact for any input, capture-safe, and the fast path carries no check at all;
//   * P overwrites its own S columns in place, which frees TMEM for the row-sum accumulator
//     (256 columns per CTA: S0/P0 [0,64) S1/P1 [64,128) | O [128,192) | L [192,208)).
// Roles (192 threads): warps 0-3 softmax (thread = query row), warp 4 TMA producer, warp 5 tcgen05.mma issuer.
#include <cstdlib>

#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;      // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;    // 8 KB
constexpr int ONES_BYTES = 16 * 128;      // [16 "n" rows][64 k] bf16, K-major, all 1.0
constexpr int KV_STAGES = 4;
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128, L_COL = 192;
constexpr int ATT_SMEM = Q_BYTES + 2 * KV_STAGES * KV_BYTES + ONES_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;

enum : int { MODE_FAST = 0, MODE_MAXPASS = 1, MODE_EXACT = 2 };

// exp2 of 32 scores against the reference -> 16 packed bf16 pairs.  PP of every 8 pairs use the FMA-pipe cubic.
template <int PP>
__device__ __forceinline__ void exp32(const uint32_t (&x)[32], uint64_t sc2, uint64_t nref2, uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a0, a1, p0, p1;
        ptx::unpack2(ptx::ffma2(ptx::pack2(__uint_as_float(x[2 * i]), __uint_as_float(x[2 * i + 1])), sc2, nref2), a0, a1);
        if (((i * PP) & 7) < PP) {
            ptx::ex2_fma2(a0, a1, p0, p1);
        } else {
            p0 = ptx::ex2_approx(a0);
            p1 = ptx::ex2_approx(a1);
        }
        pk[i] = ptx::cvt_bf16x2(p0, p1);
    }
}

// zero the packed entries of key columns >= nvalid (last, partial key tile only)
__device__ __forceinline__ void mask16(uint32_t (&pk)[16], int col0, int nvalid) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = col0 + 2 * i;
        pk[i] = c >= nvalid ? 0u : (c + 1 >= nvalid ? (pk[i] & 0xFFFFu) : pk[i]);
    }
}

__device__ __forceinline__ float max32(const uint32_t (&x)[32], int col0, int nvalid) {
    float a = -INFINITY, b = -INFINITY, c = -INFINITY, d = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        a = fmaxf(a, (col0 + i < nvalid) ? __uint_as_float(x[i]) : -INFINITY);
        b = fmaxf(b, (col0 + i + 1 < nvalid) ? __uint_as_float(x[i + 1]) : -INFINITY);
        c = fmaxf(c, (col0 + i + 2 < nvalid) ? __uint_as_float(x[i + 2]) : -INFINITY);
        d = fmaxf(d, (col0 + i + 3 < nvalid) ? __uint_as_float(x[i + 3]) : -INFINITY);
    }
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

template <int PP>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, int N, int D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + Q_BYTES;
    uint8_t* sV = smem + Q_BYTES + KV_STAGES * KV_BYTES;
    uint8_t* sOnes = smem + Q_BYTES + 2 * KV_STAGES * KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                      // [KV_STAGES]
    uint64_t* v_full = bars + 1 + KV_STAGES;          // [KV_STAGES]
    uint64_t* kv_empty = bars + 1 + 2 * KV_STAGES;    // [KV_STAGES]
    uint64_t* s_full = bars + 1 + 3 * KV_STAGES;      // [2]
    uint64_t* p_full = s_full + 2;                    // [2], 128 arrivals
    uint64_t* done = s_full + 4;                      // last P V of a pass retired
    uint64_t* verdict = s_full + 5;                   // softmax -> TMA / MMA warps: repeat the pass loop?
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
    volatile int* again = reinterpret_cast<volatile int*>(tmem_slot + 1);
    int* overflow = reinterpret_cast<int*>(tmem_slot + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int T = (N + BKV - 1) / BKV;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp < 4) {   // all-ones operand of the row-sum MMA (swizzle-invariant), visible to the async proxy
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        ptx::fence_proxy_async_smem();
    }
    if (warp == 5) {
        if (lane == 0) {
            ptx::mbar_init(q_full, 1);
            for (int i = 0; i < KV_STAGES; ++i) {
                ptx::mbar_init(&k_full[i], 1);
                ptx::mbar_init(&v_full[i], 1);
                ptx::mbar_init(&kv_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&s_full[i], 1);
                ptx::mbar_init(&p_full[i], 128);
            }
            ptx::mbar_init(done, 1);
            ptx::mbar_init(verdict, 1);
            *overflow = 0;
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

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
            ptx::tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
            int g = 0;
            for (int pass = 0;; ++pass) {
                for (int j = 0; j < T; ++j, ++g) {
                    const int s = g % KV_STAGES;
                    const uint32_t ph = (g / KV_STAGES) & 1;
                    ptx::mbar_wait(&kv_empty[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&k_full[s], KV_BYTES);
                    ptx::tma_load_3d(sK + s * KV_BYTES, &tmK, &k_full[s], h * HD, j * BKV, b);
                    ptx::mbar_arrive_expect_tx(&v_full[s], KV_BYTES);
                    ptx::tma_load_3d(sV + s * KV_BYTES, &tmV, &v_full[s], h * HD, j * BKV, b);
                }
                ptx::mbar_wait(verdict, pass & 1);
                if (!*again) break;
            }
        }
    } else if (warp == 5) {
        // -------------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
        constexpr uint32_t idesc_l = ptx::make_idesc_bf16(BQ, 16);
        constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
        const uint32_t v_lo0 = ptx::desc_lo_mn_sw128(ptx::smem_u32(sV));
        const uint32_t one_lo = ptx::desc_lo_sw128(ptx::smem_u32(sOnes));
        auto issue_qk = [&](int g) {  // S[g & 1] = Q K_g^T
            const int s = g % KV_STAGES;
            ptx::mbar_wait(&k_full[s], (g / KV_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t k_lo = k_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_bf16(tmem + S_COL + (g & 1) * BKV, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                                   ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(&s_full[g & 1]);
            }
            __syncwarp();
        };
        ptx::mbar_wait(q_full, 0);
        int g0 = 0;
        for (int pass = 0;; ++pass) {
            issue_qk(g0);
            if (T > 1) issue_qk(g0 + 1);
            for (int j = 0; j < T; ++j) {
                const int g = g0 + j;
                const int s = g % KV_STAGES;
                ptx::mbar_wait(&p_full[g & 1], (g >> 1) & 1);       // P_g written in place of S[g & 1]
                ptx::mbar_wait(&v_full[s], (g / KV_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
                const uint32_t tP = tmem + S_COL + (g & 1) * BKV;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)   // O += P V: 16 keys = 16 rows of 128 B per k-step
                        ptx::umma_bf16_ts(tmem + O_COL, tP + k * 8, ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn),
                                          idesc_pv, (j | k) != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)   // L += P 1: the row sums, on the tensor core
                        ptx::umma_bf16_ts(tmem + L_COL, tP + k * 8, ptx::make_desc(one_lo + 2 * k, ptx::kDescHiSw128),
                                          idesc_l, (j | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&kv_empty[s]);                  // K_g / V_g stage free once these retire
                    if (j == T - 1) ptx::umma_commit(done);
                }
                __syncwarp();
                if (j + 2 < T) issue_qk(g + 2);                      // executes after P V_g (in-order pipe): S[g & 1] is free
            }
            g0 += T;
            ptx::mbar_wait(verdict, pass & 1);
            if (!*again) break;
        }
        pdl_launch_dependents();
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-3)
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tO = tmem + lane_base + O_COL, tL = tmem + lane_base + L_COL;
        const uint64_t sc2 = ptx::pack2(LOG2E, LOG2E);
        uint64_t nref2 = 0;
        uint32_t v_lo[32], v_hi[32], pk[16];
        int mode = MODE_FAST;
        int g0 = 0;
        float rmax = -INFINITY, l = 1.f;
        for (int pass = 0;; ++pass) {
            ptx::mbar_wait(&s_full[g0 & 1], (g0 >> 1) & 1);
            ptx::tc_fence_after();
            ptx::tmem_ld_32x32(tS + (g0 & 1) * BKV, v_lo);
            for (int j = 0; j < T; ++j) {
                const int g = g0 + j;
                const int buf = g & 1;
                const int nvalid = min(BKV, N - j * BKV);
                ptx::tmem_ld_wait();                                   // first half of tile j is in registers
                ptx::tmem_ld_32x32(tS + buf * BKV + 32, v_hi);          // second half: in flight during the first exps
                if (mode == MODE_MAXPASS) {
                    ptx::tmem_ld_wait();
                    rmax = fmaxf(rmax, fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid)));
                } else {
                    if (j == 0 && mode == MODE_FAST) {                  // the first tile defines the reference
                        ptx::tmem_ld_wait();
                        const float m = fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid)) * LOG2E;
                        nref2 = ptx::pack2(-m, -m);
                    }
                    exp32<PP>(v_lo, sc2, nref2, pk);
                    if (nvalid < BKV) mask16(pk, 0, nvalid);
                    ptx::tmem_st_32x16(tS + buf * BKV, pk);             // P columns [0,16) <- keys [0,32) (S lo is in registers)
                    ptx::tmem_ld_wait();                                // second half arrived
                    exp32<PP>(v_hi, sc2, nref2, pk);
                    if (nvalid < BKV) mask16(pk, 32, nvalid);
                    ptx::tmem_st_32x16(tS + buf * BKV + 16, pk);
                }
                if (j + 1 < T) {                                       // request the next tile's first half before draining
                    ptx::mbar_wait(&s_full[buf ^ 1], ((g + 1) >> 1) & 1);
                    ptx::tc_fence_after();
                    ptx::tmem_ld_32x32(tS + (buf ^ 1) * BKV, v_lo);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_full[buf]);
            }
            g0 += T;
            ptx::mbar_wait(done, pass & 1);
            ptx::tc_fence_after();
            bool repeat;
            if (mode == MODE_MAXPASS) {
                const float m = rmax * LOG2E;
                nref2 = ptx::pack2(-m, -m);
                mode = MODE_EXACT;
                repeat = true;
            } else {
                uint32_t lv[8];
                ptx::tmem_ld_32x8(tL, lv);
                ptx::tmem_ld_wait();
                l = __uint_as_float(lv[0]);
                repeat = false;
                if (mode == MODE_FAST) {   // a non-finite row sum anywhere in the CTA: redo with the exact maximum
                    const bool bad = !(l < 3.0e38f);
                    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(overflow, 1);
                    ptx::named_bar_sync(1, 128);
                    repeat = *reinterpret_cast<volatile int*>(overflow) != 0;
                    if (repeat) mode = MODE_MAXPASS;
                }
            }
            if (repeat) ptx::tc_fence_before();   // TMEM reads above are complete before the next pass's MMAs overwrite
            if (threadIdx.x == 0) {
                *again = repeat ? 1 : 0;
                __threadfence_block();
            }
            if (repeat) ptx::named_bar_sync(1, 128);   // every softmax thread has read L before the verdict releases the MMAs
            if (threadIdx.x == 0) ptx::mbar_arrive(verdict);
            if (!repeat) break;
        }
        // final: O / l -> bf16 -> global (each thread owns one 128-byte row segment)
        const int row = q0 + warp * 32 + lane;
        const float inv = 1.0f / l;
        bf16* dst = out + (static_cast<long long>(b) * N + row) * D + h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(tO + c * 32, o);
            ptx::tmem_ld_wait();
            if (row < N) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 w;
                    w.x = ptx::cvt_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                    w.y = ptx::cvt_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    w.z = ptx::cvt_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                    w.w = ptx::cvt_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + c * 32 + i) = w;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, TMEM_COLS);
    }
}

template <int PP>
int launch5(const CUtensorMap* tm, bf16* out, int B, int N, int heads, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc5_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    const dim3 grid(cdiv(N, BQ), heads, B);
    DAD_CHECK_CUDA(launch_pdl(attention_tc5_kernel<PP>, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N,
                              heads * HD));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace

// qkv [B*N, 3*D] bf16 (q pre-scaled) -> out [B*N, D] bf16.  poly_pairs = pairs of every 8 on the FMA pipe (0..5).
int attention_tc5(const bf16* qkv, bf16* out, int B, int N, int heads, int poly_pairs, cudaStream_t st) {
    const int D = heads * HD;
    CUtensorMap tm[3];
    for (int i = 0; i < 3; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)3 * D * 2 * N};
        const cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)(i == 0 ? BQ : BKV), 1};
        DAD_TRY(make_tmap_bf16(&tm[i], qkv + static_cast<long long>(i) * D, 3, dims, strides, box));
    }
    switch (poly_pairs) {
        case 0: return launch5<0>(tm, out, B, N, heads, st);
        case 2: return launch5<2>(tm, out, B, N, heads, st);
        case 3: return launch5<3>(tm, out, B, N, heads, st);
        case 5: return launch5<5>(tm, out, B, N, heads, st);
        default: return launch5<4>(tm, out, B, N, heads, st);
    }
}

}  // namespace dad
