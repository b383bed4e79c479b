// Fused multi-head attention on the 5th-gen tensor cores (sm_100a), head_dim 64, non-causal, no mask
// (reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights).
//
// Round-2 kernel: PERSISTENT CTAs (two per SM), work item = (image, head, 128 queries), 64-key tiles.  Built around
// three measurements taken on a B200 (profiles/attention_r2.md):
//  (1) tools/softmax_bench.cu - exp2 results per clock per SM, 2 softmax warps per scheduler:
//        FFMA + MUFU + FADD + CVT per element (round 1)          13.3
//        FFMA2 + MUFU + CVT, part of the pairs on the FMA pipe   17.6 - 19.4   (MUFU alone: 16 by construction)
//      so the softmax threads execute nothing but  a = s * log2e - ref  (one packed FFMA2 per pair), the exponential
//      (MUFU for some pairs, a packed cubic on the FMA pipe for the others) and the bf16 pack:
//        * the row sum  l = sum_j P  is computed by the TENSOR CORE, in the SAME MMA as O: the B operand is [V | 1]
//          (N = 80; the all-ones tile is the second 64-column block of the MN-major operand, reached through the
//          descriptor's leading-dimension offset), so [O | L] are 80 adjacent accumulator columns;
//        * no running maximum and no rescale in the loop: P is taken relative to the row maximum of the FIRST key tile
//          (floating point keeps full relative precision for P up to 2^127).  A later score more than 127 log2 units
//          above that reference makes the row sum non-finite; the work item is then put on the CTA's redo list and
//          recomputed at the end of the same launch with the exact row maximum (one max-only pass + one exact pass).
//          No flag buffer, no second kernel, nothing allocated, and the fast path carries no check at all.
//  (2) ncu source page of the one-CTA-per-item version: 21 % of the softmax warps' time went to per-CTA start-up / drain
//      -> persistent CTAs whose TMA / MMA warps stream straight into the next item (Q double-buffered).
//      The next score tile's barrier is probed with a non-blocking test_wait before the second half of the exponentials,
//      so the ~90-clock latency of a successful try_wait is off the critical path.
//  (3) two further restructurings were built and measured and are NOT used (profiles/attention_r2.md): releasing the score
//      buffer as soon as S is in registers (Q K^T of tile g+2 issued at the start of tile g) with P in its own
//      single / double buffer ran 0.42 - 0.43 ms per ViT-L launch against 0.333 ms for this kernel.
//  (4) ncu source pages (attention_tc7's and this kernel's): the MMA-issuing warp is busy ~80 % of a tile - its
//      instruction stream (waits, descriptor set-up, 12 tiny MMAs, commits) is co-critical with the softmax warps.  Hence
//      8 MMAs per tile instead of 12 (the [V | 1] operand above), ONE barrier per K / V stage, all waits of a tile before
//      one elected issue block, and the last two tiles of an item peeled out of the steady loop: 0.324 -> 0.301 ms.
// TMEM (256 columns per CTA): S0/P0 [0,64) S1/P1 [64,128) | O [128,192) | L [192,208): P is written in place of S.
// Roles (192 threads): warps 0-3 softmax (thread = query row), warp 4 TMA producer, warp 5 tcgen05.mma issuer.
#include <cstdlib>
#include <type_traits>

#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;      // 16 KB (x2: the next item's Q is prefetched)
constexpr int KV_BYTES = BKV * HD * 2;    // 8 KB
constexpr int ONES_BYTES = 16 * 128;      // [16 "n" rows][64 k] bf16, K-major, all 1.0
constexpr int KV_STAGES = 4;
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128, L_COL = 192;
static_assert(L_COL == O_COL + HD, "the row sums are columns 64..79 of the N = 80 accumulator [O | L]");
constexpr int MAX_ITEMS_PER_CTA = 2048;   // redo bitmap: one bit per item of this CTA
constexpr int ATT_SMEM = 2 * Q_BYTES + 2 * KV_STAGES * KV_BYTES + ONES_BYTES + 512 + MAX_ITEMS_PER_CTA / 8 + 1024;
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

__device__ __forceinline__ float max32u(const uint32_t (&x)[32]) {   // full tile: no column masking
    float a = __uint_as_float(x[0]), b = __uint_as_float(x[1]), c = __uint_as_float(x[2]), d = __uint_as_float(x[3]);
#pragma unroll
    for (int i = 4; i < 32; i += 4) {
        a = fmaxf(a, __uint_as_float(x[i])); b = fmaxf(b, __uint_as_float(x[i + 1]));
        c = fmaxf(c, __uint_as_float(x[i + 2])); d = fmaxf(d, __uint_as_float(x[i + 3]));
    }
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

// PP: pairs of every 8 whose exponentials run on the FMA pipe.  (Giving the two co-resident CTAs of an SM different mixes -
// one MUFU-heavy, one FMA-heavy - was measured: 0.332 - 0.348 ms against 0.324 ms for the uniform 2 of 8.)
template <int PP>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, int N, int D, int heads, int total) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // [2][Q_BYTES]
    uint8_t* sK = smem + 2 * Q_BYTES;
    uint8_t* sV = sK + KV_STAGES * KV_BYTES;
    uint8_t* sOnes = sV + KV_STAGES * KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
    uint64_t* q_full = bars;                          // [2]
    uint64_t* q_empty = bars + 2;                     // [2]
    uint64_t* kv_full = bars + 4;                     // [KV_STAGES]: K_g and V_g of a stage land on ONE barrier
    uint64_t* kv_empty = kv_full + KV_STAGES;         // [KV_STAGES]
    uint64_t* s_full = kv_empty + KV_STAGES;          // [2]
    uint64_t* p_full = s_full + 2;                    // [2], 128 arrivals
    uint64_t* done = p_full + 2;                      // last P V / L of an item retired
    uint64_t* main_done = done + 1;                   // softmax -> TMA / MMA warps: the redo bitmap is final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(main_done + 1);
    uint32_t* redo = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [MAX_ITEMS_PER_CTA / 32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = (N + BKV - 1) / BKV;
    const int QT = (N + BQ - 1) / BQ;
    const int G = gridDim.x;
    const int n_mine = (total - static_cast<int>(blockIdx.x) + G - 1) / G;   // items blockIdx.x, blockIdx.x + G, ...

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp < 4) {   // all-ones operand of the row-sum MMA (swizzle-invariant), visible to the async proxy
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        if (threadIdx.x < MAX_ITEMS_PER_CTA / 32) redo[threadIdx.x] = 0u;
        ptx::fence_proxy_async_smem();
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&q_full[i], 1);
                ptx::mbar_init(&q_empty[i], 1);
                ptx::mbar_init(&s_full[i], 1);
                ptx::mbar_init(&p_full[i], 128);
            }
            for (int i = 0; i < KV_STAGES; ++i) {
                ptx::mbar_init(&kv_full[i], 1);
                ptx::mbar_init(&kv_empty[i], 1);
            }
            ptx::mbar_init(done, 1);
            ptx::mbar_init(main_done, 1);
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

    // Every role walks the same sequence of STREAMS: the main stream (all n_mine items back to back), then one
    // single-item stream per redo pass.  g = running key-tile counter, it = running item counter over all streams:
    // they index the ring stages and give every barrier its phase parity.
    const auto redo_bit = [&](int i) { return (reinterpret_cast<volatile uint32_t*>(redo)[i >> 5] >> (i & 31)) & 1u; };

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            int g = 0, it = 0;
            auto stream_item = [&](int w) {
                const int qt = w % QT, bh = w / QT, h = bh % heads, b = bh / heads;
                const int qb = it & 1;
                ptx::mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
                ptx::tma_load_3d(sQ + qb * Q_BYTES, &tmQ, &q_full[qb], h * HD, qt * BQ, b);
                for (int j = 0; j < T; ++j, ++g) {
                    const int s = g % KV_STAGES;
                    ptx::mbar_wait(&kv_empty[s], ((g / KV_STAGES) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * KV_BYTES);
                    ptx::tma_load_3d(sK + s * KV_BYTES, &tmK, &kv_full[s], h * HD, j * BKV, b);
                    ptx::tma_load_3d(sV + s * KV_BYTES, &tmV, &kv_full[s], h * HD, j * BKV, b);
                }
                ++it;
            };
            for (int i = 0; i < n_mine; ++i) stream_item(blockIdx.x + i * G);
            ptx::mbar_wait(main_done, 0);
            for (int i = 0; i < n_mine; ++i)
                if (redo_bit(i)) {
                    stream_item(blockIdx.x + i * G);   // max-only pass
                    stream_item(blockIdx.x + i * G);   // exact pass
                }
        }
    } else if (warp == 5) {
        // -------------------------------------------------------------------- MMA issuer
        // This warp's instruction stream is on the critical path of every key tile (ncu, round 2: it waited for P only
        // 22 % of the time), so it is kept short: O and the row sums L come from ONE N = 80 MMA per 16 keys (B = [V | 1]:
        // the second 64-column block of the MN-major operand is the all-ones tile, reached through the descriptor's
        // leading-dimension offset), every barrier of a tile is waited for up front, and one elected block issues
        // P V_g, Q K_{g+2}^T and their commits.
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
        constexpr uint32_t idesc_pvl = ptx::make_idesc_bf16_bmn(BQ, HD + 16);
        constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
        const uint32_t v_a0 = (ptx::smem_u32(sV) & 0x3FFFF) >> 4;
        const uint32_t ones_a = (ptx::smem_u32(sOnes) & 0x3FFFF) >> 4;
        uint32_t g0 = 0, it = 0;
        auto qk4 = [&](uint32_t g, uint32_t q_lo) {  // (elected thread) S[g & 1] = Q K_g^T; the caller waited for kv_full
            const uint32_t k_lo = k_lo0 + (g % KV_STAGES) * (KV_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
                ptx::umma_bf16(tmem + S_COL + (g & 1) * BKV, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                               ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
            ptx::umma_commit(&s_full[g & 1]);
        };
        auto pvl4 = [&](uint32_t g, bool first) {   // (elected thread) [O | L] (+)= P_g [V_g | 1]; frees the K / V stage
            const uint32_t tP = tmem + S_COL + (g & 1) * BKV;
            const uint32_t v_a = v_a0 + (g % KV_STAGES) * (KV_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) {          // 16 keys = 16 rows of 128 B per k-step
                const uint32_t a = v_a + k * (16 * 128 >> 4);
                ptx::umma_bf16_ts(tmem + O_COL, tP + k * 8, ptx::make_desc(a | ((ones_a - a) << 16), kDescHiMn), idesc_pvl,
                                  (first && k == 0) ? 0u : 1u);
            }
            ptx::umma_commit(&kv_empty[g % KV_STAGES]);
        };
        auto mma_item = [&]() {
            const uint32_t qb = it & 1;
            const uint32_t q_lo = q_lo0 + qb * (Q_BYTES >> 4);
            const uint32_t Tu = static_cast<uint32_t>(T);
            ptx::mbar_wait(&q_full[qb], (it >> 1) & 1);
            // the first two S tiles of an item are issued while the softmax warps may still be storing the previous
            // item's output: S[g & 1] only has to be past P V of tile g - 2 (in-order pipe); O / L are not touched here
            ptx::mbar_wait(&kv_full[g0 % KV_STAGES], (g0 / KV_STAGES) & 1);
            if (Tu > 1) ptx::mbar_wait(&kv_full[(g0 + 1) % KV_STAGES], ((g0 + 1) / KV_STAGES) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                qk4(g0, q_lo);
                if (Tu > 1) qk4(g0 + 1, q_lo);
            }
            __syncwarp();
            uint32_t g = g0;
            for (const uint32_t gs = g0 + Tu - 2; static_cast<int32_t>(gs - g) > 0; ++g) {   // tiles with a Q K^T two ahead
                ptx::mbar_wait(&p_full[g & 1], (g >> 1) & 1);       // P_g written in place of S[g & 1]; for the first
                                                                    // tile also: the previous item's O / L have been read out
                ptx::mbar_wait(&kv_full[(g + 2) % KV_STAGES], ((g + 2) / KV_STAGES) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    pvl4(g, g == g0);
                    qk4(g + 2, q_lo);                               // executes after P V_g (in-order pipe): S[g & 1] is free
                }
                __syncwarp();
            }
            for (; g != g0 + Tu; ++g) {                             // the last two tiles (one if T == 1)
                ptx::mbar_wait(&p_full[g & 1], (g >> 1) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    pvl4(g, g == g0);
                    if (g + 1 == g0 + Tu) {
                        ptx::umma_commit(done);
                        ptx::umma_commit(&q_empty[qb]);              // every Q K^T of this item has retired
                    }
                }
                __syncwarp();
            }
            g0 += Tu;
            ++it;
        };
        for (int i = 0; i < n_mine; ++i) mma_item();
        pdl_launch_dependents();
        ptx::mbar_wait(main_done, 0);
        for (int i = 0; i < n_mine; ++i)
            if (redo_bit(i)) {
                mma_item();
                mma_item();
            }
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-3)
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tO = tmem + lane_base + O_COL, tL = tmem + lane_base + L_COL;
        const uint64_t sc2 = ptx::pack2(LOG2E, LOG2E);
        uint32_t v_lo[32], v_hi[32], pk[16];
        int g0 = 0, it = 0;
        float rmax = -INFINITY;   // MAXPASS result, consumed by the EXACT pass that follows it

        // mode and the first / last position of a tile are compile-time: the steady-state tiles (neither first nor last) of
        // the fast path carry no mode test, no column mask and no bounds arithmetic
        auto softmax_item = [&](auto ppc, auto modec, int w, int idx) {
            constexpr int PPX = decltype(ppc)::value;
            constexpr int mode = decltype(modec)::value;
            const int qt = w % QT, bh = w / QT, h = bh % heads, b = bh / heads;
            uint64_t nref2 = 0;
            if (mode == MODE_EXACT) {
                const float m = rmax * LOG2E;
                nref2 = ptx::pack2(-m, -m);
            }
            if (mode == MODE_MAXPASS) rmax = -INFINITY;
            ptx::mbar_wait(&s_full[g0 & 1], (g0 >> 1) & 1);
            ptx::tc_fence_after();
            ptx::tmem_ld_32x32(tS + (g0 & 1) * BKV, v_lo);
            auto tile = [&](auto firstc, auto lastc, int j) {
                constexpr bool FIRST = decltype(firstc)::value, LAST = decltype(lastc)::value;
                const int g = g0 + j;
                const int buf = g & 1;
                const int nvalid = (FIRST || LAST) ? min(BKV, N - j * BKV) : BKV;
                ptx::tmem_ld_wait();                                   // first half of tile j is in registers
                ptx::tmem_ld_32x32(tS + buf * BKV + 32, v_hi);          // second half: in flight during the first exps
                uint32_t next_ready = 1u;
                if (mode == MODE_MAXPASS) {
                    ptx::tmem_ld_wait();
                    rmax = fmaxf(rmax, fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid)));
                } else {
                    if (FIRST && mode == MODE_FAST) {                   // the first tile defines the reference
                        ptx::tmem_ld_wait();
                        const float m = (nvalid == BKV ? fmaxf(max32u(v_lo), max32u(v_hi))
                                                       : fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid))) * LOG2E;
                        nref2 = ptx::pack2(-m, -m);
                    }
                    exp32<PPX>(v_lo, sc2, nref2, pk);
                    if ((FIRST || LAST) && nvalid < BKV) mask16(pk, 0, nvalid);
                    ptx::tmem_st_32x16(tS + buf * BKV, pk);             // P columns [0,16) <- keys [0,32) (S lo is in registers)
                    ptx::tmem_ld_wait();                                // second half arrived
                    // S_{j+1} follows P V_{j-1} in the tensor pipe and lands about now: probe its barrier (non-blocking);
                    // the probe's latency hides behind the second half of the exponentials
                    if (!LAST) next_ready = ptx::mbar_test_wait(&s_full[buf ^ 1], ((g + 1) >> 1) & 1);
                    if (!(FIRST || LAST) || nvalid > 32) {
                        exp32<PPX>(v_hi, sc2, nref2, pk);
                        if ((FIRST || LAST) && nvalid < BKV) mask16(pk, 32, nvalid);
                    } else {   // the last key tile ends inside its first half (N = 1370: 26 keys): no exponentials, P = 0
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
                    ptx::tmem_st_32x16(tS + buf * BKV + 16, pk);
                }
                if (!LAST) {                                           // request the next tile's first half before draining
                    if (mode == MODE_MAXPASS || !next_ready) ptx::mbar_wait(&s_full[buf ^ 1], ((g + 1) >> 1) & 1);
                    ptx::tc_fence_after();
                    ptx::tmem_ld_32x32(tS + (buf ^ 1) * BKV, v_lo);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_full[buf]);
            };
            using Yes = std::true_type;
            using No = std::false_type;
            if (T == 1) {
                tile(Yes{}, Yes{}, 0);
            } else {
                tile(Yes{}, No{}, 0);
                for (int j = 1; j < T - 1; ++j) tile(No{}, No{}, j);
                tile(No{}, Yes{}, T - 1);
            }
            g0 += T;
            ptx::mbar_wait(done, it & 1);
            ptx::tc_fence_after();
            ++it;
            if (mode == MODE_MAXPASS) return;
            uint32_t lv[8];
            ptx::tmem_ld_32x8(tL, lv);
            ptx::tmem_ld_wait();
            const float l = __uint_as_float(lv[0]);
            if (mode == MODE_FAST) {   // a non-finite row sum: the whole item goes on the redo list (rare)
                const bool bad = !(l < 3.0e38f);
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&redo[idx >> 5], 1u << (idx & 31));
            }
            // O / l -> bf16 -> global (each thread owns one 128-byte row segment); a redone item overwrites this later
            const int row = qt * BQ + warp * 32 + lane;
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
                        uint4 wv;
                        wv.x = ptx::cvt_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                        wv.y = ptx::cvt_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                        wv.z = ptx::cvt_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                        wv.w = ptx::cvt_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + c * 32 + i) = wv;
                    }
                }
            }
            ptx::tc_fence_before();   // O / L reads are complete before this thread's next p_full arrive releases P V
        };

        using Fast = std::integral_constant<int, MODE_FAST>;
        for (int i = 0; i < n_mine; ++i) softmax_item(std::integral_constant<int, PP>{}, Fast{}, blockIdx.x + i * G, i);
        ptx::named_bar_sync(1, 128);          // every softmax thread has published its redo bits
        if (threadIdx.x == 0) ptx::mbar_arrive(main_done);
        for (int i = 0; i < n_mine; ++i)
            if (redo_bit(i)) {
                softmax_item(std::integral_constant<int, PP>{}, std::integral_constant<int, MODE_MAXPASS>{}, blockIdx.x + i * G, i);
                softmax_item(std::integral_constant<int, PP>{}, std::integral_constant<int, MODE_EXACT>{}, blockIdx.x + i * G, i);
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
    static int sms = 0;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc5_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        int dev = 0;
        DAD_CHECK_CUDA(cudaGetDevice(&dev));
        DAD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const long long total = static_cast<long long>(B) * heads * cdiv(N, BQ);
    const int grid = static_cast<int>(total < 2LL * sms ? total : 2LL * sms);
    DAD_REQUIRE(cdiv(total, grid) <= MAX_ITEMS_PER_CTA, "attention: %lld work items exceed the per-CTA redo bitmap", total);
    DAD_CHECK_CUDA(launch_pdl(attention_tc5_kernel<PP>, dim3(grid), dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N,
                              heads * HD, heads, static_cast<int>(total)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace

// qkv [B*N, 3*D] bf16 (q pre-scaled) -> out [B*N, D] bf16.  poly_pairs = pairs of every 8 on the FMA pipe (0, 2..5).
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
        case 3: return launch5<3>(tm, out, B, N, heads, st);
        case 4: return launch5<4>(tm, out, B, N, heads, st);
        case 5: return launch5<5>(tm, out, B, N, heads, st);
        default: return launch5<2>(tm, out, B, N, heads, st);
    }
}

}  // namespace dad
