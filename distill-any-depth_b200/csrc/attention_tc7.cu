// Fused multi-head attention (sm_100a), head_dim 64: the two-warpgroup variant of attention_tc5.cu (DAD_ATT_VARIANT=7).
// Reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights.
//
// ONE persistent CTA per SM, work item = (image, head, 128 queries).  TWO softmax warpgroups (warps 0-3 and 4-7; warp w and
// w + 4 own the same 32 TMEM lanes / query rows and share a scheduler) alternate the 64-key tiles of the SAME item: group
// (g & 1) exponentiates tile g.  This is possible because attention_tc5's softmax has no running state: every P tile is
// taken against one fixed reference (the first key tile's row maximum, published through shared memory by the group that
// owns tile 0), the row sums are a tensor-core MMA (P . ones) and O accumulates in TMEM, so tiles are independent.  Four
// in-place score / probability buffers (two per group) let the tensor pipe run four tiles ahead.
//   TMEM (512 columns): S/P buffers [0,256) | O [256,320) | L [320,336)
// Compared with two independent CTAs per SM (attention_tc5) the two softmax warps of a scheduler work on consecutive tiles
// of one stream instead of running two identical streams in lock-step, each item finishes in half the time, and the item
// epilogue is split between the groups (32 output columns each).
// Roles (320 threads): warps 0-7 softmax, warp 8 TMA producer, warp 9 tcgen05.mma issuer.  Overflow handling: as
// attention_tc5 (redo list, max-only pass + exact pass at the end of the launch).
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
constexpr int ONES_BYTES = 16 * 128;
constexpr int KV_STAGES = 8;              // Q K^T runs four tiles ahead of P V
constexpr int NSB = 4;                    // score buffers
constexpr int ATT_THREADS = 352;
constexpr int TMEM_COLS = 512;
constexpr int S_COL = 0, O_COL = 256, L_COL = 320;
constexpr int MAX_ITEMS_PER_CTA = 4096;
constexpr int ATT_SMEM = 2 * Q_BYTES + 2 * KV_STAGES * KV_BYTES + ONES_BYTES + 512 + MAX_ITEMS_PER_CTA / 8 + 2 * 128 * 4 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

enum : int { MODE_FAST = 0, MODE_MAXPASS = 1, MODE_EXACT = 2 };

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

template <int PP, bool SPLIT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tc7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
    uint64_t* k_full = bars + 4;                      // [KV_STAGES]
    uint64_t* v_full = k_full + KV_STAGES;            // [KV_STAGES]
    uint64_t* kv_empty = v_full + KV_STAGES;          // [KV_STAGES]
    uint64_t* s_full = kv_empty + KV_STAGES;          // [NSB]
    uint64_t* p_full = s_full + NSB;                  // [NSB], 128 arrivals (the owning group)
    uint64_t* done = p_full + NSB;                    // last P V / L of an item retired
    uint64_t* main_done = done + 1;
    uint64_t* o_free = main_done + 1;                 // 256 arrivals: both groups have read O / L of the item
    uint64_t* s_free = o_free + 1;                    // [NSB] (SPLIT): P V of the tile that used the buffer has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + NSB);
    uint32_t* redo = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [MAX_ITEMS_PER_CTA / 32]
    float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(redo) + MAX_ITEMS_PER_CTA / 8);   // [2][128] row exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = (N + BKV - 1) / BKV;
    const int QT = (N + BQ - 1) / BQ;
    const int G = gridDim.x;
    const int n_mine = (total - static_cast<int>(blockIdx.x) + G - 1) / G;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp < 4) {
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        redo[threadIdx.x] = 0u;   // 128 words = MAX_ITEMS_PER_CTA bits
        ptx::fence_proxy_async_smem();
    }
    if (warp == 9) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&q_full[i], 1);
                ptx::mbar_init(&q_empty[i], 1);
            }
            for (int i = 0; i < NSB; ++i) {
                ptx::mbar_init(&s_full[i], 1);
                ptx::mbar_init(&p_full[i], 128);
                ptx::mbar_init(&s_free[i], 1);
            }
            for (int i = 0; i < KV_STAGES; ++i) {
                ptx::mbar_init(&k_full[i], 1);
                ptx::mbar_init(&v_full[i], 1);
                ptx::mbar_init(&kv_empty[i], 1);
            }
            ptx::mbar_init(done, 1);
            ptx::mbar_init(main_done, 1);
            ptx::mbar_init(o_free, 256);
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
    pdl_wait();

    const auto redo_bit = [&](int i) { return (reinterpret_cast<volatile uint32_t*>(redo)[i >> 5] >> (i & 31)) & 1u; };

    if (warp == 8) {
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
                    ptx::mbar_arrive_expect_tx(&k_full[s], KV_BYTES);
                    ptx::tma_load_3d(sK + s * KV_BYTES, &tmK, &k_full[s], h * HD, j * BKV, b);
                    ptx::mbar_arrive_expect_tx(&v_full[s], KV_BYTES);
                    ptx::tma_load_3d(sV + s * KV_BYTES, &tmV, &v_full[s], h * HD, j * BKV, b);
                }
                ++it;
            };
            for (int i = 0; i < n_mine; ++i) stream_item(blockIdx.x + i * G);
            ptx::mbar_wait(main_done, 0);
            for (int i = 0; i < n_mine; ++i)
                if (redo_bit(i)) {
                    stream_item(blockIdx.x + i * G);
                    stream_item(blockIdx.x + i * G);
                }
        }
    } else if (SPLIT && (warp == 9 || warp == 10)) {
        // -------------------------------------------------------------------- two single-thread MMA issuers
        // warp 9: S_g = Q K_g^T, up to NSB tiles ahead (also across items); warp 10: O += P_g V_g, L += P_g 1.
        // tcgen05.mma from two threads is unordered, so the reuse of a score buffer goes through s_free (a commit).
        if (lane == 0) {
            constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
            constexpr uint32_t idesc_l = ptx::make_idesc_bf16(BQ, 16);
            constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t q_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
            const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
            const uint32_t v_lo0 = ptx::desc_lo_mn_sw128(ptx::smem_u32(sV));
            const uint32_t one_lo = ptx::desc_lo_sw128(ptx::smem_u32(sOnes));
            int g0 = 0, it = 0;
            auto qk_item = [&]() {
                const int qb = it & 1;
                const uint32_t q_lo = q_lo0 + qb * (Q_BYTES >> 4);
                ptx::mbar_wait(&q_full[qb], (it >> 1) & 1);
                for (int j = 0; j < T; ++j) {
                    const int g = g0 + j, s = g % KV_STAGES, buf = g % NSB;
                    if (g >= NSB) ptx::mbar_wait(&s_free[buf], ((g / NSB) & 1) ^ 1);
                    ptx::mbar_wait(&k_full[s], (g / KV_STAGES) & 1);
                    ptx::tc_fence_after();
                    const uint32_t k_lo = k_lo0 + s * (KV_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        ptx::umma_bf16(tmem + S_COL + buf * BKV, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                                       ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
                    ptx::umma_commit(&s_full[buf]);
                    if (j == T - 1) ptx::umma_commit(&q_empty[qb]);
                }
                g0 += T;
                ++it;
            };
            auto pv_item = [&]() {
                for (int j = 0; j < T; ++j) {
                    const int g = g0 + j, s = g % KV_STAGES, buf = g % NSB;
                    if (j == 0) ptx::mbar_wait(o_free, (it & 1) ^ 1);
                    ptx::mbar_wait(&p_full[buf], (g / NSB) & 1);
                    ptx::mbar_wait(&v_full[s], (g / KV_STAGES) & 1);
                    ptx::tc_fence_after();
                    const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
                    const uint32_t tP = tmem + S_COL + buf * BKV;
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        ptx::umma_bf16_ts(tmem + O_COL, tP + k * 8, ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn),
                                          idesc_pv, (j | k) != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        ptx::umma_bf16_ts(tmem + L_COL, tP + k * 8, ptx::make_desc(one_lo + 2 * k, ptx::kDescHiSw128),
                                          idesc_l, (j | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&s_free[buf]);
                    ptx::umma_commit(&kv_empty[s]);
                    if (j == T - 1) ptx::umma_commit(done);
                }
                g0 += T;
                ++it;
            };
            if (warp == 9) {
                for (int i = 0; i < n_mine; ++i) qk_item();
                ptx::mbar_wait(main_done, 0);
                for (int i = 0; i < n_mine; ++i)
                    if (redo_bit(i)) {
                        qk_item();
                        qk_item();
                    }
            } else {
                for (int i = 0; i < n_mine; ++i) pv_item();
                pdl_launch_dependents();
                ptx::mbar_wait(main_done, 0);
                for (int i = 0; i < n_mine; ++i)
                    if (redo_bit(i)) {
                        pv_item();
                        pv_item();
                    }
            }
        }
    } else if (warp == 9) {
        // -------------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
        constexpr uint32_t idesc_l = ptx::make_idesc_bf16(BQ, 16);
        constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
        const uint32_t v_lo0 = ptx::desc_lo_mn_sw128(ptx::smem_u32(sV));
        const uint32_t one_lo = ptx::desc_lo_sw128(ptx::smem_u32(sOnes));
        int g0 = 0, it = 0;
        auto issue_qk = [&](int g, uint32_t q_lo) {  // S[g % NSB] = Q K_g^T
            const int s = g % KV_STAGES;
            ptx::mbar_wait(&k_full[s], (g / KV_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t k_lo = k_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_bf16(tmem + S_COL + (g % NSB) * BKV, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                                   ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(&s_full[g % NSB]);
            }
            __syncwarp();
        };
        auto mma_item = [&]() {
            const int qb = it & 1;
            const uint32_t q_lo = q_lo0 + qb * (Q_BYTES >> 4);
            ptx::mbar_wait(&q_full[qb], (it >> 1) & 1);
            for (int j = 0; j < NSB && j < T; ++j) issue_qk(g0 + j, q_lo);   // S[g % NSB] is past P V of tile g - NSB (in-order pipe)
            for (int j = 0; j < T; ++j) {
                const int g = g0 + j;
                const int s = g % KV_STAGES;
                if (j == 0) ptx::mbar_wait(o_free, (it & 1) ^ 1);   // both groups have read the previous item's O / L
                ptx::mbar_wait(&p_full[g % NSB], (g / NSB) & 1);    // P_g written in place
                ptx::mbar_wait(&v_full[s], (g / KV_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
                const uint32_t tP = tmem + S_COL + (g % NSB) * BKV;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        ptx::umma_bf16_ts(tmem + O_COL, tP + k * 8, ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn),
                                          idesc_pv, (j | k) != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        ptx::umma_bf16_ts(tmem + L_COL, tP + k * 8, ptx::make_desc(one_lo + 2 * k, ptx::kDescHiSw128),
                                          idesc_l, (j | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&kv_empty[s]);
                    if (j == T - 1) {
                        ptx::umma_commit(done);
                        ptx::umma_commit(&q_empty[qb]);
                    }
                }
                __syncwarp();
                if (j + NSB < T) issue_qk(g + NSB, q_lo);
            }
            g0 += T;
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
    } else if (warp < 8) {
        // -------------------------------------------------------------------- softmax (warps 0-7: two groups)
        const int grp = warp >> 2, quarter = warp & 3;
        const int rowl = quarter * 32 + lane;                       // query row inside the tile (same for w and w + 4)
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tO = tmem + lane_base + O_COL, tL = tmem + lane_base + L_COL;
        const uint64_t sc2 = ptx::pack2(LOG2E, LOG2E);
        uint32_t v_lo[32], v_hi[32], pk[16];
        int g0 = 0, it = 0;
        float rmax = -INFINITY;

        auto softmax_item = [&](auto modec, int w, int idx) {
            constexpr int mode = decltype(modec)::value;
            const int qt = w % QT, bh = w / QT, h = bh % heads, b = bh / heads;
            const int first = g0 + (((g0 & 1) == grp) ? 0 : 1);          // my first tile of this item (global index)
            const bool owner = (g0 & 1) == grp;                          // my group exponentiates tile 0
            uint64_t nref2 = 0;
            if (mode == MODE_EXACT) {
                const float m = rmax * LOG2E;
                nref2 = ptx::pack2(-m, -m);
            }
            if (mode == MODE_MAXPASS) rmax = -INFINITY;
            if (mode == MODE_FAST && !owner) {   // reference of this row: published by the group that owns tile 0
                ptx::named_bar_sync(2 + (it & 1), 256);
                const float m = xch[rowl];
                nref2 = ptx::pack2(-m, -m);
            }
            const int gend = g0 + T;
            if (first < gend) {
                ptx::mbar_wait(&s_full[first % NSB], (first / NSB) & 1);
                ptx::tc_fence_after();
                ptx::tmem_ld_32x32(tS + (first % NSB) * BKV, v_lo);
            }
            for (int g = first; g < gend; g += 2) {
                const int j = g - g0;
                const int buf = g % NSB;
                const int nvalid = min(BKV, N - j * BKV);
                const bool more = g + 2 < gend;
                ptx::tmem_ld_wait();
                ptx::tmem_ld_32x32(tS + buf * BKV + 32, v_hi);
                uint32_t next_ready = 1u;
                if (mode == MODE_MAXPASS) {
                    ptx::tmem_ld_wait();
                    rmax = fmaxf(rmax, fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid)));
                } else {
                    if (mode == MODE_FAST && j == 0) {                   // owner: the first tile defines the reference
                        ptx::tmem_ld_wait();
                        const float m = fmaxf(max32(v_lo, 0, nvalid), max32(v_hi, 32, nvalid)) * LOG2E;
                        nref2 = ptx::pack2(-m, -m);
                        xch[rowl] = m;
                        __threadfence_block();
                        asm volatile("bar.arrive %0, 256;" ::"r"(2 + (it & 1)) : "memory");
                    }
                    exp32<PP>(v_lo, sc2, nref2, pk);
                    if (nvalid < BKV) mask16(pk, 0, nvalid);
                    ptx::tmem_st_32x16(tS + buf * BKV, pk);
                    ptx::tmem_ld_wait();
                    if (more) next_ready = ptx::mbar_test_wait(&s_full[(g + 2) % NSB], ((g + 2) / NSB) & 1);
                    if (nvalid > 32) {
                        exp32<PP>(v_hi, sc2, nref2, pk);
                        if (nvalid < BKV) mask16(pk, 32, nvalid);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
                    ptx::tmem_st_32x16(tS + buf * BKV + 16, pk);
                }
                if (more) {
                    if (mode == MODE_MAXPASS || !next_ready) ptx::mbar_wait(&s_full[(g + 2) % NSB], ((g + 2) / NSB) & 1);
                    ptx::tc_fence_after();
                    ptx::tmem_ld_32x32(tS + ((g + 2) % NSB) * BKV, v_lo);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_full[buf]);
            }
            g0 += T;
            ptx::mbar_wait(done, it & 1);
            ptx::tc_fence_after();
            ++it;
            if (mode == MODE_MAXPASS) {   // the exact row maximum over BOTH groups' tiles
                xch[grp * 128 + rowl] = rmax;
                __threadfence_block();
                ptx::named_bar_sync(4, 256);
                rmax = fmaxf(xch[rowl], xch[128 + rowl]);
                ptx::named_bar_sync(4, 256);      // both groups have read before xch is reused
                ptx::mbar_arrive(o_free);
                return;
            }
            uint32_t lv[8];
            ptx::tmem_ld_32x8(tL, lv);
            ptx::tmem_ld_wait();
            const float l = __uint_as_float(lv[0]);
            if (mode == MODE_FAST) {
                const bool bad = !(l < 3.0e38f);
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&redo[idx >> 5], 1u << (idx & 31));
            }
            // group g writes output columns [32 g, 32 g + 32) of its rows
            const int row = qt * BQ + rowl;
            const float inv = 1.0f / l;
            bf16* dst = out + (static_cast<long long>(b) * N + row) * D + h * HD + grp * 32;
            uint32_t o[32];
            ptx::tmem_ld_32x32(tO + grp * 32, o);
            ptx::tmem_ld_wait();
            if (row < N) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 wv;
                    wv.x = ptx::cvt_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                    wv.y = ptx::cvt_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    wv.z = ptx::cvt_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                    wv.w = ptx::cvt_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + i) = wv;
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(o_free);
        };

        for (int i = 0; i < n_mine; ++i) softmax_item(std::integral_constant<int, MODE_FAST>{}, blockIdx.x + i * G, i);
        ptx::named_bar_sync(1, 256);          // every softmax thread has published its redo bits
        if (threadIdx.x == 0) ptx::mbar_arrive(main_done);
        for (int i = 0; i < n_mine; ++i)
            if (redo_bit(i)) {
                softmax_item(std::integral_constant<int, MODE_MAXPASS>{}, blockIdx.x + i * G, i);
                softmax_item(std::integral_constant<int, MODE_EXACT>{}, blockIdx.x + i * G, i);
            }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, TMEM_COLS);
    }
}

template <int PP, bool SPLIT>
int launch7(const CUtensorMap* tm, bf16* out, int B, int N, int heads, cudaStream_t st) {
    static bool configured = false;
    static int sms = 0;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc7_kernel<PP, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        int dev = 0;
        DAD_CHECK_CUDA(cudaGetDevice(&dev));
        DAD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const long long total = static_cast<long long>(B) * heads * cdiv(N, BQ);
    const int grid = static_cast<int>(total < sms ? total : sms);
    DAD_REQUIRE(cdiv(total, grid) <= MAX_ITEMS_PER_CTA, "attention: %lld work items exceed the per-CTA redo bitmap", total);
    DAD_CHECK_CUDA(launch_pdl(attention_tc7_kernel<PP, SPLIT>, dim3(grid), dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N,
                              heads * HD, heads, static_cast<int>(total)));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace

int attention_tc7(const bf16* qkv, bf16* out, int B, int N, int heads, int poly_pairs, int split_issuers, cudaStream_t st) {
    const int D = heads * HD;
    CUtensorMap tm[3];
    for (int i = 0; i < 3; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)3 * D * 2 * N};
        const cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)(i == 0 ? BQ : BKV), 1};
        DAD_TRY(make_tmap_bf16(&tm[i], qkv + static_cast<long long>(i) * D, 3, dims, strides, box));
    }
    if (split_issuers) {
        switch (poly_pairs) {
            case 0: return launch7<0, true>(tm, out, B, N, heads, st);
            case 3: return launch7<3, true>(tm, out, B, N, heads, st);
            case 4: return launch7<4, true>(tm, out, B, N, heads, st);
            default: return launch7<2, true>(tm, out, B, N, heads, st);
        }
    }
    switch (poly_pairs) {
        case 0: return launch7<0, false>(tm, out, B, N, heads, st);
        case 3: return launch7<3, false>(tm, out, B, N, heads, st);
        case 4: return launch7<4, false>(tm, out, B, N, heads, st);
        default: return launch7<2, false>(tm, out, B, N, heads, st);
    }
}

}  // namespace dad
