// Fused multi-head attention on the 5th-gen tensor cores (sm_100a), head_dim 64, non-causal, no mask
// (reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights).
//
// One CTA = one (image, head, 128-query tile); two CTAs are co-resident per SM.  Roles (192 threads):
//   warps 0-3  softmax: thread = query row; the 64 scores of a key tile are read from TMEM once, exponentiated
//              in registers, and P is written back to TMEM as packed bf16 (tcgen05.st) where the second MMA
//              consumes it directly (A operand from TMEM).
//   warp 4     TMA producer: Q tile once, then K / V tiles (64 keys x 64) through a 4-stage mbarrier ring,
//              3-D tensor maps over qkv [B, N, 3*D] so rows past N are zero-filled per image.
//   warp 5     tcgen05.mma issuer:  S[128x64] = Q K^T  (A, B from smem, K-major)
//                                   O[128x64] += P V     (A = P from TMEM, B = V from smem, MN-major)
// S and P are double-buffered in TMEM (256 columns per CTA: S0 S1 | P0 P1 | O), so Q K^T of tile j+2 is issued
// as soon as the softmax has drained S of tile j, and the softmax of tile j+1 never waits for the tensor pipe:
// the kernel is bounded by the exp2 throughput of the MUFU (16 / clk / SM), not by MMA latency.
// Online softmax keeps a per-row reference maximum; O / l are rescaled (TMEM round trip) only when the running
// maximum exceeds the reference by more than 2^16 ("lazy rescale": exact, P <= 2^16 stays well inside bf16 /
// fp32 range), so in the common case O is never touched until the final normalisation.
#include <cstdlib>
#include <type_traits>

#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;      // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;    // 8 KB
constexpr int KV_STAGES = 4;
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, P_COL = 128, O_COL = 192;  // S0 [0,64) S1 [64,128) | P0 [128,160) P1 [160,192) | O [192,256)
constexpr int ATT_SMEM = Q_BYTES + 2 * KV_STAGES * KV_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 16.0f;  // log2 units: P <= 2^16 relative to the reference maximum

// POLY = how many of every 8 exponentials are evaluated by ptx::ex2_fma on the FMA pipe instead of the MUFU
// (the softmax is bounded by the 16 exp2 / clk / SM of the MUFU, while the FMA pipe has slack).
template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, int N, int D) {
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
    uint64_t* p_full = s_full + 2;                    // [2]
    uint64_t* o_full = s_full + 4;                    // one completion per P V
    uint64_t* done = s_full + 5;                      // last P V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int T = (N + BKV - 1) / BKV;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
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
            ptx::mbar_init(o_full, 1);
            ptx::mbar_init(done, 1);
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
    } else if (warp == 5) {
        // -------------------------------------------------------------------- MMA issuer
        // warp-uniform control flow, one elected lane issues; descriptors = constant high word + (address >> 4)
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
            ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1);       // P_j written, S[j & 1] drained
            ptx::mbar_wait(&v_full[s], (j / KV_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)   // 16 keys = 16 rows of 128 B per k-step
                    ptx::umma_bf16_ts(tmem + O_COL, tmem + P_COL + (j & 1) * (BKV / 2) + k * 8,
                                      ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn), idesc_pv, (j | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&kv_empty[s]);                  // K_j / V_j stage free once P V_j retires
                ptx::umma_commit(o_full);
                if (j == T - 1) ptx::umma_commit(done);
            }
            __syncwarp();
            if (j + 2 < T) issue_qk(j + 2);
        }
        pdl_launch_dependents();  // last MMA issued: the next kernel's prologue may overlap this CTA's drain
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-3)
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tP = tmem + lane_base + P_COL, tO = tmem + lane_base + O_COL;
        float m_ref = -INFINITY;  // reference maximum (log2 domain) the stored P / O / l are relative to
        float l = 0.f;
        uint32_t v[BKV];          // the 64 scores of this row in the current key tile
        uint32_t (&v_lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v_hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);

        // exp2 of one 32-column half against m_ref -> packed bf16 pairs; returns the partial row sum.
        // Full tiles (all but the last) take the unmasked instantiation: FFMA + MUFU + FADD per element.
        auto exp_half = [&](auto masked, const uint32_t (&x)[32], int col0, int nvalid, uint32_t (&pk)[16]) -> float {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                // elements 1, 5 (POLY >= 2), 3 (POLY >= 3), 7 (POLY >= 4) of every 8 go to the FMA pipe
                const int e1 = (i + 1) & 7;
                const bool poly1 = (POLY >= 2 && (e1 == 1 || e1 == 5)) || (POLY >= 3 && e1 == 3) || (POLY >= 4 && e1 == 7);
                const float a0 = fmaf(__uint_as_float(x[i]), LOG2E, -m_ref);
                const float a1 = fmaf(__uint_as_float(x[i + 1]), LOG2E, -m_ref);
                float p0 = ptx::ex2_approx(a0);
                float p1 = poly1 ? ptx::ex2_fma(a1) : ptx::ex2_approx(a1);
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
        auto half_max = [&](const uint32_t (&x)[32], int col0, int nvalid) -> float {  // first tile / rare path only
            float a = -INFINITY, b2 = -INFINITY, c2 = -INFINITY, d2 = -INFINITY;  // four independent chains
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float x0 = (col0 + i < nvalid) ? __uint_as_float(x[i]) : -INFINITY;
                const float x1 = (col0 + i + 1 < nvalid) ? __uint_as_float(x[i + 1]) : -INFINITY;
                const float x2 = (col0 + i + 2 < nvalid) ? __uint_as_float(x[i + 2]) : -INFINITY;
                const float x3 = (col0 + i + 3 < nvalid) ? __uint_as_float(x[i + 3]) : -INFINITY;
                a = fmaxf(a, x0); b2 = fmaxf(b2, x1); c2 = fmaxf(c2, x2); d2 = fmaxf(d2, x3);
            }
            return fmaxf(fmaxf(a, b2), fmaxf(c2, d2));
        };
        using TrueT = std::true_type;
        using FalseT = std::false_type;
        constexpr float P_LIMIT = 65536.0f;  // 2^RESCALE_THRESHOLD

        // software pipeline: the first half of tile j+1 is requested from TMEM while P_j's stores drain
        ptx::mbar_wait(&s_full[0], 0);
        ptx::tc_fence_after();
        ptx::tmem_ld_32x32(tS, v_lo);
        for (int j = 0; j < T; ++j) {
            const int buf = j & 1;
            const int nvalid = min(BKV, N - j * BKV);
            ptx::tmem_ld_wait();                                   // first half of tile j is in registers
            ptx::tmem_ld_32x32(tS + buf * BKV + 32, v_hi);          // second half: in flight during the first exps
            if (j == 0) {                                          // the very first tile defines the reference maximum
                ptx::tmem_ld_wait();
                m_ref = fmaxf(half_max(v_lo, 0, nvalid), half_max(v_hi, 32, nvalid)) * LOG2E;
            }
            uint32_t pk[16];
            float lt;
            if (nvalid == BKV) {
                lt = exp_half(FalseT{}, v_lo, 0, nvalid, pk);
                ptx::tmem_st_32x16(tP + buf * (BKV / 2), pk);
                ptx::tmem_ld_wait();                               // second half arrived
                lt += exp_half(FalseT{}, v_hi, 32, nvalid, pk);
            } else {
                lt = exp_half(TrueT{}, v_lo, 0, nvalid, pk);
                ptx::tmem_st_32x16(tP + buf * (BKV / 2), pk);
                ptx::tmem_ld_wait();
                lt += exp_half(TrueT{}, v_hi, 32, nvalid, pk);
            }
            ptx::tmem_st_32x16(tP + buf * (BKV / 2) + 16, pk);
            // some P above 2^16 <=> some score more than 16 (log2 units) above the reference: seen in the row sum,
            // so the common path carries no max tracking at all
            const bool need = !(lt <= P_LIMIT);
            if (__any_sync(0xffffffffu, need)) {
                // rare: rescale O (TMEM) and l to the true running maximum and redo this tile's P against it.
                // P V_{j-1} must have retired: o_full is in phase j-1 or j here (P V_{j-2} retired before s_full of
                // tile j, P V_j needs this thread's arrive below).
                if (j > 0) {
                    ptx::mbar_wait(o_full, (j - 1) & 1);
                    ptx::tc_fence_after();
                }
                const float tmax = fmaxf(half_max(v_lo, 0, nvalid), half_max(v_hi, 32, nvalid)) * LOG2E;
                const float alpha = need ? ptx::ex2_approx(m_ref - tmax) : 1.0f;
                if (need) m_ref = tmax;
                l *= alpha;
                if (j > 0) {
#pragma unroll 1
                    for (int c = 0; c < HD / 16; ++c) {
                        uint32_t o[16];
                        ptx::tmem_ld_32x16(tO + c * 16, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ptx::tmem_st_32x16(tO + c * 16, o);
                    }
                }
                lt = exp_half(TrueT{}, v_lo, 0, nvalid, pk);
                ptx::tmem_st_32x16(tP + buf * (BKV / 2), pk);
                lt += exp_half(TrueT{}, v_hi, 32, nvalid, pk);
                ptx::tmem_st_32x16(tP + buf * (BKV / 2) + 16, pk);
            }
            l += lt;
            if (j + 1 < T) {                                       // request the next tile's first half before draining
                ptx::mbar_wait(&s_full[buf ^ 1], ((j + 1) >> 1) & 1);
                ptx::tc_fence_after();
                ptx::tmem_ld_32x32(tS + (buf ^ 1) * BKV, v_lo);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&p_full[buf]);
        }
        // final: O / l -> bf16 -> global (each thread owns one 128-byte row segment)
        ptx::mbar_wait(done, 0);
        ptx::tc_fence_after();
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
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                    w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
                    w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
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

}  // namespace

// qkv [B*N, 3*D] bf16 (q pre-scaled) -> out [B*N, D] bf16
int attention_tc(const bf16* qkv, bf16* out, int B, int N, int heads, cudaStream_t st) {
    const int D = heads * HD;
    static int poly = -1;
    if (poly < 0) {
        const char* e = getenv("DAD_ATT_POLY");  // A/B switch: exponentials per 8 moved from the MUFU to the FMA pipe
        // measured on B200 (ViT-L 518^2 B=32, no-rescale regime): POLY 0 / 2 / 3 = 0.358 / 0.371 / 0.392 ms per launch -
        // the softmax warps are issue-bound as much as MUFU-bound, so the extra FMA-pipe instructions do not pay: default 0
        poly = e ? atoi(e) : 0;
        if (poly != 0 && poly != 2 && poly != 3 && poly != 4) poly = 0;
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    }
    CUtensorMap tm[3];
    for (int i = 0; i < 3; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)3 * D * 2 * N};
        const cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)(i == 0 ? BQ : BKV), 1};
        DAD_TRY(make_tmap_bf16(&tm[i], qkv + static_cast<long long>(i) * D, 3, dims, strides, box));
    }
    const dim3 grid(cdiv(N, BQ), heads, B);
    switch (poly) {
        case 3: DAD_CHECK_CUDA(launch_pdl(attention_tc_kernel<3>, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N, D)); break;
        case 4: DAD_CHECK_CUDA(launch_pdl(attention_tc_kernel<4>, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N, D)); break;
        case 2: DAD_CHECK_CUDA(launch_pdl(attention_tc_kernel<2>, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N, D)); break;
        default: DAD_CHECK_CUDA(launch_pdl(attention_tc_kernel<0>, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N, D)); break;
    }
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
