// Fused multi-head attention on the 5th-gen tensor cores (sm_100a), head_dim 64, non-causal, no mask
// (reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights).
//
// One CTA = one (image, head, 128-query tile); two CTAs are co-resident per SM so one CTA's MMAs overlap the
// other's softmax.  Roles (192 threads):
//   warps 0-3  softmax: thread = query row.  S row is read from TMEM twice (max, then exp2), P is written
//              back to TMEM as packed bf16 (tcgen05.st) and consumed by the second MMA straight from TMEM.
//   warp 4     TMA producer: Q tile once, then K / V tiles (128 keys x 64) through a 2-stage mbarrier ring,
//              3-D tensor maps over qkv [B, N, 3*D] so rows past N are zero-filled per image.
//   warp 5     tcgen05.mma issuer:  S[128x128] = Q K^T  (A, B from smem, K-major)
//                                   O[128x64] += P V     (A = P from TMEM, B = V from smem, MN-major)
// TMEM (256 columns per CTA): S fp32 [0,128), P bf16x2 [128,192), O fp32 [192,256).
// Online softmax keeps a per-row reference maximum; O / l are rescaled (TMEM round trip) only when the running
// maximum exceeds the reference by more than 2^8 ("lazy rescale": exact, P <= 2^8 stays well inside bf16 /
// fp32 range), so in the common case O is never touched until the final normalisation.
#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = BKV * HD * 2;  // 16 KB
constexpr int KV_STAGES = 2;
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, P_COL = 128, O_COL = 192;
constexpr int ATT_SMEM = (1 + 2 * KV_STAGES) * TILE_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, int N, int D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;
    uint8_t* sV = smem + (1 + KV_STAGES) * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + 2 * KV_STAGES) * TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                  // [KV_STAGES]
    uint64_t* v_full = bars + 1 + KV_STAGES;      // [KV_STAGES]
    uint64_t* kv_empty = bars + 1 + 2 * KV_STAGES;  // [KV_STAGES]
    uint64_t* s_full = bars + 1 + 3 * KV_STAGES;
    uint64_t* p_full = s_full + 1;
    uint64_t* o_full = s_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

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
            ptx::mbar_init(s_full, 1);
            ptx::mbar_init(p_full, 128);
            ptx::mbar_init(o_full, 1);
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

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            ptx::mbar_arrive_expect_tx(q_full, TILE_BYTES);
            ptx::tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % KV_STAGES;
                const uint32_t ph = (j / KV_STAGES) & 1;
                ptx::mbar_wait(&kv_empty[s], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
                ptx::tma_load_3d(sK + s * TILE_BYTES, &tmK, &k_full[s], h * HD, j * BKV, b);
                ptx::mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                ptx::tma_load_3d(sV + s * TILE_BYTES, &tmV, &v_full[s], h * HD, j * BKV, b);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
            const uint32_t q_addr = ptx::smem_u32(sQ);
            auto issue_qk = [&](int j) {
                const int s = j % KV_STAGES;
                ptx::mbar_wait(&k_full[s], (j / KV_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t k_addr = ptx::smem_u32(sK + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_bf16(tmem + S_COL, ptx::make_smem_desc_sw128(q_addr + k * 32),
                                   ptx::make_smem_desc_sw128(k_addr + k * 32), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(s_full);
            };
            ptx::mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < T; ++j) {
                const int s = j % KV_STAGES;
                ptx::mbar_wait(p_full, j & 1);                       // P_j written, S_j fully read
                ptx::mbar_wait(&v_full[s], (j / KV_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t v_addr = ptx::smem_u32(sV + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    ptx::umma_bf16_ts(tmem + O_COL, tmem + P_COL + k * 8,
                                      ptx::make_smem_desc_mn_sw128(v_addr + k * 16 * 128), idesc_pv,
                                      (j | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&kv_empty[s]);                      // K_j / V_j stage free once PV_j retires
                ptx::umma_commit(o_full);
                if (j + 1 < T) issue_qk(j + 1);
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-3)
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tP = tmem + lane_base + P_COL, tO = tmem + lane_base + O_COL;
        float m_ref = -INFINITY;  // reference maximum (log2 domain) the stored P / O / l are relative to
        float l = 0.f;
        for (int j = 0; j < T; ++j) {
            ptx::mbar_wait(s_full, j & 1);
            ptx::tc_fence_after();
            const int nvalid = min(BKV, N - j * BKV);
            // pass 1: row maximum of this tile
            float tmax = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tS + c * 32, v);
                ptx::tmem_ld_wait();
                if (c * 32 + 32 <= nvalid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < nvalid) tmax = fmaxf(tmax, __uint_as_float(v[i]));
                }
            }
            tmax *= LOG2E;
            if (j == 0) {
                m_ref = tmax;
            } else {
                const bool need = tmax > m_ref + RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, need)) {
                    // rare: rescale O (TMEM) and l to the new reference; PV_{j-1} has retired (s_full(j) is
                    // committed after it) and PV_j cannot start before this thread arrives on p_full
                    const float alpha = need ? ptx::ex2_approx(m_ref - tmax) : 1.0f;
                    if (need) m_ref = tmax;
                    l *= alpha;
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
            }
            // pass 2: P = exp2(s * log2e - m_ref) -> bf16 pairs in TMEM; row sum in fp32
#pragma unroll 1
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tS + c * 32, v);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[i]), LOG2E, -m_ref));
                    float p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[i + 1]), LOG2E, -m_ref));
                    if (c * 32 + i >= nvalid) p0 = 0.f;
                    if (c * 32 + i + 1 >= nvalid) p1 = 0.f;
                    l += p0 + p1;
                    __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
                }
                ptx::tmem_st_32x16(tP + c * 16, pk);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_full);
        }
        // final: O / l -> bf16 -> global (each thread owns one 128-byte row segment)
        ptx::mbar_wait(o_full, (T - 1) & 1);
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
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    CUtensorMap tm[3];
    for (int i = 0; i < 3; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)3 * D * 2 * N};
        const cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)BKV, 1};
        DAD_TRY(make_tmap_bf16(&tm[i], qkv + static_cast<long long>(i) * D, 3, dims, strides, box));
    }
    const dim3 grid(cdiv(N, BQ), heads, B);
    attention_tc_kernel<<<grid, ATT_THREADS, ATT_SMEM, st>>>(tm[0], tm[1], tm[2], out, N, D);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
