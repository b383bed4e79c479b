// Fused multi-head attention, occupancy-pipelined variant (sm_100a), head_dim 64, non-causal, no mask
// (reference dinov2_layers/attention.py:49-62; 64^-0.5 is folded into the packed qkv weights).
//
// The softmax of this shape is bounded by the MUFU (16 exp2 / clk / SM) at twice the tensor-pipe time, so the job is
// to keep the MUFU busy.  attention_tc.cu does it inside one CTA (S / P double-buffered in TMEM, 64 scores held in
// registers, 2 CTAs per SM): ncu shows the MUFU 67 % busy with two softmax warps per scheduler.  Here every CTA is
// strictly serial - S = Q K^T -> softmax -> O += P V -> next key tile - and uses only 128 TMEM columns, 49 KB of
// shared memory and <= 80 registers, so FOUR CTAs are resident per SM and each scheduler always has four softmax
// warps from four CTAs in different phases to feed the MUFU.
//   TMEM   [0, 64)   S (fp32 scores); P (bf16 pairs, 32 columns) overwrites it in place, 16 scores behind the reads
//          [64, 128) O (fp32 accumulator)
//   warps 0-3  softmax, thread = query row: pass 1 row maximum of the tile (exact online softmax; O / l are only
//              rescaled when the maximum grows by more than 2^8), pass 2 in four 16-column pieces exp2 -> row sum ->
//              packed bf16 P written back to TMEM, consumed by the second MMA as its A operand
//   warp 4     TMA producer: Q once, K / V tiles (64 keys) through a 2-stage ring, 3-D maps (tail rows zero-filled)
//   warp 5     tcgen05.mma issuer
#include <cstdlib>

#include "elementwise.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;      // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;    // 8 KB
constexpr int KV_STAGES = 2;
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 128;
constexpr int S_COL = 0, O_COL = 64;
constexpr int ATT_SMEM = Q_BYTES + 2 * KV_STAGES * KV_BYTES + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P <= 2^8 relative to the reference maximum

__global__ void __launch_bounds__(ATT_THREADS, 4)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
    uint64_t* s_full = bars + 1 + 3 * KV_STAGES;      // S_j landed in TMEM
    uint64_t* p_full = s_full + 1;                    // P_j written (128 arrivals)
    uint64_t* o_full = s_full + 2;                    // one completion per P V
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
        // -------------------------------------------------------------------- MMA issuer (one elected lane)
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(BQ, HD);
        constexpr uint32_t kDescHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo = ptx::desc_lo_sw128(ptx::smem_u32(sQ));
        const uint32_t k_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(sK));
        const uint32_t v_lo0 = ptx::desc_lo_mn_sw128(ptx::smem_u32(sV));
        ptx::mbar_wait(q_full, 0);
        for (int j = 0; j < T; ++j) {
            const int s = j % KV_STAGES;
            const uint32_t kvph = (j / KV_STAGES) & 1;
            // S = Q K_j^T.  Issued after P V_{j-1}: the tensor pipe executes in issue order, so P_{j-1} (which lives in
            // the S columns) has been consumed before this MMA overwrites it.
            ptx::mbar_wait(&k_full[s], kvph);
            ptx::tc_fence_after();
            const uint32_t k_lo = k_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    ptx::umma_bf16(tmem + S_COL, ptx::make_desc(q_lo + 2 * k, ptx::kDescHiSw128),
                                   ptx::make_desc(k_lo + 2 * k, ptx::kDescHiSw128), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(s_full);
            }
            __syncwarp();
            // O += P_j V_j
            ptx::mbar_wait(p_full, j & 1);
            ptx::mbar_wait(&v_full[s], kvph);
            ptx::tc_fence_after();
            const uint32_t v_lo = v_lo0 + s * (KV_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)   // 16 keys = 16 rows of 128 B per k-step
                    ptx::umma_bf16_ts(tmem + O_COL, tmem + S_COL + k * 8,
                                      ptx::make_desc(v_lo + k * (16 * 128 >> 4), kDescHiMn), idesc_pv, (j | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&kv_empty[s]);
                ptx::umma_commit(o_full);
            }
            __syncwarp();
        }
        pdl_launch_dependents();  // last MMA issued: the next kernel's prologue may overlap this CTA's drain
    } else {
        // -------------------------------------------------------------------- softmax (warps 0-3), thread = query row
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t tS = tmem + lane_base + S_COL, tO = tmem + lane_base + O_COL;
        float m_ref = -INFINITY;  // reference maximum (log2 domain) that P / O / l are relative to
        float l = 0.f;
        for (int j = 0; j < T; ++j) {
            const int nvalid = min(BKV, N - j * BKV);
            ptx::mbar_wait(s_full, j & 1);
            ptx::tc_fence_after();
            // ---- pass 1: row maximum of the tile (two 32-column loads, independent max chains).  Measured: this pass is
            // free (0.371 ms with it, 0.369 ms with a sum-based overflow check instead) because the CTA is bound by
            // its serial latency chain, not by issue slots; it keeps the softmax exact with a cheap rescale path.
            float tmax;
            {
                uint32_t a[32];
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    ptx::tmem_ld_32x32(tS + hlf * 32, a);
                    ptx::tmem_ld_wait();
                    if (nvalid == BKV) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            m0 = fmaxf(m0, __uint_as_float(a[i])); m1 = fmaxf(m1, __uint_as_float(a[i + 1]));
                            m2 = fmaxf(m2, __uint_as_float(a[i + 2])); m3 = fmaxf(m3, __uint_as_float(a[i + 3]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (hlf * 32 + i < nvalid) m0 = fmaxf(m0, __uint_as_float(a[i]));
                    }
                }
                tmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * LOG2E;
            }
            // ---- lazy rescale: only when the maximum outgrows the reference by more than 2^8 (exact otherwise too:
            // P <= 2^8 is far inside the bf16 / fp32 range)
            const bool need = tmax > m_ref + RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
                float alpha = 1.0f;
                if (need) {
                    alpha = ptx::ex2_approx(m_ref - tmax);  // 0 for the very first tile (m_ref = -inf)
                    m_ref = tmax;
                    l *= alpha;
                }
                if (j > 0) {  // O holds P V_0..j-1: wait for P V_{j-1} to retire, then scale this row
                    ptx::mbar_wait(o_full, (j - 1) & 1);
                    ptx::tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < HD / 16; ++c) {
                        uint32_t o[16];
                        ptx::tmem_ld_32x16(tO + c * 16, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ptx::tmem_st_32x16(tO + c * 16, o);
                    }
                    ptx::tmem_st_wait();
                }
            }
            // ---- pass 2: P = exp2(S * log2e - m_ref) in 16-column pieces; the packed P piece k (8 columns) lands on
            // S columns [8k, 8k + 8), all of which have been read by then
            float s0 = 0.f, s1 = 0.f;
            uint32_t x[2][16];
            ptx::tmem_ld_32x16(tS, x[0]);
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) {
                ptx::tmem_ld_wait();
                if (k + 1 < BKV / 16) ptx::tmem_ld_32x16(tS + (k + 1) * 16, x[(k + 1) & 1]);
                const uint32_t (&v)[16] = x[k & 1];
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[i]), LOG2E, -m_ref));
                    float p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[i + 1]), LOG2E, -m_ref));
                    if (nvalid != BKV) {
                        if (k * 16 + i >= nvalid) p0 = 0.f;
                        if (k * 16 + i + 1 >= nvalid) p1 = 0.f;
                    }
                    s0 += p0;
                    s1 += p1;
                    __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
                }
                // columns [8k, 8k + 8) belong to S pieces <= k / 2, which are in registers already; the piece in flight
                // (k + 1) starts at column 16k + 16 and is never touched by this store
                ptx::tmem_st_32x8(tS + k * 8, pk);
            }
            l += s0 + s1;
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
        for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            ptx::tmem_ld_32x16(tO + c * 16, o);
            ptx::tmem_ld_wait();
            if (row < N) {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint4 w;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                    w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
                    w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
                    *reinterpret_cast<uint4*>(dst + c * 16 + i) = w;
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
int attention_tc3(const bf16* qkv, bf16* out, int B, int N, int heads, cudaStream_t st) {
    const int D = heads * HD;
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
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
    DAD_CHECK_CUDA(launch_pdl(attention_tc3_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, st, tm[0], tm[1], tm[2], out, N, D));
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
