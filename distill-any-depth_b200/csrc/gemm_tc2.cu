// 2-CTA (cta_group::2) variant of the tcgen05 GEMM for the wide encoder GEMMs (N % 256 == 0).
//
// A cluster of two CTAs (one SM pair) computes a 256 x 256 tile with UMMA 256 x 256 x 16: each CTA stages its own
// 128 rows of A and HALF of the B tile (128 of the 256 N rows), so per-SM shared-memory fill traffic drops from
// 48 KB to 32 KB per k-block and the ring deepens from 4 to 6 stages.  The leader CTA's MMA thread issues for
// both SMs; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; both CTAs' TMA loads
// complete on the leader's mbarrier; each CTA drains its own 128 accumulator rows from its TMEM with the same
// TMA-store / reduce-add epilogues as the 1-CTA kernel (bias->bf16, bias+GELU->bf16, x += gamma*(acc+bias)).
#include <cuda.h>

#include "epilogue.cuh"
#include "gemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BNH = 128;  // B rows staged per CTA
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int A_STAGE = BM * BK * 2, B_STAGE = BNH * BK * 2, STAGE_BYTES = A_STAGE + B_STAGE;  // 32 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int STAGING_OFF = STAGES * STAGE_BYTES + 1024;
constexpr int SMEM_BYTES = STAGING_OFF + 2 * BM * 128 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

struct Tc2Args {
    Epilogue epi;
    int M, N, num_k_blocks, num_m_tiles, num_n_tiles;  // tiles of 256 x 256
    // Tail splitting: the last `tiles % pairs` tiles (the partial wave of the persistent grid) are cut along N into `split`
    // sub-items of 256 x (256 / split) each, so that the partial wave costs 1 / split of a tile time.  Same K loop per
    // output element, so results are bit-identical to the unsplit schedule.
    int full_items, split, num_items;
};

// item -> (tile, first column inside the tile, item width)
__device__ __forceinline__ void decode_item(const Tc2Args& g, int item, int& tile, int& sub0, int& width) {
    if (item < g.full_items) { tile = item; sub0 = 0; width = BN; return; }
    const int t = item - g.full_items;
    tile = g.full_items + t / g.split;
    width = BN / g.split;
    sub0 = (t % g.split) * width;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmBs, const __grid_constant__ CUtensorMap tmC,
                const __grid_constant__ Tc2Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // used in the leader CTA only
    uint64_t* empty = bars + STAGES;        // per CTA (multicast commit)
    uint64_t* tfull = bars + 2 * STAGES;    // per CTA (multicast commit)
    uint64_t* tempty = bars + 2 * STAGES + 2;  // leader's copy collects both CTAs' epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        ptx::prefetch_tmap(&tmBs);
        ptx::prefetch_tmap(&tmC);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                ptx::mbar_init(&full[i], 1);
                ptx::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&tfull[i], 1);
                ptx::mbar_init(&tempty[i], 2 * NUM_EPI_WARPS);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / complete_tx
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // prologue above overlaps the previous kernel's tail; global memory is touched only below

    const int num_items = g.num_items;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int nkb = g.num_k_blocks;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer (both CTAs)
            int stage = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                int tile, sub0, width;
                decode_item(g, item, tile, sub0, width);
                const int mt = tile / g.num_n_tiles;
                const int nt = tile - mt * g.num_n_tiles;
                const int row0 = mt * (2 * BM) + static_cast<int>(rank) * BM;
                const int nrow0 = nt * BN + sub0 + static_cast<int>(rank) * (width >> 1);   // each CTA stages half of the B rows
                const bool fullw = width == BN;
                const uint32_t tx = 2 * (A_STAGE + (width >> 1) * BK * 2);
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    const uint32_t lbar = ptx::smem_u32(&full[stage]) & PEER_MASK;
                    if (leader) ptx::mbar_arrive_expect_tx(&full[stage], tx);
                    tma_load_2d_cg2(sA + stage * A_STAGE, &tmA, lbar, kb * BK, row0);
                    tma_load_2d_cg2(sB + stage * B_STAGE, fullw ? &tmB : &tmBs, lbar, kb * BK, nrow0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();  // reconverge before the (warp-aligned) cluster barrier below
    } else if (warp == 1) {
        if (leader) {
            // ------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop, one elected lane
            // issues for both SMs; descriptors = constant high word + (address >> 4))
            constexpr uint32_t idesc_full = ptx::make_idesc_bf16(2 * BM, BN);
            const uint32_t sA_lo = ptx::desc_lo_sw128(ptx::smem_u32(sA)), sB_lo = ptx::desc_lo_sw128(ptx::smem_u32(sB));
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                int tile, sub0, width;
                decode_item(g, item, tile, sub0, width);
                const uint32_t idesc = width == BN ? idesc_full : ptx::make_idesc_bf16(2 * BM, static_cast<uint32_t>(width));
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = sA_lo + stage * (A_STAGE >> 4), b_lo = sB_lo + stage * (B_STAGE >> 4);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_cg2(d_tmem, ptx::make_desc(a_lo + 2 * k, ptx::kDescHiSw128),
                                          ptx::make_desc(b_lo + 2 * k, ptx::kDescHiSw128), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_mc(&empty[stage], 3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) umma_commit_mc(&tfull[as], 3);
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
        pdl_launch_dependents();  // last MMA issued (leader) / nothing to issue (peer): let the next kernel's prologue start
        __syncwarp();
    } else {
        // ---------------------------------------------------- epilogue (warps 2..9, both CTAs): TMA store / reduce-add
        constexpr bool F32 = KIND == EK_RES_F32;
        constexpr int CW = F32 ? 32 : 64;
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tile_stg = ptx::smem_u32(smem + STAGING_OFF) + half * (BM * 128);
        const int r_tile = quarter * 32 + lane;
        const bool issuer = (quarter == 0) && (lane == 0);
        const uint32_t row_addr = tile_stg + r_tile * 128;
        int as = 0;
        uint32_t aphase = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            int tile, sub0, width;
            decode_item(g, item, tile, sub0, width);
            const int mt = tile / g.num_n_tiles;
            const int n0 = (tile - mt * g.num_n_tiles) * BN + sub0;
            const int row0 = mt * (2 * BM) + static_cast<int>(rank) * BM;
            // chunks of this item, split over the two column halves (a narrow item may leave half 1 without work)
            const int nch = width / CW;
            const int c_lo = half == 0 ? 0 : (nch + 1) / 2, c_hi = half == 0 ? (nch + 1) / 2 : nch;
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            if (c_lo >= c_hi) {   // nothing to convert for this half: still release the TMEM stage (16 arrivals expected)
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ptx::smem_u32(&tempty[as]) & PEER_MASK);
            }
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                const int col0 = n0 + c * CW;
                uint32_t v[CW];
                {
                    uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
                    ptx::tmem_ld_32x32(t_row + c * CW, lo);
                    if constexpr (!F32) {
                        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
                        ptx::tmem_ld_32x32(t_row + c * CW + 32, hi);
                    }
                    ptx::tmem_ld_wait();
                }
                if (c == c_hi - 1) {  // accumulator share is in registers: release the TMEM stage to the leader's MMA thread
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ptx::smem_u32(&tempty[as]) & PEER_MASK);
                }
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < CW; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(g.epi.bias + col0 + i);  // warp-uniform address
                    float f0 = __uint_as_float(v[i]) + b4.x, f1 = __uint_as_float(v[i + 1]) + b4.y;
                    float f2 = __uint_as_float(v[i + 2]) + b4.z, f3 = __uint_as_float(v[i + 3]) + b4.w;
                    if constexpr (KIND == EK_GELU_BF16) {
                        f0 = gelu_fast(f0); f1 = gelu_fast(f1); f2 = gelu_fast(f2); f3 = gelu_fast(f3);
                    }
                    if constexpr (F32) {
                        const float4 g4 = *reinterpret_cast<const float4*>(g.epi.gamma + col0 + i);
                        w[i] = __float_as_uint(f0 * g4.x); w[i + 1] = __float_as_uint(f1 * g4.y);
                        w[i + 2] = __float_as_uint(f2 * g4.z); w[i + 3] = __float_as_uint(f3 * g4.w);
                    } else {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(f0, f1), p1 = __floats2bfloat162_rn(f2, f3);
                        w[i >> 1] = *reinterpret_cast<uint32_t*>(&p0);
                        w[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
                    }
                }
                if (issuer) ptx::bulk_wait_read0();
                ptx::named_bar_sync(1 + half, 128);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    ptx::sts128u(row_addr + ((j ^ (r_tile & 7)) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                ptx::fence_proxy_async_smem();
                ptx::named_bar_sync(1 + half, 128);
                if (issuer) {
                    if constexpr (F32) ptx::tma_reduce_add_2d(&tmC, tile_stg, col0, row0);
                    else ptx::tma_store_2d(&tmC, tile_stg, col0, row0);
                    ptx::bulk_commit();
                }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (issuer) ptx::bulk_wait0();
        __syncwarp();
    }

    ptx::tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the peer can still touch its smem / barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int KIND>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBs, const CUtensorMap& tmC, const Tc2Args& a,
            double flops, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int pairs = num_sms() / 2;
    const int grid = 2 * (a.num_items < pairs ? a.num_items : pairs);
    ProfScope prof(PROF_GEMM_TC, flops, stream);
    DAD_CHECK_CUDA(launch_pdl(gemm_tc2_kernel<KIND>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, tmBs, tmC, a));
    return DAD_OK;
}

}  // namespace

bool gemm_tc2_eligible(const GemmProblem& p) {
    if (p.conv || p.N % BN != 0 || p.M < 2 * BM || p.K % 8 != 0) return false;
    const int kind = epilogue_kind(p.epi);
    return kind == EK_BIAS_BF16 || kind == EK_GELU_BF16 || kind == EK_RES_F32;
}

int gemm_tc2(const GemmProblem& p, cudaStream_t stream) {
    DAD_REQUIRE(gemm_tc2_eligible(p), "gemm_tc2: problem not eligible for the 2-CTA kernel");
    const int kind = epilogue_kind(p.epi);
    Tc2Args a{};
    a.epi = p.epi;
    a.M = p.M; a.N = p.N;
    a.num_k_blocks = cdiv(p.K, BK);
    a.num_m_tiles = cdiv(p.M, 2 * BM);
    a.num_n_tiles = p.N / BN;
    {   // tail splitting (see Tc2Args): the partial last wave of `rem` tiles becomes rem * split narrower items
        const int tiles = a.num_m_tiles * a.num_n_tiles, pairs = num_sms() / 2;
        const int rem = tiles % pairs;
        static const bool off = getenv("DAD_NO_TAIL_SPLIT") != nullptr;   // A/B switch
        int split = 1;
        if (!off && rem > 0 && tiles > pairs) {
            if (4 * rem <= pairs) split = 4;
            else if (2 * rem <= pairs) split = 2;
        }
        a.split = split;
        a.full_items = split > 1 ? tiles - rem : tiles;
        a.num_items = a.full_items + (split > 1 ? rem * split : 0);
    }
    CUtensorMap tmA, tmB, tmBs, tmC;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
        const cuuint64_t strides[1] = {(cuuint64_t)p.lda * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
        DAD_TRY(make_tmap_bf16(&tmA, p.A, 2, dims, strides, box));
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p.Kp, (cuuint64_t)p.N};
        const cuuint64_t strides[1] = {(cuuint64_t)p.Kp * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BNH};
        DAD_TRY(make_tmap_bf16(&tmB, p.Wt, 2, dims, strides, box));
        const cuuint32_t boxs[2] = {(cuuint32_t)BK, (cuuint32_t)(BNH / a.split)};   // B rows per CTA of a split item
        DAD_TRY(make_tmap_bf16(&tmBs, p.Wt, 2, dims, strides, boxs));
    }
    {
        const bool f32 = kind == EK_RES_F32;
        const int es = f32 ? 4 : 2, cw = f32 ? 32 : 64;
        DAD_REQUIRE((p.epi.ldc * es) % 16 == 0, "gemm_tc2: output row pitch must be a multiple of 16 bytes");
        const cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.M};
        const cuuint64_t strides[1] = {(cuuint64_t)p.epi.ldc * es};
        const cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)BM};
        DAD_TRY(make_tmap(&tmC, f32 ? 1 : 0, p.epi.out, 2, dims, strides, box));
    }
    const double flops = 2.0 * p.M * p.N * static_cast<double>(p.K);
    switch (kind) {
        case EK_BIAS_BF16: return launch2<EK_BIAS_BF16>(tmA, tmB, tmBs, tmC, a, flops, stream);
        case EK_GELU_BF16: return launch2<EK_GELU_BF16>(tmA, tmB, tmBs, tmC, a, flops, stream);
        default: return launch2<EK_RES_F32>(tmA, tmB, tmBs, tmC, a, flops, stream);
    }
}

}  // namespace dad
