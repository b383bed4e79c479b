// 2-CTA (cta_group::2) halo-mode 3x3 convolution on the 5th-gen tensor cores: the decoder's wide convs
// (N = 128 / 256 output channels; RefineNet residual units, layer_rn, output_conv1).
//
// The 1-CTA implicit GEMM (gemm_tc.cu) re-fetches the input patch once per tap and the whole weight tile per
// 128-pixel tile: 48 KB of L2 -> smem traffic per 512 MMA cycles (96 B/clk/SM), which is what bounds it (ncu:
// tensor pipe 51 % active, DRAM 9 %).  Here
//   * a cluster of two CTAs computes 2 x (16 x 8 pixels) x BN channels with UMMA 256 x BN x 16: each CTA stages
//     only HALF of every weight tile (the pair's MMA reads both halves), and
//   * each CTA fetches ONE 18 x 16-pixel halo box per 64-channel chunk and all nine taps read it through
//     row-shifted UMMA descriptors (the 128B swizzle is a function of the absolute smem address for TMA and UMMA
//     alike, so shifted views need no fix-up),
// so the fill traffic drops to (36 KB + 9 x BN/2 x 128 B) per 9 k-blocks: 39 B/clk/SM at BN = 256.
// Warp roles per CTA as in gemm_tc2.cu: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2-9 the
// fused epilogue (bias / ReLU / up to two bf16 residuals / ReLU copy) on the CTA's own 128 accumulator rows.
#include <cuda.h>

#include <cstdlib>

#include "epilogue.cuh"
#include "gemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int TW_LOG2 = 3, TW = 8, TH = 16;                         // output pixels per CTA tile
constexpr int HALO_W = 16, HALO_H = 18, HALO_BYTES = HALO_W * HALO_H * 128;  // 36 KB (box width 16: 2048 B row pitch)
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int STG_LD = 20;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

template <int BN>
struct Cfg {
    static constexpr int B_STAGE = (BN / 2) * BK * 2;               // this CTA's half of a (tap, chunk) weight tile
    static constexpr int SA = (BN == 256) ? 2 : 3;                  // halo ring
    static constexpr int SB = (BN == 256) ? 8 : 10;                 // weight ring
    static constexpr int RING_BYTES = SA * HALO_BYTES + SB * B_STAGE;
    static constexpr int STAGING_OFF = RING_BYTES + 1024;           // barriers live in the 1 KB before it
    static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 32 * STG_LD * 4;
    static constexpr int SMEM_BYTES = STAGING_OFF + STAGING_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct Conv2Args {
    Epilogue epi;
    int N, B, H, W;
    int cchunks, tiles_x, tiles_y, num_m_tiles, num_n_tiles;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1,
                                                int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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

struct TileCoord {
    int b, y0, x0;
    bool valid;
};
__device__ __forceinline__ TileCoord tile_coord(const Conv2Args& g, int mt) {
    TileCoord t;
    const int per_img = g.tiles_x * g.tiles_y;
    t.valid = mt < g.num_m_tiles;
    t.b = mt / per_img;             // an invalid (odd tail) tile lands at b >= B: TMA zero-fills, nothing is stored
    const int r = mt - t.b * per_img;
    const int ty = r / g.tiles_x;
    t.y0 = ty * TH;
    t.x0 = (r - ty * g.tiles_x) << TW_LOG2;
    return t;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ Conv2Args g) {
    using C = Cfg<BN>;
    constexpr int SA = C::SA, SB = C::SB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + SA * HALO_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
    uint64_t* full = bars;                   // [SB] weights landed (leader's copy is used)
    uint64_t* empty = bars + SB;             // [SB] per CTA (multicast commit)
    uint64_t* afull = bars + 2 * SB;         // [SA] halos landed (leader's copy is used)
    uint64_t* aempty = afull + SA;           // [SA] per CTA (multicast commit)
    uint64_t* tfull = aempty + SA;           // [2]  per CTA (multicast commit)
    uint64_t* tempty = tfull + 2;            // [2]  leader's copy collects both CTAs' epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < SB; ++i) {
                ptx::mbar_init(&full[i], 1);
                ptx::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < SA; ++i) {
                ptx::mbar_init(&afull[i], 1);
                ptx::mbar_init(&aempty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&tfull[i], 1);
                ptx::mbar_init(&tempty[i], 2 * NUM_EPI_WARPS);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                     "r"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / complete_tx
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // prologue above overlaps the previous kernel's tail; global memory is touched only below

    const int num_pair_tiles = ((g.num_m_tiles + 1) >> 1) * g.num_n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer (both CTAs)
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
            for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters) {
                const int mp = pt / g.num_n_tiles;
                const int n0 = (pt - mp * g.num_n_tiles) * BN;
                const TileCoord t = tile_coord(g, 2 * mp + static_cast<int>(rank));
                for (int cc = 0; cc < g.cchunks; ++cc) {
                    ptx::mbar_wait(&aempty[sa], pa ^ 1);
                    if (leader) ptx::mbar_arrive_expect_tx(&afull[sa], 2 * HALO_BYTES);
                    tma_load_4d_cg2(sA_u + sa * HALO_BYTES, &tmA, ptx::smem_u32(&afull[sa]) & PEER_MASK, cc * BK, t.x0 - 1,
                                    t.y0 - 1, t.b);
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                    for (int tap = 0; tap < 9; ++tap) {
                        ptx::mbar_wait(&empty[sb], pb ^ 1);
                        if (leader) ptx::mbar_arrive_expect_tx(&full[sb], 2 * C::B_STAGE);
                        tma_load_2d_cg2(sB_u + sb * C::B_STAGE, &tmB, ptx::smem_u32(&full[sb]) & PEER_MASK,
                                        (tap * g.cchunks + cc) * BK, n0 + static_cast<int>(rank) * (BN / 2));
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader) {
            // ------------------------------------------------ MMA issuer (leader CTA, one elected lane, both SMs)
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN);
            const uint32_t sA_lo = ptx::desc_lo_sw128(ptx::smem_u32(sA)), sB_lo = ptx::desc_lo_sw128(ptx::smem_u32(sB));
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int cc = 0; cc < g.cchunks; ++cc) {
                    ptx::mbar_wait(&afull[sa], pa);
                    const uint32_t halo_lo = sA_lo + sa * (HALO_BYTES >> 4);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        ptx::mbar_wait(&full[sb], pb);
                        ptx::tc_fence_after();
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t a_lo = halo_lo + (dy * HALO_W + dx) * (128 >> 4);
                        const uint32_t b_lo = sB_lo + sb * (C::B_STAGE >> 4);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_bf16_cg2(d_tmem, ptx::make_desc(a_lo + 2 * k, ptx::kDescHiSw128Halo),
                                              ptx::make_desc(b_lo + 2 * k, ptx::kDescHiSw128), idesc,
                                              (cc | tap | k) != 0 ? 1u : 0u);
                            umma_commit_mc(&empty[sb], 3);
                            if (tap == 8) umma_commit_mc(&aempty[sa], 3);  // the nine taps have read this halo
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                    if (++sa == SA) { sa = 0; pa ^= 1; }
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
        // ---------------------------------------------------- epilogue (warps 2..9, both CTAs)
        // TMEM -> registers (thread = row, 16 columns) -> padded smem transpose -> (row, 4 columns) groups: each
        // global access of a warp covers 8 rows x 32 contiguous bytes; two warps share a TMEM lane quarter.
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t stg = ptx::smem_u32(smem + C::STAGING_OFF) + (warp - 2) * (32 * STG_LD * 4);
        const int cg = (lane & 3) * 4;
        constexpr int NPC = BN / 16;
        const int c_lo = half * (NPC / 2), c_hi = c_lo + NPC / 2;
        int as = 0;
        uint32_t aphase = 0;
        for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters) {
            const int mp = pt / g.num_n_tiles;
            const int n0 = (pt - mp * g.num_n_tiles) * BN;
            const TileCoord t = tile_coord(g, 2 * mp + static_cast<int>(rank));
            bool ok[4];
            long long grow4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = quarter * 32 + i * 8 + (lane >> 2);
                const int y = t.y0 + (r >> TW_LOG2), x = t.x0 + (r & (TW - 1));
                ok[i] = t.valid && y < g.H && x < g.W;
                grow4[i] = (static_cast<long long>(t.b) * g.H + y) * g.W + x;
            }
            // residual rows of this CTA's NEXT tile -> L2 while the current tile's MMAs run
            if (g.epi.res1 != nullptr && pt + num_clusters < num_pair_tiles) {
                const int npt = pt + num_clusters;
                const int nmp = npt / g.num_n_tiles;
                const int nn0 = (npt - nmp * g.num_n_tiles) * BN;
                const TileCoord nt = tile_coord(g, 2 * nmp + static_cast<int>(rank));
                const int es1 = g.epi.res1_bf16 ? 2 : 4;
                const int lines = (BN * es1) >> 7;
                if (nt.valid) {
                    for (int l = (warp - 2) * 32 + lane; l < BM * lines; l += 32 * NUM_EPI_WARPS) {
                        const int r = l / lines, ln = l - r * lines;
                        const int y = nt.y0 + (r >> TW_LOG2), x = nt.x0 + (r & (TW - 1));
                        if (y < g.H && x < g.W && nn0 + ln * (128 / es1) < g.N) {
                            const long long nrow = (static_cast<long long>(nt.b) * g.H + y) * g.W + x;
                            const long long boff = (nrow * g.epi.ldc + nn0) * es1 + ln * 128;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(g.epi.res1) + boff));
                            if (g.epi.res2 != nullptr)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(g.epi.res2) + boff));
                        }
                    }
                }
            }
            // bf16 residual rows stay RAW (8 bytes = 4 channels per row) in a rotating register window PD pieces deep:
            // the loads of piece c + PD are issued while piece c is finished, so their L2 / HBM latency is covered by
            // PD pieces of work instead of being exposed once per piece (measured: the residual convs ran epilogue-
            // bound at 1.5 ms against 0.63 ms for the same convolution without residuals).
            constexpr int PD = 3;
            constexpr int NP = NPC / 2;  // pieces per warp
            const bool has_r1 = g.epi.res1 != nullptr, has_r2 = g.epi.res2 != nullptr;
            long long off4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) off4[i] = grow4[i] * g.epi.ldc + n0 + cg;
            uint2 r1raw[PD][4], r2raw[PD][4];
            auto issue = [&](int pc, int slot) {
                const int col = n0 + (c_lo + pc) * 16 + cg;
                if (col < g.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (ok[i]) {
                            const long long o = off4[i] + (c_lo + pc) * 16;
                            if (has_r1) r1raw[slot][i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(g.epi.res1) + o));
                            if (has_r2) r2raw[slot][i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(g.epi.res2) + o));
                        }
                    }
                }
            };
            if (has_r1 || has_r2) {
#pragma unroll
                for (int d = 0; d < PD; ++d)
                    if (d < NP) issue(d, d);
            }
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll
            for (int pc = 0; pc < NP; ++pc) {
                const int c = c_lo + pc;
                const int slot = pc % PD;
                const int col = n0 + c * 16 + cg;
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col < g.N && g.epi.bias != nullptr) b4 = *reinterpret_cast<const float4*>(g.epi.bias + col);
                uint32_t v[16];
                ptx::tmem_ld_32x16(t_row + c * 16, v);
                ptx::tmem_ld_wait();
                if (pc == NP - 1) {  // this warp's share of the accumulator is in registers: release the TMEM stage
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ptx::smem_u32(&tempty[as]) & PEER_MASK);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    ptx::sts128(stg + (lane * STG_LD + 4 * j) * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                __syncwarp();
                if (col < g.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (ok[i]) {
                            const int rr = i * 8 + (lane >> 2);
                            const float4 q = ptx::lds128(stg + (rr * STG_LD + cg) * 4);
                            float f[4] = {q.x + b4.x, q.y + b4.y, q.z + b4.z, q.w + b4.w};
                            if (g.epi.act == ACT_RELU) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) f[e] = fmaxf(f[e], 0.f);
                            }
                            if (has_r1) {
                                const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&r1raw[slot][i].x);
                                const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&r1raw[slot][i].y);
                                f[0] += __low2float(lo); f[1] += __high2float(lo); f[2] += __low2float(hi); f[3] += __high2float(hi);
                            }
                            if (has_r2) {
                                const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&r2raw[slot][i].x);
                                const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&r2raw[slot][i].y);
                                f[0] += __low2float(lo); f[1] += __high2float(lo); f[2] += __low2float(hi); f[3] += __high2float(hi);
                            }
                            const long long o = off4[i] + c * 16;
                            if (g.epi.out_bf16) {
                                store4_bf16(reinterpret_cast<bf16*>(g.epi.out), o, f);
                            } else {
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.epi.out) + o) = make_float4(f[0], f[1], f[2], f[3]);
                            }
                            if (g.epi.out_relu != nullptr) {
                                float r[4] = {fmaxf(f[0], 0.f), fmaxf(f[1], 0.f), fmaxf(f[2], 0.f), fmaxf(f[3], 0.f)};
                                if (g.epi.out_bf16) store4_bf16(reinterpret_cast<bf16*>(g.epi.out_relu), o, r);
                                else *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.epi.out_relu) + o) = make_float4(r[0], r[1], r[2], r[3]);
                            }
                        }
                    }
                }
                if ((has_r1 || has_r2) && pc + PD < NP) issue(pc + PD, slot);
                __syncwarp();
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    ptx::tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the peer can still touch its smem / barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    }
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Conv2Args& a, double flops, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg<BN>::SMEM_BYTES));
        configured = true;
    }
    const int pair_tiles = ((a.num_m_tiles + 1) / 2) * a.num_n_tiles;
    const int pairs = num_sms() / 2;
    const int grid = 2 * (pair_tiles < pairs ? pair_tiles : pairs);
    ProfScope prof(PROF_GEMM_TC, flops, stream);
    DAD_CHECK_CUDA(launch_pdl(conv_tc2_kernel<BN>, dim3(grid), dim3(NUM_THREADS), Cfg<BN>::SMEM_BYTES, stream, tmA, tmB, a));
    return DAD_OK;
}

}  // namespace

bool conv_tc2_eligible(const GemmProblem& p) {
    static const bool off = getenv("DAD_NO_CONV2") != nullptr;  // A/B switch
    if (off || !p.conv || p.stride != 1 || p.taps != 9 || (p.N % 128) != 0 || p.C % 8 != 0 || p.ldp % 8 != 0) return false;
    const Epilogue& e = p.epi;
    if (e.scat_k || e.rowtab || e.head_out || e.gamma || e.act == ACT_GELU || !e.out) return false;
    if ((e.res1 && !e.res1_bf16) || (e.res2 && !e.res2_bf16)) return false;  // the pipelined epilogue keeps bf16 residuals raw
    return p.Kp == 9 * cdiv(p.C, BK) * BK;
}

int conv_tc2(const GemmProblem& p, cudaStream_t stream) {
    DAD_REQUIRE(conv_tc2_eligible(p), "conv_tc2: problem not eligible for the 2-CTA halo convolution");
    const int bn = (p.N % 256 == 0) ? 256 : 128;
    Conv2Args a{};
    a.epi = p.epi;
    a.N = p.N; a.B = p.B; a.H = p.H; a.W = p.W;
    a.cchunks = cdiv(p.C, BK);
    a.tiles_x = cdiv(p.W, TW);
    a.tiles_y = cdiv(p.H, TH);
    a.num_m_tiles = p.B * a.tiles_x * a.tiles_y;
    a.num_n_tiles = p.N / bn;
    CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
        const cuuint64_t strides[3] = {(cuuint64_t)p.ldp * 2, (cuuint64_t)p.ldp * 2 * p.W, (cuuint64_t)p.ldp * 2 * p.W * p.H};
        const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1};
        DAD_TRY(make_tmap_bf16(&tmA, p.A, 4, dims, strides, box));
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p.Kp, (cuuint64_t)p.N};
        const cuuint64_t strides[1] = {(cuuint64_t)p.Kp * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(bn / 2)};
        DAD_TRY(make_tmap_bf16(&tmB, p.Wt, 2, dims, strides, box));
    }
    const double flops = 2.0 * p.B * p.H * p.W * p.N * (9.0 * p.C);
    return bn == 256 ? launch<256>(tmA, tmB, a, flops, stream) : launch<128>(tmA, tmB, a, flops, stream);
}

}  // namespace dad
