// bf16 GEMM / implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 (one lane)  TMA producer: A tile (128 x 64 bf16) and B tile (BN x 64 bf16) per k-block,
//                      SWIZZLE_128B, through a STAGES-deep mbarrier ring
//   warp 1 (one lane)  tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16, fp32 accumulate in TMEM);
//                      tcgen05.commit frees smem stages and publishes finished accumulators
//   warps 2-9          epilogue: tcgen05.ld (32 lanes x 32 columns) -> smem transpose -> fused epilogue ->
//                      global, two warps per TMEM lane quarter (one column half each)
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// Convolution (3x3 / 1x1, stride 1, zero "same" padding) is an implicit GEMM: the M tile is a
// TH x TW patch of output pixels and each k-block is one (tap, 64-channel chunk); its A tile is a
// 4-D TMA box of the NHWC input at the tap-shifted coordinate, with out-of-bounds rows/channels
// zero-filled by the TMA unit (that is the padding).
#include <cuda.h>

#include <cstdlib>

#include "epilogue.cuh"
#include "gemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace dad {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quarter (column halves)
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int STG_LD = 20;        // floats per staged 16-column row: 16-byte aligned, (near) conflict-free float4 access

struct TcArgs {
    Epilogue epi;
    int M, N;
    int num_k_blocks;
    int conv, taps, cchunks, pad, stride;
    int B, H, W;
    int tw_log2, th;        // spatial tile (conv): TW = 1 << tw_log2, TH = 128 / TW
    int tiles_x, tiles_y;
    int num_m_tiles, num_n_tiles;
    int ksplit, kb_per_split;  // split-K (linear + TMA reduce-add epilogue only): work item = (k slice, tile)
    int batch_h, mt_per_batch, rows_per_batch, c_col_h;   // batched linear problems (batch_h > 0): m tile -> (image, head, local tile)
    long long c_row_b, c_row_h;
    int shift_ld, shift_off[9], shift_row[9];   // B rows are shifted views of one matrix (GemmProblem::shift_*); 0 = off
    int c_store;            // batched fp32 epilogue: plain clipped TMA store instead of reduce-add (GemmProblem::c_store)
    int mn;                 // MN-major operands (GemmProblem::mn): 0 off, 1 linear, 2 conv weight gradient over 8 x 8 patches
    double flops;           // algorithmic 2*M*N*K of this launch (host-side bookkeeping only)
};

// HALO = true: small-N 3x3 convolution with shared-memory resident weights and input patch.  The M tile is 16 x 8
// output pixels; per 64-channel chunk ONE 18 x 16-pixel halo box is fetched (36 KB) and all nine taps read it
// through row-shifted UMMA descriptors (row pitch 16 pixels = 2048 B; the 128B swizzle is a function of the
// absolute smem address for both TMA and UMMA, so shifted views need no fix-up), so the activations cross
// L2 -> smem once instead of nine times and one mbarrier round trip feeds 36 MMAs instead of 4.  The persistent
// CTA loads the whole weight set once.
constexpr int HALO_W = 16, HALO_H = 18, HALO_BYTES = HALO_W * HALO_H * 128;  // 36 KB
constexpr int WRES_BYTES = 73728;  // resident weights: 9 taps x Cp x N x 2 B (e.g. 128 -> 32 channels, or 64 -> 64)

template <int BN, bool HALO = false>
struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
    // halo mode: a ring of SA halo stages; the whole weight set (<= WRES_BYTES) stays resident in shared memory
    static constexpr int SA = 3;
    static constexpr int SB = WRES_BYTES / B_STAGE_BYTES;  // resident (tap, chunk) weight tiles
    static constexpr int RING_BYTES = HALO ? SA * HALO_BYTES + WRES_BYTES : STAGES * STAGE_BYTES;
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int STAGING_OFF = RING_BYTES + 1024;            // barriers live in the 1 KB before it
    static constexpr int STAGING_BYTES = 2 * BM * 128;               // TMA-store path: one 128-row x 128-byte tile per column half
    static constexpr int SMEM_BYTES = STAGING_OFF + STAGING_BYTES + 1024 /*align slack*/;
    static_assert(NUM_EPI_WARPS * 32 * STG_LD * 4 <= STAGING_BYTES, "generic staging must fit");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BN, int KIND, bool HALO = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ TcArgs g) {
    using C = Cfg<BN, HALO>;
    constexpr int STAGES = C::STAGES;
    constexpr int NA = HALO ? C::SA : STAGES, NB = HALO ? C::SB : STAGES;  // ring depths (A / B share one ring unless HALO)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + NA * (HALO ? HALO_BYTES : A_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
    uint64_t* full = bars;                  // [NB]  (B ring; carries A too unless HALO)
    uint64_t* empty = bars + NB;            // [NB]
    uint64_t* tfull = bars + 2 * NB;
    uint64_t* tempty = bars + 2 * NB + 2;
    uint64_t* afull = bars + 2 * NB + 4;    // [NA]  halo ring (HALO only)
    uint64_t* aempty = afull + NA;          // [NA]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + NA);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < NB; ++i) {
                ptx::mbar_init(&full[i], 1);
                ptx::mbar_init(&empty[i], 1);
            }
            if constexpr (HALO) {
                for (int i = 0; i < NA; ++i) {
                    ptx::mbar_init(&afull[i], 1);
                    ptx::mbar_init(&aempty[i], 1);
                }
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&tfull[i], 1);
                ptx::mbar_init(&tempty[i], 32 * NUM_EPI_WARPS);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched from here on

    const int num_tiles = g.num_m_tiles * g.num_n_tiles;
    const int num_items = num_tiles * g.ksplit;   // ksplit == 1 everywhere except the weight-gradient GEMMs
    const int nkb = g.num_k_blocks;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            int stage = 0, sa = 0;
            uint32_t phase = 0, pa = 0;
            bool weights_loaded = false;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int ks = item / num_tiles, tile = item - ks * num_tiles;
                const int kb0 = ks * g.kb_per_split, kb1 = min(nkb, kb0 + g.kb_per_split);
                const int mt = tile / g.num_n_tiles;
                const int n0 = (tile - mt * g.num_n_tiles) * BN;
                int brow = n0, bk = 0;   // B operand: first row / K-coordinate offset of this tile
                if (g.shift_ld) {
                    const int tap = n0 / g.shift_ld;
                    brow = g.shift_row[tap] + n0 - tap * g.shift_ld;
                    bk = g.shift_off[tap];
                }
                int b = 0, y0 = 0, x0 = 0;
                if (g.conv) {
                    const int per_img = g.tiles_x * g.tiles_y;
                    b = mt / per_img;
                    const int r = mt - b * per_img;
                    const int ty = r / g.tiles_x;
                    y0 = ty * g.th;
                    x0 = (r - ty * g.tiles_x) << g.tw_log2;
                }
                if constexpr (HALO) {
                    if (!weights_loaded) {  // once per CTA: all (chunk, tap) weight tiles, one barrier
                        weights_loaded = true;
                        ptx::mbar_arrive_expect_tx(&full[0], static_cast<uint32_t>(g.cchunks * 9 * C::B_STAGE_BYTES));
                        for (int cc = 0; cc < g.cchunks; ++cc)
                            for (int tap = 0; tap < 9; ++tap)
                                ptx::tma_load_2d(sB + (cc * 9 + tap) * C::B_STAGE_BYTES, &tmB, &full[0],
                                                 (tap * g.cchunks + cc) * BK, n0);
                    }
                    for (int cc = 0; cc < g.cchunks; ++cc) {
                        ptx::mbar_wait(&aempty[sa], pa ^ 1);
                        ptx::mbar_arrive_expect_tx(&afull[sa], HALO_BYTES);
                        ptx::tma_load_4d(sA + sa * HALO_BYTES, &tmA, &afull[sa], cc * BK, x0 - 1, y0 - 1, b);
                        if (++sa == NA) { sa = 0; pa ^= 1; }
                    }
                } else {
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    if (g.mn) {
                        // MN-major: a stage holds 64 contraction rows; per 64 columns of M / N one [64 rows][128 B] atom
                        uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
                        uint8_t* b_dst = sB + stage * C::B_STAGE_BYTES;
                        if (g.mn != 2) {
                            int zh = 0, zb = 0, ml = mt * BM;
                            if (g.batch_h) {
                                const int z = mt / g.mt_per_batch;
                                zb = z / g.batch_h;
                                zh = z - zb * g.batch_h;
                                ml = (mt - z * g.mt_per_batch) * BM;
                            }
                            if (g.mn == 1) {
#pragma unroll
                                for (int j = 0; j < BM / 64; ++j) {
                                    if (g.batch_h) ptx::tma_load_4d(a_dst + j * 8192, &tmA, &full[stage], ml + j * 64, kb * BK, zh, zb);
                                    else ptx::tma_load_2d(a_dst + j * 8192, &tmA, &full[stage], ml + j * 64, kb * BK);
                                }
                            } else if (g.batch_h) {   // mn == 3: A is K-major
                                ptx::tma_load_4d(a_dst, &tmA, &full[stage], kb * BK, ml, zh, zb);
                            } else {
                                ptx::tma_load_2d(a_dst, &tmA, &full[stage], kb * BK, ml);
                            }
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) {
                                if (g.batch_h) ptx::tma_load_4d(b_dst + j * 8192, &tmB, &full[stage], n0 + j * 64, kb * BK, zh, zb);
                                else ptx::tma_load_2d(b_dst + j * 8192, &tmB, &full[stage], n0 + j * 64, kb * BK);
                            }
                        } else {
                            const int per_img = g.tiles_x * g.tiles_y;
                            const int pb = kb / per_img, pr = kb - pb * per_img;
                            const int py = pr / g.tiles_x, px = pr - py * g.tiles_x;
                            const int tap = n0 / g.shift_ld, c0 = n0 - tap * g.shift_ld;
                            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                ptx::tma_load_4d(a_dst + j * 8192, &tmA, &full[stage], mt * BM + j * 64, px * 8, py * 8, pb);
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                ptx::tma_load_4d(b_dst + j * 8192, &tmB, &full[stage], c0 + j * 64, px * 8 + dx, py * 8 + dy, pb);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (g.conv) {
                        const int tap = kb / g.cchunks;
                        const int cc = kb - tap * g.cchunks;
                        int dy = 0, dx = 0;
                        if (g.taps == 9) { dy = tap / 3; dx = tap - dy * 3; }
                        ptx::tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], cc * BK,
                                         x0 * g.stride + dx - g.pad, y0 * g.stride + dy - g.pad, b);
                    } else if (g.batch_h) {
                        const int z = mt / g.mt_per_batch, zb = z / g.batch_h;
                        ptx::tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, (mt - z * g.mt_per_batch) * BM,
                                         z - zb * g.batch_h, zb);
                        ptx::tma_load_4d(sB + stage * C::B_STAGE_BYTES, &tmB, &full[stage], kb * BK, n0, z - zb * g.batch_h, zb);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    } else {
                        ptx::tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, mt * BM);
                    }
                    ptx::tma_load_2d(sB + stage * C::B_STAGE_BYTES, &tmB, &full[stage], kb * BK + bk, brow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer
        // The whole warp walks the pipeline with uniform control flow and ONE elected lane issues: descriptors are
        // (constant high word, low word = address >> 4), so each tcgen05.mma costs two integer adds to set up.
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
        const uint32_t sA_lo = ptx::desc_lo_sw128(ptx::smem_u32(sA)), sB_lo = ptx::desc_lo_sw128(ptx::smem_u32(sB));
        int stage = 0, sa = 0;
        uint32_t phase = 0, pa = 0;
        bool weights_ready = false;
        int as = 0;
        uint32_t aphase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int ks = item / num_tiles;
            const int kb0 = ks * g.kb_per_split, kb1 = min(nkb, kb0 + g.kb_per_split);
            ptx::mbar_wait(&tempty[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            if constexpr (HALO) {
                if (!weights_ready) {
                    weights_ready = true;
                    ptx::mbar_wait(&full[0], 0);
                }
                for (int cc = 0; cc < g.cchunks; ++cc) {
                    ptx::mbar_wait(&afull[sa], pa);
                    ptx::tc_fence_after();
                    const uint32_t halo_lo = sA_lo + sa * (HALO_BYTES >> 4);
                    const uint32_t w_lo = sB_lo + cc * 9 * (C::B_STAGE_BYTES >> 4);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t a_lo = halo_lo + ((tap / 3) * HALO_W + (tap % 3)) * (128 >> 4);
                            const uint32_t b_lo = w_lo + tap * (C::B_STAGE_BYTES >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                ptx::umma_bf16(d_tmem, ptx::make_desc(a_lo + 2 * k, ptx::kDescHiSw128Halo),
                                               ptx::make_desc(b_lo + 2 * k, ptx::kDescHiSw128), idesc,
                                               (cc | tap | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&aempty[sa]);  // the nine taps of this chunk have read the halo
                    }
                    __syncwarp();
                    if (++sa == NA) { sa = 0; pa ^= 1; }
                }
            } else {
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = sA_lo + stage * (A_STAGE_BYTES >> 4);
                    const uint32_t b_lo = sB_lo + stage * (C::B_STAGE_BYTES >> 4);
                    if (ptx::elect_one()) {
                        if (g.mn) {
                            // MN-major operands: 64-column atoms 8 KB apart (leading offset), 8 contraction rows = 1024 B
                            // (stride offset); a K step of 16 rows advances the start address by 2048 B
                            constexpr uint32_t kHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);
                            constexpr uint32_t kLbo = (8192u >> 4) << 16;
                            const uint32_t am = ((a_lo & 0xFFFFu) | kLbo), bm = ((b_lo & 0xFFFFu) | kLbo);
                            if (g.mn == 3) {   // A K-major, B MN-major
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    ptx::umma_bf16(d_tmem, ptx::make_desc(a_lo + 2 * k, ptx::kDescHiSw128), ptx::make_desc(bm + k * (2048 >> 4), kHiMn),
                                                   idesc | (1u << 16), ((kb - kb0) | k) != 0 ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    ptx::umma_bf16(d_tmem, ptx::make_desc(am + k * (2048 >> 4), kHiMn), ptx::make_desc(bm + k * (2048 >> 4), kHiMn),
                                                   idesc | (1u << 15) | (1u << 16), ((kb - kb0) | k) != 0 ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            ptx::umma_bf16(d_tmem, ptx::make_desc(a_lo + 2 * k, ptx::kDescHiSw128),
                                           ptx::make_desc(b_lo + 2 * k, ptx::kDescHiSw128), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tfull[as]);
            __syncwarp();
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        // this CTA has issued its last MMA: only the last tile's epilogue remains, so the next kernel may start its
        // prologue now (triggering at kernel start instead lets dependents squat on SM resources: measured 1.7 % slower)
        pdl_launch_dependents();
    } else {
        // ---------------------------------------------------- epilogue (warps 2..9)
        // TMEM -> registers (thread = row, 32 columns) -> padded smem transpose -> (row, 4 columns)
        // groups: each global access of a warp covers 4 rows x 128 contiguous bytes.  Two warps share a
        // TMEM lane quarter and drain one half of the tile's columns each.
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;      // column half
        const uint32_t stg = ptx::smem_u32(smem + C::STAGING_OFF) + (warp - 2) * (32 * STG_LD * 4);
        const uint32_t tile_stg = ptx::smem_u32(smem + C::STAGING_OFF) + half * (BM * 128);  // TMA-store staging tile
        const int tw_mask = (1 << g.tw_log2) - 1;
        int as = 0;
        uint32_t aphase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int tile = item % num_tiles;
            const int mt = tile / g.num_n_tiles;
            const int n0 = (tile - mt * g.num_n_tiles) * BN;
            int cb = 0, cy0 = 0, cx0 = 0;
            if (g.conv) {
                const int per_img = g.tiles_x * g.tiles_y;
                cb = mt / per_img;
                const int rr = mt - cb * per_img;
                const int ty = rr / g.tiles_x;
                cy0 = ty * g.th;
                cx0 = (rr - ty * g.tiles_x) << g.tw_log2;
            }
            // logical output row of tile row r (and whether it exists)
            auto row_of = [&](int r, long long& grow) -> bool {
                if (g.conv) {
                    const int y = cy0 + (r >> g.tw_log2), x = cx0 + (r & tw_mask);
                    grow = (static_cast<long long>(cb) * g.H + y) * g.W + x;
                    return (y < g.H) && (x < g.W);
                }
                if (g.batch_h) {
                    const int z = mt / g.mt_per_batch, zb = z / g.batch_h;
                    const int lr = (mt - z * g.mt_per_batch) * BM + r;
                    grow = zb * g.c_row_b + (z - zb * g.batch_h) * g.c_row_h + lr;
                    return lr < g.rows_per_batch;
                }
                grow = static_cast<long long>(mt) * BM + r;
                return grow < g.M;
            };
            int col_shift = 0;   // batched problems: per-head column offset of the output
            if (g.batch_h) {
                const int z = mt / g.mt_per_batch;
                col_shift = (z - (z / g.batch_h) * g.batch_h) * g.c_col_h;
            }
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            if constexpr (KIND == EK_BIAS_BF16 || KIND == EK_GELU_BF16 || KIND == EK_RES_F32) {
                // ---- TMA-store epilogue: thread = row; each column half stages 128 rows x 128 bytes (SWIZZLE_128B)
                // and one thread hands the tile to the TMA unit (plain store, or fp32 reduce-add into the residual
                // stream: x += gamma * (acc + bias) without ever reading x into the SM).
                constexpr bool F32 = KIND == EK_RES_F32;
                constexpr int CW = F32 ? 32 : 64;              // columns per staged tile (128 bytes per row)
                constexpr int NCHUNK = BN / CW;
                const int c_lo = half * (NCHUNK / 2), c_hi = c_lo + NCHUNK / 2;
                const int r_tile = quarter * 32 + lane;
                const bool issuer = (quarter == 0) && (lane == 0);
                const uint32_t row_addr = tile_stg + r_tile * 128;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    const int col0 = n0 + c * CW;
                    // ConvTranspose scatter: column = (ky * k + kx) * Co + co, a 64-column chunk lies in one (ky, kx)
                    const int skk = g.epi.scat_k ? col0 / g.epi.scat_Co : 0;
                    const int bcol0 = col0 - skk * g.epi.scat_Co;
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
                    if (c == c_hi - 1) {  // this warp's share of the accumulator is in registers: release the TMEM stage
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&tempty[as]);
                    }
                    uint32_t w[32];  // 128 bytes of output per row
#pragma unroll
                    for (int i = 0; i < CW; i += 4) {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (col0 + i < g.N) {
                            b4 = *reinterpret_cast<const float4*>(g.epi.bias + bcol0 + i);  // warp-uniform address
                            if constexpr (F32) g4 = *reinterpret_cast<const float4*>(g.epi.gamma + col0 + i);
                        }
                        float f0 = __uint_as_float(v[i]) + b4.x, f1 = __uint_as_float(v[i + 1]) + b4.y;
                        float f2 = __uint_as_float(v[i + 2]) + b4.z, f3 = __uint_as_float(v[i + 3]) + b4.w;
                        if constexpr (KIND == EK_GELU_BF16) {
                            f0 = gelu_fast(f0); f1 = gelu_fast(f1); f2 = gelu_fast(f2); f3 = gelu_fast(f3);
                        }
                        if constexpr (F32) {
                            w[i] = __float_as_uint(f0 * g4.x); w[i + 1] = __float_as_uint(f1 * g4.y);
                            w[i + 2] = __float_as_uint(f2 * g4.z); w[i + 3] = __float_as_uint(f3 * g4.w);
                        } else {
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(f0, f1), p1 = __floats2bfloat162_rn(f2, f3);
                            w[i >> 1] = *reinterpret_cast<uint32_t*>(&p0);
                            w[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&p1);
                        }
                    }
                    if (issuer) ptx::bulk_wait_read0();      // previous tile of this half has been read by the TMA unit
                    ptx::named_bar_sync(1 + half, 128);
#pragma unroll
                    for (int j = 0; j < 8; ++j)              // 16-byte chunk j of the row lands at chunk (j ^ (row & 7))
                        ptx::sts128u(row_addr + ((j ^ (r_tile & 7)) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                    ptx::fence_proxy_async_smem();
                    ptx::named_bar_sync(1 + half, 128);
                    if (issuer) {
                        if (g.conv) {
                            if (g.epi.scat_k) {  // out[b, k*y + ky, k*x + kx, co] as a 5-D box {64 co, TW x, 1 ky, TH y, 1 b}
                                const int ky = skk / g.epi.scat_k, kx = skk - ky * g.epi.scat_k;
                                ptx::tma_store_5d(&tmC, tile_stg, kx * g.epi.scat_Co + bcol0, cx0, ky, cy0, cb);
                            } else {
                                ptx::tma_store_4d(&tmC, tile_stg, col0, cx0, cy0, cb);
                            }
                        } else if constexpr (F32) {
                            int row0 = mt * BM;
                            if (g.batch_h) {   // batched problems: rows of problem (b, h) start at b * c_row_b + h * c_row_h
                                const int z = mt / g.mt_per_batch, zb = z / g.batch_h;
                                row0 = static_cast<int>(zb * g.c_row_b + (z - zb * g.batch_h) * g.c_row_h) + (mt - z * g.mt_per_batch) * BM;
                                if (g.c_store) {   // {columns, rows of the problem, problem}: rows past the problem are clipped
                                    ptx::tma_store_4d(&tmC, tile_stg, col0, (mt - z * g.mt_per_batch) * BM, z, 0);
                                    ptx::bulk_commit();
                                    continue;
                                }
                            }
                            ptx::tma_reduce_add_2d(&tmC, tile_stg, col0, row0);
                        } else {
                            ptx::tma_store_2d(&tmC, tile_stg, col0, mt * BM);
                        }
                        ptx::bulk_commit();
                    }
                }
                as ^= 1;
                if (as == 0) aphase ^= 1;
                continue;
            }
            if (KIND == EK_GENERIC_NOGELU && g.epi.head_out != nullptr) {
                if constexpr (BN == 32 && KIND == EK_GENERIC_NOGELU) {  // fused output head: relu(dot(relu(acc + b1), w2) + b2), one row per thread
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_row, v);
                    ptx::tmem_ld_wait();
                    float sacc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float a = __uint_as_float(v[j]) + g.epi.bias[j];
                        sacc = fmaf(fmaxf(a, 0.f), g.epi.head_w[j], sacc);
                    }
                    long long grow;
                    if (half == 0 && row_of(quarter * 32 + lane, grow)) g.epi.head_out[grow] = fmaxf(sacc + __ldg(g.epi.head_b), 0.f);
                }
            } else {
                // 16-column pieces: lane -> (row = pass * 8 + lane / 4, 4 columns = (lane % 4) * 4)
                const int cg = (lane & 3) * 4;
                constexpr int NPC = BN / 16;
                const int c_lo = half * (NPC / 2), c_hi = c_lo + NPC / 2;
                // the four rows this lane finishes are the same for every piece of the tile: resolve them once
                bool ok[4];
                long long grow4[4], sbase4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ok[i] = row_of(quarter * 32 + i * 8 + (lane >> 2), grow4[i]);
                    sbase4[i] = (g.epi.scat_k && ok[i]) ? epilogue_scatter_base(g.epi, grow4[i]) : 0;
                }
                // residual rows of the NEXT tile -> L2 (they stream from HBM; without this every piece eats a DRAM latency)
                if (g.epi.res1 != nullptr && tile + static_cast<int>(gridDim.x) < num_tiles) {
                    const int nt = tile + gridDim.x;
                    const int nmt = nt / g.num_n_tiles;
                    const int nn0 = (nt - nmt * g.num_n_tiles) * BN;
                    int nb = 0, ny0 = 0, nx0 = 0;
                    if (g.conv) {
                        const int per_img = g.tiles_x * g.tiles_y;
                        nb = nmt / per_img;
                        const int rr = nmt - nb * per_img;
                        const int ty = rr / g.tiles_x;
                        ny0 = ty * g.th;
                        nx0 = (rr - ty * g.tiles_x) << g.tw_log2;
                    }
                    const int es1 = g.epi.res1_bf16 ? 2 : 4;
                    const int lines = (BN * es1) >> 7;  // 128-byte lines per tile row
                    for (int l = (warp - 2) * 32 + lane; l < BM * lines; l += 32 * NUM_EPI_WARPS) {
                        const int r = l / lines, ln = l - r * lines;
                        long long nrow;
                        bool rok;
                        if (g.conv) {
                            const int y = ny0 + (r >> g.tw_log2), x = nx0 + (r & tw_mask);
                            nrow = (static_cast<long long>(nb) * g.H + y) * g.W + x;
                            rok = (y < g.H) && (x < g.W);
                        } else {
                            nrow = static_cast<long long>(nmt) * BM + r;
                            rok = nrow < g.M;
                        }
                        if (rok && nn0 + ln * (128 / es1) < g.N) {
                            const long long boff = (nrow * g.epi.ldc + nn0) * es1 + ln * 128;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(g.epi.res1) + boff));
                            if (g.epi.res2 != nullptr)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(g.epi.res2) + boff));
                        }
                    }
                }
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    const int col = n0 + c * 16 + cg;
                    EpiCols cols;
                    if (col < g.N) epilogue_load_cols<KIND>(g.epi, col + col_shift, cols);  // in flight while TMEM is read
                    uint32_t v[16];
                    ptx::tmem_ld_32x16(t_row + c * 16, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        ptx::sts128(stg + (lane * STG_LD + 4 * j) * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    __syncwarp();
                    if (col < g.N) {
                        EpiPre pre[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)  // issue the residual / table loads of the 4 passes first
                            if (ok[i]) epilogue_prefetch<KIND>(g.epi, g.N, grow4[i], grow4[i], col + col_shift, pre[i], sbase4[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (ok[i]) {
                                const int rr = i * 8 + (lane >> 2);
                                const float4 q = ptx::lds128(stg + (rr * STG_LD + cg) * 4);
                                float f[4] = {q.x, q.y, q.z, q.w};
                                epilogue_finish<KIND>(g.epi, pre[i], cols, f);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // this warp's share of the accumulator has left TMEM: hand the stage back to the MMA warp
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    if constexpr (KIND == EK_BIAS_BF16 || KIND == EK_GELU_BF16 || KIND == EK_RES_F32) {
        if (warp >= 2 && (warp & 3) == 0 && lane == 0) ptx::bulk_wait0();  // the issuers drain their TMA stores
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
template <int BN, int KIND, bool HALO = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcArgs& a, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, KIND, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg<BN, HALO>::SMEM_BYTES));
        configured = true;
    }
    const int tiles = a.num_m_tiles * a.num_n_tiles * a.ksplit;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    ProfScope prof(PROF_GEMM_TC, a.flops, stream);
    DAD_CHECK_CUDA(launch_pdl(gemm_tc_kernel<BN, KIND, HALO>, dim3(grid), dim3(NUM_THREADS), Cfg<BN, HALO>::SMEM_BYTES, stream,
                              tmA, tmB, tmC, a));
    return DAD_OK;
}

template <int BN>
int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcArgs& a,
                cudaStream_t stream) {
    return launch<BN, EK_GENERIC_NOGELU, true>(tmA, tmB, tmC, a, stream);
}

template <int BN>
int launch_kind(int kind, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcArgs& a,
                cudaStream_t stream) {
    if constexpr (BN >= 128) {  // the specialised encoder epilogues only occur with wide N
        switch (kind) {
            case EK_BIAS_BF16: return launch<BN, EK_BIAS_BF16>(tmA, tmB, tmC, a, stream);
            case EK_GELU_BF16: return launch<BN, EK_GELU_BF16>(tmA, tmB, tmC, a, stream);
            case EK_RES_F32: return launch<BN, EK_RES_F32>(tmA, tmB, tmC, a, stream);
            default: break;
        }
    }
    return launch<BN, EK_GENERIC_NOGELU>(tmA, tmB, tmC, a, stream);
}

int pick_bn(int N) {
    const int cands[4] = {256, 128, 64, 32};
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        if (N % bn == 0 || N >= 8 * bn) return bn;
    }
    return 32;
}

}  // namespace

int gemm_tc(const GemmProblem& p, cudaStream_t stream) {
    DAD_REQUIRE(p.A && p.Wt && p.N > 0, "gemm_tc: null operand or N<=0");
    if (p.ksplit > 1 || p.batch_h > 0) debug_label(p.batch_h > 0 ? "gemm_tc batched" : "gemm_tc split-K");
    {
        static const bool no_2cta = getenv("DAD_NO_2CTA") != nullptr;  // A/B switch while the 2-CTA kernel is validated
        if (!no_2cta && p.ksplit <= 1 && p.batch_h == 0 && gemm_tc2_eligible(p)) return gemm_tc2(p, stream);
        if (!no_2cta && conv_tc2_eligible(p)) return conv_tc2(p, stream);
    }
    DAD_REQUIRE(p.N % 8 == 0, "gemm_tc: N=%d must be a multiple of 8", p.N);
    DAD_REQUIRE(p.Kp % 8 == 0, "gemm_tc: Kp=%d must be a multiple of 8", p.Kp);
    TcArgs a{};
    a.epi = p.epi;
    a.N = p.N;
    a.conv = p.conv;
    a.batch_h = 0;
    a.ksplit = 1;
    int bn = pick_bn(p.N);
    if (p.ksplit > 1 && bn < 128) bn = 128;   // split-K needs the TMA reduce-add epilogue (wide tiles only)
    if (p.shift_taps) {
        DAD_REQUIRE(!p.conv && p.batch_h == 0 && p.ksplit > 1 && p.shift_taps <= 9 && p.shift_rows > 0 && p.shift_ld % 128 == 0 &&
                        p.N == p.shift_taps * p.shift_ld,
                    "gemm_tc: bad shifted-view problem");
        bn = 128;   // a tile must not straddle two taps
        a.shift_ld = p.shift_ld;
        for (int t = 0; t < 9; ++t) {
            DAD_REQUIRE(p.shift_off[t] % 8 == 0 && p.shift_row[t] >= 0, "gemm_tc: shifted views need 16-byte aligned offsets");
            a.shift_off[t] = p.shift_off[t];
            a.shift_row[t] = p.shift_row[t];
        }
    }
    if (p.epi.head_out) {
        DAD_REQUIRE(p.N == 32 && p.epi.bias && p.epi.head_w, "gemm_tc: fused head needs N == 32, bias and head_w");
        bn = 32;
    }
    if (p.epi.scat_k) DAD_REQUIRE(p.epi.scat_CoP % 32 == 0, "gemm_tc: scatter needs CoP %% 32 == 0");
    a.num_n_tiles = cdiv(p.N, bn);

    CUtensorMap tmA, tmB;
    bool halo = false;
    if (p.mn) {
        // MN-major operands (weight gradients straight from the activation layouts): 64 x 64 boxes, contraction rows outer
        DAD_REQUIRE(!p.conv && (p.batch_h > 0 ? p.ksplit <= 1 && p.mn != 2 : p.ksplit > 1) && p.M > 0 && p.lda % 8 == 0 &&
                        p.ldw % 8 == 0 && p.ldw > 0,
                    "gemm_tc: bad MN-major problem");
        a.mn = p.mn;
        a.M = p.M;
        a.num_m_tiles = cdiv(p.M, BM);
        const cuuint32_t box2[2] = {64u, 64u};
        const cuuint32_t box4[4] = {64u, 8u, 8u, 1u};
        if (p.mn == 1 || p.mn == 3) {
            DAD_REQUIRE(p.K > 0 && !p.shift_taps, "gemm_tc: bad MN-major linear problem");
            a.num_k_blocks = cdiv(p.K, BK);
            a.flops = 2.0 * p.M * p.N * static_cast<double>(p.K);
            const cuuint32_t boxk2[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
            if (p.batch_h > 0) {
                DAD_REQUIRE(p.batch_b > 0 && p.w_rows > 0 && p.w_rows <= p.N && p.a_sh % 8 == 0 && p.a_sb % 8 == 0 && p.w_sh % 8 == 0 &&
                                p.w_sb % 8 == 0 && !p.epi.bias && !p.epi.res1 && !p.epi.scat_k && !p.epi.rowtab && !p.epi.res2 &&
                                !p.epi.head_out,
                            "gemm_tc: bad batched MN-major problem");
                const int nz = p.batch_h * p.batch_b;
                a.batch_h = p.batch_h;
                a.mt_per_batch = cdiv(p.M, BM);
                a.rows_per_batch = p.M;
                a.num_m_tiles = a.mt_per_batch * nz;
                a.c_row_b = p.c_row_b; a.c_row_h = p.c_row_h; a.c_col_h = p.c_col_h;
                a.flops *= nz;
                const cuuint32_t box44[4] = {64u, 64u, 1u, 1u}, boxk4[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1u, 1u};
                const cuuint64_t sA3[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.a_sh * 2, (cuuint64_t)p.a_sb * 2};
                const cuuint64_t sB3[3] = {(cuuint64_t)p.ldw * 2, (cuuint64_t)p.w_sh * 2, (cuuint64_t)p.w_sb * 2};
                const cuuint64_t dAm[4] = {(cuuint64_t)p.M, (cuuint64_t)p.K, (cuuint64_t)p.batch_h, (cuuint64_t)p.batch_b};
                const cuuint64_t dAk[4] = {(cuuint64_t)p.K, (cuuint64_t)p.M, (cuuint64_t)p.batch_h, (cuuint64_t)p.batch_b};
                const cuuint64_t dB[4] = {(cuuint64_t)p.w_rows, (cuuint64_t)p.K, (cuuint64_t)p.batch_h, (cuuint64_t)p.batch_b};
                if (p.mn == 1) DAD_TRY(make_tmap(&tmA, 0, p.A, 4, dAm, sA3, box44));
                else DAD_TRY(make_tmap(&tmA, 0, p.A, 4, dAk, sA3, boxk4));
                DAD_TRY(make_tmap(&tmB, 0, p.Wt, 4, dB, sB3, box44));
            } else {
                DAD_REQUIRE(p.ldw >= p.N && p.lda >= (p.mn == 1 ? p.M : p.K), "gemm_tc: bad MN-major row pitch");
                const cuuint64_t dAm[2] = {(cuuint64_t)p.M, (cuuint64_t)p.K}, dAk[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
                const cuuint64_t sA1[1] = {(cuuint64_t)p.lda * 2};
                const cuuint64_t dB[2] = {(cuuint64_t)p.N, (cuuint64_t)p.K}, sB1[1] = {(cuuint64_t)p.ldw * 2};
                if (p.mn == 1) DAD_TRY(make_tmap_bf16(&tmA, p.A, 2, dAm, sA1, box2));
                else DAD_TRY(make_tmap_bf16(&tmA, p.A, 2, dAk, sA1, boxk2));
                DAD_TRY(make_tmap_bf16(&tmB, p.Wt, 2, dB, sB1, box2));
            }
        } else {
            DAD_REQUIRE(p.mn == 2 && p.B > 0 && p.H > 0 && p.W > 0 && p.shift_rows > 0 && p.shift_ld % 128 == 0 &&
                            p.shift_rows <= p.shift_ld && p.N == 9 * p.shift_ld && p.lda >= p.M && p.ldw >= p.shift_rows,
                        "gemm_tc: bad convolution weight-gradient problem");
            bn = 128;   // a tile must not straddle two taps
            a.num_n_tiles = cdiv(p.N, bn);
            a.shift_ld = p.shift_ld;
            a.tiles_x = cdiv(p.W, 8);
            a.tiles_y = cdiv(p.H, 8);
            a.num_k_blocks = p.B * a.tiles_x * a.tiles_y;
            a.flops = 2.0 * p.M * 9.0 * p.shift_rows * (static_cast<double>(p.B) * p.H * p.W);
            const cuuint64_t dA[4] = {(cuuint64_t)p.M, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
            const cuuint64_t sA3[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.lda * 2 * p.W, (cuuint64_t)p.lda * 2 * p.W * p.H};
            const cuuint64_t dB[4] = {(cuuint64_t)p.shift_rows, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
            const cuuint64_t sB3[3] = {(cuuint64_t)p.ldw * 2, (cuuint64_t)p.ldw * 2 * p.W, (cuuint64_t)p.ldw * 2 * p.W * p.H};
            DAD_TRY(make_tmap(&tmA, 0, p.A, 4, dA, sA3, box4));
            DAD_TRY(make_tmap(&tmB, 0, p.Wt, 4, dB, sB3, box4));
        }
    } else if (p.conv) {
        DAD_REQUIRE(p.taps == 1 || p.taps == 9, "gemm_tc: taps must be 1 or 9");
        DAD_REQUIRE(p.C % 8 == 0 && p.ldp % 8 == 0, "gemm_tc: conv C/ldp must be multiples of 8");
        DAD_REQUIRE(p.stride == 1 || (p.stride == 2 && p.taps == 9), "gemm_tc: stride %d unsupported", p.stride);
        a.taps = p.taps;
        a.stride = p.stride;
        a.pad = p.taps == 9 ? 1 : 0;
        // output extent ("same" padding): H x W at stride 1, ((H - 1) / 2 + 1) x ((W - 1) / 2 + 1) at stride 2
        const int Ho = (p.H - 1) / p.stride + 1, Wo = (p.W - 1) / p.stride + 1;
        a.cchunks = cdiv(p.C, BK);
        a.num_k_blocks = a.taps * a.cchunks;
        DAD_REQUIRE(p.Kp == a.num_k_blocks * BK, "gemm_tc: conv weights must be packed to Kp=%d (got %d)",
                    a.num_k_blocks * BK, p.Kp);
        a.B = p.B; a.H = Ho; a.W = Wo;   // the kernel tiles and stores OUTPUT pixels
        a.M = p.B * Ho * Wo;
        a.flops = 2.0 * a.M * p.N * (static_cast<double>(p.taps) * p.C);
        DAD_REQUIRE(p.stride == 1 || !p.epi.scat_k, "gemm_tc: strided conv with a scatter epilogue is unsupported");
        static const bool halo_off = getenv("DAD_NO_HALO") != nullptr;      // bring-up A/B switches
        // (measured on B200: UMMA applies the 128B swizzle to absolute shared-memory address bits, exactly like TMA,
        // so the row-shifted halo views need base_offset = 0 in their descriptors)
        halo = !halo_off && p.stride == 1 && p.taps == 9 && bn <= 64 && cdiv(p.N, bn) == 1 &&
               static_cast<long long>(9) * cdiv(p.C, BK) * bn * BK * 2 <= WRES_BYTES;  // weights fit in smem
        if (halo) {  // 16 x 8 output pixels per tile, one 18 x 16 halo box per channel chunk
            a.tw_log2 = 3;
            a.th = 16;
        } else {
            // choose the spatial tile (TH x TW = 128) with the least padded area
            long long best = -1;
            for (int l2 = 3; l2 <= 7; ++l2) {
                const int tw = 1 << l2, th = BM / tw;
                const long long area = static_cast<long long>(cdiv(Wo, tw)) * tw * cdiv(Ho, th) * th;
                if (best < 0 || area < best) { best = area; a.tw_log2 = l2; a.th = th; }
            }
        }
        const int tw = 1 << a.tw_log2;
        a.tiles_x = cdiv(Wo, tw);
        a.tiles_y = cdiv(Ho, a.th);
        a.num_m_tiles = p.B * a.tiles_x * a.tiles_y;
        const cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
        const cuuint64_t strides[3] = {(cuuint64_t)p.ldp * 2, (cuuint64_t)p.ldp * 2 * p.W,
                                       (cuuint64_t)p.ldp * 2 * p.W * p.H};
        // stride 2: the box walks 2 * t - 1 input pixels and picks every second one (t elements land in shared memory)
        const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(halo ? HALO_W : (tw - 1) * p.stride + 1),
                                   (cuuint32_t)(halo ? HALO_H : (a.th - 1) * p.stride + 1), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
        DAD_TRY(make_tmap(&tmA, 0, p.A, 4, dims, strides, box, estr));
    } else {
        DAD_REQUIRE(p.M > 0 && p.K > 0 && p.lda >= p.K && p.lda % 8 == 0, "gemm_tc: bad linear dims M=%d K=%d lda=%lld",
                    p.M, p.K, p.lda);
        a.M = p.M;
        a.flops = 2.0 * p.M * p.N * static_cast<double>(p.K);
        a.num_k_blocks = cdiv(p.K, BK);
        a.num_m_tiles = cdiv(p.M, BM);
        if (p.batch_h > 0) {
            DAD_REQUIRE(p.batch_b > 0 && p.ksplit <= 1 && p.w_rows > 0 && p.w_rows <= p.N && p.a_sh % 8 == 0 && p.a_sb % 8 == 0 &&
                            p.w_sh % 8 == 0 && p.w_sb % 8 == 0 && !p.epi.scat_k && !p.epi.rowtab && !p.epi.res2 && !p.epi.head_out,
                        "gemm_tc: bad batched problem");
            // with a residual the only supported form is out += gamma * (acc + bias) through the TMA reduce-add epilogue: rows
            // past a problem's M are zero (A is zero-filled there), so tiles may overhang into the next problem's rows
            DAD_REQUIRE((!p.epi.bias && !p.epi.res1) || (epilogue_kind(p.epi) == EK_RES_F32 && p.c_col_h == 0 && pick_bn(p.N) >= 128),
                        "gemm_tc: batched problems take either no bias / residual or the reduce-add epilogue with N %% 128 == 0");
            const int nz = p.batch_h * p.batch_b;
            a.batch_h = p.batch_h;
            a.mt_per_batch = cdiv(p.M, BM);
            a.rows_per_batch = p.M;
            a.num_m_tiles = a.mt_per_batch * nz;
            a.c_row_b = p.c_row_b; a.c_row_h = p.c_row_h; a.c_col_h = p.c_col_h;
            a.flops *= nz;
            const cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.M, (cuuint64_t)p.batch_h, (cuuint64_t)p.batch_b};
            const cuuint64_t strides[3] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.a_sh * 2, (cuuint64_t)p.a_sb * 2};
            const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1, 1};
            DAD_TRY(make_tmap(&tmA, 0, p.A, 4, dims, strides, box));
        } else {
        const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
        const cuuint64_t strides[1] = {(cuuint64_t)p.lda * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
        DAD_TRY(make_tmap_bf16(&tmA, p.A, 2, dims, strides, box));
        }
    }
    if (p.mn) {
        // (both maps were built above)
    } else if (a.batch_h) {
        const cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.w_rows, (cuuint64_t)p.batch_h, (cuuint64_t)p.batch_b};
        const cuuint64_t strides[3] = {(cuuint64_t)(p.ldw ? p.ldw : p.Kp) * 2, (cuuint64_t)p.w_sh * 2, (cuuint64_t)p.w_sb * 2};
        const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bn, 1, 1};
        DAD_TRY(make_tmap(&tmB, 0, p.Wt, 4, dims, strides, box));
    } else {
        const cuuint64_t dims[2] = {(cuuint64_t)p.Kp, (cuuint64_t)(p.shift_taps ? p.shift_rows : p.N)};
        const cuuint64_t strides[1] = {(cuuint64_t)p.Kp * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
        DAD_TRY(make_tmap_bf16(&tmB, p.Wt, 2, dims, strides, box));
    }
    int kind = epilogue_kind(p.epi);
    // ConvTranspose (k = s) as a 1x1 "conv" whose bf16 output tile is scattered by ONE 5-D TMA store per 64-channel
    // chunk instead of 8-byte scattered stores: needs whole 64-channel chunks per (ky, kx) and no padded channels
    const Epilogue& e = p.epi;
    const bool scat_tma = e.scat_k && p.conv && p.taps == 1 && e.scat_CoP == e.scat_Co && e.scat_Co % 64 == 0 &&
                          e.scat_H == p.H && e.scat_W == p.W && e.bias && e.out && e.out_bf16 && !e.gamma && !e.res1 &&
                          !e.res2 && !e.out_relu && !e.rowtab && !e.head_out && e.act == ACT_NONE && e.ldc == e.scat_Co &&
                          bn >= 128;
    if (scat_tma) kind = EK_BIAS_BF16;
    DAD_REQUIRE(!(e.scat_k && p.conv && !scat_tma), "gemm_tc: conv-mode ConvTranspose scatter needs Co %% 64 == 0, bf16 out");
    if (bn < 128 || (a.batch_h && kind != EK_RES_F32)) kind = EK_GENERIC;  // the specialised (TMA-store) epilogues exist for the wide tiles only
    if (p.ksplit > 1) {
        // split-K: each (k slice, tile) work item reduce-adds its partial product into the fp32 output through TMA
        DAD_REQUIRE(!p.conv && kind == EK_RES_F32, "gemm_tc: split-K needs a linear problem with the out += gamma * (acc + bias) epilogue");
        a.kb_per_split = cdiv(a.num_k_blocks, p.ksplit);
        a.ksplit = cdiv(a.num_k_blocks, a.kb_per_split);
    } else {
        a.kb_per_split = a.num_k_blocks;
    }
    DAD_REQUIRE(!(kind == EK_GENERIC && p.epi.act == ACT_GELU), "gemm_tc: GELU is only fused as bias+GELU->bf16");
    CUtensorMap tmC = tmA;  // placeholder for the generic epilogue (never dereferenced)
    if (kind != EK_GENERIC) {
        const bool f32 = kind == EK_RES_F32;
        const int es = f32 ? 4 : 2, cw = f32 ? 32 : 64;
        DAD_REQUIRE((p.epi.ldc * es) % 16 == 0, "gemm_tc: output row pitch must be a multiple of 16 bytes");
        if (scat_tma) {
            const int tw = 1 << a.tw_log2, k = e.scat_k, Co = e.scat_Co;
            const cuuint64_t dims[5] = {(cuuint64_t)k * Co, (cuuint64_t)p.W, (cuuint64_t)k, (cuuint64_t)p.H, (cuuint64_t)p.B};
            const cuuint64_t sx = (cuuint64_t)k * Co * 2, sky = sx * p.W, sy = sky * k, sb = sy * p.H;  // stride 1 only
            const cuuint64_t strides[4] = {sx, sky, sy, sb};
            const cuuint32_t box[5] = {64u, (cuuint32_t)tw, 1u, (cuuint32_t)a.th, 1u};
            DAD_TRY(make_tmap(&tmC, 0, p.epi.out, 5, dims, strides, box));
        } else if (p.conv) {
            const int tw = 1 << a.tw_log2;
            const cuuint64_t dims[4] = {(cuuint64_t)p.N, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)p.B};
            const cuuint64_t strides[3] = {(cuuint64_t)p.epi.ldc * es, (cuuint64_t)p.epi.ldc * es * a.W,
                                           (cuuint64_t)p.epi.ldc * es * a.W * a.H};
            const cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)tw, (cuuint32_t)a.th, 1};
            DAD_TRY(make_tmap(&tmC, f32 ? 1 : 0, p.epi.out, 4, dims, strides, box));
        } else if (p.c_store) {
            DAD_REQUIRE(a.batch_h && f32 && p.ksplit <= 1 && p.c_row_h == p.M && p.c_row_b == static_cast<long long>(p.batch_h) * p.M &&
                            p.c_col_h == 0,
                        "gemm_tc: c_store needs contiguous batched problems and the fp32 epilogue");
            a.c_store = 1;
            const cuuint64_t dims[4] = {(cuuint64_t)p.N, (cuuint64_t)p.M, (cuuint64_t)(p.batch_h * p.batch_b), 1};
            const cuuint64_t strides[3] = {(cuuint64_t)p.epi.ldc * 4, (cuuint64_t)p.epi.ldc * 4 * p.M, (cuuint64_t)p.epi.ldc * 4 * p.M * p.batch_h * p.batch_b};
            const cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)BM, 1, 1};
            DAD_TRY(make_tmap(&tmC, 1, p.epi.out, 4, dims, strides, box));
        } else {
            const cuuint64_t out_rows = a.batch_h ? (cuuint64_t)(p.batch_b * p.c_row_b) : (cuuint64_t)p.M;
            const cuuint64_t dims[2] = {(cuuint64_t)p.N, out_rows};
            const cuuint64_t strides[1] = {(cuuint64_t)p.epi.ldc * es};
            const cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)BM};
            DAD_TRY(make_tmap(&tmC, f32 ? 1 : 0, p.epi.out, 2, dims, strides, box));
        }
    }
    if (halo) return bn == 64 ? launch_halo<64>(tmA, tmB, tmC, a, stream) : launch_halo<32>(tmA, tmB, tmC, a, stream);
    switch (bn) {
        case 256: return launch_kind<256>(kind, tmA, tmB, tmC, a, stream);
        case 128: return launch_kind<128>(kind, tmA, tmB, tmC, a, stream);
        case 64: return launch_kind<64>(kind, tmA, tmB, tmC, a, stream);
        default: return launch_kind<32>(kind, tmA, tmB, tmC, a, stream);
    }
}

}  // namespace dad
