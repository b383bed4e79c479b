// Distillation-loss reductions (fp32, exact selection), all on the caller's stream with no host
// sync and no per-image host loop.  Reference: tools/train_distillation.py:173-707.
//
// Exact lower medians (nanmedian / torch.median semantics) use an MSB-first 4 x 8-bit radix select
// on the order-preserving uint32 image of fp32: each pass histograms one digit of every member
// element whose higher digits match the prefix found so far (shared-memory histograms with
// warp-aggregated atomics, flushed to a per-row global histogram), then a one-warp-per-row scan
// picks the bin holding the wanted rank.  A "row" is (array in {pred, gt}) x image x context; the
// HDN-DR path computes each pixel's context membership on the fly from the per-image min/max
// (no [7,B,1,H,W] mask tensor, no 7x replicated maps).
#include <cfloat>
#include <cstdlib>

#include "common.h"
#include "losses.h"

namespace dad {

namespace {

constexpr int MODE_MASK = 0;  // K = 1: optional u8 mask [B, L]
constexpr int MODE_DR = 1;    // K = 2^level - 1 depth-range contexts from gt + optional mask
constexpr int MODE_CTX = 2;   // K explicit u8 contexts [K, B, L]
constexpr int MODE_RANK = 3;  // K order statistics of ONE array over the optional mask (quantile bins, HDN-DP)
constexpr int MAX_K = 21;
constexpr int THREADS = 256;
constexpr int UNROLL = 4;     // pixels per thread whose loads are issued together in the streaming passes

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(u);
}

struct SelArgs {
    const float* pred;     // [B, L]
    const float* gt;       // [B, L]
    const uint8_t* mask;   // [B, L] or null
    const uint8_t* ctx;    // [K, B, L] (MODE_CTX)
    int B, K, level, narr; // narr = 2 (pred, gt) or 1 (pred only)
    int nq;                // MODE_RANK: quantile points q_j = j / (nq - 1); row 2j = floor rank, 2j + 1 = ceil rank
    long long L;
    int chunk;             // pixels per CTA
    // workspace
    uint32_t* minmax;      // [B][2] keys over valid gt (MODE_DR)
    uint32_t* hist;        // [4][R][256]
    uint32_t* prefix;      // [R]
    uint32_t* krank;       // [R]
    uint32_t* count;       // [R]
    float* t;              // [R] medians
    double* madsum;        // [R]
    float* s;              // [R] scales
    double* acc;           // [0] numerator, [1] (as double) denominator count
    int R;
    // range-bin selection (run_select_lin): value-linear histogram -> candidates of the median's bin -> exact select
    uint32_t* imm;         // [B][4] keys: pred min / max, gt min / max over the valid pixels of the image
    uint32_t* lhist;       // [R][NB]
    uint32_t* tbin;        // [R] bin that holds the wanted rank (0xFFFFFFFF: empty row)
    uint32_t* trank;       // [R] rank inside that bin
    uint32_t* cbin;        // [R] population of that bin
    uint32_t* ccount;      // [R] append cursor of the candidate list
    uint32_t* cand;        // [R][CAP] keys of the candidates
    int NB, CAP;
    // backward (gradient w.r.t. pred; gt is the detached teacher map)
    int mean_all;          // global_normalize: s = sum |x - t| / L instead of / (n + 1)
    double* racc;          // [B*K][3]  E = sum w*sgn, G = sum w*sgn*(p - t), S = sum sign(p - t) over the members of the row
    unsigned int* jstar;   // [B*K]     lowest member index holding the median value (where d median / d p lives)
};

__device__ __forceinline__ int row_of(const SelArgs& a, int arr, int b, int k) { return (arr * a.B + b) * a.K + k; }

// Depth-range thresholds, reference op order (tools/train_distillation.py:562-569):
//   lo = min + ((max - min) * i) * bin ;  hi = (min + ((max - min) * (i + 1)) * bin) + 1e-30
__device__ __forceinline__ void dr_thresholds(int level, int k, float mn, float mx, float& lo, float& hi) {
    int lvl = 0, first = 0;  // contexts ordered finest level first: 2^(level-1) bins, ..., 1 bin
    int nb = 1 << (level - 1);
    while (k >= first + nb) { first += nb; nb >>= 1; ++lvl; }
    const int i = k - first;
    const float bin = 1.0f / static_cast<float>(nb);
    const float range = __fsub_rn(mx, mn);
    lo = __fadd_rn(mn, __fmul_rn(__fmul_rn(range, static_cast<float>(i)), bin));
    hi = __fadd_rn(__fadd_rn(mn, __fmul_rn(__fmul_rn(range, static_cast<float>(i + 1)), bin)), 1e-30f);
    (void)lvl;
}

template <int MODE>
__device__ __forceinline__ uint32_t member_bits(const SelArgs& a, int b, long long i, float g, const float* lo,
                                                const float* hi, bool has_valid) {
    if (MODE == MODE_MASK) {
        return a.mask ? (a.mask[static_cast<long long>(b) * a.L + i] != 0 ? 1u : 0u) : 1u;
    } else if (MODE == MODE_RANK) {
        const bool valid = a.mask ? a.mask[static_cast<long long>(b) * a.L + i] != 0 : true;
        return valid ? ((1u << a.K) - 1u) : 0u;
    } else if (MODE == MODE_DR) {
        const bool valid = a.mask ? a.mask[static_cast<long long>(b) * a.L + i] != 0 : true;
        if (!valid || !has_valid) return 0u;
        uint32_t bits = 0;
        for (int k = 0; k < a.K; ++k) bits |= (g >= lo[k] && g < hi[k]) ? (1u << k) : 0u;
        return bits;
    } else {
        uint32_t bits = 0;
        for (int k = 0; k < a.K; ++k)
            bits |= a.ctx[(static_cast<long long>(k) * a.B + b) * a.L + i] != 0 ? (1u << k) : 0u;
        return bits;
    }
}

template <int MODE>
__device__ __forceinline__ bool setup_thresholds(const SelArgs& a, int b, float* lo, float* hi) {
    bool has_valid = true;
    if (MODE == MODE_DR) {
        const uint32_t kmn = a.minmax[2 * b], kmx = a.minmax[2 * b + 1];
        has_valid = kmn <= kmx;
        if (threadIdx.x < a.K && has_valid)
            dr_thresholds(a.level, threadIdx.x, key2f(kmn), key2f(kmx), lo[threadIdx.x], hi[threadIdx.x]);
        __syncthreads();
    }
    return has_valid;
}

// ---------------------------------------------------------------- min / max over valid gt
__global__ void __launch_bounds__(THREADS) minmax_kernel(const float* x, const uint8_t* mask, long long L, int chunk,
                                                         uint32_t* minmax) {
    const int b = blockIdx.y;
    const long long start = static_cast<long long>(blockIdx.x) * chunk;
    const long long end = min(start + chunk, L);
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        if (mask && mask[b * L + i] == 0) continue;
        const uint32_t k = f2key(x[b * L + i]);
        mn = min(mn, k);
        mx = max(mx, k);
    }
    for (int o = 16; o; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mn != 0xFFFFFFFFu) atomicMin(&minmax[2 * b], mn);
        if (mx != 0u || mn != 0xFFFFFFFFu) atomicMax(&minmax[2 * b + 1], mx);
    }
}

__global__ void init_minmax_kernel(uint32_t* minmax, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { minmax[2 * i] = 0xFFFFFFFFu; minmax[2 * i + 1] = 0u; }
}

// ---------------------------------------------------------------- radix-select histogram pass
__device__ __forceinline__ void agg_inc(uint32_t* h, bool ok, uint32_t digit, int lane) {
    if (__ballot_sync(0xffffffffu, ok) == 0u) return;
    const uint32_t m = __match_any_sync(0xffffffffu, ok ? digit : 0xFFFFFFFFu);
    if (ok && (__ffs(m) - 1) == lane) atomicAdd(&h[digit], __popc(m));
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) sel_hist_kernel(const SelArgs a, int pass) {
    extern __shared__ uint32_t sm[];
    const int nrow = a.narr * a.K;
    uint32_t* h = sm;                                        // [narr*K][256]
    uint32_t* pref = sm + nrow * 256;                        // [narr*K]
    float* lo = reinterpret_cast<float*>(pref + nrow);       // [K]
    float* hi = lo + a.K;                                    // [K]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nrow * 256; i += THREADS) h[i] = 0;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        pref[threadIdx.x] = a.prefix[row_of(a, arr, b, k)];
    }
    __syncthreads();
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    const int shift = 24 - 8 * pass;
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long base = start + (threadIdx.x & ~31); base < end; base += THREADS) {
        const long long i = base + lane;
        const bool inb = i < end;
        float p = 0.f, g = 0.f;
        if (inb) {
            p = a.pred[b * a.L + i];
            if (a.narr == 2 || MODE == MODE_DR) g = a.gt[b * a.L + i];
        }
        const uint32_t bits = inb ? member_bits<MODE>(a, b, i, g, lo, hi, has_valid) : 0u;
        if (__ballot_sync(0xffffffffu, bits != 0u) == 0u) continue;
        const uint32_t kp = f2key(p), kg = f2key(g);
        if (MODE == MODE_DR) {
            // Depth-range contexts of one level are disjoint unless the image is degenerate (range 0 at depth 0,
            // SURVEY A.4 iii): a pixel then has at most ONE context per level, and (context, digit) can share one
            // warp-aggregated atomic per level and array - 2 * level aggregations per pixel instead of 2 * K.
            bool fast = true;
            int lvl_first = 0;
            for (int nb = 1 << (a.level - 1); nb >= 1; lvl_first += nb, nb >>= 1)
                fast = fast && __popc(bits & (((1u << nb) - 1u) << lvl_first)) <= 1;
            if (__all_sync(0xffffffffu, fast)) {
                lvl_first = 0;
                for (int nb = 1 << (a.level - 1); nb >= 1; lvl_first += nb, nb >>= 1) {
                    const uint32_t sel = bits & (((1u << nb) - 1u) << lvl_first);
                    const int k = sel ? __ffs(sel) - 1 : 0;
                    {
                        const bool ok = sel != 0u && (pass == 0 || (kp >> (shift + 8)) == (pref[k] >> (shift + 8)));
                        agg_inc(h, ok, (static_cast<uint32_t>(k) << 8) | ((kp >> shift) & 255u), lane);
                    }
                    {
                        const bool ok = sel != 0u && (pass == 0 || (kg >> (shift + 8)) == (pref[a.K + k] >> (shift + 8)));
                        agg_inc(h + a.K * 256, ok, (static_cast<uint32_t>(k) << 8) | ((kg >> shift) & 255u), lane);
                    }
                }
                continue;
            }
        }
        for (int k = 0; k < a.K; ++k) {
            const bool in = (bits >> k) & 1u;
            {
                const bool ok = in && (pass == 0 || (kp >> (shift + 8)) == (pref[k] >> (shift + 8)));
                agg_inc(h + k * 256, ok, (kp >> shift) & 255u, lane);
            }
            if (a.narr == 2) {
                const bool ok = in && (pass == 0 || (kg >> (shift + 8)) == (pref[a.K + k] >> (shift + 8)));
                agg_inc(h + (a.K + k) * 256, ok, (kg >> shift) & 255u, lane);
            }
        }
    }
    __syncthreads();
    uint32_t* gh = a.hist + static_cast<long long>(pass) * a.R * 256;
    for (int i = threadIdx.x; i < nrow * 256; i += THREADS) {
        const uint32_t v = h[i];
        if (v) {
            const int r = i >> 8, arr = r / a.K, k = r - arr * a.K;
            atomicAdd(&gh[static_cast<long long>(row_of(a, arr, b, k)) * 256 + (i & 255)], v);
        }
    }
}

// rank of quantile point j among n non-NaN values as ATen computes it (aten/native/Sorting.cpp quantile_compute):
// fp32 q = j / (nq - 1) times fp32(n - 1), one rounding.
__device__ __forceinline__ float quantile_rank(int j, int nq, uint32_t n) {
    const float q = static_cast<float>(j) / static_cast<float>(nq - 1);  // powers of two: exact
    return __fmul_rn(q, static_cast<float>(n - 1u));
}

// one warp per row: locate the bin that holds rank k, extend the prefix
__global__ void __launch_bounds__(THREADS) sel_scan_kernel(const SelArgs a, int pass) {
    const int r = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= a.R) return;
    const uint32_t* h = a.hist + (static_cast<long long>(pass) * a.R + r) * 256;
    uint32_t c[8], local = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = h[lane * 8 + j]; local += c[j]; }
    uint32_t incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t k;
    if (pass == 0) {
        k = total ? (total - 1) / 2 : 0;  // lower median rank
        if (a.nq > 0 && total) {          // quantile order statistic (torch.nanquantile, 'linear' interpolation)
            const int kk = r % a.K;
            const float rk = quantile_rank(kk >> 1, a.nq, total);
            k = static_cast<uint32_t>((kk & 1) ? ceilf(rk) : floorf(rk));
        }
        if (lane == 0) a.count[r] = total;
    } else {
        k = a.krank[r];
    }
    const uint32_t excl = incl - local;
    const bool mine = total != 0 && k >= excl && k < incl;
    if (mine) {
        uint32_t cum = excl;
        int bin = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k >= cum + c[j]) { cum += c[j]; bin = j + 1; }
            else break;
        }
        const int shift = 24 - 8 * pass;
        const uint32_t np = (pass == 0 ? 0u : a.prefix[r]) | (static_cast<uint32_t>(lane * 8 + bin) << shift);
        a.prefix[r] = np;
        a.krank[r] = k - cum;
        if (pass == 3) a.t[r] = key2f(np);
    }
    if (total == 0 && lane == 0) {
        a.prefix[r] = 0;
        a.krank[r] = 0;
        if (pass == 3) a.t[r] = 0.f;  // nanmedian of an all-NaN row -> NaN -> 0 (reference :490)
    }
}

// ================================================================ range-bin selection
// The 4 x 8-bit radix passes above spend their time in warp-aggregated atomics (every member element takes part in
// every one of the first passes).  For the medians of the loss path a cheaper exact scheme is used instead:
//   1. per image: min / max of pred and gt over the valid pixels                                   (imgminmax_kernel)
//   2. ONE histogram pass over NB value-linear bins of each row's range - bin(x) = int((x - lo) * NB / (hi - lo)) is
//      monotone in x, so bins are value-ordered; occupancy is spread, plain shared-memory atomics do   (lin_hist_kernel)
//   3. per row: the bin that holds the median rank, and the rank inside it                           (lin_scan_kernel)
//   4. ONE pass that appends the keys of that bin's members to a small per-row list (~ n / NB of them) (compact_kernel)
//   5. per row: exact radix select inside the list, in shared memory.  A list that overflows CAP (a degenerate
//      distribution: constant images, heavy ties) is not used: the row's CTA streams the image instead - slow, exact
//                                                                                                     (select_rows_kernel)
__device__ __forceinline__ int lin_bin(float x, float lo, float scale, int NB) {
    const int bq = static_cast<int>(__fmul_rn(__fsub_rn(x, lo), scale));
    return min(max(bq, 0), NB - 1);
}

__global__ void __launch_bounds__(THREADS) imgminmax_kernel(const SelArgs a) {
    const int b = blockIdx.y;
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    uint32_t pmn = 0xFFFFFFFFu, pmx = 0u, gmn = 0xFFFFFFFFu, gmx = 0u;
    for (long long base = start + threadIdx.x; base < end; base += UNROLL * THREADS) {
        float p[UNROLL], g[UNROLL];
        bool in[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {  // all loads of the group are issued before any use
            const long long i = base + u * THREADS;
            in[u] = i < end && !(a.mask && a.mask[b * a.L + i] == 0);
            p[u] = in[u] ? a.pred[b * a.L + i] : 0.f;
            g[u] = (in[u] && a.narr == 2) ? a.gt[b * a.L + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!in[u]) continue;
            const uint32_t kp = f2key(p[u]);
            pmn = min(pmn, kp); pmx = max(pmx, kp);
            if (a.narr == 2) {
                const uint32_t kg = f2key(g[u]);
                gmn = min(gmn, kg); gmx = max(gmx, kg);
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        pmn = min(pmn, __shfl_xor_sync(0xffffffffu, pmn, o)); pmx = max(pmx, __shfl_xor_sync(0xffffffffu, pmx, o));
        gmn = min(gmn, __shfl_xor_sync(0xffffffffu, gmn, o)); gmx = max(gmx, __shfl_xor_sync(0xffffffffu, gmx, o));
    }
    if ((threadIdx.x & 31) == 0 && pmn <= pmx) {
        atomicMin(&a.imm[4 * b], pmn); atomicMax(&a.imm[4 * b + 1], pmx);
        if (a.narr == 2) {
            atomicMin(&a.imm[4 * b + 2], gmn); atomicMax(&a.imm[4 * b + 3], gmx);
            atomicMin(&a.minmax[2 * b], gmn); atomicMax(&a.minmax[2 * b + 1], gmx);
        }
    }
}

__global__ void init_imm_kernel(uint32_t* imm, uint32_t* minmax, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        imm[4 * i] = 0xFFFFFFFFu; imm[4 * i + 1] = 0u; imm[4 * i + 2] = 0xFFFFFFFFu; imm[4 * i + 3] = 0u;
        minmax[2 * i] = 0xFFFFFFFFu; minmax[2 * i + 1] = 0u;
    }
}

// binning range of every row of image b -> rlo / rsc (shared); DR gt rows use their context's own [lo, hi)
template <int MODE>
__device__ __forceinline__ void setup_ranges(const SelArgs& a, int b, const float* lo, const float* hi, float* rlo, float* rsc) {
    const int nrow = a.narr * a.K;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        float l, h;
        const uint32_t kmn = a.imm[4 * b + 2 * arr], kmx = a.imm[4 * b + 2 * arr + 1];
        if (MODE == MODE_DR && arr == 1 && kmn <= kmx) { l = lo[k]; h = hi[k]; }
        else { l = kmn <= kmx ? key2f(kmn) : 0.f; h = kmn <= kmx ? key2f(kmx) : 0.f; }
        const float w = __fsub_rn(h, l);
        rlo[threadIdx.x] = l;
        const float sc = (w > 0.f && w < 3.0e38f) ? __fdiv_rn(static_cast<float>(a.NB), w) : 0.f;
        rsc[threadIdx.x] = sc < 3.0e38f ? sc : 0.f;   // a denormal width would give an infinite scale
    }
    __syncthreads();
}

// MODE: hist (PHASE 0) or compaction (PHASE 1) over the members of every row
template <int MODE, int PHASE>
__global__ void __launch_bounds__(THREADS) lin_pass_kernel(const SelArgs a) {
    extern __shared__ uint32_t sm[];
    const int nrow = a.narr * a.K;
    float* lo = reinterpret_cast<float*>(sm);          // [K]
    float* hi = lo + a.K;                              // [K]
    float* rlo = hi + a.K;                             // [nrow]
    float* rsc = rlo + nrow;                           // [nrow]
    uint32_t* tb = reinterpret_cast<uint32_t*>(rsc + nrow);  // [nrow] target bins (PHASE 1)
    uint32_t* h = tb + nrow;                           // [nrow][NB] (PHASE 0)
    const int b = blockIdx.y;
    if (PHASE == 0)
        for (int i = threadIdx.x; i < nrow * a.NB; i += THREADS) h[i] = 0;
    if (PHASE == 1 && threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        tb[threadIdx.x] = a.tbin[row_of(a, arr, b, k)];
    }
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    setup_ranges<MODE>(a, b, lo, hi, rlo, rsc);
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    auto visit = [&](int row, int k, int arr, float x) {
        const int bq = lin_bin(x, rlo[row], rsc[row], a.NB);
        if (PHASE == 0) {
            atomicAdd(&h[row * a.NB + bq], 1u);
        } else if (static_cast<uint32_t>(bq) == tb[row]) {
            const int r = row_of(a, arr, b, k);
            const uint32_t pos = atomicAdd(&a.ccount[r], 1u);
            if (pos < static_cast<uint32_t>(a.CAP)) a.cand[static_cast<long long>(r) * a.CAP + pos] = f2key(x);
        }
    };
    for (long long base = start + threadIdx.x; base < end; base += UNROLL * THREADS) {
        float p[UNROLL], g[UNROLL];
        uint32_t mb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {  // all loads of the group are issued before any use
            const long long i = base + u * THREADS;
            const bool in = i < end;
            p[u] = in ? a.pred[b * a.L + i] : 0.f;
            g[u] = (in && (a.narr == 2 || MODE == MODE_DR)) ? a.gt[b * a.L + i] : 0.f;
            mb[u] = in ? 1u : 0u;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!mb[u]) continue;
            uint32_t bits = member_bits<MODE>(a, b, base + u * THREADS, g[u], lo, hi, has_valid);
            while (bits) {
                const int k = __ffs(bits) - 1;
                bits &= bits - 1;
                visit(k, k, 0, p[u]);
                if (a.narr == 2) visit(a.K + k, k, 1, g[u]);
            }
        }
    }
    if (PHASE == 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < nrow * a.NB; i += THREADS) {
            const uint32_t v = h[i];
            if (v) {
                const int row = i / a.NB, arr = row / a.K, k = row - arr * a.K;
                atomicAdd(&a.lhist[static_cast<long long>(row_of(a, arr, b, k)) * a.NB + (i - row * a.NB)], v);
            }
        }
    }
}

// one warp per row: the bin that holds the lower-median rank
__global__ void __launch_bounds__(THREADS) lin_scan_kernel(const SelArgs a) {
    const int r = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= a.R) return;
    const uint32_t* h = a.lhist + static_cast<long long>(r) * a.NB;
    const int per = a.NB / 32;
    uint32_t local = 0;
    for (int j = 0; j < per; ++j) local += h[lane * per + j];
    uint32_t incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) a.count[r] = total;
    if (total == 0) {
        if (lane == 0) { a.tbin[r] = 0xFFFFFFFFu; a.trank[r] = 0; a.cbin[r] = 0; a.t[r] = 0.f; }  // all-NaN row -> 0 (:490)
        return;
    }
    const uint32_t k = (total - 1) / 2;
    const uint32_t excl = incl - local;
    if (k >= excl && k < incl) {
        uint32_t cum = excl;
        for (int j = 0; j < per; ++j) {
            const uint32_t c = h[lane * per + j];
            if (k < cum + c) { a.tbin[r] = lane * per + j; a.trank[r] = k - cum; a.cbin[r] = c; break; }
            cum += c;
        }
    }
}

// one CTA per row: exact select of rank trank[r] among the members of bin tbin[r]
template <int MODE>
__global__ void __launch_bounds__(THREADS) select_rows_kernel(const SelArgs a) {
    extern __shared__ uint32_t sm[];
    float* lo = reinterpret_cast<float*>(sm);          // [K]
    float* hi = lo + a.K;
    float* rlo = hi + a.K;                             // [nrow]
    float* rsc = rlo + a.narr * a.K;
    uint32_t* hist = reinterpret_cast<uint32_t*>(rsc + a.narr * a.K);  // [256]
    uint32_t* sel = hist + 256;                        // [2] prefix, rank
    uint32_t* keys = sel + 2;                          // [CAP]
    const int r = blockIdx.x;
    const int arr = r / (a.B * a.K), b = (r / a.K) % a.B, k = r % a.K;
    const uint32_t tbin = a.tbin[r];
    if (tbin == 0xFFFFFFFFu) return;                   // empty row: t = 0 was written by the scan
    const uint32_t n = a.cbin[r];
    const bool listed = n <= static_cast<uint32_t>(a.CAP);
    const int row = arr * a.K + k;
    bool has_valid = true;
    if (listed) {
        for (uint32_t i = threadIdx.x; i < n; i += THREADS) keys[i] = a.cand[static_cast<long long>(r) * a.CAP + i];
    } else {
        has_valid = setup_thresholds<MODE>(a, b, lo, hi);
        setup_ranges<MODE>(a, b, lo, hi, rlo, rsc);
    }
    if (threadIdx.x == 0) { sel[0] = 0u; sel[1] = a.trank[r]; }
    __syncthreads();
    const float* x = arr == 0 ? a.pred : a.gt;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[threadIdx.x] = 0;   // THREADS == 256
        __syncthreads();
        const uint32_t prefix = sel[0];
        auto take = [&](uint32_t key) {
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        };
        if (listed) {
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) take(keys[i]);
        } else {
            for (long long i = threadIdx.x; i < a.L; i += THREADS) {
                const float g = (a.narr == 2 || MODE == MODE_DR) ? a.gt[b * a.L + i] : 0.f;
                const uint32_t bits = member_bits<MODE>(a, b, i, g, lo, hi, has_valid);
                if (!((bits >> k) & 1u)) continue;
                const float v = x[b * a.L + i];
                if (static_cast<uint32_t>(lin_bin(v, rlo[row], rsc[row], a.NB)) == tbin) take(f2key(v));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            uint32_t c[8], local = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; local += c[j]; }
            uint32_t incl = local;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t kk = sel[1];
            __syncwarp();  // every lane has read the rank before its owner overwrites it
            const uint32_t excl = incl - local;
            if (kk >= excl && kk < incl) {
                uint32_t cum = excl;
                int bin = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (kk >= cum + c[j]) { cum += c[j]; bin = j + 1; }
                    else break;
                }
                sel[0] = prefix | (static_cast<uint32_t>(lane * 8 + bin) << shift);
                sel[1] = kk - cum;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.t[r] = key2f(sel[0]);
}

// ---------------------------------------------------------------- sum |x - t| over members
// KCAP bounds the per-thread accumulator arrays at compile time (8 covers SSI and HDN level <= 3; 21 the largest
// explicit-context case), so the common cases do not pay for 21 predicated accumulations per pixel.
template <int MODE, int KCAP>
__global__ void __launch_bounds__(THREADS) mad_kernel(const SelArgs a) {
    extern __shared__ uint32_t sm[];
    const int nrow = a.narr * a.K;
    float* tt = reinterpret_cast<float*>(sm);  // [nrow]
    float* lo = tt + nrow;
    float* hi = lo + a.K;
    float* red = hi + a.K;                     // [nrow][8 warps]
    const int b = blockIdx.y;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        tt[threadIdx.x] = a.t[row_of(a, arr, b, k)];
    }
    __syncthreads();
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    float accp[KCAP], accg[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) { accp[j] = 0.f; accg[j] = 0.f; }
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long base = start + threadIdx.x; base < end; base += UNROLL * THREADS) {
        float pv[UNROLL], gv[UNROLL];
        bool in[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {  // all loads of the group are issued before any use
            const long long i = base + u * THREADS;
            in[u] = i < end;
            pv[u] = in[u] ? a.pred[b * a.L + i] : 0.f;
            gv[u] = (in[u] && (a.narr == 2 || MODE == MODE_DR)) ? a.gt[b * a.L + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!in[u]) continue;
            const float p = pv[u], g = gv[u];
            const uint32_t bits = member_bits<MODE>(a, b, base + u * THREADS, g, lo, hi, has_valid);
            if (!bits) continue;
#pragma unroll
            for (int k = 0; k < KCAP; ++k) {
                if (k < a.K && ((bits >> k) & 1u)) {
                    accp[k] += fabsf(p - tt[k]);
                    if (a.narr == 2) accg[k] += fabsf(g - tt[a.K + k]);
                }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KCAP; ++k) {
        if (k < a.K) {
            float v = accp[k], w = accg[k];
            for (int o = 16; o; o >>= 1) {
                v += __shfl_xor_sync(0xffffffffu, v, o);
                w += __shfl_xor_sync(0xffffffffu, w, o);
            }
            if (lane == 0) {
                red[k * 8 + warp] = v;
                if (a.narr == 2) red[(a.K + k) * 8 + warp] = w;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < nrow) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += static_cast<double>(red[threadIdx.x * 8 + w]);
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        if (v != 0.0) atomicAdd(&a.madsum[row_of(a, arr, b, k)], v);
    }
}

// s = sum / (n + 1)  (reference :470,:495)   or   sum / L  (global_normalize :177)
__global__ void scale_kernel(const SelArgs a, int mean_over_all) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const float sum = static_cast<float>(a.madsum[r]);
    a.s[r] = mean_over_all ? sum / static_cast<float>(a.L) : sum / static_cast<float>(a.count[r] + 1u);
}

// ---------------------------------------------------------------- final pass
// per pixel: v = sum_k in_k * |pa_k - ga_k| / (#contexts containing the pixel)
struct FinalArgs {
    float* aligned_pred;  // optional [B, L] (K == 1)
    float* aligned_gt;    // optional [B, L]
    float* dense;         // optional [B, L] dense map (K == 1: in * |pa - ga|)
    int accumulate;       // add sum / valid count into acc[0] / acc[1]
    int all_pixels;       // 1: every pixel is a member for the loss (global-normalised L1: mean over all)
};

template <int MODE>
__global__ void __launch_bounds__(THREADS) final_kernel(const SelArgs a, const FinalArgs f) {
    extern __shared__ uint32_t sm[];
    const int nrow = 2 * a.K;
    float* tt = reinterpret_cast<float*>(sm);  // [2K]
    float* inv = tt + nrow;                    // [2K] 1 / (s + 1e-6) is NOT used: keep the division exact
    float* lo = inv + nrow;
    float* hi = lo + a.K;
    __shared__ double red_sum[8];
    __shared__ unsigned long long red_cnt[8];
    const int b = blockIdx.y;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        tt[threadIdx.x] = a.t[row_of(a, arr, b, k)];
        inv[threadIdx.x] = a.s[row_of(a, arr, b, k)] + 1e-6f;
    }
    __syncthreads();
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    float sum = 0.f;
    unsigned long long cnt = 0;
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long base = start + threadIdx.x; base < end; base += UNROLL * THREADS) {
        float pv[UNROLL], gv[UNROLL];
        bool in[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {  // all loads of the group are issued before any use
            const long long i = base + u * THREADS;
            in[u] = i < end;
            pv[u] = in[u] ? a.pred[b * a.L + i] : 0.f;
            gv[u] = in[u] ? a.gt[b * a.L + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!in[u]) continue;
            const long long i = base + u * THREADS;
            const float p = pv[u], g = gv[u];
            const uint32_t bits = f.all_pixels ? 1u : member_bits<MODE>(a, b, i, g, lo, hi, has_valid);
            if (a.K == 1) {
                const float pa = (p - tt[0]) / inv[0];
                const float ga = (g - tt[1]) / inv[1];
                if (f.aligned_pred) f.aligned_pred[b * a.L + i] = pa;
                if (f.aligned_gt) f.aligned_gt[b * a.L + i] = ga;
                const float e = bits ? fabsf(pa - ga) : 0.f;
                if (f.dense) f.dense[b * a.L + i] = e;
                sum += e;
                cnt += bits ? 1u : 0u;
            } else if (bits) {
                float e = 0.f;
                uint32_t rest = bits;
                while (rest) {  // member contexts in ascending order (3 of 7 for HDN-DR)
                    const int k = __ffs(rest) - 1;
                    rest &= rest - 1;
                    const float pa = (p - tt[k]) / inv[k];
                    const float ga = (g - tt[a.K + k]) / inv[a.K + k];
                    e += fabsf(pa - ga);
                }
                sum += e / static_cast<float>(__popc(bits));
                cnt += 1u;
            }
        }
    }
    if (!f.accumulate) return;
    double ds = static_cast<double>(sum);
    for (int o = 16; o; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red_sum[warp] = ds; red_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        unsigned long long c = 0;
        for (int w = 0; w < 8; ++w) { s += red_sum[w]; c += red_cnt[w]; }
        atomicAdd(&a.acc[0], s);
        atomicAdd(&a.acc[1], static_cast<double>(c));
    }
}

// out = num / (den + eps); partials (num, den) exported for the multi-GPU all-reduce
__global__ void ratio_kernel(const double* acc, double eps, float* out, double* partials, int one_minus) {
    const double v = acc[0] / (acc[1] + eps);
    if (out) *out = static_cast<float>(one_minus ? 1.0 - v : v);
    if (partials) { partials[0] = acc[0]; partials[1] = acc[1]; }
}

// ---------------------------------------------------------------- workspace carving
struct Carver {
    uint8_t* p;
    size_t used = 0, cap;
    Carver(void* base, size_t cap_) : p(reinterpret_cast<uint8_t*>(base)), cap(cap_) {}
    template <typename T>
    T* take(size_t n) {
        used = (used + 255) & ~size_t(255);
        T* r = reinterpret_cast<T*>(p + used);
        used += n * sizeof(T);
        return r;
    }
};

constexpr int LIN_CAP = 4096;                      // candidate-list capacity per row
inline int lin_nb(int K) { return K <= 8 ? 512 : 128; }  // value-linear bins per row (shared-memory histogram budget)

size_t select_ws_bytes(int B, int K) {
    const size_t R = static_cast<size_t>(2) * B * K;
    const size_t radix = R * (4 * 256 * 4 + 4 + 4 + 4 + 4 + 8 + 4 + 6 * 256);
    const size_t lin = R * (static_cast<size_t>(lin_nb(K)) * 4 + LIN_CAP * 4 + 16 + 256 + 32) + static_cast<size_t>(B) * 16 + 2048;
    return 4096 + radix + lin + static_cast<size_t>(B) * 8 + 256;
}

int carve(SelArgs& a, void* ws, size_t ws_bytes, size_t* zero_bytes, size_t* zero_small = nullptr) {
    const size_t need = select_ws_bytes(a.B, a.K);
    if (!ws || ws_bytes < need)
        return set_error(DAD_ERR_WORKSPACE, "loss workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return set_error(DAD_ERR_INVALID, "workspace must be 256-byte aligned");
    Carver c(ws, ws_bytes);
    a.R = a.narr * a.B * a.K;
    // zero-initialised region first
    a.acc = c.take<double>(2);
    a.madsum = c.take<double>(a.R);
    if (zero_small) *zero_small = c.used;   // all the range-bin path needs zeroed up front
    a.hist = c.take<uint32_t>(static_cast<size_t>(4) * a.R * 256);
    a.prefix = c.take<uint32_t>(a.R);
    a.krank = c.take<uint32_t>(a.R);
    a.count = c.take<uint32_t>(a.R);
    *zero_bytes = c.used;
    a.t = c.take<float>(a.R);
    a.s = c.take<float>(a.R);
    a.minmax = c.take<uint32_t>(static_cast<size_t>(2) * a.B);
    // range-bin selection state (lhist / ccount are zeroed by run_select_lin itself)
    a.NB = lin_nb(a.K);
    a.CAP = LIN_CAP;
    a.imm = c.take<uint32_t>(static_cast<size_t>(4) * a.B);
    a.tbin = c.take<uint32_t>(a.R);
    a.trank = c.take<uint32_t>(a.R);
    a.cbin = c.take<uint32_t>(a.R);
    a.lhist = c.take<uint32_t>(static_cast<size_t>(a.R) * a.NB);
    a.ccount = c.take<uint32_t>(a.R);
    a.cand = c.take<uint32_t>(static_cast<size_t>(a.R) * a.CAP);
    a.racc = c.take<double>(static_cast<size_t>(a.B) * a.K * 3);
    a.jstar = c.take<unsigned int>(static_cast<size_t>(a.B) * a.K);
    return DAD_OK;
}

int pick_chunk(long long L, int B) {
    // enough CTAs to fill 148 SMs a few times, at least 2048 pixels each
    long long want = cdivl(L * B, 148LL * 8);
    if (want < 2048) want = 2048;
    want = cdivl(want, THREADS) * THREADS;
    return static_cast<int>(want > L ? cdivl(L, THREADS) * THREADS : want);
}

template <int MODE>
int run_select(SelArgs& a, int mean_over_all, cudaStream_t st) {
    const dim3 grid(static_cast<unsigned>(cdivl(a.L, a.chunk)), a.B);
    const int nrow = a.narr * a.K;
    const size_t sm_hist = static_cast<size_t>(nrow) * 256 * 4 + nrow * 4 + 2 * a.K * 4;
    if (sm_hist > 48 * 1024) {
        static bool done = false;
        if (!done) {
            DAD_CHECK_CUDA(cudaFuncSetAttribute(sel_hist_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            done = true;
        }
    }
    if (MODE == MODE_DR) {
        init_minmax_kernel<<<cdiv(a.B, 128), 128, 0, st>>>(a.minmax, a.B);
        minmax_kernel<<<grid, THREADS, 0, st>>>(a.gt, a.mask, a.L, a.chunk, a.minmax);
        DAD_CHECK_LAUNCH();
    }
    for (int pass = 0; pass < 4; ++pass) {
        sel_hist_kernel<MODE><<<grid, THREADS, sm_hist, st>>>(a, pass);
        sel_scan_kernel<<<cdiv(a.R, THREADS / 32), THREADS, 0, st>>>(a, pass);
        DAD_CHECK_LAUNCH();
    }
    const size_t sm_mad = static_cast<size_t>(nrow) * 4 + 2 * a.K * 4 + static_cast<size_t>(nrow) * 8 * 4;
    if (a.K <= 8) mad_kernel<MODE, 8><<<grid, THREADS, sm_mad, st>>>(a);
    else mad_kernel<MODE, MAX_K><<<grid, THREADS, sm_mad, st>>>(a);
    scale_kernel<<<cdiv(a.R, 128), 128, 0, st>>>(a, mean_over_all);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

// medians by range-bin selection (see the kernels above), then the same MAD / scale passes
template <int MODE>
int run_select_lin(SelArgs& a, int mean_over_all, cudaStream_t st) {
    const int nrow = a.narr * a.K;
    // every CTA flushes nrow * NB bins: keep the flush well below the pixel work (SSI: 1 K bins, HDN-DR: 7 K bins)
    int chunk = a.chunk * (nrow * a.NB > 2048 ? 2 : 1);
    if (chunk > a.L) chunk = static_cast<int>(cdivl(a.L, THREADS) * THREADS);
    SelArgs h = a;
    h.chunk = chunk;
    const dim3 grid(static_cast<unsigned>(cdivl(a.L, a.chunk)), a.B);
    const dim3 gridh(static_cast<unsigned>(cdivl(a.L, chunk)), a.B);
    const size_t sm_pass = (2 * a.K + 3 * nrow) * 4 + static_cast<size_t>(nrow) * a.NB * 4;
    const size_t sm_sel = (2 * a.K + 2 * nrow + 256 + 2 + a.CAP) * 4;
    static bool done = false;
    if (!done) {
        DAD_CHECK_CUDA(cudaFuncSetAttribute(lin_pass_kernel<MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        done = true;
    }
    DAD_REQUIRE(sm_pass <= 64 * 1024, "loss: K=%d contexts need too much shared memory", a.K);
    DAD_CHECK_CUDA(cudaMemsetAsync(a.lhist, 0, reinterpret_cast<uint8_t*>(a.ccount + a.R) - reinterpret_cast<uint8_t*>(a.lhist),
                                   st));  // lhist and the ccount cursors behind it
    init_imm_kernel<<<cdiv(a.B, 128), 128, 0, st>>>(a.imm, a.minmax, a.B);
    imgminmax_kernel<<<grid, THREADS, 0, st>>>(a);
    lin_pass_kernel<MODE, 0><<<gridh, THREADS, sm_pass, st>>>(h);
    lin_scan_kernel<<<cdiv(a.R, THREADS / 32), THREADS, 0, st>>>(a);
    lin_pass_kernel<MODE, 1><<<grid, THREADS, (2 * a.K + 3 * nrow) * 4, st>>>(a);
    select_rows_kernel<MODE><<<a.R, THREADS, sm_sel, st>>>(a);
    DAD_CHECK_LAUNCH();
    const size_t sm_mad = static_cast<size_t>(nrow) * 4 + 2 * a.K * 4 + static_cast<size_t>(nrow) * 8 * 4;
    if (a.K <= 8) mad_kernel<MODE, 8><<<grid, THREADS, sm_mad, st>>>(a);
    else mad_kernel<MODE, MAX_K><<<grid, THREADS, sm_mad, st>>>(a);
    scale_kernel<<<cdiv(a.R, 128), 128, 0, st>>>(a, mean_over_all);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

template <int MODE>
int run_final(const SelArgs& a, const FinalArgs& f, cudaStream_t st) {
    const dim3 grid(static_cast<unsigned>(cdivl(a.L, a.chunk)), a.B);
    const size_t sm = static_cast<size_t>(4) * a.K * 4 + 2 * a.K * 4;
    final_kernel<MODE><<<grid, THREADS, sm, st>>>(a, f);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int ssi_common(int mode, const float* pred, const float* gt, const uint8_t* mask, const uint8_t* ctx, int K, int level,
               int B, long long L, float* aligned_pred, float* aligned_gt, float* dense, float* out_scalar,
               double* partials, int mean_over_all, int all_pixels, void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(pred && gt, "loss: null input");
    DAD_REQUIRE(B > 0 && L > 0, "loss: empty input (B=%d, L=%lld)", B, L);
    DAD_REQUIRE(K >= 1 && K <= MAX_K, "loss: K=%d contexts unsupported (max %d)", K, MAX_K);
    SelArgs a{};
    a.pred = pred; a.gt = gt; a.mask = mask; a.ctx = ctx;
    a.B = B; a.K = K; a.level = level; a.narr = 2; a.L = L;
    a.chunk = pick_chunk(L, B);
    size_t zero_bytes = 0, zero_small = 0;
    DAD_TRY(carve(a, ws, ws_bytes, &zero_bytes, &zero_small));
    static const bool radix = getenv("DAD_LOSS_RADIX") != nullptr;  // A/B switch: the 4 x 8-bit radix passes
    if (!radix) zero_bytes = zero_small;
    // algorithmic bytes: pred + gt fp32 (+ 1-byte mask / K-byte contexts) per pixel
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * L * (8.0 + (mask ? 1 : 0) + (ctx ? K : 0)), st,
                   (mode == MODE_DR ? 2 : 0) + 8 + 2 + 1 + ((out_scalar || partials) ? 1 : 0));
    DAD_CHECK_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, st));
    FinalArgs f{};
    f.aligned_pred = aligned_pred; f.aligned_gt = aligned_gt; f.dense = dense;
    f.accumulate = (out_scalar || partials) ? 1 : 0;
    f.all_pixels = all_pixels;
    if (mode == MODE_MASK) {
        DAD_TRY(radix ? run_select<MODE_MASK>(a, mean_over_all, st) : run_select_lin<MODE_MASK>(a, mean_over_all, st));
        DAD_TRY(run_final<MODE_MASK>(a, f, st));
    } else if (mode == MODE_DR) {
        DAD_TRY(radix ? run_select<MODE_DR>(a, mean_over_all, st) : run_select_lin<MODE_DR>(a, mean_over_all, st));
        DAD_TRY(run_final<MODE_DR>(a, f, st));
    } else {
        DAD_TRY(radix ? run_select<MODE_CTX>(a, mean_over_all, st) : run_select_lin<MODE_CTX>(a, mean_over_all, st));
        DAD_TRY(run_final<MODE_CTX>(a, f, st));
    }
    if (f.accumulate) {
        ratio_kernel<<<1, 1, 0, st>>>(a.acc, all_pixels ? 0.0 : 1e-6, out_scalar, partials, 0);
        DAD_CHECK_LAUNCH();
    }
    return DAD_OK;
}

// ================================================================ backward of SSI / HDN w.r.t. pred
// Forward (tools/train_distillation.py:449-542, 686-707), per row r = (image, context k) with members M_r:
//   t = lower median of p over M_r,  s = sum_{M_r} |p - t| / (n + 1),  pa_i = (p_i - t) / (s + 1e-6)
//   L = (1 / (N + 1e-6)) * sum_i (1 / c_i) * sum_{k contains i} |pa_ik - ga_ik|      (c_i = 1, N = sum mask for SSI)
// PyTorch's autograd through nanmedian routes d t / d p to the selected element j*, and |.| has sign(0) = 0, so with
//   e_ik = sgn(pa_ik - ga_ik) / c_i,  E = sum e,  G = sum e * (p - t),  S = sum_{M_r} sign(p - t):
//   dL/dp_j * (N + 1e-6) = sum_{k contains j} [ e_jk / (s+eps) - G / (s+eps)^2 * sign(p_j - t) / (n+1) ]
//                          + [j == j*_k] * ( -E / (s+eps) + G / (s+eps)^2 * S / (n+1) )
// Pass 1 reduces E, G, S per row and finds j* (lowest member index holding the median value); pass 2 writes the map.
struct BwdArgs {
    const float* gout;   // upstream gradient of the scalar loss (device)
    float* grad;         // [B, L]
};

__device__ __forceinline__ float fsign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

template <int MODE, int KCAP>
__global__ void __launch_bounds__(THREADS) bwd_reduce_kernel(const SelArgs a) {
    extern __shared__ uint32_t sm[];
    const int nrow = 2 * a.K;
    float* tt = reinterpret_cast<float*>(sm);  // [2K] medians (pred rows, then gt rows)
    float* den = tt + nrow;                    // [2K] s + 1e-6
    float* lo = den + nrow;
    float* hi = lo + a.K;
    float* red = hi + a.K;                     // [3K][8 warps]
    __shared__ unsigned long long red_cnt[8];
    const int b = blockIdx.y;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        tt[threadIdx.x] = a.t[row_of(a, arr, b, k)];
        den[threadIdx.x] = a.s[row_of(a, arr, b, k)] + 1e-6f;
    }
    __syncthreads();
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    float accE[KCAP], accG[KCAP], accS[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) { accE[j] = 0.f; accG[j] = 0.f; accS[j] = 0.f; }
    unsigned long long cnt = 0;
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float p = a.pred[b * a.L + i], g = a.gt[b * a.L + i];
        const uint32_t bits = member_bits<MODE>(a, b, i, g, lo, hi, has_valid);
        if (!bits) continue;
        cnt += 1u;
        const float w = 1.0f / static_cast<float>(__popc(bits));
#pragma unroll
        for (int k = 0; k < KCAP; ++k) {
            if (k < a.K && ((bits >> k) & 1u)) {
                const float dp = p - tt[k];
                const float pa = dp / den[k], ga = (g - tt[a.K + k]) / den[a.K + k];
                const float e = w * fsign(pa - ga);
                accE[k] += e;
                accG[k] += e * dp;
                accS[k] += fsign(dp);
                if (dp == 0.f) atomicMin(&a.jstar[b * a.K + k], static_cast<unsigned int>(i));
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KCAP; ++k) {
        if (k < a.K) {
            float e = accE[k], gq = accG[k], sq = accS[k];
            for (int o = 16; o; o >>= 1) {
                e += __shfl_xor_sync(0xffffffffu, e, o);
                gq += __shfl_xor_sync(0xffffffffu, gq, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            if (lane == 0) { red[(3 * k) * 8 + warp] = e; red[(3 * k + 1) * 8 + warp] = gq; red[(3 * k + 2) * 8 + warp] = sq; }
        }
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) red_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x < 3 * a.K) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += static_cast<double>(red[threadIdx.x * 8 + w]);
        const int k = threadIdx.x / 3, q = threadIdx.x - 3 * k;
        if (v != 0.0) atomicAdd(&a.racc[(static_cast<long long>(b) * a.K + k) * 3 + q], v);
    }
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
        for (int w = 0; w < 8; ++w) c += red_cnt[w];
        atomicAdd(&a.acc[1], static_cast<double>(c));
    }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) bwd_apply_kernel(const SelArgs a, const BwdArgs w) {
    extern __shared__ uint32_t sm[];
    const int nrow = 2 * a.K;
    float* tt = reinterpret_cast<float*>(sm);   // [2K]
    float* den = tt + nrow;                     // [2K]
    float* lo = den + nrow;
    float* hi = lo + a.K;
    float* cS = hi + a.K;                       // [K]  G / (s+eps)^2 / (n+1)       (coefficient of sign(p - t))
    float* cJ = cS + a.K;                       // [K]  -E / (s+eps) + cS * S       (extra term at j*)
    unsigned int* js = reinterpret_cast<unsigned int*>(cJ + a.K);  // [K]
    const int b = blockIdx.y;
    if (threadIdx.x < nrow) {
        const int arr = threadIdx.x / a.K, k = threadIdx.x - arr * a.K;
        tt[threadIdx.x] = a.t[row_of(a, arr, b, k)];
        den[threadIdx.x] = a.s[row_of(a, arr, b, k)] + 1e-6f;
    }
    if (threadIdx.x < a.K) {
        const int k = threadIdx.x;
        const double* r = a.racc + (static_cast<long long>(b) * a.K + k) * 3;
        const double d = static_cast<double>(a.s[row_of(a, 0, b, k)]) + 1e-6;
        const double n1 = a.mean_all ? static_cast<double>(a.L) : static_cast<double>(a.count[row_of(a, 0, b, k)]) + 1.0;
        const double cs = r[1] / (d * d) / n1;
        cS[k] = static_cast<float>(cs);
        cJ[k] = static_cast<float>(-r[0] / d + cs * r[2]);
        js[k] = a.jstar[b * a.K + k];
    }
    __syncthreads();
    const bool has_valid = setup_thresholds<MODE>(a, b, lo, hi);
    const float scale = *w.gout / static_cast<float>(a.acc[1] + 1e-6);
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float p = a.pred[b * a.L + i], g = a.gt[b * a.L + i];
        uint32_t bits = member_bits<MODE>(a, b, i, g, lo, hi, has_valid);
        float gsum = 0.f;
        if (bits) {
            const float wgt = 1.0f / static_cast<float>(__popc(bits));
            while (bits) {
                const int k = __ffs(bits) - 1;
                bits &= bits - 1;
                const float dp = p - tt[k];
                const float pa = dp / den[k], ga = (g - tt[a.K + k]) / den[a.K + k];
                gsum += wgt * fsign(pa - ga) / den[k] - cS[k] * fsign(dp);
                if (static_cast<unsigned int>(i) == js[k]) gsum += cJ[k];
            }
        }
        w.grad[b * a.L + i] = gsum * scale;
    }
}

__global__ void init_jstar_kernel(unsigned int* jstar, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) jstar[i] = 0xFFFFFFFFu;
}

template <int MODE>
int run_backward(SelArgs& a, const BwdArgs& w, cudaStream_t st) {
    DAD_TRY(run_select_lin<MODE>(a, a.mean_all, st));   // medians, scales, counts (the forward's statistics, recomputed)
    const dim3 grid(static_cast<unsigned>(cdivl(a.L, a.chunk)), a.B);
    DAD_CHECK_CUDA(cudaMemsetAsync(a.racc, 0, static_cast<size_t>(a.B) * a.K * 3 * 8, st));
    init_jstar_kernel<<<cdiv(a.B * a.K, 128), 128, 0, st>>>(a.jstar, a.B * a.K);
    const size_t sm_r = (static_cast<size_t>(4) * a.K + 2 * a.K + static_cast<size_t>(3) * a.K * 8) * 4;
    if (a.K <= 8) bwd_reduce_kernel<MODE, 8><<<grid, THREADS, sm_r, st>>>(a);
    else bwd_reduce_kernel<MODE, MAX_K><<<grid, THREADS, sm_r, st>>>(a);
    const size_t sm_a = (static_cast<size_t>(4) * a.K + 2 * a.K + 3 * a.K) * 4;
    bwd_apply_kernel<MODE><<<grid, THREADS, sm_a, st>>>(a, w);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int ssi_backward_common(int mode, const float* pred, const float* gt, const uint8_t* mask, const uint8_t* ctx, int K, int level,
                        int B, long long L, const float* gout, float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st,
                        int mean_all = 0) {
    DAD_REQUIRE(pred && gt && gout && grad_pred, "loss backward: null argument");
    DAD_REQUIRE(B > 0 && L > 0 && L < (1LL << 32) - 1, "loss backward: bad size (B=%d, L=%lld)", B, L);
    DAD_REQUIRE(K >= 1 && K <= MAX_K, "loss backward: K=%d contexts unsupported (max %d)", K, MAX_K);
    SelArgs a{};
    a.pred = pred; a.gt = gt; a.mask = mask; a.ctx = ctx;
    a.B = B; a.K = K; a.level = level; a.narr = 2; a.L = L; a.mean_all = mean_all;
    a.chunk = pick_chunk(L, B);
    size_t zero_bytes = 0, zero_small = 0;
    DAD_TRY(carve(a, ws, ws_bytes, &zero_bytes, &zero_small));
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * L * (12.0 + (mask ? 1 : 0) + (ctx ? K : 0)), st, 14);
    DAD_CHECK_CUDA(cudaMemsetAsync(ws, 0, zero_small, st));
    BwdArgs w{gout, grad_pred};
    if (mode == MODE_MASK) return run_backward<MODE_MASK>(a, w, st);
    if (mode == MODE_DR) return run_backward<MODE_DR>(a, w, st);
    return run_backward<MODE_CTX>(a, w, st);
}

// ---------------------------------------------------------------- Sobel gradient loss backward
// L = mean exp(-m), m = sqrt(gx^2 + gy^2 + 1e-6):  a = dL/dg = -exp(-m) * g / (m * N);  dL/dd = corr^T(ax, kx) + corr^T(ay, ky)
__global__ void __launch_bounds__(THREADS) sobel_bwd_kernel(const float* d, int H, int W, const float* gout, double n_inv,
                                                            float* grad) {
    constexpr int TX = 32, TY = 8;
    __shared__ float sd[TY + 4][TX + 4];
    __shared__ float sax[TY + 2][TX + 2], say[TY + 2][TX + 2];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const float* img = d + static_cast<long long>(b) * H * W;
    for (int i = threadIdx.x; i < (TY + 4) * (TX + 4); i += THREADS) {
        const int ly = i / (TX + 4), lx = i - ly * (TX + 4);
        const int y = y0 + ly - 2, x = x0 + lx - 2;
        sd[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? img[static_cast<long long>(y) * W + x] : 0.f;
    }
    __syncthreads();
    const float c = -(*gout) * static_cast<float>(n_inv);
    for (int i = threadIdx.x; i < (TY + 2) * (TX + 2); i += THREADS) {
        const int ly = i / (TX + 2), lx = i - ly * (TX + 2);
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        float ax = 0.f, ay = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {   // gradient magnitude exists only at real pixels
            const float (*r)[TX + 4] = reinterpret_cast<const float (*)[TX + 4]>(&sd[ly][lx]);  // window origin (y-1, x-1)
            const float gx = (r[0][2] - r[0][0]) + 2.f * (r[1][2] - r[1][0]) + (r[2][2] - r[2][0]);
            const float gy = (r[2][0] - r[0][0]) + 2.f * (r[2][1] - r[0][1]) + (r[2][2] - r[0][2]);
            const float m = sqrtf(gx * gx + gy * gy + 1e-6f);
            const float f = c * expf(-m) / m;
            ax = f * gx;
            ay = f * gy;
        }
        sax[ly][lx] = ax;
        say[ly][lx] = ay;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x < W && y < H) {
        // d gx(y', x') / d d(y, x) = kx[y - y' + 1][x - x' + 1]; sum over the 3 x 3 neighbours (y', x')
        float gsum = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int u = -dy + 1, v = -dx + 1;                           // kernel tap seen from the neighbour
                const float kx = (v == 2 ? 1.f : (v == 0 ? -1.f : 0.f)) * (u == 1 ? 2.f : 1.f);
                const float ky = (u == 2 ? 1.f : (u == 0 ? -1.f : 0.f)) * (v == 1 ? 2.f : 1.f);
                gsum += kx * sax[ly + 1 + dy][lx + 1 + dx] + ky * say[ly + 1 + dy][lx + 1 + dx];
            }
        grad[(static_cast<long long>(b) * H + y) * W + x] = gsum;
    }
}

// ---------------------------------------------------------------- HDN-DR context export (bool [K,B,L])
__global__ void __launch_bounds__(THREADS) contexts_dr_kernel(const SelArgs a, uint8_t* out) {
    __shared__ float lo[MAX_K], hi[MAX_K];
    const int b = blockIdx.y;
    const bool has_valid = setup_thresholds<MODE_DR>(a, b, lo, hi);
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const uint32_t bits = member_bits<MODE_DR>(a, b, i, a.gt[b * a.L + i], lo, hi, has_valid);
        for (int k = 0; k < a.K; ++k) out[(static_cast<long long>(k) * a.B + b) * a.L + i] = (bits >> k) & 1u;
    }
}

// ---------------------------------------------------------------- HDN-DP context export (bool [K,B,L])
// get_contexts_dp (tools/train_distillation.py:578-644): bins between nanquantile(i * bin) and nanquantile((i + 1) * bin)
// of the valid gt values; 'linear' interpolation = lerp(sorted[floor r], sorted[ceil r], r - floor r) with ATen's
// fused form (fma(w, d, lo) for w < 0.5, fma(w - 1, d, hi) otherwise); an image without valid pixels has NaN
// quantiles, i.e. empty contexts.
__global__ void __launch_bounds__(THREADS) contexts_dp_kernel(const SelArgs a, uint8_t* out) {
    __shared__ float Q[MAX_K];
    const int b = blockIdx.y;
    const uint32_t n = a.count[row_of(a, 0, b, 0)];
    if (threadIdx.x < a.nq && n) {
        const int j = threadIdx.x;
        const float lo = a.t[row_of(a, 0, b, 2 * j)], hi = a.t[row_of(a, 0, b, 2 * j + 1)];
        const float rk = quantile_rank(j, a.nq, n);
        const float w = __fsub_rn(rk, floorf(rk));
        const float d = __fsub_rn(hi, lo);
        Q[j] = (w < 0.5f) ? __fmaf_rn(w, d, lo) : __fmaf_rn(__fsub_rn(w, 1.0f), d, hi);
    }
    __syncthreads();
    const int nb = a.nq - 1;  // finest level: nb bins
    const long long start = static_cast<long long>(blockIdx.x) * a.chunk;
    const long long end = min(start + a.chunk, a.L);
    const int K = (nb << 1) - 1;
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float g = a.gt[b * a.L + i];
        const bool valid = n && (a.mask ? a.mask[b * a.L + i] != 0 : true);
        int k = 0;
        for (int step = 1; step <= nb; step <<= 1)        // bins of `step` finest bins each, finest level first
            for (int j = 0; j < nb; j += step, ++k)
                out[(static_cast<long long>(k) * a.B + b) * a.L + i] = valid && g >= Q[j] && g < Q[j + step];
        (void)K;
    }
}

// ---------------------------------------------------------------- HDN-DS context export (bool [K,B,H,W])
// get_contexts_ds / init_temp_masks_ds (:646-673): an n x n grid of [int(h*size/n), int((h+1)*size/n)) squares per
// level (size = W; the reference needs square maps), AND-ed with the valid mask.
__global__ void __launch_bounds__(THREADS) contexts_ds_kernel(const uint8_t* mask, int B, int H, int W, int level,
                                                              uint8_t* out) {
    const int b = blockIdx.y;
    const long long L = static_cast<long long>(H) * W;
    const int nb = 1 << (level - 1);
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < L;
         i += static_cast<long long>(gridDim.x) * THREADS) {
        const int y = static_cast<int>(i / W), x = static_cast<int>(i - static_cast<long long>(y) * W);
        const bool valid = mask ? mask[b * L + i] != 0 : true;
        int k = 0;
        for (int n = nb; n >= 1; n >>= 1)
            for (int h = 0; h < n; ++h) {
                const bool iny = y >= (h * W) / n && y < ((h + 1) * W) / n;
                for (int w = 0; w < n; ++w, ++k)
                    out[(static_cast<long long>(k) * B + b) * L + i] =
                        valid && iny && x >= (w * W) / n && x < ((w + 1) * W) / n;
            }
    }
}

// ---------------------------------------------------------------- Sobel gradient loss
__global__ void __launch_bounds__(THREADS) sobel_kernel(const float* d, int H, int W, double* acc) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y0 = blockIdx.y * 64 + (threadIdx.x >> 5) * 8;
    const float* img = d + static_cast<long long>(b) * H * W;
    float sum = 0.f;
    if (x < W) {
        auto at = [&](int yy, int xx) -> float {
            return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[static_cast<long long>(yy) * W + xx] : 0.f;
        };
        // sliding 3-row window down 8 rows
        float r0[3], r1[3], r2[3];
        for (int j = 0; j < 3; ++j) { r0[j] = at(y0 - 1, x - 1 + j); r1[j] = at(y0, x - 1 + j); }
        for (int yy = y0; yy < y0 + 8 && yy < H; ++yy) {
            for (int j = 0; j < 3; ++j) r2[j] = at(yy + 1, x - 1 + j);
            const float gx = (r0[2] - r0[0]) + 2.f * (r1[2] - r1[0]) + (r2[2] - r2[0]);
            const float gy = (r2[0] - r0[0]) + 2.f * (r2[1] - r0[1]) + (r2[2] - r0[2]);
            sum += expf(-sqrtf(gx * gx + gy * gy + 1e-6f));
            for (int j = 0; j < 3; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
        }
    }
    __shared__ float red[8];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(&acc[0], s);
    }
}

__global__ void set_den_kernel(double* acc, double den) { acc[1] = den; }

// ---------------------------------------------------------------- feature cosine loss
// s [B,N,Ds], t [B,N,Dt]; D = min(Ds, Dt); wider tensor nearest-resized along channels.
__global__ void __launch_bounds__(THREADS) featcos_kernel(const float* s, const float* t, int N, int Ds, int Dt, int D,
                                                          double* acc) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5;
    float st = 0.f, ss = 0.f, tt = 0.f;
    if (c < D) {
        // F.interpolate(mode='nearest'): src = min(floor(dst * (in / out)), in - 1), scale in fp32
        const int cs = (Ds == D) ? c : min(static_cast<int>(floorf(c * (static_cast<float>(Ds) / D))), Ds - 1);
        const int ct = (Dt == D) ? c : min(static_cast<int>(floorf(c * (static_cast<float>(Dt) / D))), Dt - 1);
        const float* sp = s + static_cast<long long>(b) * N * Ds + cs;
        const float* tp = t + static_cast<long long>(b) * N * Dt + ct;
        for (int n = warp; n < N; n += 8) {
            const float a = sp[static_cast<long long>(n) * Ds], v = tp[static_cast<long long>(n) * Dt];
            st = fmaf(a, v, st); ss = fmaf(a, a, ss); tt = fmaf(v, v, tt);
        }
    }
    __shared__ float red[3][8][32];
    red[0][warp][threadIdx.x & 31] = st; red[1][warp][threadIdx.x & 31] = ss; red[2][warp][threadIdx.x & 31] = tt;
    __syncthreads();
    if (warp == 0) {
        float a = 0.f, q = 0.f, r = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w][threadIdx.x]; q += red[1][w][threadIdx.x]; r += red[2][w][threadIdx.x]; }
        float cosv = 0.f;
        if (c < D) {
            // F.normalize(eps=1e-12) on both, then cosine_similarity(eps=1e-8) of the unit vectors
            const float ns = fmaxf(sqrtf(q), 1e-12f), nt = fmaxf(sqrtf(r), 1e-12f);
            const float dot = a / (ns * nt);
            const float n1 = sqrtf(q) / ns, n2 = sqrtf(r) / nt;
            cosv = dot / fmaxf(n1 * n2, 1e-8f);
        }
        for (int o = 16; o; o >>= 1) cosv += __shfl_xor_sync(0xffffffffu, cosv, o);
        if (threadIdx.x == 0) atomicAdd(&acc[0], static_cast<double>(cosv));
    }
}

// ---------------------------------------------------------------- plain / hybrid-normalised L1
__global__ void __launch_bounds__(THREADS) l1_kernel(const float* a, const float* b, long long n, double* acc) {
    float sum = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * THREADS)
        sum += fabsf(a[i] - b[i]);
    __shared__ float red[8];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(&acc[0], s);
    }
}

constexpr int MAX_SEG = 8;
struct HybArgs {
    const float* x[2];   // student, teacher [B, L]
    int B, nseg;
    long long L;
    int chunk;
    uint32_t* minmax;    // [2][B][2]
    double* segsum;      // [2][B][nseg]
    double* segcnt;      // [2][B][nseg]
    double* segmad;      // [2][B][nseg]
    double* acc;
    float* norm_out[2];  // optional normalised maps
};

// segment bounds, reference op order (:198): bound_i = min + (i / nseg) * range
__device__ __forceinline__ void seg_bounds(float mn, float mx, int nseg, float* bnd) {
    const float range = __fsub_rn(mx, mn);
    for (int i = 0; i <= nseg; ++i)
        bnd[i] = __fadd_rn(mn, __fmul_rn(static_cast<float>(static_cast<double>(i) / nseg), range));
}

__global__ void __launch_bounds__(THREADS) hyb_stats_kernel(const HybArgs h, int phase) {
    // phase 0: per-segment sum and count; phase 1: per-segment sum |d - mean|
    const int b = blockIdx.y, arr = blockIdx.z;
    __shared__ float bnd[MAX_SEG + 1], mean[MAX_SEG];
    __shared__ float red[2][MAX_SEG][8];
    if (threadIdx.x == 0) {
        seg_bounds(key2f(h.minmax[(arr * h.B + b) * 2]), key2f(h.minmax[(arr * h.B + b) * 2 + 1]), h.nseg, bnd);
        for (int s = 0; s < h.nseg; ++s) {
            const long long o = (static_cast<long long>(arr) * h.B + b) * h.nseg + s;
            mean[s] = static_cast<float>(h.segsum[o]) / (static_cast<float>(h.segcnt[o]) + 1e-6f);
        }
    }
    __syncthreads();
    float a0[MAX_SEG], a1[MAX_SEG];
#pragma unroll
    for (int s = 0; s < MAX_SEG; ++s) { a0[s] = 0.f; a1[s] = 0.f; }
    const float* x = h.x[arr] + static_cast<long long>(b) * h.L;
    const long long start = static_cast<long long>(blockIdx.x) * h.chunk;
    const long long end = min(start + h.chunk, h.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float d = x[i];
#pragma unroll
        for (int s = 0; s < MAX_SEG; ++s) {
            if (s < h.nseg && d >= bnd[s] && d <= bnd[s + 1]) {
                if (phase == 0) { a0[s] += d; a1[s] += 1.f; }
                else a0[s] += fabsf(d - mean[s]);
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < MAX_SEG; ++s) {
        if (s < h.nseg) {
            float v0 = a0[s], v1 = a1[s];
            for (int o = 16; o; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
            if (lane == 0) { red[0][s][warp] = v0; red[1][s][warp] = v1; }
        }
    }
    __syncthreads();
    if (threadIdx.x < h.nseg) {
        double v0 = 0.0, v1 = 0.0;
        for (int w = 0; w < 8; ++w) { v0 += red[0][threadIdx.x][w]; v1 += red[1][threadIdx.x][w]; }
        const long long o = (static_cast<long long>(arr) * h.B + b) * h.nseg + threadIdx.x;
        if (phase == 0) { atomicAdd(&h.segsum[o], v0); atomicAdd(&h.segcnt[o], v1); }
        else atomicAdd(&h.segmad[o], v0);
    }
}

__global__ void __launch_bounds__(THREADS) hyb_final_kernel(const HybArgs h) {
    const int b = blockIdx.y;
    __shared__ float bnd[2][MAX_SEG + 1], mean[2][MAX_SEG], den[2][MAX_SEG];
    __shared__ int present[2][MAX_SEG];
    if (threadIdx.x < 2) {
        const int arr = threadIdx.x;
        seg_bounds(key2f(h.minmax[(arr * h.B + b) * 2]), key2f(h.minmax[(arr * h.B + b) * 2 + 1]), h.nseg, bnd[arr]);
        for (int s = 0; s < h.nseg; ++s) {
            const long long o = (static_cast<long long>(arr) * h.B + b) * h.nseg + s;
            const float cnt = static_cast<float>(h.segcnt[o]) + 1e-6f;
            mean[arr][s] = static_cast<float>(h.segsum[o]) / cnt;
            den[arr][s] = static_cast<float>(h.segmad[o]) / cnt + 1e-6f;
        }
    }
    __syncthreads();
    float sum = 0.f;
    const long long start = static_cast<long long>(blockIdx.x) * h.chunk;
    const long long end = min(start + h.chunk, h.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        float nv[2];
#pragma unroll
        for (int arr = 0; arr < 2; ++arr) {
            const float d = h.x[arr][static_cast<long long>(b) * h.L + i];
            float v = 0.f;
            for (int s = 0; s < h.nseg; ++s)  // later segment wins on shared boundaries (:247)
                if (d >= bnd[arr][s] && d <= bnd[arr][s + 1]) v = (d - mean[arr][s]) / den[arr][s];
            nv[arr] = v;
            if (h.norm_out[arr]) h.norm_out[arr][static_cast<long long>(b) * h.L + i] = v;
        }
        sum += fabsf(nv[0] - nv[1]);
    }
    __shared__ float red[8];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(&h.acc[0], s);
    }
    (void)present;
}

// ---------------------------------------------------------------- backward: plain / hybrid-normalised L1, feature cosine
__global__ void __launch_bounds__(THREADS) l1_bwd_kernel(const float* a, const float* b, long long n, const float* gout,
                                                         float* grad) {
    const float c = *gout / static_cast<float>(n);
    for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * THREADS)
        grad[i] = c * fsign(a[i] - b[i]);
}

// hybrid_normalize (:217-249) backward w.r.t. x[0]; x[1] is normalised by its own statistics and detached.
//   ns_i = (d_i - mu_g) / (sigma_g + 1e-6) with g = the LAST segment containing i;  mu_g, sigma_g over ALL members of g
//   e_i = sgn(ns_i - nt_i) / (B L);  A_g = sum_{g(i)=g} e_i;  Gq_g = sum_{g(i)=g} e_i (d_i - mu_g);  Sg_g = sum_{M_g} sign(d - mu_g)
//   dL/dd_j = e_j / den_g(j) + sum_{g contains j} [ -A_g / (den_g cnt_g) - Gq_g / den_g^2 * (sign(d_j - mu_g) - Sg_g / cnt_g) / cnt_g ]
struct HybBwd {
    double* racc;        // [B][nseg][3]  A, Gq, Sg
    const float* gout;
    float* grad;
};

__device__ __forceinline__ void hyb_load_stats(const HybArgs& h, int b, float (*bnd)[MAX_SEG + 1], float (*mean)[MAX_SEG],
                                               float (*den)[MAX_SEG], float (*cnt)[MAX_SEG]) {
    if (threadIdx.x < 2) {
        const int arr = threadIdx.x;
        seg_bounds(key2f(h.minmax[(arr * h.B + b) * 2]), key2f(h.minmax[(arr * h.B + b) * 2 + 1]), h.nseg, bnd[arr]);
        for (int s = 0; s < h.nseg; ++s) {
            const long long o = (static_cast<long long>(arr) * h.B + b) * h.nseg + s;
            const float c = static_cast<float>(h.segcnt[o]) + 1e-6f;
            cnt[arr][s] = c;
            mean[arr][s] = static_cast<float>(h.segsum[o]) / c;
            den[arr][s] = static_cast<float>(h.segmad[o]) / c + 1e-6f;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS) hyb_bwd_reduce_kernel(const HybArgs h, const HybBwd w) {
    const int b = blockIdx.y;
    __shared__ float bnd[2][MAX_SEG + 1], mean[2][MAX_SEG], den[2][MAX_SEG], cnt[2][MAX_SEG];
    __shared__ float red[3][MAX_SEG][8];
    hyb_load_stats(h, b, bnd, mean, den, cnt);
    float aA[MAX_SEG], aG[MAX_SEG], aS[MAX_SEG];
#pragma unroll
    for (int s = 0; s < MAX_SEG; ++s) { aA[s] = 0.f; aG[s] = 0.f; aS[s] = 0.f; }
    const long long start = static_cast<long long>(blockIdx.x) * h.chunk;
    const long long end = min(start + h.chunk, h.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float d = h.x[0][static_cast<long long>(b) * h.L + i], t = h.x[1][static_cast<long long>(b) * h.L + i];
        float ns = 0.f, nt = 0.f;
        int gs = -1;
        for (int s = 0; s < h.nseg; ++s) {
            if (d >= bnd[0][s] && d <= bnd[0][s + 1]) { ns = (d - mean[0][s]) / den[0][s]; gs = s; }
            if (t >= bnd[1][s] && t <= bnd[1][s + 1]) nt = (t - mean[1][s]) / den[1][s];
        }
        const float e = fsign(ns - nt);
#pragma unroll
        for (int s = 0; s < MAX_SEG; ++s) {
            if (s < h.nseg && d >= bnd[0][s] && d <= bnd[0][s + 1]) {
                aS[s] += fsign(d - mean[0][s]);
                if (s == gs) { aA[s] += e; aG[s] += e * (d - mean[0][s]); }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < MAX_SEG; ++s) {
        if (s < h.nseg) {
            float v0 = aA[s], v1 = aG[s], v2 = aS[s];
            for (int o = 16; o; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            if (lane == 0) { red[0][s][warp] = v0; red[1][s][warp] = v1; red[2][s][warp] = v2; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * h.nseg) {
        const int s = threadIdx.x / 3, q = threadIdx.x - 3 * s;
        double v = 0.0;
        for (int k = 0; k < 8; ++k) v += red[q][s][k];
        if (v != 0.0) atomicAdd(&w.racc[(static_cast<long long>(b) * h.nseg + s) * 3 + q], v);
    }
}

__global__ void __launch_bounds__(THREADS) hyb_bwd_apply_kernel(const HybArgs h, const HybBwd w) {
    const int b = blockIdx.y;
    __shared__ float bnd[2][MAX_SEG + 1], mean[2][MAX_SEG], den[2][MAX_SEG], cnt[2][MAX_SEG];
    __shared__ float cA[MAX_SEG], cG[MAX_SEG], cS[MAX_SEG];
    hyb_load_stats(h, b, bnd, mean, den, cnt);
    if (threadIdx.x < h.nseg) {
        const int s = threadIdx.x;
        const double* r = w.racc + (static_cast<long long>(b) * h.nseg + s) * 3;
        const double dn = den[0][s], c = cnt[0][s];
        cA[s] = static_cast<float>(-r[0] / (dn * c));           // mean term, per member
        cG[s] = static_cast<float>(r[1] / (dn * dn) / c);      // coefficient of (sign(d - mu) - Sg / cnt)
        cS[s] = static_cast<float>(r[2] / c);
    }
    __syncthreads();
    const float scale = *w.gout / (static_cast<float>(h.B) * static_cast<float>(h.L));
    const long long start = static_cast<long long>(blockIdx.x) * h.chunk;
    const long long end = min(start + h.chunk, h.L);
    for (long long i = start + threadIdx.x; i < end; i += THREADS) {
        const float d = h.x[0][static_cast<long long>(b) * h.L + i], t = h.x[1][static_cast<long long>(b) * h.L + i];
        float ns = 0.f, nt = 0.f, gsum = 0.f;
        int gs = -1;
        for (int s = 0; s < h.nseg; ++s) {
            if (d >= bnd[0][s] && d <= bnd[0][s + 1]) {
                ns = (d - mean[0][s]) / den[0][s];
                gs = s;
                gsum += cA[s] - cG[s] * (fsign(d - mean[0][s]) - cS[s]);
            }
            if (t >= bnd[1][s] && t <= bnd[1][s + 1]) nt = (t - mean[1][s]) / den[1][s];
        }
        if (gs >= 0) gsum += fsign(ns - nt) / den[0][gs];
        w.grad[static_cast<long long>(b) * h.L + i] = gsum * scale;
    }
}

// feature cosine backward w.r.t. s:  d(1 - mean cos) / d a = -(1 / (B D)) * (v / n_b - cos * a / n_a) / n_a  per column
__global__ void __launch_bounds__(THREADS) featcos_bwd_kernel(const float* s, const float* t, int N, int Ds, int Dt, int D, int B,
                                                              const float* gout, float* grad) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5;
    float st = 0.f, ss = 0.f, tt = 0.f;
    int cs = 0, ct = 0;
    if (c < D) {
        cs = (Ds == D) ? c : min(static_cast<int>(floorf(c * (static_cast<float>(Ds) / D))), Ds - 1);
        ct = (Dt == D) ? c : min(static_cast<int>(floorf(c * (static_cast<float>(Dt) / D))), Dt - 1);
        const float* sp = s + static_cast<long long>(b) * N * Ds + cs;
        const float* tp = t + static_cast<long long>(b) * N * Dt + ct;
        for (int n = warp; n < N; n += 8) {
            const float a = sp[static_cast<long long>(n) * Ds], v = tp[static_cast<long long>(n) * Dt];
            st = fmaf(a, v, st); ss = fmaf(a, a, ss); tt = fmaf(v, v, tt);
        }
    }
    __shared__ float red[3][8][32];
    __shared__ float col[3][32];   // 1 / n_a, 1 / n_b, cos
    red[0][warp][threadIdx.x & 31] = st; red[1][warp][threadIdx.x & 31] = ss; red[2][warp][threadIdx.x & 31] = tt;
    __syncthreads();
    if (warp == 0) {
        float a = 0.f, q = 0.f, r = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w][threadIdx.x]; q += red[1][w][threadIdx.x]; r += red[2][w][threadIdx.x]; }
        const float na = sqrtf(q), nb = sqrtf(r);
        const bool ok = na > 1e-12f && nb > 1e-12f;   // a zero column has no direction: its gradient is left at 0
        col[0][threadIdx.x] = ok ? 1.0f / na : 0.f;
        col[1][threadIdx.x] = ok ? 1.0f / nb : 0.f;
        col[2][threadIdx.x] = ok ? a / (na * nb) : 0.f;
    }
    __syncthreads();
    if (c < D) {
        const float ia = col[0][threadIdx.x & 31], ib = col[1][threadIdx.x & 31], cosv = col[2][threadIdx.x & 31];
        const float k = -(*gout) / (static_cast<float>(B) * static_cast<float>(D));
        const float* sp = s + static_cast<long long>(b) * N * Ds + cs;
        const float* tp = t + static_cast<long long>(b) * N * Dt + ct;
        float* gp = grad + static_cast<long long>(b) * N * Ds + cs;
        for (int n = warp; n < N; n += 8) {
            const float a = sp[static_cast<long long>(n) * Ds], v = tp[static_cast<long long>(n) * Dt];
            gp[static_cast<long long>(n) * Ds] = k * (v * ib - cosv * a * ia) * ia;
        }
    }
}

}  // namespace

// ================================================================== public (internal C++) API
size_t loss_workspace_bytes(int B, int K) {
    const size_t sel = select_ws_bytes(B, K < 1 ? 1 : K);
    const size_t hyb = 4096 + static_cast<size_t>(B) * (2 * 2 * 4 + 3 * 2 * MAX_SEG * 8 + 3 * MAX_SEG * 8) + 2048;
    return (sel > hyb ? sel : hyb) + 1024;
}

int masked_shift_and_scale(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L,
                           float* pred_aligned, float* gt_aligned, void* ws, size_t ws_bytes, cudaStream_t st) {
    return ssi_common(MODE_MASK, pred, gt, mask, nullptr, 1, 0, rows, L, pred_aligned, gt_aligned, nullptr, nullptr,
                      nullptr, 0, 0, ws, ws_bytes, st);
}

int ssi_loss(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L, float* dense_out,
             float* out_scalar, double* partials, void* ws, size_t ws_bytes, cudaStream_t st) {
    return ssi_common(MODE_MASK, pred, gt, mask, nullptr, 1, 0, rows, L, nullptr, nullptr, dense_out, out_scalar,
                      partials, 0, 0, ws, ws_bytes, st);
}

int hdn_loss_dr(int level, const float* pred, const float* gt, const uint8_t* mask, int B, long long L, float* out_scalar,
                double* partials, void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(level >= 1 && level <= 4, "hdn_loss_dr: level=%d unsupported (1..4)", level);
    return ssi_common(MODE_DR, pred, gt, mask, nullptr, (1 << level) - 1, level, B, L, nullptr, nullptr, nullptr,
                      out_scalar, partials, 0, 0, ws, ws_bytes, st);
}

int hdn_loss_ctx(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, long long L, float* out_scalar,
                 double* partials, void* ws, size_t ws_bytes, cudaStream_t st) {
    return ssi_common(MODE_CTX, pred, gt, nullptr, ctx, K, 0, B, L, nullptr, nullptr, nullptr, out_scalar, partials, 0,
                      0, ws, ws_bytes, st);
}

int contexts_dr(int level, const float* gt, const uint8_t* mask, int B, long long L, uint8_t* ctx_out, void* ws,
                size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(level >= 1 && level <= 4, "contexts_dr: level=%d unsupported (1..4)", level);
    DAD_REQUIRE(gt && ctx_out && B > 0 && L > 0, "contexts_dr: bad arguments");
    SelArgs a{};
    a.gt = gt; a.pred = gt; a.mask = mask; a.B = B; a.K = (1 << level) - 1; a.level = level; a.narr = 2; a.L = L;
    a.chunk = pick_chunk(L, B);
    size_t zero_bytes = 0;
    DAD_TRY(carve(a, ws, ws_bytes, &zero_bytes));
    const dim3 grid(static_cast<unsigned>(cdivl(L, a.chunk)), B);
    init_minmax_kernel<<<cdiv(B, 128), 128, 0, st>>>(a.minmax, B);
    minmax_kernel<<<grid, THREADS, 0, st>>>(gt, mask, L, a.chunk, a.minmax);
    contexts_dr_kernel<<<grid, THREADS, 0, st>>>(a, ctx_out);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int ssi_loss_bwd(const float* pred, const float* gt, const uint8_t* mask, int rows, long long L, const float* gout,
                 float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st) {
    return ssi_backward_common(MODE_MASK, pred, gt, mask, nullptr, 1, 0, rows, L, gout, grad_pred, ws, ws_bytes, st);
}

int hdn_loss_dr_bwd(int level, const float* pred, const float* gt, const uint8_t* mask, int B, long long L, const float* gout,
                    float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(level >= 1 && level <= 4, "hdn_loss_dr_bwd: level=%d unsupported (1..4)", level);
    return ssi_backward_common(MODE_DR, pred, gt, mask, nullptr, (1 << level) - 1, level, B, L, gout, grad_pred, ws, ws_bytes, st);
}

int hdn_loss_ctx_bwd(const float* pred, const float* gt, const uint8_t* ctx, int K, int B, long long L, const float* gout,
                     float* grad_pred, void* ws, size_t ws_bytes, cudaStream_t st) {
    return ssi_backward_common(MODE_CTX, pred, gt, nullptr, ctx, K, 0, B, L, gout, grad_pred, ws, ws_bytes, st);
}

int grad_loss_bwd(const float* depth, int B, int H, int W, const float* gout, float* grad_depth, cudaStream_t st) {
    DAD_REQUIRE(depth && gout && grad_depth && B > 0 && H > 0 && W > 0 && B <= 65535, "grad_loss_bwd: bad arguments");
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * H * W * 8, st);
    const dim3 grid(cdiv(W, 32), cdiv(H, 8), B);
    sobel_bwd_kernel<<<grid, THREADS, 0, st>>>(depth, H, W, gout, 1.0 / (static_cast<double>(B) * H * W), grad_depth);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int contexts_dp(int level, const float* gt, const uint8_t* mask, int B, long long L, uint8_t* ctx_out, void* ws,
                size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(level >= 1 && level <= 4, "contexts_dp: level=%d unsupported (1..4)", level);
    DAD_REQUIRE(gt && ctx_out && B > 0 && L > 0, "contexts_dp: bad arguments");
    SelArgs a{};
    a.nq = (1 << (level - 1)) + 1;
    a.gt = gt; a.pred = gt; a.mask = mask; a.B = B; a.K = 2 * a.nq; a.level = level; a.narr = 1; a.L = L;
    a.chunk = pick_chunk(L, B);
    size_t zero_bytes = 0;
    DAD_TRY(carve(a, ws, ws_bytes, &zero_bytes));
    DAD_CHECK_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, st));
    const dim3 grid(static_cast<unsigned>(cdivl(L, a.chunk)), B);
    const size_t sm_hist = static_cast<size_t>(a.K) * 256 * 4 + a.K * 4 + 2 * a.K * 4;
    DAD_REQUIRE(sm_hist <= 48 * 1024, "contexts_dp: histogram does not fit");
    for (int pass = 0; pass < 4; ++pass) {
        sel_hist_kernel<MODE_RANK><<<grid, THREADS, sm_hist, st>>>(a, pass);
        sel_scan_kernel<<<cdiv(a.R, THREADS / 32), THREADS, 0, st>>>(a, pass);
    }
    contexts_dp_kernel<<<grid, THREADS, 0, st>>>(a, ctx_out);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int contexts_ds(int level, const uint8_t* mask, int B, int H, int W, uint8_t* ctx_out, cudaStream_t st) {
    DAD_REQUIRE(level >= 1 && level <= 3, "contexts_ds: level=%d unsupported (1..3: at most 21 contexts)", level);
    DAD_REQUIRE(ctx_out && B > 0 && H > 0 && W > 0, "contexts_ds: bad arguments");
    DAD_REQUIRE(H == W, "contexts_ds: the reference's template masks are size x size with size = W; H=%d != W=%d", H, W);
    const long long L = static_cast<long long>(H) * W;
    const long long want = cdivl(L, THREADS);
    const dim3 grid(static_cast<unsigned>(want < 148 * 4 ? want : 148 * 4), B);
    contexts_ds_kernel<<<grid, THREADS, 0, st>>>(mask, B, H, W, level, ctx_out);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int grad_loss(const float* depth, int B, int H, int W, float* out_scalar, double* partials, void* ws, size_t ws_bytes,
              cudaStream_t st) {
    DAD_REQUIRE(depth && B > 0 && H > 0 && W > 0, "grad_loss: bad arguments");
    DAD_REQUIRE(ws && ws_bytes >= 256, "grad_loss: workspace too small");
    double* acc = reinterpret_cast<double*>(ws);
    DAD_CHECK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    const dim3 grid(cdiv(W, 32), cdiv(H, 64), B);
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * H * W * 4, st, 3);
    sobel_kernel<<<grid, THREADS, 0, st>>>(depth, H, W, acc);
    set_den_kernel<<<1, 1, 0, st>>>(acc, static_cast<double>(B) * H * W);
    ratio_kernel<<<1, 1, 0, st>>>(acc, 0.0, out_scalar, partials, 0);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int feat_cos_loss(const float* s, const float* t, int B, int N, int Ds, int Dt, float* out_scalar, double* partials,
                  void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(s && t && B > 0 && N > 0 && Ds > 0 && Dt > 0, "feat_cos_loss: bad arguments");
    DAD_REQUIRE(ws && ws_bytes >= 256, "feat_cos_loss: workspace too small");
    const int D = Ds < Dt ? Ds : Dt;
    double* acc = reinterpret_cast<double*>(ws);
    DAD_CHECK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    ProfScope prof(PROF_LOSS, 4.0 * B * N * (static_cast<double>(Ds) + Dt), st, 3);
    featcos_kernel<<<dim3(cdiv(D, 32), B), THREADS, 0, st>>>(s, t, N, Ds, Dt, D, acc);
    set_den_kernel<<<1, 1, 0, st>>>(acc, static_cast<double>(B) * D);
    ratio_kernel<<<1, 1, 0, st>>>(acc, 0.0, out_scalar, partials, 1);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int distill_loss(const float* student, const float* teacher, int strategy, int num_segments, int B, long long L,
                 float* out_scalar, double* partials, float* norm_student, float* norm_teacher, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(student && teacher && B > 0 && L > 0, "distill_loss: bad arguments");
    DAD_REQUIRE(ws && ws_bytes >= loss_workspace_bytes(B, 1), "distill_loss: workspace too small");
    if (strategy == 1) {  // global: (d - median) / (mean|d - median| + 1e-6), then mean L1 over all pixels
        DAD_TRY(ssi_common(MODE_MASK, student, teacher, nullptr, nullptr, 1, 0, B, L, norm_student, norm_teacher,
                           nullptr, out_scalar, partials, 1, 1, ws, ws_bytes, st));
        return DAD_OK;
    }
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * L * 8, st, strategy == 0 ? 3 : 8);
    if (strategy == 0) {  // none
        double* acc = reinterpret_cast<double*>(ws);
        DAD_CHECK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
        const long long n = static_cast<long long>(B) * L;
        const int grid = static_cast<int>(cdivl(n, THREADS * 8) < 148 * 8 ? cdivl(n, THREADS * 8) : 148 * 8);
        l1_kernel<<<grid < 1 ? 1 : grid, THREADS, 0, st>>>(student, teacher, n, acc);
        set_den_kernel<<<1, 1, 0, st>>>(acc, static_cast<double>(n));
        ratio_kernel<<<1, 1, 0, st>>>(acc, 0.0, out_scalar, partials, 0);
        DAD_CHECK_LAUNCH();
        return DAD_OK;
    }
    DAD_REQUIRE(strategy == 2, "distill_loss: unknown strategy %d", strategy);
    DAD_REQUIRE(num_segments >= 1 && num_segments <= MAX_SEG, "distill_loss: num_segments=%d unsupported (1..%d)",
                num_segments, MAX_SEG);
    HybArgs h{};
    h.x[0] = student; h.x[1] = teacher; h.B = B; h.nseg = num_segments; h.L = L;
    h.chunk = pick_chunk(L, B);
    h.norm_out[0] = norm_student; h.norm_out[1] = norm_teacher;
    Carver c(ws, ws_bytes);
    h.acc = c.take<double>(2);
    h.segsum = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    h.segcnt = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    h.segmad = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    const size_t zero_bytes = c.used;
    h.minmax = c.take<uint32_t>(static_cast<size_t>(4) * B);
    DAD_CHECK_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, st));
    const dim3 grid(static_cast<unsigned>(cdivl(L, h.chunk)), B);
    init_minmax_kernel<<<cdiv(2 * B, 128), 128, 0, st>>>(h.minmax, 2 * B);
    minmax_kernel<<<grid, THREADS, 0, st>>>(student, nullptr, L, h.chunk, h.minmax);
    minmax_kernel<<<grid, THREADS, 0, st>>>(teacher, nullptr, L, h.chunk, h.minmax + 2 * B);
    const dim3 grid2(grid.x, B, 2);
    hyb_stats_kernel<<<grid2, THREADS, 0, st>>>(h, 0);
    hyb_stats_kernel<<<grid2, THREADS, 0, st>>>(h, 1);
    hyb_final_kernel<<<grid, THREADS, 0, st>>>(h);
    set_den_kernel<<<1, 1, 0, st>>>(h.acc, static_cast<double>(B) * L);
    ratio_kernel<<<1, 1, 0, st>>>(h.acc, 0.0, out_scalar, partials, 0);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int feat_cos_loss_bwd(const float* s, const float* t, int B, int N, int Ds, int Dt, const float* gout, float* grad_s,
                      cudaStream_t st) {
    DAD_REQUIRE(s && t && gout && grad_s && B > 0 && N > 0 && Ds > 0 && Dt > 0, "feat_cos_loss_bwd: bad arguments");
    const int D = Ds < Dt ? Ds : Dt;
    ProfScope prof(PROF_LOSS, 4.0 * B * N * (2.0 * Ds + Dt), st, 2);
    if (Ds != D) DAD_CHECK_CUDA(cudaMemsetAsync(grad_s, 0, static_cast<size_t>(B) * N * Ds * 4, st));  // unselected channels
    featcos_bwd_kernel<<<dim3(cdiv(D, 32), B), THREADS, 0, st>>>(s, t, N, Ds, Dt, D, B, gout, grad_s);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

int distill_loss_bwd(const float* student, const float* teacher, int strategy, int num_segments, int B, long long L,
                     const float* gout, float* grad_student, void* ws, size_t ws_bytes, cudaStream_t st) {
    DAD_REQUIRE(student && teacher && gout && grad_student && B > 0 && L > 0, "distill_loss_bwd: bad arguments");
    DAD_REQUIRE(ws && ws_bytes >= loss_workspace_bytes(B, 1), "distill_loss_bwd: workspace too small");
    if (strategy == 1)  // global: the SSI machinery with mean-over-all statistics and a plain mean
        return ssi_backward_common(MODE_MASK, student, teacher, nullptr, nullptr, 1, 0, B, L, gout, grad_student, ws, ws_bytes,
                                   st, 1);
    if (strategy == 0) {
        const long long n = static_cast<long long>(B) * L;
        const int grid = static_cast<int>(cdivl(n, THREADS * 8) < 148 * 8 ? cdivl(n, THREADS * 8) : 148 * 8);
        ProfScope prof(PROF_LOSS, static_cast<double>(n) * 12, st);
        l1_bwd_kernel<<<grid < 1 ? 1 : grid, THREADS, 0, st>>>(student, teacher, n, gout, grad_student);
        DAD_CHECK_LAUNCH();
        return DAD_OK;
    }
    DAD_REQUIRE(strategy == 2, "distill_loss_bwd: unknown strategy %d", strategy);
    DAD_REQUIRE(num_segments >= 1 && num_segments <= MAX_SEG, "distill_loss_bwd: num_segments=%d unsupported (1..%d)",
                num_segments, MAX_SEG);
    ProfScope prof(PROF_LOSS, static_cast<double>(B) * L * 12, st, 8);
    HybArgs h{};
    h.x[0] = student; h.x[1] = teacher; h.B = B; h.nseg = num_segments; h.L = L;
    h.chunk = pick_chunk(L, B);
    Carver c(ws, ws_bytes);
    h.acc = c.take<double>(2);
    h.segsum = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    h.segcnt = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    h.segmad = c.take<double>(static_cast<size_t>(2) * B * num_segments);
    HybBwd w{};
    w.racc = c.take<double>(static_cast<size_t>(B) * num_segments * 3);
    w.gout = gout; w.grad = grad_student;
    const size_t zero_bytes = c.used;
    h.minmax = c.take<uint32_t>(static_cast<size_t>(4) * B);
    DAD_CHECK_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, st));
    const dim3 grid(static_cast<unsigned>(cdivl(L, h.chunk)), B);
    init_minmax_kernel<<<cdiv(2 * B, 128), 128, 0, st>>>(h.minmax, 2 * B);
    minmax_kernel<<<grid, THREADS, 0, st>>>(student, nullptr, L, h.chunk, h.minmax);
    minmax_kernel<<<grid, THREADS, 0, st>>>(teacher, nullptr, L, h.chunk, h.minmax + 2 * B);
    const dim3 grid2(grid.x, B, 2);
    hyb_stats_kernel<<<grid2, THREADS, 0, st>>>(h, 0);
    hyb_stats_kernel<<<grid2, THREADS, 0, st>>>(h, 1);
    hyb_bwd_reduce_kernel<<<grid, THREADS, 0, st>>>(h, w);
    hyb_bwd_apply_kernel<<<grid, THREADS, 0, st>>>(h, w);
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
