// fp32 verification engine: the same GEMM / implicit-GEMM-convolution contract as gemm_tc.cu,
// computed with plain FFMA in fp32 (no tensor cores, no TF32) so the whole forward can be checked
// against the fp32 reference to <=1e-4.  Classic 64x64x16 shared-memory tiling, 4x4 micro-tiles.
#include "epilogue.cuh"
#include "gemm.h"

namespace dad {

namespace {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtArgs {
    Epilogue epi;
    const float* A;
    const float* Wt;
    int M, N, K, Kp;
    long long lda;
    int conv, taps, C, cpad;  // cpad = cchunks * 64 (per-tap padded channels)
    int B, H, W;
    long long ldp;
};

__device__ __forceinline__ float load_a(const SimtArgs& g, long long row, int k, int b, int y, int x) {
    if (!g.conv) return (k < g.K) ? g.A[row * g.lda + k] : 0.f;
    const int tap = k / g.cpad;
    const int c = k - tap * g.cpad;
    if (c >= g.C) return 0.f;
    int yy = y, xx = x;
    if (g.taps == 9) {
        const int dy = tap / 3;
        yy = y + dy - 1;
        xx = x + (tap - dy * 3) - 1;
    }
    if (yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) return 0.f;
    return g.A[((static_cast<long long>(b) * g.H + yy) * g.W + xx) * g.ldp + c];
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs g) {
    __shared__ float sA[SBK][SBM + 4];
    __shared__ float sB[SBK][SBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * SBM;
    const int n0 = blockIdx.x * SBN;

    // each thread loads 4 A and 4 B elements per k-step: row = tid / 4, k = (tid % 4) * 4 + i
    const int lrow = tid >> 2;
    const int lk = (tid & 3) * 4;
    const long long arow = m0 + lrow;
    int ab = 0, ay = 0, ax = 0;
    if (g.conv) {
        const long long hw = static_cast<long long>(g.H) * g.W;
        ab = static_cast<int>(arow / hw);
        const int rem = static_cast<int>(arow - ab * hw);
        ay = rem / g.W;
        ax = rem - ay * g.W;
    }
    const bool arow_ok = arow < g.M;
    const int brow = n0 + lrow;
    const bool brow_ok = brow < g.N;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < g.Kp; k0 += SBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + lk + i;
            sA[lk + i][lrow] = (arow_ok && k < g.Kp) ? load_a(g, arow, k, ab, ay, ax) : 0.f;
            sB[lk + i][lrow] = (brow_ok && k < g.Kp) ? g.Wt[static_cast<long long>(brow) * g.Kp + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int col = n0 + tx * 4;
    if (col >= g.N) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = m0 + ty * 4 + i;
        if (row < g.M) {
            float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
            epilogue_store4(g.epi, g.N, row, row, col, v);
        }
    }
}

}  // namespace

int gemm_simt(const GemmProblem& p, cudaStream_t stream) {
    DAD_REQUIRE(p.A && p.Wt && p.N > 0 && p.N % 4 == 0, "gemm_simt: bad operands (N=%d)", p.N);
    DAD_REQUIRE(p.epi.head_out == nullptr, "gemm_simt: fused head epilogue is a tensor-core-path feature");
    SimtArgs a{};
    a.epi = p.epi;
    a.A = reinterpret_cast<const float*>(p.A);
    a.Wt = reinterpret_cast<const float*>(p.Wt);
    a.N = p.N;
    a.Kp = p.Kp;
    a.conv = p.conv;
    if (p.conv) {
        DAD_REQUIRE(p.taps == 1 || p.taps == 9, "gemm_simt: taps must be 1 or 9");
        a.taps = p.taps;
        a.C = p.C;
        a.cpad = cdiv(p.C, 64) * 64;
        DAD_REQUIRE(p.Kp == p.taps * a.cpad, "gemm_simt: conv weights must be packed to Kp=%d (got %d)",
                    p.taps * a.cpad, p.Kp);
        a.B = p.B; a.H = p.H; a.W = p.W;
        a.ldp = p.ldp;
        a.M = p.B * p.H * p.W;
    } else {
        DAD_REQUIRE(p.M > 0 && p.K > 0 && p.Kp >= p.K, "gemm_simt: bad linear dims");
        a.M = p.M;
        a.K = p.K;
        a.lda = p.lda;
    }
    dim3 grid(cdiv(p.N, SBN), cdiv(a.M, SBM));
    DAD_REQUIRE(grid.y <= 65535u * 64u, "gemm_simt: M too large");
    // grid.y limit is 65535; fold larger M by looping launches
    const int max_rows = 65535 * SBM;
    if (a.M <= max_rows) {
        ProfScope prof(PROF_GEMM_SIMT, 2.0 * a.M * p.N * (p.conv ? static_cast<double>(p.taps) * p.C : p.K), stream);
        gemm_simt_kernel<<<grid, 256, 0, stream>>>(a);
        DAD_CHECK_LAUNCH();
        return DAD_OK;
    }
    return set_error(DAD_ERR_UNSUPPORTED, "gemm_simt: M=%d exceeds %d rows (verification mode is for small batches)",
                     a.M, max_rows);
}

}  // namespace dad
