// Multi-head self-attention, head_dim 64, non-causal, no mask (reference
// dinov2_layers/attention.py:49-62; q is pre-scaled: 64^-0.5 is folded into the packed qkv weights).
//
//   attention_bf16 : flash-style single pass (online softmax in fp32), bf16 operands on tensor cores.
//                    Round-1 version issues mma.sync m16n8k16 (legacy tensor path); the tcgen05/TMEM
//                    version replaces it once the GEMM path is validated on hardware.
//   attention_f32  : verification mode, one query per thread, fp32 FFMA, full-precision expf.
//
// qkv layout: [B*N, 3*D] rows = tokens, columns = (3, heads, 64) as produced by the qkv GEMM.
#include "common.h"
#include "elementwise.h"

namespace dad {

namespace {

// ------------------------------------------------------------------ bf16 tensor-core kernel
constexpr int BQ = 64, BKV = 64, HD = 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 64 x 64 bf16 tile, 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ bf16* tile_ptr(bf16* tile, int row, int chunk) {
    return tile + row * HD + ((chunk ^ (row & 7)) << 3);
}

__device__ __forceinline__ void load_tile(bf16* tile, const bf16* src, long long ld, int row0, int nrows_total) {
    // 64 rows x 8 chunks = 512 x 16 B over 128 threads
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 128;
        const int row = idx >> 3, chunk = idx & 7;
        const bool valid = (row0 + row) < nrows_total;
        const bf16* g = src + static_cast<long long>(valid ? row0 + row : 0) * ld + chunk * 8;
        cp_async16(tile_ptr(tile, row, chunk), g, valid);
    }
}

__global__ void __launch_bounds__(128) attention_mma_kernel(const bf16* qkv, bf16* out, int N, int D) {
    __shared__ __align__(128) bf16 sQ[BQ * HD];
    __shared__ __align__(128) bf16 sK[2][BKV * HD];
    __shared__ __align__(128) bf16 sV[2][BKV * HD];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
    const long long ld = 3LL * D;
    const bf16* qb = qkv + static_cast<long long>(b) * N * ld + h * HD;
    const bf16* kb = qb + D;
    const bf16* vb = qb + 2 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (N + BKV - 1) / BKV;

    load_tile(sQ, qb, ld, q0, N);
    load_tile(sK[0], kb, ld, 0, N);
    load_tile(sV[0], vb, ld, 0, N);
    cp_async_commit();
    if (ntiles > 1) {
        load_tile(sK[1], kb, ld, BKV, N);
        load_tile(sV[1], vb, ld, BKV, N);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // Q fragments: 4 k-steps x 4 regs
    uint32_t qa[4][4];
    {
        const int mi = lane >> 3, r = warp * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], tile_ptr(sQ, r, ks * 2 + (mi >> 1)));
    }

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    constexpr float LOG2E = 1.4426950408889634f;

    for (int j = 0; j < ntiles; ++j) {
        const int buf = j & 1;
        if (j > 0) {
            cp_async_wait<1>();  // tile j landed; tile j+1 may still be in flight
            __syncthreads();
        }
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
        {
            const int mi = lane >> 3;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int nt = 0; nt < 8; nt += 2) {
                    uint32_t kf[4];
                    ldsm_x4(kf, tile_ptr(sK[buf], (nt + (mi >> 1)) * 8 + (lane & 7), ks * 2 + (mi & 1)));
                    mma_bf16(s[nt], qa[ks], kf[0], kf[1]);
                    mma_bf16(s[nt + 1], qa[ks], kf[2], kf[3]);
                }
            }
        }
        if (j == ntiles - 1) {
            const int kbase = j * BKV + (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int k0 = kbase + nt * 8;
                if (k0 >= N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                if (k0 + 1 >= N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float sc0 = exp2f((m0 - mn0) * LOG2E), sc1 = exp2f((m1 - mn1) * LOG2E);
        m0 = mn0; m1 = mn1;
        l0 *= sc0; l1 *= sc1;
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= sc0; o[i][1] *= sc0; o[i][2] *= sc1; o[i][3] *= sc1; }
        const float mb0 = m0 * LOG2E, mb1 = m1 * LOG2E;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(fmaf(s[nt][0], LOG2E, -mb0));
            s[nt][1] = exp2f(fmaf(s[nt][1], LOG2E, -mb0));
            s[nt][2] = exp2f(fmaf(s[nt][2], LOG2E, -mb1));
            s[nt][3] = exp2f(fmaf(s[nt][3], LOG2E, -mb1));
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        {
            const int mi = lane >> 3;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t pa[4];
                pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
                pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
                pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
                for (int dt = 0; dt < 8; dt += 2) {
                    uint32_t vf[4];
                    ldsm_x4_t(vf, tile_ptr(sV[buf], kk * 16 + (mi & 1) * 8 + (lane & 7), dt + (mi >> 1)));
                    mma_bf16(o[dt], pa, vf[0], vf[1]);
                    mma_bf16(o[dt + 1], pa, vf[2], vf[3]);
                }
            }
        }
        __syncthreads();  // everyone is done with buffer `buf`
        if (j + 2 < ntiles) {
            load_tile(sK[buf], kb, ld, (j + 2) * BKV, N);
            load_tile(sV[buf], vb, ld, (j + 2) * BKV, N);
        }
        cp_async_commit();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    bf16* ob = out + static_cast<long long>(b) * N * D + h * HD + (lane & 3) * 2;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        if (r0 < N) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * D + dt * 8) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
        if (r1 < N) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * D + dt * 8) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
    }
}

// ------------------------------------------------------------------ fp32 verification kernel
constexpr int FQ = 128, FK = 32;

__global__ void __launch_bounds__(FQ) attention_f32_kernel(const float* qkv, float* out, int N, int D) {
    __shared__ float sK[FK][HD];
    __shared__ float sV[FK][HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * FQ + threadIdx.x;
    const long long ld = 3LL * D;
    const float* base = qkv + static_cast<long long>(b) * N * ld + h * HD;
    float q[HD], acc[HD];
    const bool qok = qi < N;
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = qok ? base[static_cast<long long>(qi) * ld + d] : 0.f; acc[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < N; k0 += FK) {
        __syncthreads();
        for (int i = threadIdx.x; i < FK * HD; i += FQ) {
            const int r = i / HD, d = i - r * HD;
            const bool ok = (k0 + r) < N;
            sK[r][d] = ok ? base[static_cast<long long>(k0 + r) * ld + D + d] : 0.f;
            sV[r][d] = ok ? base[static_cast<long long>(k0 + r) * ld + 2 * D + d] : 0.f;
        }
        __syncthreads();
        const int kn = min(FK, N - k0);
        float sc[FK];
        float cmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < FK; ++r) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], sK[r][d], s);
            sc[r] = (r < kn) ? s : -INFINITY;
            cmax = fmaxf(cmax, sc[r]);
        }
        const float mn = fmaxf(m, cmax);
        const float scale = expf(m - mn);
        m = mn;
        l *= scale;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= scale;
#pragma unroll
        for (int r = 0; r < FK; ++r) {
            const float p = expf(sc[r] - m);
            l += p;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, sV[r][d], acc[d]);
        }
    }
    if (qok) {
        float* o = out + (static_cast<long long>(b) * N + qi) * D + h * HD;
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = acc[d] * inv;
    }
}

}  // namespace

int attention(const void* qkv, void* out, int is_bf16, int B, int N, int heads, cudaStream_t st) {
    DAD_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0, "attention: bad arguments");
    const int D = heads * HD;
    ProfScope prof(PROF_ATTN, 4.0 * B * static_cast<double>(N) * N * D, st);
    if (is_bf16) {
        const dim3 grid(cdiv(N, BQ), heads, B);
        attention_mma_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), N, D);
    } else {
        const dim3 grid(cdiv(N, FQ), heads, B);
        attention_f32_kernel<<<grid, FQ, 0, st>>>(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(out), N, D);
    }
    DAD_CHECK_LAUNCH();
    return DAD_OK;
}

}  // namespace dad
