// Micro-benchmark of the attention kernel's softmax inner loop WITHOUT the MMAs: what exp2 rate per SM can a given
// instruction mix / warp count / tile size sustain?  (148 x OCC CTAs of 128 threads; thread = one S row in TMEM.)
//   per tile of TILE columns: tcgen05.ld (32-column pieces) -> exp2 -> bf16 pack -> tcgen05.st (in place), optional
//   tcgen05.wait::st + mbarrier try_wait / arrive per tile (the real kernel's hand-shake with the MMA warp).
// MODE 0: FFMA + MUFU + FADD (row sum) + CVT     (the round-1 kernel's mix)
// MODE 1: FFMA + MUFU + CVT                      (row sum moved to the tensor core)
// MODE 2: FFMA2 (fma.rn.f32x2) + MUFU + CVT
// POLY  = how many of every 8 exponentials use the FMA-pipe polynomial instead of the MUFU (MODE 2 only: packed f32x2)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/softmax_bench tools/softmax_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// packed polynomial 2^x for two elements (x >= -126 assumed clamped by the caller)
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    const uint64_t x = pack2(x0, x1);
    const uint64_t magic = pack2(12582912.0f, 12582912.0f), nmagic = pack2(-12582912.0f, -12582912.0f);
    const uint64_t t = fadd2(x, magic);
    const uint64_t n = fadd2(t, nmagic);                                  // round(x)
    const uint64_t r = ffma2(n, pack2(-1.f, -1.f), x);                    // x - n
    uint64_t p = ffma2(r, pack2(0.0551716648f, 0.0551716648f), pack2(0.2426111251f, 0.2426111251f));
    p = ffma2(p, r, pack2(0.6932609677f, 0.6932609677f));
    p = ffma2(p, r, pack2(0.9999280572f, 0.9999280572f));
    float p0, p1, t0, t1;
    unpack2(p, p0, p1);
    unpack2(t, t0, t1);
    y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

template <int MODE, int POLY>
__device__ __forceinline__ float exp_piece(const uint32_t (&x)[32], float m_ref, uint32_t (&pk)[16]) {
    constexpr float LOG2E = 1.4426950408889634f;
    float s0 = 0.f, s1 = 0.f;
    if (MODE <= 1) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(x[i]), LOG2E, -m_ref));
            const float p1 = ex2(fmaf(__uint_as_float(x[i + 1]), LOG2E, -m_ref));
            if (MODE == 0) { s0 += p0; s1 += p1; }
            pk[i >> 1] = cvt_bf16x2(p0, p1);
        }
    } else {
        const uint64_t sc = pack2(LOG2E, LOG2E), mr = pack2(-m_ref, -m_ref);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float a0, a1, p0, p1;
            unpack2(ffma2(pack2(__uint_as_float(x[i]), __uint_as_float(x[i + 1])), sc, mr), a0, a1);
            // pairs (i, i+1) with ((i >> 1) & 3) < POLY / 2 go to the polynomial: POLY of every 8 elements
            if (((i >> 1) & 3) < POLY / 2) {
                ex2_poly2(a0, a1, p0, p1);
            } else {
                p0 = ex2(a0);
                p1 = ex2(a1);
            }
            pk[i >> 1] = cvt_bf16x2(p0, p1);
        }
    }
    return s0 + s1;
}

// SYNC 0: nothing per tile; 1: tcgen05.wait::st per tile; 2: + mbarrier try_wait (complete barrier) + arrive per tile
template <int MODE, int POLY, int TILE, int SYNC>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int tiles) {
    extern __shared__ uint8_t dyn[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar_done, bar_sink;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_sink)), "r"(128));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_done)) : "memory");  // phase 0 complete
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    {   // defined contents: zeros
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0;
        for (int c = 0; c < 128; c += 16) tmem_st16(base + c, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    float l = 0.f;
    const float m_ref = 0.25f + threadIdx.x * 1e-3f;
    uint32_t va[32], vb[32], pk[16];
    tmem_ld32(base, va);
    for (int j = 0; j < tiles; ++j) {
#pragma unroll
        for (int c = 0; c < TILE; c += 64) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_ld32(base + c + 32, vb);
            l += exp_piece<MODE, POLY>(va, m_ref, pk);
            tmem_st16(base + (c >> 1), pk);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c + 64 < TILE) tmem_ld32(base + c + 64, va);
            l += exp_piece<MODE, POLY>(vb, m_ref, pk);
            tmem_st16(base + (c >> 1) + 16, pk);
        }
        if (SYNC >= 2) {
            while (!mbar_try_wait(&bar_done, 0)) {}
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        tmem_ld32(base, va);   // next tile's first piece (same columns; values differ after the in-place P write - irrelevant)
        if (SYNC >= 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (SYNC >= 2) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_sink)) : "memory");
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(va[0]) + __uint_as_float(pk[3]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
}

template <int MODE, int POLY, int TILE, int SYNC>
void run(int occ, float* out, long long* cyc) {
    const int tiles = 4096 * 64 / TILE;
    const int smem = (occ == 1 ? 200 : occ == 2 ? 100 : occ == 3 ? 70 : 50) * 1024;   // pins the CTAs-per-SM count
    auto kern = k<MODE, POLY, TILE, SYNC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid = 148 * occ;
    kern<<<grid, 128, smem>>>(out, cyc, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, 128, smem>>>(out, cyc, tiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[148 * 4];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double elems_sm = double(occ) * 128 * 4096 * 64;
    printf("mode %d poly %d tile %3d sync %d warps/sched %d: %7.3f ms  %5.2f exp/clk/SM  (%.1f exp/ns/SM)  %s\n", MODE, POLY, TILE, SYNC, occ,
           ms, elems_sm / mx, elems_sm / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 4 * 128 * 4); cudaMalloc(&cyc, 148 * 4 * 8);
    for (int occ : {1, 2, 4}) {
        run<0, 0, 64, 0>(occ, out, cyc);
        run<0, 0, 64, 2>(occ, out, cyc);
        run<0, 0, 128, 2>(occ, out, cyc);
        run<1, 0, 64, 0>(occ, out, cyc);
        run<1, 0, 64, 2>(occ, out, cyc);
        run<1, 0, 128, 2>(occ, out, cyc);
        run<2, 0, 64, 0>(occ, out, cyc);
        run<2, 0, 128, 2>(occ, out, cyc);
        run<2, 2, 64, 0>(occ, out, cyc);
        run<2, 2, 128, 2>(occ, out, cyc);
        run<2, 4, 64, 0>(occ, out, cyc);
        run<2, 4, 128, 2>(occ, out, cyc);
        run<2, 6, 128, 2>(occ, out, cyc);
    }
    return 0;
}
