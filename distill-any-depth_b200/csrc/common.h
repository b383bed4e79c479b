// Shared host-side helpers: error codes, thread-local error string, CUDA checks.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define DAD_OK 0
#define DAD_ERR_INVALID (-1)
#define DAD_ERR_UNSUPPORTED (-2)
#define DAD_ERR_CUDA (-3)
#define DAD_ERR_WORKSPACE (-4)

namespace dad {

typedef __nv_bfloat16 bf16;

int set_error(int code, const char* fmt, ...);
const char* last_error();
int num_sms();
void profile_enable(int on);
int profile_get(int cls, double* ms, double* work, long long* launches);
long long launch_count();
void debug_label(const char* s);  // no-op unless DAD_DEBUG_SYNC is set

#define DAD_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return ::dad::set_error(DAD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                              \
    } while (0)

#define DAD_CHECK_LAUNCH()                                                                            \
    do {                                                                                              \
        cudaError_t _e = cudaGetLastError();                                                          \
        if (_e != cudaSuccess)                                                                        \
            return ::dad::set_error(DAD_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);                      \
    } while (0)

#define DAD_REQUIRE(cond, ...)                                                                        \
    do {                                                                                              \
        if (!(cond)) return ::dad::set_error(DAD_ERR_INVALID, __VA_ARGS__);                           \
    } while (0)

#define DAD_TRY(expr)                                                                                 \
    do {                                                                                              \
        int _r = (expr);                                                                              \
        if (_r != DAD_OK) return _r;                                                                  \
    } while (0)

// ---- launch counting and optional per-kernel-class device timing (bench.py's roofline numbers)
enum { PROF_GEMM_TC = 0, PROF_GEMM_SIMT, PROF_ATTN, PROF_LN, PROF_ELEM, PROF_LOSS, PROF_NCLASS };
// RAII around one (or `launches`) kernel launch(es): always bumps the launch counter; when profiling
// is enabled it also brackets the launch with CUDA events on `st` and books `work`
// (algorithmic FLOPs for tensor-bound classes, algorithmic bytes for HBM-bound ones).
struct ProfScope {
    int idx;
    cudaStream_t st;
    ProfScope(int cls, double work, cudaStream_t stream, int launches = 1);
    ~ProfScope();
};

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may begin (barrier init, TMEM
// allocation, tensor-map prefetch) while its predecessor in the stream drains; it must execute pdl_wait() before it
// touches global memory.  Kernels call pdl_launch_dependents() early so that the NEXT launch can do the same.  Without
// the launch attribute both instructions are no-ops, so kernels stay valid under plain <<<>>> launches.
bool pdl_enabled();  // false when DAD_NO_PDL is set (A/B switch)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivl(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace dad
