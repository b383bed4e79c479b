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

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivl(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace dad
