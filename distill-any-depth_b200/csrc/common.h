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

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivl(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace dad
