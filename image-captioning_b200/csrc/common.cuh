// Shared host/device helpers for libdcap.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dcap.h"

namespace dcap {

// thread-local error message (dc_last_error)
char *err_buf();
int set_error(int code, const char *fmt, ...);

#define DC_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::dcap::set_error(DC_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__,  \
                                     __LINE__, cudaGetErrorString(_e));                       \
    } while (0)

#define DC_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) return ::dcap::set_error(DC_ERR_INVALID, __VA_ARGS__);                   \
    } while (0)

#define DC_CHECK_LAUNCH() DC_CHECK_CUDA(cudaGetLastError())

inline int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace dcap
