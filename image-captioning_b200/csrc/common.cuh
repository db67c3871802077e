// Shared host/device helpers for libdcap.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

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

// cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs through this library must opt every kernel
// into its large dynamic shared memory on each of them.  `seen` is one static per call site; f() is idempotent, so
// two threads racing on the first call may both run it.
template <typename F>
inline cudaError_t once_per_device(std::atomic<unsigned long long> &seen, F &&f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) seen.fetch_or(bit, std::memory_order_release);
    return e;
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may become resident while
// the previous kernel of the stream is still draining; it must execute pdl_wait() before it touches
// anything that kernel wrote.  Everything before the wait (barrier init, TMEM allocation, descriptor
// prefetch) overlaps the predecessor's tail.  Both calls are no-ops under a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();          // api.cu: true inside a PdlScope unless DCAP_PDL=0
// Enables PDL launches on this thread for its lifetime.  Used by the training step (eager launches of >100
// short dependent kernels: -2 % per step); NOT by the CUDA-graph-captured greedy loop, where programmatic
// edges measured slower (+2 %).
struct PdlScope {
    bool prev;
    PdlScope();
    ~PdlScope();
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace dcap
