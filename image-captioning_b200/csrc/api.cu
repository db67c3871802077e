// Error reporting and device introspection for libdcap.so.
#include "common.cuh"
#include <string.h>

namespace dcap {

char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace dcap

extern "C" const char *dc_last_error(void) { return dcap::err_buf(); }

extern "C" int dc_device_info(int *sm_count, int *compute_capability) {
    int dev = 0;
    DC_CHECK_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (compute_capability) *compute_capability = major * 10 + minor;
    // keep stream-ordered allocations cached between calls (host-buffer entry points)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return DC_OK;
}
