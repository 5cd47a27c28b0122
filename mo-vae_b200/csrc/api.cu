// Library-level entry points: version, error text, device info.
#include <stdarg.h>

#include "common.cuh"

namespace movae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail_cuda(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return MOVAE_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace movae

extern "C" {

int movae_abi_version(void) { return MOVAE_ABI_VERSION; }

const char* movae_last_error(void) { return movae::g_err; }

int movae_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    int n = 0, ma = 0, mi = 0;
    MOVAE_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    MOVAE_CUDA_TRY(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    MOVAE_CUDA_TRY(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return MOVAE_OK;
}

}  // extern "C"
