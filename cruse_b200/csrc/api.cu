// api.cu -- version / error / device queries of the C ABI.
#include "common.cuh"

namespace cruse {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    // read-only per-device value, cached per thread (no global mutable state across devices)
    static thread_local int cached_dev = -1, cached_n = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached_dev = dev;
        cached_n = n;
    }
    return cached_n;
}

}  // namespace cruse

extern "C" {

int cruse_version(void) { return 100; }  // 0.1.0

const char* cruse_last_error(void) { return cruse::g_err; }

int cruse_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cruse::set_error("cudaGetDevice failed (no CUDA device?)");
        return -2;
    }
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cruse::set_error("cudaDeviceGetAttribute failed");
        return -2;
    }
    return n;
}

}  // extern "C"
