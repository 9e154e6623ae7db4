// jg_api.cu -- library-level entry points: version, error string, device info, workspace size.
#include <stdarg.h>

#include "jg_common.cuh"

namespace jg {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace jg

extern "C" {

int jg_version(void) { return 100; }   // 0.1.0

const char *jg_last_error(void) { return jg::g_error; }

int jg_device_info(int *sms, int *major, int *minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        jg::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return -1;
    }
    int a = 0, b = 0, c = 0;
    cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev);
    if (sms) *sms = a;
    if (major) *major = b;
    if (minor) *minor = c;
    return 0;
}

// Scratch for any call on n events: either one look-back descriptor per scan tile, or (mask select, arg
// compaction) two int64 arrays with one entry per 512-event tile plus the descriptors of the scan over them.
size_t jg_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    size_t tiles = static_cast<size_t>(n / 512 + 4);
    return 3 * (tiles * 8 + 256) + 8192;
}

}  // extern "C"
