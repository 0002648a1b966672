// wae_lib.cu -- library core: error state, device check, launch accounting.
#include "wae_common.cuh"

namespace wae {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int require_sm100() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = 0;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(WAE_ERR_DEVICE, "no CUDA device: %s (libwae_b200 has no CPU fallback)",
                         cudaGetErrorString(e));
    }
    if (dev == cached_dev) return cached_rc;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(WAE_ERR_DEVICE, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    }
    cached_dev = dev;
    cached_rc = (major == 10) ? WAE_OK
                              : set_error(WAE_ERR_DEVICE,
                                          "device %d is sm_%d0; libwae_b200 kernels are sm_100a only", dev,
                                          major);
    return cached_rc;
}

}  // namespace wae

extern "C" {

int wae_version(void) { return 100; }

const char* wae_last_error(void) { return wae::last_error_buf(); }

int64_t wae_launch_count(void) { return wae::g_launches.load(std::memory_order_relaxed); }

int wae_device_check(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return wae::set_error(WAE_ERR_DEVICE, "no CUDA device visible (%s); libwae_b200 has no CPU fallback",
                              e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return wae::set_error(WAE_ERR_ARG, "device %d out of range [0,%d)", device, n);
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) return wae::set_error(WAE_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10)
        return wae::set_error(WAE_ERR_DEVICE, "device %d is sm_%d0; libwae_b200 kernels are sm_100a only", device,
                              major);
    return WAE_OK;
}

}  // extern "C"
