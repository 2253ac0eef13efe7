// common.cpp -- error reporting / device checks shared by the MATCH and BA entry points.
#include "common.h"
#include <cstring>
#include <cstdlib>

namespace bsfm {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_kernel_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (getenv("BSFM_VERBOSE")) fprintf(stderr, "[bsfm_b200] error: %s\n", g_err);
}
void clear_error() { g_err[0] = 0; }
const char *last_error() { return g_err; }

int require_device()
{
    // the verdict is cached per device: cudaGetDeviceProperties costs a fraction of a millisecond, which is visible in
    // per-pair MatchKeys calls and in a 10 ms bundle adjustment
    static std::atomic<int> verified[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev >= 0 && dev < 64 && verified[dev].load(std::memory_order_relaxed) == 1) return BSFM_OK;
    int ndev = 0;
    e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s): libbsfm_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return BSFM_ERR_NO_DEVICE;
    }
    e = cudaGetDevice(&dev);
    int major = 0, minor = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return BSFM_ERR_NO_DEVICE;
    }
    if (major != 10) {
        set_error("device %d is sm_%d%d; this library ships sm_100a kernels only", dev, major, minor);
        return BSFM_ERR_NO_DEVICE;
    }
    if (dev >= 0 && dev < 64) verified[dev].store(1, std::memory_order_relaxed);
    return BSFM_OK;
}

}  // namespace bsfm

extern "C" {

const char *bsfm_last_error(void) { return bsfm::last_error(); }
const char *bsfm_version(void) { return "bsfm_b200 0.1 (sm_100a)"; }
int64_t bsfm_kernel_launches(void) { return (int64_t) bsfm::g_kernel_launches.load(); }
int bsfm_set_device(int device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        bsfm::set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return BSFM_ERR_CUDA;
    }
    return BSFM_OK;
}
int bsfm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
