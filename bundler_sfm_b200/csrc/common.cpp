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
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s): libbsfm_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return BSFM_ERR_NO_DEVICE;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return BSFM_ERR_NO_DEVICE;
    }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library ships sm_100a kernels only", dev, prop.major,
                  prop.minor);
        return BSFM_ERR_NO_DEVICE;
    }
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

}  // extern "C"
