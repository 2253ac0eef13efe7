// common.cpp -- error reporting / device checks shared by the MATCH and BA entry points.
#include "common.h"
#include <vector>
#include <thread>
#include <cstdint>
#include <algorithm>
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

namespace bsfm {
// The reference interface hands over a DENSE n x m visibility mask (sba.h: char *vmask): 500 MB at config 3 for 3M set bytes.
// Large host-resident masks are compressed on the host instead of being uploaded: a few threads scan disjoint point ranges
// (64 bytes per test, the mask is > 99 % zero) into the CRS form sba builds anyway (sba_levmar.c:652-663: row pointers per point,
// camera index per observation, ascending), and only that travels: 4 (n + 1) + 4 nvis bytes.
void host_scan_vmask(const char *vmask, int n, int m, std::vector<int> &rowptr, std::vector<int> &obs_cam)
{
    unsigned hw = std::thread::hardware_concurrency();
    const char *env = getenv("BSFM_BA_MASK_THREADS");
    int T = env ? atoi(env) : (int) std::min(16u, std::max(1u, hw));
    T = std::max(1, std::min(T, n));
    std::vector<std::vector<int>> cams(T);
    rowptr.assign((size_t) n + 1, 0);
    auto work = [&](int t) {
        const int i0 = (int) ((long long) n * t / T), i1 = (int) ((long long) n * (t + 1) / T);
        std::vector<int> &out = cams[t];
        for (int i = i0; i < i1; i++) {
            const char *row = vmask + (size_t) i * m;
            int cnt = 0, j = 0;
            for (; j + 64 <= m; j += 64) {           // 64 bytes per branch: the mask is > 99 % zero
                uint64_t w[8];
                memcpy(w, row + j, 64);
                if (!(w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7])) continue;
                for (int u = 0; u < 64; u++)
                    if (row[j + u]) { out.push_back(j + u); cnt++; }
            }
            for (; j < m; j++)
                if (row[j]) { out.push_back(j); cnt++; }
            rowptr[(size_t) i + 1] = cnt;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    for (int i = 0; i < n; i++) rowptr[(size_t) i + 1] += rowptr[i];
    obs_cam.resize((size_t) rowptr[n]);
    size_t pos = 0;
    for (int t = 0; t < T; t++) {
        if (!cams[t].empty()) memcpy(obs_cam.data() + pos, cams[t].data(), cams[t].size() * sizeof(int));
        pos += cams[t].size();
    }
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
/* dev / test hook: the host-side mask compression on plain host arrays; returns nvis, or -1 when obs_cam (cap entries) is too small */
int bsfm_debug_scan_vmask(const char *vmask, int n, int m, int *rowptr, int *obs_cam, int cap)
{
    std::vector<int> rp, oc;
    bsfm::host_scan_vmask(vmask, n, m, rp, oc);
    if ((int) oc.size() > cap) return -1;
    memcpy(rowptr, rp.data(), rp.size() * sizeof(int));
    if (!oc.empty()) memcpy(obs_cam, oc.data(), oc.size() * sizeof(int));
    return (int) oc.size();
}
int bsfm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
