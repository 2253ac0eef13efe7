// common.h -- shared host-side helpers of libbsfm_b200.so (error reporting, launch counting).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/bsfm_b200.h"

#include <vector>
namespace bsfm {

void set_error(const char *fmt, ...);
void clear_error();
const char *last_error();
extern std::atomic<long long> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// Fails loudly when no sm_100 device is usable.  Returns BSFM_OK or a negative error.
int require_device();

// dense n x m visibility mask (sba.h: char *vmask) -> CRS: rowptr[n + 1], obs_cam[nvis] ascending per point (common.cpp)
void host_scan_vmask(const char *vmask, int n, int m, std::vector<int> &rowptr, std::vector<int> &obs_cam);
}  // namespace bsfm

#define BSFM_CUDA_TRY(expr)                                                                       \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            bsfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,    \
                            __LINE__);                                                            \
            return BSFM_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define BSFM_CUDA_TRY_PTR(expr)                                                                   \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            bsfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,    \
                            __LINE__);                                                            \
            return nullptr;                                                                       \
        }                                                                                         \
    } while (0)

#define BSFM_KERNEL_CHECK()                                                                       \
    do {                                                                                          \
        bsfm::count_launch();                                                                     \
        BSFM_CUDA_TRY(cudaGetLastError());                                                        \
    } while (0)
