// ba_outliers.cu -- the reprojection statistics / outlier pass that BundlerApp::RunSFM_SBA runs after every
// run_sfm (src/Bundle.cpp:659-856), on the arguments run_sfm just used.
//   per observation : dist = || sfm_project_rd(camera, point) - key ||            (Bundle.cpp:721-754, sfm.c:302-380)
//   per camera      : med = kth_element_copy(n, iround(0.8 n), dists)             (Bundle.cpp:761-764; qsort.c:152-204:
//                     the k-th smallest, 0.0 when k >= n), thresh = CLAMP(1.2 * 2.0 * med, min, max) (:768-771),
//                     mean and median (iround(0.5 n)) for the log line (:776-788)
//   outliers        : points with dist > thresh in some camera, except protected (constrained) points (:793-823);
//                     the recorded error is the one of the FIRST camera that flags the point (:809-821)
// Compiled with -fmad=false like the other BA kernels: the projection keeps the reference's operation order.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "common.h"
#include "../../include/bsfm_b200_ba.h"

namespace bsfm {
namespace ba {

__global__ void vmask_count_kernel(const char *, int, int, int *);
__global__ void vmask_fill_kernel(const char *, int, int, const int *, int *, int *);

struct CamCompact { double R[9], t[3], f, k0, k1; };

// one thread per observation (point-major order == the order of `projections`)
__global__ void outlier_dist_kernel(int nvis, const int *obs_cam, const int *obs_pt, const CamCompact *cams, const double *pts,
                                    const double *proj, int undistort, double *dist)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nvis) return;
    const CamCompact &C = cams[obs_cam[o]];
    const double *b = pts + 3 * (size_t) obs_pt[o];
    // sfm_project_rd with explicit camera centres (Bundle.cpp:741-744 passes `true`), sfm.c:326-338
    const double b0 = b[0] - C.t[0], b1 = b[1] - C.t[1], b2 = b[2] - C.t[2];
    const double c0 = C.R[0] * b0 + C.R[1] * b1 + C.R[2] * b2;
    const double c1 = C.R[3] * b0 + C.R[4] * b1 + C.R[5] * b2;
    const double c2 = C.R[6] * b0 + C.R[7] * b1 + C.R[8] * b2;
    double p0 = -c0 * C.f / c2;
    double p1 = -c1 * C.f / c2;
    if (undistort) {   // sfm.c:362-377
        const double rsq = (p0 * p0 + p1 * p1) / (C.f * C.f);
        const double factor = 1.0 + C.k0 * rsq + C.k1 * rsq * rsq;
        p0 *= factor;
        p1 *= factor;
    }
    const double dx = p0 - proj[2 * (size_t) o], dy = p1 - proj[2 * (size_t) o + 1];
    dist[o] = sqrt(dx * dx + dy * dy);     // Bundle.cpp:753-756
}

__global__ void outlier_iota_kernel(int *v, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) v[i] = i;
}
__global__ void outlier_gather_kernel(int count, const int *idx, const double *src, double *dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[idx[i]];
}
// cam_ptr[j] = first position of camera j in the camera-sorted key array
__global__ void outlier_cam_ptr_kernel(const int *sorted_cam, int nvis, int m, int *cam_ptr)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > m) return;
    int lo = 0, hi = nvis;     // first index with sorted_cam[idx] >= j
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_cam[mid] < j) lo = mid + 1; else hi = mid; }
    cam_ptr[j] = lo;
}

__device__ __forceinline__ int iround_dev(double x) { return (x < 0.0) ? (int) (x - 0.5) : (int) (x + 0.5); }   // util.c:75-81

// one warp per camera: mean (sum in ascending point order), k-th elements from the sorted distances, threshold
__global__ void outlier_cam_stats_kernel(int m, const int *cam_ptr, const double *dist_cam, const double *dist_sorted,
                                         double min_thresh, double max_thresh, double *stats)
{
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= m) return;
    const int s = cam_ptr[j], n = cam_ptr[j + 1] - s;
    double sum = 0.0;
    for (int q = lane; q < n; q += 32) sum += dist_cam[s + q];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane != 0) return;
    const int k80 = iround_dev(0.8 * n), k50 = iround_dev(0.5 * n);
    const double med80 = (k80 >= n) ? 0.0 : dist_sorted[s + k80];      // kth_element: "k should be < n" -> 0.0 (qsort.c:192-194)
    const double med50 = (k50 >= n) ? 0.0 : dist_sorted[s + k50];
    double thresh = 1.2 * 2.0 * med80;                                   // NUM_STDDEV 2.0 (Bundle.cpp:767-768)
    thresh = (thresh < min_thresh) ? min_thresh : ((thresh > max_thresh) ? max_thresh : thresh);   // CLAMP (defines.h)
    double *o = stats + 5 * (size_t) j;
    o[0] = (double) n; o[1] = sum / n; o[2] = med50; o[3] = med80; o[4] = thresh;
}

// first (lowest-index) camera that flags every point, then that camera's distance
__global__ void outlier_flag_kernel(int nvis, const int *obs_cam, const int *obs_pt, const double *dist, const double *stats,
                                    const char *pt_protected, int *first_cam)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nvis) return;
    const int i = obs_pt[o], j = obs_cam[o];
    if (pt_protected && pt_protected[i]) return;                          // Bundle.cpp:801-806
    if (dist[o] > stats[5 * (size_t) j + 4]) atomicMin(&first_cam[i], j);
}
__global__ void outlier_err_kernel(int nvis, const int *obs_cam, const int *obs_pt, const double *dist, const int *first_cam, double *err)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nvis) return;
    const int i = obs_pt[o];
    if (first_cam[i] == obs_cam[o]) err[i] = dist[o];
}

namespace {
struct DevBuf {
    std::vector<void *> ptrs;
    ~DevBuf() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> cudaError_t alloc(T **out, size_t count)
    {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) { ptrs.push_back(p); *out = (T *) p; }
        return e;
    }
};
}  // namespace

}  // namespace ba
}  // namespace bsfm

using namespace bsfm;
using namespace bsfm::ba;

extern "C" int bsfm_reprojection_outliers(int num_pts, int num_cameras, const char *vmask, const double *projections,
                                          const bsfm_camera_params_t *cams, const bsfm_v3_t *pts, int estimate_distortion,
                                          double min_proj_error_threshold, double max_proj_error_threshold,
                                          const char *pt_protected, double *cam_stats, double *obs_dist,
                                          int32_t *outliers, double *outlier_errors, int cap, double *global_mean)
{
    clear_error();
    { int rc = require_device(); if (rc != BSFM_OK) return rc; }
    const int n = num_pts, m = num_cameras;
    if (n <= 0 || m <= 0 || !vmask || !projections || !cams || !pts || cap < 0 || (cap > 0 && (!outliers || !outlier_errors))) {
        set_error("bsfm_reprojection_outliers: bad arguments");
        return BSFM_ERR_ARG;
    }
    for (int j = 0; j < m; j++)
        if (cams[j].known_intrinsics || cams[j].fisheye) { set_error("bsfm_reprojection_outliers: known_intrinsics / fisheye cameras are outside the GPU path"); return BSFM_ERR_UNSUPPORTED; }

    DevBuf D;
    char *d_vmask, *d_prot = nullptr;
    int *d_rowcnt, *d_rowptr, *d_obs_cam, *d_obs_pt;
    BSFM_CUDA_TRY(D.alloc(&d_vmask, (size_t) n * m));
    BSFM_CUDA_TRY(D.alloc(&d_rowcnt, (size_t) n + 1));
    BSFM_CUDA_TRY(D.alloc(&d_rowptr, (size_t) n + 1));
    BSFM_CUDA_TRY(cudaMemcpy(d_vmask, vmask, (size_t) n * m, cudaMemcpyDefault));
    BSFM_CUDA_TRY(cudaMemset(d_rowcnt, 0, ((size_t) n + 1) * sizeof(int)));
    vmask_count_kernel<<<(n * 32 + 255) / 256, 256>>>(d_vmask, n, m, d_rowcnt);
    BSFM_KERNEL_CHECK();
    void *d_tmp = nullptr; size_t tmp_bytes = 0;
    auto ensure_tmp = [&](size_t need) -> cudaError_t {
        if (need <= tmp_bytes) return cudaSuccess;
        char *q; cudaError_t e = D.alloc(&q, need); if (e == cudaSuccess) { d_tmp = q; tmp_bytes = need; } return e;
    };
    {
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, d_rowcnt, d_rowptr, n + 1);
        BSFM_CUDA_TRY(ensure_tmp(need));
        BSFM_CUDA_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_rowcnt, d_rowptr, n + 1));
        count_launch(2);
    }
    int nvis = 0;
    BSFM_CUDA_TRY(cudaMemcpy(&nvis, d_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost));
    if (nvis <= 0) { set_error("bsfm_reprojection_outliers: empty visibility mask"); return BSFM_ERR_ARG; }
    BSFM_CUDA_TRY(D.alloc(&d_obs_cam, (size_t) nvis));
    BSFM_CUDA_TRY(D.alloc(&d_obs_pt, (size_t) nvis));
    vmask_fill_kernel<<<(n * 32 + 255) / 256, 256>>>(d_vmask, n, m, d_rowptr, d_obs_cam, d_obs_pt);
    BSFM_KERNEL_CHECK();

    // cameras (what the loop reads: R, t = centre, f, k), points, measurements
    std::vector<CamCompact> hc(m);
    for (int j = 0; j < m; j++) {
        memcpy(hc[j].R, cams[j].R, sizeof hc[j].R); memcpy(hc[j].t, cams[j].t, sizeof hc[j].t);
        hc[j].f = cams[j].f; hc[j].k0 = cams[j].k[0]; hc[j].k1 = cams[j].k[1];
    }
    CamCompact *d_cams; double *d_pts, *d_proj, *d_dist, *d_dist_cam, *d_dist_sorted, *d_stats, *d_err;
    int *d_iota, *d_cam_obs, *d_cam_sorted, *d_cam_ptr, *d_first;
    BSFM_CUDA_TRY(D.alloc(&d_cams, (size_t) m));
    BSFM_CUDA_TRY(D.alloc(&d_pts, (size_t) n * 3));
    BSFM_CUDA_TRY(D.alloc(&d_proj, (size_t) nvis * 2));
    BSFM_CUDA_TRY(D.alloc(&d_dist, (size_t) nvis)); BSFM_CUDA_TRY(D.alloc(&d_dist_cam, (size_t) nvis)); BSFM_CUDA_TRY(D.alloc(&d_dist_sorted, (size_t) nvis));
    BSFM_CUDA_TRY(D.alloc(&d_stats, (size_t) m * 5)); BSFM_CUDA_TRY(D.alloc(&d_err, (size_t) n));
    BSFM_CUDA_TRY(D.alloc(&d_iota, (size_t) nvis)); BSFM_CUDA_TRY(D.alloc(&d_cam_obs, (size_t) nvis)); BSFM_CUDA_TRY(D.alloc(&d_cam_sorted, (size_t) nvis));
    BSFM_CUDA_TRY(D.alloc(&d_cam_ptr, (size_t) m + 1)); BSFM_CUDA_TRY(D.alloc(&d_first, (size_t) n));
    BSFM_CUDA_TRY(cudaMemcpy(d_cams, hc.data(), (size_t) m * sizeof(CamCompact), cudaMemcpyHostToDevice));
    BSFM_CUDA_TRY(cudaMemcpy(d_pts, pts, (size_t) n * 3 * sizeof(double), cudaMemcpyDefault));
    BSFM_CUDA_TRY(cudaMemcpy(d_proj, projections, (size_t) nvis * 2 * sizeof(double), cudaMemcpyDefault));
    if (pt_protected) {
        BSFM_CUDA_TRY(D.alloc(&d_prot, (size_t) n));
        BSFM_CUDA_TRY(cudaMemcpy(d_prot, pt_protected, (size_t) n, cudaMemcpyDefault));
    }
    const int gb = (nvis + 255) / 256;
    outlier_dist_kernel<<<gb, 256>>>(nvis, d_obs_cam, d_obs_pt, d_cams, d_pts, d_proj, estimate_distortion ? 1 : 0, d_dist);
    BSFM_KERNEL_CHECK();

    // camera-major view (stable sort by camera keeps ascending point order), per-camera sorted distances
    outlier_iota_kernel<<<gb, 256>>>(d_iota, nvis);
    BSFM_KERNEL_CHECK();
    {
        int bits = 1; while ((1LL << bits) < m) bits++;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, d_obs_cam, d_cam_sorted, d_iota, d_cam_obs, nvis, 0, bits);
        BSFM_CUDA_TRY(ensure_tmp(need));
        BSFM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_obs_cam, d_cam_sorted, d_iota, d_cam_obs, nvis, 0, bits));
        count_launch(3);
    }
    outlier_cam_ptr_kernel<<<(m + 1 + 255) / 256, 256>>>(d_cam_sorted, nvis, m, d_cam_ptr);
    BSFM_KERNEL_CHECK();
    outlier_gather_kernel<<<gb, 256>>>(nvis, d_cam_obs, d_dist, d_dist_cam);
    BSFM_KERNEL_CHECK();
    {
        size_t need = 0;
        cub::DeviceSegmentedSort::SortKeys(nullptr, need, d_dist_cam, d_dist_sorted, nvis, m, d_cam_ptr, d_cam_ptr + 1);
        BSFM_CUDA_TRY(ensure_tmp(need));
        BSFM_CUDA_TRY(cub::DeviceSegmentedSort::SortKeys(d_tmp, need, d_dist_cam, d_dist_sorted, nvis, m, d_cam_ptr, d_cam_ptr + 1));
        count_launch(3);
    }
    outlier_cam_stats_kernel<<<(m * 32 + 255) / 256, 256>>>(m, d_cam_ptr, d_dist_cam, d_dist_sorted, min_proj_error_threshold, max_proj_error_threshold, d_stats);
    BSFM_KERNEL_CHECK();
    BSFM_CUDA_TRY(cudaMemset(d_first, 0x7f, (size_t) n * sizeof(int)));      // 0x7f7f7f7f: larger than any camera index
    outlier_flag_kernel<<<gb, 256>>>(nvis, d_obs_cam, d_obs_pt, d_dist, d_stats, d_prot, d_first);
    BSFM_KERNEL_CHECK();
    outlier_err_kernel<<<gb, 256>>>(nvis, d_obs_cam, d_obs_pt, d_dist, d_first, d_err);
    BSFM_KERNEL_CHECK();

    // results
    std::vector<double> stats((size_t) m * 5);
    BSFM_CUDA_TRY(cudaMemcpy(stats.data(), d_stats, stats.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (cam_stats) memcpy(cam_stats, stats.data(), stats.size() * sizeof(double));
    if (obs_dist) BSFM_CUDA_TRY(cudaMemcpy(obs_dist, d_dist, (size_t) nvis * sizeof(double), cudaMemcpyDefault));
    if (global_mean) {      // Bundle.cpp:790-791, 852-856: sum of the per-camera sums / number of observations
        double tot = 0.0; double cnt = 0.0;
        for (int j = 0; j < m; j++) if (stats[5 * (size_t) j] > 0) { tot += stats[5 * (size_t) j + 1] * stats[5 * (size_t) j]; cnt += stats[5 * (size_t) j]; }
        *global_mean = tot / cnt;
    }
    std::vector<int> first(n);
    std::vector<double> err(n);
    BSFM_CUDA_TRY(cudaMemcpy(first.data(), d_first, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost));
    BSFM_CUDA_TRY(cudaMemcpy(err.data(), d_err, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<std::pair<int, int>> found;      // (first flagging camera, point)
    for (int i = 0; i < n; i++) if (first[i] < m) found.push_back(std::make_pair(first[i], i));
    std::sort(found.begin(), found.end());
    const int total = (int) found.size();
    for (int q = 0; q < total && q < cap; q++) { outliers[q] = found[q].second; outlier_errors[q] = err[found[q].second]; }
    return total;
}
