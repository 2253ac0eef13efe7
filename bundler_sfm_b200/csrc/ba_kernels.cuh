// ba_kernels.cuh -- device data model of the BA hot path.
//
// All arithmetic is fp64.  Parameter vector p = (a_1..a_m, b_1..b_n): camera j has `cnp` values
// [c(3), w(3), f*f_scale (if est_focal), k1*k_scale, k2*k_scale (if undistort)], point i has 3
// (lib/sfm-driver/sfm.c:652-703).  Observations are stored point-major (the order of `projections`
// / idxij.val, lib/sba-1.5/sba_levmar.c:652-663) with a camera-major permutation beside it.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace bsfm {
namespace ba {

constexpr int MAX_CNP = 9;
constexpr int PNP = 3;
constexpr int MNP = 2;
constexpr int SCHUR_CHUNK = 64;        // tuples per partial-sum warp
constexpr int SCHUR_PART_STRIDE = 96;  // doubles per partial (81 block entries + 9 E + pad)

struct Model {
    int cnp, est_focal, undistort, explicit_centers;
    int focal_idx;   // 6 or -1
    int k_idx;       // 7, 6 or -1
    double f_scale, k_scale;
};

// scalars produced on the device, copied to pinned host memory once per LM phase
struct Scalars {
    double e_L2;          // sum e^2 of the last residual evaluation
    double dp_L2;         // sum dp^2
    double dL;            // sum dp (mu dp + J^T e)
    double max_pct;       // stop-8 statistic (sba_levmar.c:1552-1561)
    double eab_inf;       // ||J^T e||_inf
    double p_L2;          // sum p^2
    double max_diag;      // max diagonal of U (j >= mcon) and V
    double penalty;       // constraint penalty at p (sba_levmar.c:808-842)
    int singular_v;       // some V*_i could not be inverted
    int chol_fail;        // dense Cholesky hit a non-positive pivot
    int nonfinite;        // residual evaluation produced NaN/Inf
    int pad;
};

struct Problem {
    int n, m, mcon, nvis, nvars, Sdim;
    int nlm;              // LM unknowns: nvars (motion+structure) or m*cnp (motion only, sba_mot_levmar_x: points stay fixed)
    Model M;
    // structure (device)
    const int *rowptr;    // n+1, point-major CRS (idxij.rowptr)
    const int *obs_cam;   // nvis (idxij.colidx)
    const int *obs_pt;    // nvis
    const int *cam_ptr;   // m+1
    const int *cam_obs;   // nvis, observation ids grouped by camera, ascending point
    // Schur structure: upper blocks (j<=k) in ascending (j,k); tuples (obs of j, obs of k) per block
    int nblocks;
    const uint32_t *blk_key;   // j*m + k
    const int *blk_start;      // nblocks+1
    const int4 *tuples;        // (obs_a, obs_b, point, 0)
    int nchunks;
    const int *chunk_off;      // nblocks+1: first chunk of every block
    double *schur_part;        // nchunks * SCHUR_PART_STRIDE
    // data
    const double *x;       // 2*nvis measurements
    const double *R_init;  // m*9
    const double *f_fixed; // m
    // constraints (device copies; null when unused)
    const char *cam_constrained;    // m*cnp
    const double *cam_constraints;  // m*cnp
    const double *cam_weights;      // m*cnp
    const char *pt_constrained;     // n
    const double *pt_constraints;   // n*3
    const double *pt_weights;       // n
    // work arrays
    double *camR;     // m*4*9 : R(w), R(w + d e_0), R(w + d e_1), R(w + d e_2)
    double *jacA;     // nvis*2*cnp
    double *jacB;     // nvis*6
    double *W;        // nvis*cnp*3
    double *u_part;   // m*nseg*54 partial U/ea sums
    double *U;        // m*cnp*cnp (undamped)
    double *V;        // n*9 (undamped, full symmetric)
    double *Vinv;     // n*9 ((V+mu I)^-1, full symmetric)
    double *eab;      // nvars
    double *S;        // Sdim*Sdim
    double *E;        // Sdim (+ solution workspace)
    double *dp;       // nvars
    double *partial;  // reduction scratch
    unsigned int *ticket;  // last-block counters
    Scalars *sc;      // device scalars
    const double *mu; // damping term of the current try (device scalar, written by the host)
};

// launch with the programmatic-stream-serialization attribute (PDL): the launch latency of a kernel overlaps
// the tail of its predecessor; every kernel launched this way calls cudaGridDependencySynchronize() before it
// reads anything its predecessors wrote
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace ba
}  // namespace bsfm
