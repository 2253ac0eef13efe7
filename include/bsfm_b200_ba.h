/* bsfm_b200_ba.h -- C ABI of the BA hot path of libbsfm_b200.so (sparse Levenberg-Marquardt bundle
 * adjustment on one B200).  Replaces, behind the reference's own call boundary:
 *
 *   run_sfm(...)                 lib/sfm-driver/sfm.h:68-86,  lib/sfm-driver/sfm.c:592-1003
 *   sba_motstr_levmar_x(...)     lib/sba-1.5/sba.h:127-138,   lib/sba-1.5/sba_levmar.c:457-2081
 *
 * The structs below restate the reference's interface types field-for-field so that a caller
 * compiled against the reference headers can pass its own objects unchanged (same size, same
 * offsets; checked by tests/test_abi.py against the compiled reference when it is available).
 * There is no CPU fallback; unsupported reference options return BSFM_ERR_UNSUPPORTED.
 */
#ifndef BSFM_B200_BA_H
#define BSFM_B200_BA_H

#include <stdint.h>
#include "bsfm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BSFM_NUM_CAMERA_PARAMS 9     /* lib/sfm-driver/sfm.h:29 */
#define BSFM_POLY_INVERSE_DEGREE 6   /* lib/sfm-driver/sfm.h:30 */

/* == camera_params_t, lib/sfm-driver/sfm.h:32-51 */
typedef struct {
    double R[9];
    double t[3];
    double f;
    double k[2];
    double k_inv[BSFM_POLY_INVERSE_DEGREE];
    char constrained[BSFM_NUM_CAMERA_PARAMS];
    double constraints[BSFM_NUM_CAMERA_PARAMS];
    double weights[BSFM_NUM_CAMERA_PARAMS];
    double K_known[9];
    double k_known[5];
    char fisheye;
    char known_intrinsics;
    double f_cx, f_cy;
    double f_rad, f_angle;
    double f_focal;
    double f_scale, k_scale;
} bsfm_camera_params_t;

/* == v3_t, lib/matrix/vector.h:65-67 */
typedef struct { double p[3]; } bsfm_v3_t;

/* == camera_constraints_t / point_constraints_t, lib/sba-1.5/sba.h:80-90 */
typedef struct { char *constrained; double *constraints; double *weights; } bsfm_camera_constraints_t;
typedef struct { char constrained; double constraints[3]; double weight; } bsfm_point_constraints_t;

/* Jacobian mode of the solver (env BSFM_BA_JAC overrides: "fd" | "analytic")
 *   0 = forward finite differences with the reference's step rule d = max(1e-4*|p|, 1e-6)
 *       (lib/sba-1.5/sba_levmar_wrap.c:203-256, sba.h:52-53) -- what run_sfm uses (projac = NULL,
 *       lib/sfm-driver/sfm.c:820-828); the parity mode and the default.
 *   1 = analytic 2x9 / 2x3 Jacobians of the same camera model (include/snavely_reprojection_error.h:58-92
 *       restated in the SBA parameterisation)                            */
#define BSFM_BA_JAC_FD        0
#define BSFM_BA_JAC_ANALYTIC  1

/* Camera model descriptor: what sfm_project_point3 (lib/sfm-driver/sfm.c:503-552) reads from its
 * `adata` (sfm_global_t) -- made explicit because a GPU solver cannot call an opaque callback.   */
typedef struct {
    int est_focal_length;         /* aj[6] = f * f_scale is a parameter (else f_fixed[j] is used)   */
    int undistort;                /* two radial terms k1,k2 (scaled by k_scale) are parameters       */
    int explicit_camera_centers;  /* aj[0..2] is the camera centre c (P = R (X - c)), else P = R X + t */
    double f_scale, k_scale;      /* 0.001 and 5.0 in run_sfm (sfm.c:634-635)                        */
    const double *R_init;         /* m x 9 row-major initial rotations (init_params[j].R)            */
    const double *f_fixed;        /* m focal lengths used when est_focal_length == 0                 */
} bsfm_sfm_model_t;

/* Core solver == sba_motstr_levmar_x (lib/sba-1.5/sba_levmar.c:457-2081) with func/fjac/adata
 * replaced by the explicit camera model.  Arguments keep the reference's meaning:
 *   n points, m cameras, mcon leading cameras held fixed, vmask n x m, p = (a_1..a_m, b_1..b_n)
 *   in/out, cnp in {6,7,8,9}, pnp = 3, x = measurements (2 per visible projection, point-major),
 *   covx must be NULL, mnp = 2, itmax, verbose, opts[6] (reads opts[5], sba_levmar.c:610),
 *   info[10] (sba_levmar.c:2028-2049), camera / point constraints as in sba.h:80-90.
 * vmask, p and x may be host pointers or device pointers (unified addressing decides the copy kind);
 * with device pointers no bulk host<->device copy happens inside the call (bench `value` leg).
 * Vout/Sout/Uout/Wout (nullable) export the undamped V (n x 9), reduced camera matrix S ((m cnp)^2), U (m x cnp^2)
 * and W (dense (m cnp) x (3n)) like sba_levmar.c:1633-2026 (extra Jacobian + Schur pass when Sout != NULL; mcon == 0).
 * Returns the number of iterations (>=0) like the reference, SBA_ERROR (-1) where the reference
 * returns it, or another negative BSFM_ERR_* code.                                              */
int bsfm_sba_motstr_levmar_model(int n, int m, int mcon, const char *vmask, double *p, int cnp, int pnp,
                                 const double *x, const double *covx, int mnp,
                                 const bsfm_sfm_model_t *model, int jac_mode,
                                 int itmax, int verbose, const double opts[6], double info[10],
                                 int use_constraints, const bsfm_camera_constraints_t *constraints,
                                 int use_point_constraints, const bsfm_point_constraints_t *point_constraints,
                                 double *Vout, double *Sout, double *Uout, double *Wout);

/* Motion-only solver == sba_mot_levmar_x (lib/sba-1.5/sba.h:157-166, lib/sba-1.5/sba_levmar.c:2090-2690) with
 * func/fjac/adata replaced by the explicit camera model: only the cameras are unknowns, the n points are read from
 * `points` (n x 3, host or device; what sfm_project_point3_mot takes from its adata, lib/sfm-driver/sfm.c:553-560)
 * and never change.  p = (a_1..a_m) in/out.  The augmented normal equations are block diagonal,
 * (U_j + mu I) da_j = ea_j per camera (sba_levmar.c:2487-2514).  The LM controller is the stock SBA one: it has no
 * stop-8 rule and prints no max_pct_change; info[9] counts one linear system per camera and try (:2513).
 * Returns like bsfm_sba_motstr_levmar_model.                                                                    */
int bsfm_sba_mot_levmar_model(int n, int m, int mcon, const char *vmask, double *p, int cnp,
                              const double *x, const double *covx, int mnp,
                              const bsfm_sfm_model_t *model, const double *points, int jac_mode,
                              int itmax, int verbose, const double opts[6], double info[10],
                              int use_constraints, const bsfm_camera_constraints_t *constraints);

/* == run_sfm (lib/sfm-driver/sfm.h:68-86): same arguments, same in/out semantics
 * (init_camera_params and init_pts are overwritten with the solution, sfm.c:876-929; prints
 * "[run_sfm] Number of iterations" / "info[6]" like sfm.c:872-873), plus `info_out` (nullable,
 * 10 doubles) and an int return (0 or negative error) which the void reference lacks.
 * fix_points == 1 runs the motion-only solver (sfm.c:843-849; init_pts untouched).  GPU path covers
 * optimize_for_fisheye == 0, const_focal_length == 0 and cameras with known_intrinsics == 0; anything else
 * returns BSFM_ERR_UNSUPPORTED.                                                                 */
int bsfm_run_sfm(int num_pts, int num_cameras, int ncons, char *vmask, double *projections,
                 int est_focal_length, int const_focal_length, int undistort, int explicit_camera_centers,
                 bsfm_camera_params_t *init_camera_params, bsfm_v3_t *init_pts,
                 int use_constraints, int use_point_constraints, bsfm_v3_t *points_constraints,
                 double point_constraint_weight, int fix_points, int optimize_for_fisheye, double eps2,
                 double *Vout, double *Sout, double *Uout, double *Wout, double *info_out);

/* The reprojection statistics / outlier pass BundlerApp::RunSFM_SBA runs after every run_sfm on the same
 * vmask / projections / cameras / points (src/Bundle.cpp:659-856), as one call:
 *   dist(obs)   = || sfm_project_rd(camera, point, estimate_distortion, explicit centres) - measurement ||  (:721-754)
 *   per camera  : med80 = kth_element_copy(n, iround(0.8 n), dists) (0.0 when that index is >= n, lib/imagelib/
 *                 qsort.c:191-194), thresh = CLAMP(1.2 * 2.0 * med80, min_thresh, max_thresh)  (:761-771)
 *   outliers    : every point with dist > thresh in some camera, unless pt_protected[point] != 0 (the caller's
 *                 "constrained point" rule, :801-806); the reported error is the distance in the first (lowest
 *                 index) camera that flags the point, like the reference's first-found rule (:809-821)
 * cameras: the `R`, `t` (= centre), `f`, `k` fields of run_sfm's output are read.
 * cam_stats (nullable): num_cameras x 5 = { #observations, mean distance, median (iround(0.5 n)-th), med80, thresh }
 * obs_dist  (nullable): one distance per observation in the order of `projections`
 * outliers / outlier_errors: up to `cap` entries, ordered by (first flagging camera, point index) -- the
 *   reference lists them by (camera, key index); same set, same errors
 * global_mean (nullable): the "[RunSFM] Global mean reprojection error" value (:852-856)
 * Returns the number of outliers (may exceed cap) or a negative error.  Host or device pointers for vmask,
 * projections, pts.                                                                                          */
int bsfm_reprojection_outliers(int num_pts, int num_cameras, const char *vmask, const double *projections,
                               const bsfm_camera_params_t *cams, const bsfm_v3_t *pts, int estimate_distortion,
                               double min_proj_error_threshold, double max_proj_error_threshold,
                               const char *pt_protected, double *cam_stats, double *obs_dist,
                               int32_t *outliers, double *outlier_errors, int cap, double *global_mean);

/* == sba_Axb_Chol (lib/sba-1.5/sba_lapack.c:374-485, called at sba_levmar.c:1343): x = A^-1 B for a symmetric
 * positive definite m x m matrix (dpotrf + dpotrs in the reference).  A, B, x: host or device pointers; A and B are
 * NOT modified (the reference overwrites them when iscolmaj == 1).  Returns 1 on success, 0 when A is not positive
 * definite -- the reference's own return values -- and a negative BSFM_ERR_* on a library error.
 * Systems above 1536 unknowns run the 256-column panel factorisation whose trailing update is the tcgen05 int8-slice
 * kernel (csrc/ba_chol_tc.cu); smaller ones the latency-optimised path (csrc/ba_chol.cu).                          */
int bsfm_sba_Axb_Chol(const double *A, const double *B, double *x, int m, int iscolmaj);
/* the same solve repeated `reps` times on a fresh copy of A; *ms_per_solve = mean device time (CUDA events) of one
 * factorisation + both triangular solves (measurement entry used by bench.py and the tests)                          */
int bsfm_sba_Axb_Chol_timed(const double *A, const double *B, double *x, int m, int reps, float *ms_per_solve);

/* Per-kernel-class device timing of the large-system Cholesky (systems > 1536 unknowns), for roofline reporting:
 * bsfm_ba_chol_profile(1) resets the counters of the calling thread and records CUDA events around every launch of
 * the following solves; bsfm_ba_chol_profile_read returns accumulated ms / launch counts for
 * [0] diagonal-block factorisation  [1] panel solve (+ int8 slicing)  [2] trailing update  [3] back substitution,
 * the int8 operations the tensor-core trailing update executed and its fp64-equivalent FLOPs (2 per multiply-add of
 * the lower triangle).  Call it after the solves have completed.                                                  */
int bsfm_ba_chol_profile(int enable);
int bsfm_ba_chol_profile_read(float ms[4], int launches[4], double *tensor_int8_ops, double *fp64_equiv_flops);

/* Per-phase device time (ms, CUDA events) of the last solve of this thread:
 * [0] setup (H2D, CSR, Schur structure)  [1] residual/Jacobian/U/V/W  [2] Schur S assembly
 * [3] dense Cholesky + solves  [4] back-substitution/update/function eval  [5] whole solve      */
int bsfm_ba_last_timing(float ms[6], int *iterations, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* BSFM_B200_BA_H */
