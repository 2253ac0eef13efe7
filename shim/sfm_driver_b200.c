/* shim/sfm_driver_b200.c -- link-time drop-in for the reference's libsfmdrv.a (lib/sfm-driver).
 *
 * Exports `run_sfm` with exactly the reference signature (lib/sfm-driver/sfm.h:68-86, extern "C"
 * :23-25) and forwards to bsfm_run_sfm in libbsfm_b200.so.  src/Bundle.cpp (RunSFM_SBA, :645-652)
 * links against this object instead of libsfmdrv.a + libsba.v1.5.a and is otherwise unchanged:
 *
 *     gcc -c shim/sfm_driver_b200.c -Iinclude -o sfm_driver_b200.o
 *     ... -o bundler ... sfm_driver_b200.o -L<repo>/bundler_sfm_b200 -lbsfm_b200   (instead of -lsfmdrv -lsba.v1.5)
 *
 * The struct arguments are layout-identical (include/bsfm_b200_ba.h restates camera_params_t / v3_t).
 * Options the GPU path does not cover (fisheye, known_intrinsics) terminate like the reference's own fatal
 * paths do (printf + exit(1), sfm.c:56-73): there is deliberately no silent CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include "bsfm_b200_ba.h"

void run_sfm(int num_pts, int num_cameras, int ncons, char *vmask, double *projections,
             int est_focal_length, int const_focal_length, int undistort, int explicit_camera_centers,
             bsfm_camera_params_t *init_camera_params, bsfm_v3_t *init_pts,
             int use_constraints, int use_point_constraints, bsfm_v3_t *points_constraints,
             double point_constraint_weight, int fix_points, int optimize_for_fisheye, double eps2,
             double *Vout, double *Sout, double *Uout, double *Wout)
{
    int rc = bsfm_run_sfm(num_pts, num_cameras, ncons, vmask, projections, est_focal_length, const_focal_length,
                          undistort, explicit_camera_centers, init_camera_params, init_pts, use_constraints,
                          use_point_constraints, points_constraints, point_constraint_weight, fix_points,
                          optimize_for_fisheye, eps2, Vout, Sout, Uout, Wout, NULL);
    if (rc < 0) {
        printf("[run_sfm/b200] error %d: %s\n", rc, bsfm_last_error());
        fflush(stdout);
        exit(1);
    }
}
