/* oracle/ref_sba_wrap.c -- TEST INFRASTRUCTURE ONLY.
 * Glue for the UNMODIFIED reference BA (lib/sba-1.5 + lib/sfm-driver) built into
 * oracle/_ref/libref_sba.so:
 *   - lmdif_/lmdif1_ stubs (minpack; only reached from camera_refine, never from run_sfm,
 *     lib/sfm-driver/sfm.c:1016-1100)
 *   - ref_hook_sba_motstr_levmar: sfm.c is compiled with -Dsba_motstr_levmar=<this> so the
 *     info[10] vector that run_sfm only prints (sfm.c:872-873) can be read back by tests.
 *   - ref_hook_sba_mot_levmar: the same for the motion-only call (fix_points = 1, sfm.c:843-856).
 *   - REF_SBA_ITMAX / REF_SBA_EPS5 (environment): cap the iteration count / replace opts[5] that run_sfm hard-codes
 *     (sfm.c:705-714, :814-815), for equal-iteration-count comparisons and bounded timing samples
 *     of the big configuration.  The reference sources stay unmodified; only the arguments of the call change.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sba.h"

void lmdif_(void)  { fprintf(stderr, "lmdif_ stub reached\n");  abort(); }
void lmdif1_(void) { fprintf(stderr, "lmdif1_ stub reached\n"); abort(); }

static double g_last_info[SBA_INFOSZ];
static int g_last_ret = -2;

int ref_hook_sba_motstr_levmar(const int n, const int m, const int mcon, char *vmask,
        double *p, const int cnp, const int pnp, double *x, double *covx, const int mnp,
        void (*proj)(int j, int i, double *aj, double *bi, double *xij, void *adata),
        void (*projac)(int j, int i, double *aj, double *bi, double *Aij, double *Bij, void *adata),
        void *adata, const int itmax, const int verbose, const double opts[SBA_OPTSSZ],
        double info[SBA_INFOSZ], int use_constraints, camera_constraints_t *constraints,
        int use_point_constraints, point_constraints_t *point_constraints,
        double *Vout, double *Sout, double *Uout, double *Wout)
{
    int verb = verbose, it = itmax;
    double o[6];      /* run_sfm passes 6 entries; sba_motstr_levmar_x reads opts[5] (sba_levmar.c:610) although SBA_OPTSSZ is 5 */
    const char *q = getenv("REF_SBA_VERBOSE");
    if (q) verb = atoi(q);
    memcpy(o, opts, 6 * sizeof(double));
    q = getenv("REF_SBA_ITMAX");
    if (q) it = atoi(q);
    q = getenv("REF_SBA_EPS5");
    if (q) o[5] = atof(q);
    g_last_ret = sba_motstr_levmar(n, m, mcon, vmask, p, cnp, pnp, x, covx, mnp, proj, projac,
                                   adata, it, verb, o, info, use_constraints, constraints,
                                   use_point_constraints, point_constraints, Vout, Sout, Uout, Wout);
    memcpy(g_last_info, info, sizeof(g_last_info));
    return g_last_ret;
}

int ref_hook_sba_mot_levmar(const int n, const int m, const int mcon, char *vmask,
        double *p, const int cnp, double *x, double *covx, const int mnp,
        void (*proj)(int j, int i, double *aj, double *xij, void *adata),
        void (*projac)(int j, int i, double *aj, double *Aij, void *adata),
        void *adata, const int itmax, const int verbose, const double opts[SBA_OPTSSZ],
        double info[SBA_INFOSZ], int use_constraints, camera_constraints_t *constraints)
{
    int verb = verbose;
    const char *q = getenv("REF_SBA_VERBOSE");
    if (q) verb = atoi(q);
    g_last_ret = sba_mot_levmar(n, m, mcon, vmask, p, cnp, x, covx, mnp, proj, projac,
                                adata, itmax, verb, opts, info, use_constraints, constraints);
    memcpy(g_last_info, info, sizeof(g_last_info));
    return g_last_ret;
}

void ref_last_info(double *info10, int *ret)
{
    memcpy(info10, g_last_info, sizeof(g_last_info));
    if (ret) *ret = g_last_ret;
}
