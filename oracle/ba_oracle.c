/* oracle/ba_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; never shipped, never timed as the
 * product).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Restates the BA hot path of snavely/bundler_sfm in plain C, single-threaded, in the reference's own
 * order of operations:
 *   run_sfm                 lib/sfm-driver/sfm.c:592-1003   (packing :652-703, constraints :721-781,
 *                                                            options :705-714, unpacking :876-929)
 *   sfm_project_point3      lib/sfm-driver/sfm.c:503-552
 *   sfm_project_rd          lib/sfm-driver/sfm.c:302-380    (known_intrinsics = 0 branch)
 *   rot_update              lib/sfm-driver/sfm.c:77-116
 *   sba_motstr_Qs_fdjac     lib/sba-1.5/sba_levmar_wrap.c:163-259  (forward differences, sba.h:52-53)
 *   sba_motstr_levmar_x     lib/sba-1.5/sba_levmar.c:457-2081      (every controller quirk of SURVEY.md A.3)
 *   sba_symat_invert_BK     lib/sba-1.5/sba_lapack.c:1053-1140     -> closed-form symmetric 3x3 inverse
 *   sba_Axb_Chol            lib/sba-1.5/sba_lapack.c:374-485       -> plain dense Cholesky (dpotrf/dpotrs
 *                            are LAPACK, outside /root/reference; results agree to rounding)
 *
 * Pinning: tests/test_oracle_ba.py checks this file against (a) the committed outputs of the UNMODIFIED
 * reference (tests/golden/ba_golden.npz: kermit example + synthetic scenes, incl. constraints) and
 * (b) oracle/_ref/libref_sba.so directly when it is built: same iteration count and stop reason, RMSE
 * within 1e-9, parameters within 1e-5 of each group's magnitude (differences come only from LAPACK's vs
 * this file's Cholesky rounding, amplified by the conditioning of the reduced camera system).
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/bsfm_b200_ba.h"

#define SBA_EPSILON_SQ (1E-12 * 1E-12)     /* sba_levmar.c:34-35 */
#define SBA_ONE_THIRD 0.3333333334         /* sba_levmar.c:37 */

typedef struct {
    int cnp, est_focal, undistort, explicit_centers;
    double f_scale, k_scale;
    const double *R_init;   /* m*9 */
    const double *f_fixed;  /* m */
} omodel_t;

static void o_rot_update(const double *R, const double *w, double *Rnew)   /* sfm.c:77-116 */
{
    double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double n[3], nx[9], nxsq[9], dR[9], sinth, costh;
    int r, c, q;
    if (theta == 0.0) { memcpy(Rnew, R, 9 * sizeof(double)); return; }
    n[0] = w[0] / theta; n[1] = w[1] / theta; n[2] = w[2] / theta;
    nx[0] = 0.0; nx[1] = -n[2]; nx[2] = n[1];
    nx[3] = n[2]; nx[4] = 0.0; nx[5] = -n[0];
    nx[6] = -n[1]; nx[7] = n[0]; nx[8] = 0.0;
    for (r = 0; r < 3; r++)
        for (c = 0; c < 3; c++) nxsq[3 * r + c] = nx[3 * r] * nx[c] + nx[3 * r + 1] * nx[3 + c] + nx[3 * r + 2] * nx[6 + c];
    sinth = sin(theta); costh = cos(theta);
    for (q = 0; q < 9; q++) dR[q] = (((q % 4 == 0) ? 1.0 : 0.0) + nx[q] * sinth) + nxsq[q] * (1.0 - costh);
    for (r = 0; r < 3; r++)
        for (c = 0; c < 3; c++) Rnew[3 * r + c] = dR[3 * r] * R[c] + dR[3 * r + 1] * R[3 + c] + dR[3 * r + 2] * R[6 + c];
}

/* sfm_project_point3 + sfm_project_rd; the per-camera rotation cache of sfm.c:539-547 is replaced by
 * recomputing rot_update (same value) */
static void o_project(const omodel_t *M, int j, const double *aj, const double *b, double *xij)
{
    double R[9], bc[3], K0, p0, p1;
    o_rot_update(M->R_init + 9 * j, aj + 3, R);
    if (M->explicit_centers) {
        double b2[3];
        b2[0] = b[0] - aj[0]; b2[1] = b[1] - aj[1]; b2[2] = b[2] - aj[2];
        bc[0] = R[0] * b2[0] + R[1] * b2[1] + R[2] * b2[2];
        bc[1] = R[3] * b2[0] + R[4] * b2[1] + R[5] * b2[2];
        bc[2] = R[6] * b2[0] + R[7] * b2[1] + R[8] * b2[2];
    } else {
        bc[0] = R[0] * b[0] + R[1] * b[1] + R[2] * b[2];
        bc[1] = R[3] * b[0] + R[4] * b[1] + R[5] * b[2];
        bc[2] = R[6] * b[0] + R[7] * b[1] + R[8] * b[2];
        bc[0] += aj[0]; bc[1] += aj[1]; bc[2] += aj[2];
    }
    K0 = M->est_focal ? aj[6] / M->f_scale : M->f_fixed[j];
    p0 = -bc[0] * K0 / bc[2];
    p1 = -bc[1] * K0 / bc[2];
    if (M->undistort) {
        const double *k = M->est_focal ? aj + 7 : aj + 6;
        double k1 = k[0] / M->k_scale, k2 = k[1] / M->k_scale;
        double rsq = (p0 * p0 + p1 * p1) / (K0 * K0);
        double factor = 1.0 + k1 * rsq + k2 * rsq * rsq;
        p0 *= factor; p1 *= factor;
    }
    xij[0] = p0; xij[1] = p1;
}

static double o_fd_step(double v)   /* sba_levmar_wrap.c:207-212 */
{
    double d = 1E-04 * v;
    d = fabs(d);
    if (d < 1E-06) d = 1E-06;
    return d;
}

/* dense SPD solve A x = b (A n x n row-major, overwritten); returns 0 when not positive definite */
static int o_chol_solve(double *A, double *b, double *x, int n)
{
    int i, j, k;
    for (j = 0; j < n; j++) {
        double d = A[(size_t) j * n + j];
        for (k = 0; k < j; k++) d -= A[(size_t) j * n + k] * A[(size_t) j * n + k];
        if (!(d > 0.0)) return 0;
        d = sqrt(d);
        A[(size_t) j * n + j] = d;
        for (i = j + 1; i < n; i++) {
            double s = A[(size_t) i * n + j];
            for (k = 0; k < j; k++) s -= A[(size_t) i * n + k] * A[(size_t) j * n + k];
            A[(size_t) i * n + j] = s / d;
        }
    }
    for (i = 0; i < n; i++) {
        double s = b[i];
        for (k = 0; k < i; k++) s -= A[(size_t) i * n + k] * x[k];
        x[i] = s / A[(size_t) i * n + i];
    }
    for (i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (k = i + 1; k < n; k++) s -= A[(size_t) k * n + i] * x[k];
        x[i] = s / A[(size_t) i * n + i];
    }
    return 1;
}

/* symmetric 3x3 inverse (upper triangle in, full out); 0 when singular */
static int o_inv3(const double *V, double mu, double *O)
{
    double a00 = V[0] + mu, a01 = V[1], a02 = V[2], a11 = V[4] + mu, a12 = V[5], a22 = V[8] + mu;
    double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    double det = a00 * c00 + a01 * c01 + a02 * c02, id;
    if (!(det != 0.0) || !isfinite(det)) return 0;
    id = 1.0 / det;
    O[0] = c00 * id; O[1] = c01 * id; O[2] = c02 * id;
    O[3] = c01 * id; O[4] = (a00 * a22 - a02 * a02) * id; O[5] = (a01 * a02 - a00 * a12) * id;
    O[6] = c02 * id; O[7] = O[5]; O[8] = (a00 * a11 - a01 * a01) * id;
    return 1;
}

/* sba_motstr_levmar_x restated for the sfm camera model (covx == NULL, mnp = 2, pnp = 3) */
int oracle_sba_motstr_levmar(int n, int m, int mcon, const char *vmask, double *p, int cnp, const double *x,
                             const omodel_t *M, int itmax, int verbose, const double *opts, double *info,
                             int use_constraints, const bsfm_camera_constraints_t *constraints,
                             int use_point_constraints, const bsfm_point_constraints_t *pcons)
{
    int i, j, k, ii, jj, l, nvis = 0, nobs, nvars = m * cnp + n * 3, Sdim = (m - mcon) * cnp;
    int *rowptr, *colidx, *camptr, *camobs, *fill;
    double *e, *hx, *jac, *W, *U, *V, *Vinv, *eab, *S, *E, *dp, *pdp, *Y, *da;
    double mu = 0.0, eab_inf = 0.0, p_eL2, pdp_eL2, p_L2 = 0.0, dp_L2 = DBL_MAX, dF, dL, init_p_eL2, maxdiag = DBL_MIN;
    double tau = fabs(opts[0]), eps1 = fabs(opts[1]), eps2 = fabs(opts[2]), eps2_sq = opts[2] * opts[2],
           eps3_sq = opts[3] * opts[3], eps4_sq = opts[4] * opts[4], eps5 = opts[5];
    int nu = 2, nu2, stop = 0, nfev = 0, njev = 0, nlss = 0, itno, retval;
    const int Asz = 2 * cnp, ABsz = 2 * cnp + 6, Wsz = cnp * 3, Usz = cnp * cnp;
    double *pa = p, *pb = p + m * cnp;

    for (i = 0; i < n * m; i++) nvis += (vmask[i] != 0);
    nobs = 2 * nvis;
    if (nobs < nvars) {
        fprintf(stderr, "SBA: sba_motstr_levmar_x() cannot solve a problem with fewer measurements [%d] than unknowns [%d]\n", nobs, nvars);
        return -1;
    }
    rowptr = malloc((n + 1) * sizeof(int)); colidx = malloc(nvis * sizeof(int));
    camptr = calloc(m + 2, sizeof(int)); camobs = malloc(nvis * sizeof(int)); fill = calloc(m + 1, sizeof(int));
    for (i = k = 0; i < n; i++) {   /* sba_levmar.c:652-663 */
        rowptr[i] = k;
        for (j = 0; j < m; j++) if (vmask[(size_t) i * m + j]) { colidx[k++] = j; camptr[j + 1]++; }
    }
    rowptr[n] = nvis;
    for (j = 0; j < m; j++) camptr[j + 1] += camptr[j];
    for (i = 0; i < n; i++) for (k = rowptr[i]; k < rowptr[i + 1]; k++) { j = colidx[k]; camobs[camptr[j] + fill[j]++] = k; }
    e = malloc(nobs * sizeof(double)); hx = malloc(nobs * sizeof(double));
    jac = malloc((size_t) nvis * ABsz * sizeof(double)); W = malloc((size_t) nvis * Wsz * sizeof(double));
    U = malloc((size_t) m * Usz * sizeof(double)); V = malloc((size_t) n * 9 * sizeof(double)); Vinv = malloc((size_t) n * 9 * sizeof(double));
    eab = malloc(nvars * sizeof(double)); S = malloc((size_t) Sdim * Sdim * sizeof(double)); E = malloc((size_t) m * cnp * sizeof(double));
    dp = malloc(nvars * sizeof(double)); pdp = malloc(nvars * sizeof(double)); Y = malloc((size_t) nvis * Wsz * sizeof(double));
    da = malloc((size_t) (Sdim > 0 ? Sdim : 1) * sizeof(double));

#define POINT_OF(kk, ivar) do { int lo_ = 0, hi_ = n - 1; while (lo_ < hi_) { int mid_ = (lo_ + hi_ + 1) >> 1; if (rowptr[mid_] <= (kk)) lo_ = mid_; else hi_ = mid_ - 1; } ivar = lo_; } while (0)
#define EVAL(pp, out) do { int i_, k_; for (i_ = 0; i_ < n; i_++) for (k_ = rowptr[i_]; k_ < rowptr[i_ + 1]; k_++) \
        o_project(M, colidx[k_], (pp) + colidx[k_] * cnp, (pp) + m * cnp + i_ * 3, (out) + 2 * k_); } while (0)
#define PENALTY(acc) do { if (use_constraints) for (j = 0; j < m; j++) for (jj = 0; jj < cnp; jj++) if (constraints[j].constrained[jj]) { \
            double diff = constraints[j].constraints[jj] - p[j * cnp + jj]; acc += constraints[j].weights[jj] * diff * diff; } \
        if (use_point_constraints) for (i = 0; i < n; i++) if (pcons[i].constrained) for (ii = 0; ii < 3; ii++) { \
            double diff = pcons[i].constraints[ii] - p[m * cnp + i * 3 + ii]; acc += nvis * pcons[i].weight * diff * diff; } } while (0)

    EVAL(p, hx); nfev = 1;
    p_eL2 = 0.0;
    for (i = 0; i < nobs; i++) { e[i] = x[i] - hx[i]; p_eL2 += e[i] * e[i]; }   /* nrmL2xmy :159-207 (sum order differs) */
    PENALTY(p_eL2);                                                             /* :808-842 */
    if (verbose) printf("initial motstr-SBA error %g [%g]\n", p_eL2, p_eL2 / nvis);
    init_p_eL2 = p_eL2;
    if (!isfinite(p_eL2)) stop = 7;

    for (itno = 0; itno < itmax && !stop; ++itno) {
        /* forward-difference Jacobian, sba_levmar_wrap.c:203-256 */
        for (i = 0; i < n; i++)
            for (k = rowptr[i]; k < rowptr[i + 1]; k++) {
                double a[9], b[3], h0[2], h1[2], *pAB = jac + (size_t) k * ABsz;
                j = colidx[k];
                memcpy(a, pa + j * cnp, cnp * sizeof(double)); memcpy(b, pb + i * 3, 3 * sizeof(double));
                o_project(M, j, a, b, h0);
                for (jj = 0; jj < cnp; jj++) {
                    double d = o_fd_step(a[jj]), d1 = 1.0 / d, tmp = a[jj];
                    a[jj] += d; o_project(M, j, a, b, h1); a[jj] = tmp;
                    pAB[jj] = (h1[0] - h0[0]) * d1; pAB[cnp + jj] = (h1[1] - h0[1]) * d1;
                }
                for (jj = 0; jj < 3; jj++) {
                    double d = o_fd_step(b[jj]), d1 = 1.0 / d, tmp = b[jj];
                    b[jj] += d; o_project(M, j, a, b, h1); b[jj] = tmp;
                    pAB[Asz + jj] = (h1[0] - h0[0]) * d1; pAB[Asz + 3 + jj] = (h1[1] - h0[1]) * d1;
                }
            }
        ++njev;
        /* U_j, ea_j :919-964 */
        memset(U, 0, (size_t) m * Usz * sizeof(double)); memset(eab, 0, nvars * sizeof(double));
        for (j = mcon; j < m; j++) {
            double *Uj = U + (size_t) j * Usz, *eaj = eab + j * cnp;
            for (l = camptr[j]; l < camptr[j + 1]; l++) {
                const double *A = jac + (size_t) camobs[l] * ABsz, *ee = e + 2 * camobs[l];
                for (ii = 0; ii < cnp; ii++) {
                    for (jj = ii; jj < cnp; jj++) { double sum = 0.0; for (k = 0; k < 2; k++) sum += A[k * cnp + ii] * A[k * cnp + jj]; Uj[ii * cnp + jj] += sum; }
                    for (jj = 0; jj < ii; jj++) Uj[ii * cnp + jj] = Uj[jj * cnp + ii];
                }
                for (ii = 0; ii < cnp; ii++) { double sum = 0.0; for (jj = 0; jj < 2; jj++) sum += A[jj * cnp + ii] * ee[jj]; eaj[ii] += sum; }
            }
            if (use_constraints)
                for (jj = 0; jj < cnp; jj++) if (constraints[j].constrained[jj]) {
                    double diff = constraints[j].constraints[jj] - p[j * cnp + jj];
                    Uj[jj * cnp + jj] += constraints[j].weights[jj]; eaj[jj] += constraints[j].weights[jj] * diff;
                }
        }
        /* V_i, eb_i :987-1030 ; W_ij :1053-1082 */
        for (i = 0; i < n; i++) {
            double *Vi = V + (size_t) i * 9, *ebi = eab + m * cnp + i * 3;
            memset(Vi, 0, 9 * sizeof(double));
            for (k = rowptr[i]; k < rowptr[i + 1]; k++) {
                const double *A = jac + (size_t) k * ABsz, *B = A + Asz, *ee = e + 2 * k;
                double *Wk = W + (size_t) k * Wsz;
                for (ii = 0; ii < 3; ii++) {
                    for (jj = ii; jj < 3; jj++) { double sum = 0.0; for (l = 0; l < 2; l++) sum += B[l * 3 + ii] * B[l * 3 + jj]; Vi[ii * 3 + jj] += sum; }
                    { double sum = 0.0; for (jj = 0; jj < 2; jj++) sum += B[jj * 3 + ii] * ee[jj]; ebi[ii] += sum; }
                }
                for (ii = 0; ii < cnp; ii++) for (jj = 0; jj < 3; jj++) {
                    double sum = 0.0; for (l = 0; l < 2; l++) sum += A[l * cnp + ii] * B[l * 3 + jj];
                    Wk[ii * 3 + jj] = (colidx[k] < mcon) ? 0.0 : sum;
                }
            }
            if (use_point_constraints && pcons[i].constrained)
                for (ii = 0; ii < 3; ii++) {
                    double diff = pcons[i].constraints[ii] - p[m * cnp + i * 3 + ii];
                    Vi[ii * 3 + ii] += nvis * pcons[i].weight; ebi[ii] += nvis * pcons[i].weight * diff;
                }
            Vi[3] = Vi[1]; Vi[6] = Vi[2]; Vi[7] = Vi[5];
        }
        for (i = 0, p_L2 = eab_inf = 0.0; i < nvars; i++) { double t = fabs(eab[i]); if (eab_inf < t) eab_inf = t; p_L2 += p[i] * p[i]; }   /* :1084-1088 */
        maxdiag = DBL_MIN;
        for (j = mcon; j < m; j++) for (ii = 0; ii < cnp; ii++) if (U[(size_t) j * Usz + ii * cnp + ii] > maxdiag) maxdiag = U[(size_t) j * Usz + ii * cnp + ii];
        for (i = 0; i < n; i++) for (ii = 0; ii < 3; ii++) if (V[(size_t) i * 9 + ii * 4] > maxdiag) maxdiag = V[(size_t) i * 9 + ii * 4];
        if (eab_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
        if (itno == 0) mu = tau * maxdiag;

        while (1) {
            int singular = 0, issolved, accepted = 0;
            for (i = 0; i < n && !singular; i++) if (!o_inv3(V + (size_t) i * 9, mu, Vinv + (size_t) i * 9)) singular = 1;   /* :1137-1162 */
            if (singular) {
                fprintf(stderr, "SBA: singular matrix V*_i in sba_motstr_levmar_x(), increasing damping\n");
            } else {
                /* Y_ij = W_ij V*_i^-1 ; S_jk = delta_jk U*_j - sum_i Y_ij W_ik^T ; E_j = ea_j - sum_i Y_ij eb_i  :1170-1339 */
                memset(S, 0, (size_t) Sdim * Sdim * sizeof(double));
                for (i = 0; i < n; i++)
                    for (k = rowptr[i]; k < rowptr[i + 1]; k++) {
                        const double *Wk = W + (size_t) k * Wsz, *Vi = Vinv + (size_t) i * 9;
                        double *Yk = Y + (size_t) k * Wsz;
                        for (ii = 0; ii < cnp; ii++) for (jj = 0; jj < 3; jj++) {
                            double sum = 0.0; for (l = 0; l < 3; l++) sum += Wk[ii * 3 + l] * Vi[l * 3 + jj];
                            Yk[ii * 3 + jj] = sum;
                        }
                    }
                for (j = mcon; j < m; j++) {
                    double *Ej = E + j * cnp;
                    for (ii = 0; ii < cnp; ii++) Ej[ii] = 0.0;
                    for (ii = 0; ii < cnp; ii++) for (jj = 0; jj < cnp; jj++) {
                        double u = U[(size_t) j * Usz + ii * cnp + jj]; if (ii == jj) u += mu;
                        S[(size_t) ((j - mcon) * cnp + ii) * Sdim + (j - mcon) * cnp + jj] = u;
                    }
                    for (l = camptr[j]; l < camptr[j + 1]; l++) {      /* points seen by camera j, ascending */
                        int ka = camobs[l], kb, pt;
                        const double *Ya = Y + (size_t) ka * Wsz;
                        POINT_OF(ka, pt);
                        for (ii = 0; ii < cnp; ii++) { double sum = 0.0; for (jj = 0; jj < 3; jj++) sum += Ya[ii * 3 + jj] * eab[m * cnp + pt * 3 + jj]; Ej[ii] += sum; }
                        for (kb = rowptr[pt]; kb < rowptr[pt + 1]; kb++) {
                            int kc = colidx[kb];
                            const double *Wb = W + (size_t) kb * Wsz;
                            if (kc < mcon) continue;
                            for (ii = 0; ii < cnp; ii++) for (jj = 0; jj < cnp; jj++) {
                                double sum = 0.0; for (k = 0; k < 3; k++) sum += Ya[ii * 3 + k] * Wb[jj * 3 + k];
                                S[(size_t) ((j - mcon) * cnp + ii) * Sdim + (kc - mcon) * cnp + jj] -= sum;
                            }
                        }
                    }
                    for (ii = 0; ii < cnp; ii++) Ej[ii] = eab[j * cnp + ii] - Ej[ii];
                }
                issolved = o_chol_solve(S, E + mcon * cnp, da, Sdim);   /* :1368 */
                ++nlss;
                if (issolved) {
                    for (i = 0; i < m * cnp; i++) dp[i] = (i < mcon * cnp) ? 0.0 : da[i - mcon * cnp];
                    for (i = 0; i < n; i++) {   /* db_i :1393-1433 */
                        double Wt[3] = {0, 0, 0}; const double *Vi = Vinv + (size_t) i * 9;
                        for (k = rowptr[i]; k < rowptr[i + 1]; k++) {
                            const double *Wk = W + (size_t) k * Wsz, *daj = dp + colidx[k] * cnp;
                            if (colidx[k] < mcon) continue;
                            for (ii = 0; ii < 3; ii++) { double sum = 0.0; for (jj = 0; jj < cnp; jj++) sum += Wk[jj * 3 + ii] * daj[jj]; Wt[ii] += sum; }
                        }
                        for (ii = 0; ii < 3; ii++) Wt[ii] = eab[m * cnp + i * 3 + ii] - Wt[ii];
                        for (ii = 0; ii < 3; ii++) { double sum = 0.0; for (jj = 0; jj < 3; jj++) sum += Vi[ii * 3 + jj] * Wt[jj]; dp[m * cnp + i * 3 + ii] = sum; }
                    }
                    for (i = 0, dp_L2 = 0.0; i < nvars; i++) { pdp[i] = p[i] + dp[i]; dp_L2 += dp[i] * dp[i]; }
                    if (dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                    if (dp_L2 >= (p_L2 + eps2) / SBA_EPSILON_SQ) { retval = -1; goto done; }
                    EVAL(pdp, hx); ++nfev;
                    pdp_eL2 = 0.0;
                    for (i = 0; i < nobs; i++) { hx[i] = x[i] - hx[i]; pdp_eL2 += hx[i] * hx[i]; }
                    if (!isfinite(pdp_eL2)) { stop = 7; break; }
                    PENALTY(pdp_eL2);                               /* evaluated at the OLD p, :1487-1522 */
                    for (i = 0, dL = 0.0; i < nvars; i++) dL += dp[i] * (mu * dp[i] + eab[i]);
                    dF = p_eL2 - pdp_eL2;
                    if (dL > 0.0 && dF > 0.0) {
                        double max_pct_change = 0.0, tmp = (2.0 * dF / dL - 1.0);
                        tmp = 1.0 - tmp * tmp * tmp;
                        mu = mu * ((tmp >= SBA_ONE_THIRD) ? tmp : SBA_ONE_THIRD);
                        nu = 2;
                        for (i = 0; i < nobs; i++)                    /* :1552-1561 */
                            if (!(e[i] < eps5 && hx[i] < eps5)) { double pc = fabs((e[i] - hx[i]) / e[i]); if (pc > max_pct_change) max_pct_change = pc; }
                        if (verbose) printf("max_pct_change: %0.3e\n", max_pct_change);
                        if (pdp_eL2 - 2.0 * sqrt(p_eL2 * pdp_eL2) < (eps4_sq - 1.0) * p_eL2) stop = 4;
                        if (max_pct_change < eps5 && itno >= 4) { stop = 8; break; }
                        memcpy(p, pdp, nvars * sizeof(double)); memcpy(e, hx, nobs * sizeof(double));
                        p_eL2 = pdp_eL2;
                        accepted = 1;
                    }
                }
            }
            if (accepted) break;
            mu *= nu; nu2 = nu << 1;                                 /* :1584-1597 */
            if (nu2 <= nu) { stop = 6; break; }
            nu = nu2;
        }
        if (p_eL2 <= eps3_sq) stop = 5;
    }
    if (itno >= itmax) stop = 3;
    if (info) {
        info[0] = init_p_eL2; info[1] = p_eL2; info[2] = eab_inf; info[3] = dp_L2; info[4] = mu / maxdiag;
        info[5] = itno; info[6] = stop; info[7] = nfev; info[8] = njev; info[9] = nlss;
    }
    retval = (stop != 7) ? itno : -1;
done:
    free(rowptr); free(colidx); free(camptr); free(camobs); free(fill); free(e); free(hx); free(jac); free(W); free(U); free(V);
    free(Vinv); free(eab); free(S); free(E); free(dp); free(pdp); free(Y); free(da);
    return retval;
}

/* run_sfm restated (sfm.c:592-1003) + info_out; GPU-path subset: fix_points = 0, no fisheye */
int oracle_run_sfm(int num_pts, int num_cameras, int ncons, char *vmask, double *projections,
                   int est_focal_length, int const_focal_length, int undistort, int explicit_camera_centers,
                   bsfm_camera_params_t *cams, bsfm_v3_t *init_pts, int use_constraints, int use_point_constraints,
                   bsfm_v3_t *pt_constraints, double pt_constraint_weight, int fix_points, int optimize_for_fisheye,
                   double eps2, double *Vout, double *Sout, double *Uout, double *Wout, double *info_out)
{
    const double f_scale = 0.001, k_scale = 5.0;
    int cnp = (est_focal_length ? 7 : 6) + (undistort ? 2 : 0), i, j, c, rc;
    int ncp = cnp * num_cameras;
    double *params = malloc(((size_t) ncp + 3 * (size_t) num_pts) * sizeof(double));
    double *R_init = malloc((size_t) num_cameras * 9 * sizeof(double)), *f_fixed = malloc(num_cameras * sizeof(double));
    double opts[6], info[10];
    bsfm_camera_constraints_t *cons = NULL;
    bsfm_point_constraints_t *pcons = NULL;
    omodel_t M;
    (void) const_focal_length; (void) Vout; (void) Sout; (void) Uout; (void) Wout;
    if (fix_points || optimize_for_fisheye) return -6;
    for (j = 0; j < num_cameras; j++) {
        cams[j].f_scale = f_scale; cams[j].k_scale = k_scale;
        params[cnp * j + 0] = cams[j].t[0]; params[cnp * j + 1] = cams[j].t[1]; params[cnp * j + 2] = cams[j].t[2];
        params[cnp * j + 3] = params[cnp * j + 4] = params[cnp * j + 5] = 0.0;
        if (est_focal_length) { params[cnp * j + 6] = cams[j].f * cams[j].f_scale; c = 7; } else c = 6;
        if (undistort) { params[cnp * j + c] = cams[j].k[0] * k_scale; params[cnp * j + c + 1] = cams[j].k[1] * k_scale; }
        memcpy(R_init + 9 * j, cams[j].R, 9 * sizeof(double));
        f_fixed[j] = cams[j].f;
    }
    for (i = 0; i < num_pts; i++) { params[ncp + 3 * i] = init_pts[i].p[0]; params[ncp + 3 * i + 1] = init_pts[i].p[1]; params[ncp + 3 * i + 2] = init_pts[i].p[2]; }
    opts[0] = 1.0e-3; opts[1] = 1.0e-10; opts[2] = eps2; opts[3] = 1.0e-12; opts[4] = 0.0; opts[5] = 4.0e-2;
    if (use_constraints) {
        cons = malloc(num_cameras * sizeof(*cons));
        for (i = 0; i < num_cameras; i++) {
            int k0 = est_focal_length ? 7 : 6;
            cons[i].constrained = malloc(cnp); cons[i].constraints = malloc(cnp * sizeof(double)); cons[i].weights = malloc(cnp * sizeof(double));
            memcpy(cons[i].constrained, cams[i].constrained, cnp);
            memcpy(cons[i].constraints, cams[i].constraints, cnp * sizeof(double));
            memcpy(cons[i].weights, cams[i].weights, cnp * sizeof(double));
            if (est_focal_length) { cons[i].constraints[6] *= f_scale; cons[i].weights[6] *= (1.0 / (f_scale * f_scale)); }
            if (undistort) {
                cons[i].constraints[k0] *= k_scale; cons[i].weights[k0] *= (1.0 / (k_scale * k_scale));
                cons[i].constraints[k0 + 1] *= k_scale; cons[i].weights[k0 + 1] *= (1.0 / (k_scale * k_scale));
            }
        }
    }
    if (use_point_constraints) {
        pcons = malloc(num_pts * sizeof(*pcons));
        for (i = 0; i < num_pts; i++) {
            const double *q = pt_constraints[i].p;
            if (q[0] == 0.0 && q[1] == 0.0 && q[2] == 0.0) { memset(&pcons[i], 0, sizeof(pcons[i])); }
            else { pcons[i].constrained = 1; pcons[i].weight = pt_constraint_weight; memcpy(pcons[i].constraints, q, 3 * sizeof(double)); }
        }
    }
    M.cnp = cnp; M.est_focal = est_focal_length; M.undistort = undistort; M.explicit_centers = explicit_camera_centers;
    M.f_scale = f_scale; M.k_scale = k_scale; M.R_init = R_init; M.f_fixed = f_fixed;
    rc = oracle_sba_motstr_levmar(num_pts, num_cameras, ncons, vmask, params, cnp, projections, &M, 150,
                                  getenv("ORACLE_SBA_VERBOSE") ? atoi(getenv("ORACLE_SBA_VERBOSE")) : 0, opts, info,
                                  use_constraints, cons, use_point_constraints, pcons);
    if (info_out) memcpy(info_out, info, sizeof info);
    for (j = 0; j < num_cameras; j++) {
        double Rnew[9];
        cams[j].t[0] = params[cnp * j]; cams[j].t[1] = params[cnp * j + 1]; cams[j].t[2] = params[cnp * j + 2];
        o_rot_update(cams[j].R, params + cnp * j + 3, Rnew);
        memcpy(cams[j].R, Rnew, sizeof Rnew);
        if (est_focal_length) { c = 7; cams[j].f = params[cnp * j + 6] / cams[j].f_scale; } else c = 6;
        if (undistort) { cams[j].k[0] = params[cnp * j + c] / k_scale; cams[j].k[1] = params[cnp * j + c + 1] / k_scale; }
        cams[j].f_scale = 1.0; cams[j].k_scale = 1.0;
    }
    for (i = 0; i < num_pts; i++) { init_pts[i].p[0] = params[ncp + 3 * i]; init_pts[i].p[1] = params[ncp + 3 * i + 1]; init_pts[i].p[2] = params[ncp + 3 * i + 2]; }
    if (cons) { for (i = 0; i < num_cameras; i++) { free(cons[i].constrained); free(cons[i].constraints); free(cons[i].weights); } free(cons); }
    free(pcons); free(params); free(R_init); free(f_fixed);
    return rc < 0 ? rc : 0;
}
