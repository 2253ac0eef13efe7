// ba_chol.cu -- dense SPD solve of the reduced camera system S da = E.
// Reference: sba_Axb_Chol = LAPACK dpotrf + dpotrs, lib/sba-1.5/sba_lapack.c:374-485.
//
// fp64 blocked right-looking Cholesky (lower, row-major) with the right-hand side carried as an
// extra matrix row (row n), so the forward substitution L y = E falls out of the factorisation;
// back substitution uses the explicitly inverted 32x32 diagonal blocks.  Two-level blocking: inner
// panels of 32 columns, outer panels of NBO columns whose trailing update is one large SYRK-shaped
// GEMM (A22 -= L21 L21^T) -- the dense contraction of this path.
//
//   per inner panel k : diag_kernel   (1 CTA)  potf2 of A_kk, L_kk^-1
//                       trsm_kernel   (rows)   A_rk <- A_rk L_kk^-T   for all rows below (incl. RHS row)
//                       syrk_kernel   (tiles)  columns inside the outer panel
//   per outer panel   : syrk_kernel   (tiles)  trailing matrix, K = NBO
//   end               : backsolve_kernel (1 CTA) L^T x = y
#include "ba_kernels.cuh"
#include "common.h"

namespace bsfm {
namespace ba {

constexpr int NB = 32;

// ---- diagonal block: potf2 + triangular inverse ------------------------------------------------
__global__ void __launch_bounds__(256) chol_diag_kernel(double *A, int ld, int k0, int nb, double *Linv, Scalars *sc)
{
    __shared__ double L[NB][NB + 1];
    __shared__ double Z[NB][NB + 1];
    __shared__ int fail;
    const int tid = threadIdx.x;
    if (tid == 0) fail = 0;
    for (int q = tid; q < NB * NB; q += 256) {
        const int r = q / NB, c = q % NB;
        L[r][c] = (r < nb && c <= r) ? A[(size_t) (k0 + r) * ld + (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            const double d = L[j][j];
            if (!(d > 0.0) || !isfinite(d)) { fail = 1; L[j][j] = 1.0; }
            else L[j][j] = sqrt(d);
        }
        __syncthreads();
        const double djj = L[j][j];
        if (tid > j && tid < nb) L[tid][j] = L[tid][j] / djj;   // column scale (thread = row)
        __syncthreads();
        // trailing rank-1 update of the lower triangle: rows r > j, cols j < c <= r
        const int rem = nb - j - 1;
        for (int q = tid; q < rem * rem; q += 256) {
            const int r = j + 1 + q / rem, c = j + 1 + q % rem;
            if (c <= r) L[r][c] -= L[r][j] * L[c][j];
        }
        __syncthreads();
    }
    // inverse of the lower-triangular L: thread c solves L z = e_c (forward substitution)
    if (tid < NB) {
        const int c = tid;
        for (int r = 0; r < NB; r++) {
            double s = (r == c) ? 1.0 : 0.0;
            for (int t = c; t < r; t++) s -= L[r][t] * Z[t][c];
            Z[r][c] = (r < c) ? 0.0 : s / L[r][r];
        }
    }
    __syncthreads();
    for (int q = tid; q < NB * NB; q += 256) {
        const int r = q / NB, c = q % NB;
        if (r < nb && c <= r) A[(size_t) (k0 + r) * ld + (k0 + c)] = L[r][c];
        Linv[q] = Z[r][c];
    }
    if (tid == 0 && fail) sc->chol_fail = 1;
}

// ---- panel rows: X = B L^-T,  X[r][c] = sum_{t<=c} B[r][t] Linv[c][t] ---------------------------
__global__ void __launch_bounds__(256) chol_trsm_kernel(double *A, int ld, int nrows, int k0, int nb, const double *Linv)
{
    __shared__ double Li[NB][NB + 1];
    __shared__ double Bt[64][NB + 1];
    const int tid = threadIdx.x;
    const int r0 = k0 + nb + blockIdx.x * 64;
    for (int q = tid; q < NB * NB; q += 256) Li[q / NB][q % NB] = Linv[q];
    for (int q = tid; q < 64 * NB; q += 256) {
        const int r = q / NB, c = q % NB;
        Bt[r][c] = (r0 + r < nrows && c < nb) ? A[(size_t) (r0 + r) * ld + (k0 + c)] : 0.0;
    }
    __syncthreads();
    for (int q = tid; q < 64 * NB; q += 256) {
        const int r = q / NB, c = q % NB;
        if (r0 + r < nrows && c < nb) {
            double s = 0.0;
            for (int t = 0; t <= c; t++) s += Bt[r][t] * Li[c][t];
            A[(size_t) (r0 + r) * ld + (k0 + c)] = s;
        }
    }
}

// ---- trailing update: A[r][c] -= sum_{t in [kb,ke)} A[r][t] A[c][t]  for c in [cb,ce), r >= c -----
// square tiles BT x BT, (BT/TT)^2 = 256 threads, TT x TT outputs per thread, K chunks of 16.
template <int BT, int TT>
__global__ void __launch_bounds__(256) chol_syrk_kernel(double *A, int ld, int nrows, int cb, int ce, int kb, int ke)
{
    constexpr int BK = 16;
    __shared__ double As[BK][BT + 4];
    __shared__ double Bs[BK][BT + 4];
    const int c0 = cb + blockIdx.x * BT;
    const int r0 = cb + blockIdx.y * BT;
    if (r0 + BT <= c0) return;            // tile strictly above the diagonal band
    if (c0 >= ce || r0 >= nrows) return;
    const int tid = threadIdx.x;
    const int tx = tid % (BT / TT), ty = tid / (BT / TT);
    double acc[TT][TT];
#pragma unroll
    for (int a = 0; a < TT; a++)
#pragma unroll
        for (int b = 0; b < TT; b++) acc[a][b] = 0.0;
    for (int k0 = kb; k0 < ke; k0 += BK) {
        for (int q = tid; q < BT * BK; q += 256) {
            const int r = q / BK, t = q % BK;
            const int kk = k0 + t;
            As[t][r] = (r0 + r < nrows && kk < ke) ? A[(size_t) (r0 + r) * ld + kk] : 0.0;
            Bs[t][r] = (c0 + r < ce && kk < ke) ? A[(size_t) (c0 + r) * ld + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < BK; t++) {
            double a[TT], b[TT];
#pragma unroll
            for (int q = 0; q < TT; q++) { a[q] = As[t][ty * TT + q]; b[q] = Bs[t][tx * TT + q]; }
#pragma unroll
            for (int qa = 0; qa < TT; qa++)
#pragma unroll
                for (int qb = 0; qb < TT; qb++) acc[qa][qb] = fma(a[qa], b[qb], acc[qa][qb]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int qa = 0; qa < TT; qa++) {
        const int r = r0 + ty * TT + qa;
        if (r >= nrows) continue;
#pragma unroll
        for (int qb = 0; qb < TT; qb++) {
            const int c = c0 + tx * TT + qb;
            if (c < ce && c <= r) A[(size_t) r * ld + c] -= acc[qa][qb];
        }
    }
}

// ---- small systems: whole block-column panel factorised by ONE CTA in shared memory ---------------
// rows [k0, nrows) x cols [k0, k0+nb): potf2 of the diagonal block fused with the TRSM of every row
// below it (right-looking, unblocked inside the panel).  Used when the panel fits in shared memory.
constexpr int PANEL_LD = NB + 1;
__global__ void __launch_bounds__(1024) chol_panel_smem_kernel(double *A, int ld, int nrows, int k0, int nb, Scalars *sc)
{
    extern __shared__ double Pn[];   // [(nrows-k0)][PANEL_LD]
    __shared__ double djj_s;
    __shared__ int fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nr = nrows - k0;
    if (tid == 0) fail = 0;
    for (int q = tid; q < nr * nb; q += nt) {
        const int r = q / nb, c = q % nb;
        Pn[r * PANEL_LD + c] = (r >= nb || c <= r) ? A[(size_t) (k0 + r) * ld + (k0 + c)] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            const double d = Pn[j * PANEL_LD + j];
            if (!(d > 0.0) || !isfinite(d)) { fail = 1; djj_s = 1.0; }
            else djj_s = sqrt(d);
            Pn[j * PANEL_LD + j] = djj_s;
        }
        __syncthreads();
        const double djj = djj_s;
        for (int r = j + 1 + tid; r < nr; r += nt) Pn[r * PANEL_LD + j] = Pn[r * PANEL_LD + j] / djj;
        __syncthreads();
        const int ncols = nb - j - 1;
        if (ncols > 0) {
            const int total = (nr - j - 1) * ncols;
            for (int q = tid; q < total; q += nt) {
                const int r = j + 1 + q / ncols, c = j + 1 + q % ncols;
                if (c <= r) Pn[r * PANEL_LD + c] = fma(-Pn[r * PANEL_LD + j], Pn[c * PANEL_LD + j], Pn[r * PANEL_LD + c]);
            }
        }
        __syncthreads();
    }
    for (int q = tid; q < nr * nb; q += nt) {
        const int r = q / nb, c = q % nb;
        if (r >= nb || c <= r) A[(size_t) (k0 + r) * ld + (k0 + c)] = Pn[r * PANEL_LD + c];
    }
    if (tid == 0 && fail) sc->chol_fail = 1;
}

// ---- back substitution L^T x = y (y = row n of A), single CTA, row-oriented -----------------------
// thread c owns y_c (columns strided over the CTA); rows are consumed from the last to the first:
// x_i = y_i / L_ii, then y_c -= L_ic x_i for c < i (row i of L is contiguous => coalesced).
constexpr int BS_COLS = 9;   // columns per thread: supports n <= 9216 with 1024 threads
__global__ void __launch_bounds__(1024) chol_backsolve_kernel(const double *A, int ld, int n, double *x)
{
    __shared__ double xi_s[2];
    const int tid = threadIdx.x;
    const double *yrow = A + (size_t) n * ld;
    double y[BS_COLS];
#pragma unroll
    for (int q = 0; q < BS_COLS; q++) { const int c = tid + q * 1024; y[q] = (c < n) ? yrow[c] : 0.0; }
    for (int i = n - 1; i >= 0; i--) {
        const double *Li = A + (size_t) i * ld;
        // issue this row's loads before the dependent broadcast
        double l[BS_COLS];
#pragma unroll
        for (int q = 0; q < BS_COLS; q++) { const int c = tid + q * 1024; l[q] = (c < i) ? Li[c] : 0.0; }
        const int owner = i & 1023, oq = i >> 10;
        if (tid == owner) {
            double yi = 0.0;
#pragma unroll
            for (int q = 0; q < BS_COLS; q++) if (q == oq) yi = y[q];
            const double xi = yi / Li[i];
            xi_s[i & 1] = xi;
            x[i] = xi;
        }
        __syncthreads();
        const double xi = xi_s[i & 1];
#pragma unroll
        for (int q = 0; q < BS_COLS; q++) y[q] = fma(-l[q], xi, y[q]);
    }
}

// symmetric S (both triangles filled by the Schur kernel) + E in row n.  x receives the solution.
// linv_ws: ceil(n/32) * 1024 doubles.
int chol_solve(cudaStream_t st, double *A, int n, double *linv_ws, double *x, Scalars *sc)
{
    const int ld = n, nrows = n + 1;
    if (n > 1024 * BS_COLS) { set_error("reduced camera system of dimension %d exceeds the supported %d", n, 1024 * BS_COLS); return BSFM_ERR_UNSUPPORTED; }
    const size_t panel_bytes = (size_t) nrows * PANEL_LD * sizeof(double);
    const bool small = panel_bytes <= 200 * 1024;
    if (small) {
        static bool attr = false;
        if (!attr) { BSFM_CUDA_TRY(cudaFuncSetAttribute(chol_panel_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    }
    const int NBO = (n > 2048) ? 256 : NB;
    const int BT = (n <= 1024) ? 32 : (n <= 4096 ? 64 : 128);
    auto syrk = [&](int cb, int ce, int kb, int ke) -> int {
        if (cb >= ce) return BSFM_OK;
        const int tiles_c = (ce - cb + BT - 1) / BT;
        const int tiles_r = (nrows - cb + BT - 1) / BT;
        dim3 grid(tiles_c, tiles_r);
        if (BT == 32) chol_syrk_kernel<32, 2><<<grid, 256, 0, st>>>(A, ld, nrows, cb, ce, kb, ke);
        else if (BT == 64) chol_syrk_kernel<64, 4><<<grid, 256, 0, st>>>(A, ld, nrows, cb, ce, kb, ke);
        else chol_syrk_kernel<128, 8><<<grid, 256, 0, st>>>(A, ld, nrows, cb, ce, kb, ke);
        BSFM_KERNEL_CHECK();
        return BSFM_OK;
    };
    for (int K0 = 0; K0 < n; K0 += NBO) {
        const int K1 = min(n, K0 + NBO);
        for (int k0 = K0; k0 < K1; k0 += NB) {
            const int nb = min(NB, n - k0);
            if (small) {
                const int nr = nrows - k0;
                const int threads = nr * nb >= 8192 ? 1024 : (nr * nb >= 2048 ? 512 : 256);
                chol_panel_smem_kernel<<<1, threads, (size_t) nr * PANEL_LD * sizeof(double), st>>>(A, ld, nrows, k0, nb, sc);
                BSFM_KERNEL_CHECK();
            } else {
                double *Li = linv_ws + (size_t) (k0 / NB) * NB * NB;
                chol_diag_kernel<<<1, 256, 0, st>>>(A, ld, k0, nb, Li, sc);
                BSFM_KERNEL_CHECK();
                const int rows_below = nrows - (k0 + nb);
                if (rows_below > 0) {
                    chol_trsm_kernel<<<(rows_below + 63) / 64, 256, 0, st>>>(A, ld, nrows, k0, nb, Li);
                    BSFM_KERNEL_CHECK();
                }
            }
            // inner update restricted to the columns of the outer panel
            int rc = syrk(k0 + nb, K1, k0, k0 + nb);
            if (rc != BSFM_OK) return rc;
        }
        if (NBO != NB || true) {
            // outer trailing update with the whole outer panel (K = K1 - K0); when NBO == NB the inner
            // update above had an empty column range, so this is the only update.
            int rc = syrk(K1, n, K0, K1);
            if (rc != BSFM_OK) return rc;
            // RHS row (row n) against columns >= K1 is part of the tiles (nrows = n + 1)
        }
    }
    chol_backsolve_kernel<<<1, 1024, 0, st>>>(A, ld, n, x);
    BSFM_KERNEL_CHECK();
    return BSFM_OK;
}

}  // namespace ba
}  // namespace bsfm
