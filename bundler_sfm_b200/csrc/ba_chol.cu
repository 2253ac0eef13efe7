// ba_chol.cu -- dense SPD solve of the reduced camera system S da = E.
// Reference: sba_Axb_Chol = LAPACK dpotrf + dpotrs, lib/sba-1.5/sba_lapack.c:374-485.
//
// fp64 blocked right-looking Cholesky (lower, row-major) with the right-hand side carried as an
// extra matrix row (row n), so the forward substitution L y = E falls out of the factorisation.
//
// Small systems (n <= 1536, i.e. up to ~170 cameras; latency-bound):
//   per 32-column step : chol_step_kernel  one launch; every CTA redundantly factors the diagonal
//                        block fused with the row solves of the two panel blocks its 32x32 trailing
//                        tile needs, then applies the rank-32 update (critical path: 32 pivots)
//   end                : chol_backsolve_blocked_kernel (1 CTA) L^T x = y with inverted diagonal blocks
// Large systems (throughput-bound; the dense contraction of this path):
//   two-level blocking: inner panels of 32 columns, outer panels of 256 columns whose trailing update is
//   one large SYRK-shaped GEMM (A22 -= L21 L21^T)
//   per inner panel k : diag_kernel   (1 CTA)  potf2 of A_kk, L_kk^-1
//                       trsm_kernel   (rows)   A_rk <- A_rk L_kk^-T   for all rows below (incl. RHS row)
//                       syrk_kernel   (tiles)  columns inside the outer panel
//   per outer panel   : syrk_kernel   (tiles)  trailing matrix, K = 256
//   end               : chol_backsolve_kernel (1 CTA) row-oriented L^T x = y
#include "ba_kernels.cuh"
#include "common.h"

namespace bsfm {
namespace ba {

constexpr int NB = 32;
#ifdef BSFM_DEBUG_CLOCKS
__device__ long long g_dbg[64];
__device__ __forceinline__ long long dbg_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#define DBG_T(i) do { if (blockIdx.x == 1 && k == 0 && (threadIdx.x & 31) == 0) g_dbg[(i) + 5 * (threadIdx.x >> 5)] = dbg_clock(); } while (0)
#else
#define DBG_T(i) do { } while (0)
#endif

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

// ---- small systems, one launch per 32-column step ---------------------------------------------------
// Every CTA redundantly factors the 32x32 diagonal block (one warp, rows in registers, pivots and
// multipliers exchanged with shuffles), solves the two 32-row panel blocks its tile needs (one warp
// each, row in registers, L_kk broadcast from shared memory) and applies the rank-32 update to its
// 32x32 tile of the trailing matrix.  Redundant panel work costs no latency and removes the separate
// panel / TRSM launches: the critical path per step is potf2 -> trsm -> update inside one kernel.
// Row block index nbk (= one row) is the right-hand side carried along as matrix row n.
// (loops over shared memory, not unrolled register code: a kernel that runs ~10 us must not spend it
// fetching tens of KB of straight-line instructions)
__global__ void __launch_bounds__(256) chol_step_kernel(double *A, int ld, int n, int k, double *Linv_all, Scalars *sc)
{
    __shared__ double Lk[NB][NB + 1];
    __shared__ double Xr[NB][NB + 1];
    __shared__ double Xc[NB][NB + 1];
    __shared__ double Z[NB][NB + 1];
    __shared__ double dinv[NB];
    __shared__ int fail_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nbk = (n + NB - 1) / NB;           // matrix row blocks; block nbk = RHS row
    const int k0 = k * NB, nb = min(NB, n - k0);
    // tile of this CTA: blockIdx 0 = panel owner (no tile); else (cb, rb), k < cb <= rb <= nbk, cb < nbk
    int cb = -1, rb = nbk;                        // the owner solves the RHS row segment
    if (blockIdx.x > 0) {
        int t = blockIdx.x - 1;
        for (int c = k + 1; c < nbk; c++) {
            const int cnt = nbk - c + 1;
            if (t < cnt) { cb = c; rb = c + t; break; }
            t -= cnt;
        }
    }
    const int crow0 = cb * NB;
    const int rrows = (rb == nbk) ? 1 : min(NB, n - rb * NB);        // valid rows in the rb block
    const int rbase = (rb == nbk) ? n : rb * NB;                      // first matrix row of the rb block
    const int crows = (cb >= 0) ? min(NB, n - crow0) : 0;
    const bool need_c = cb >= 0 && cb != rb;

    DBG_T(0);
    if (tid == 0) fail_s = 0;
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e >> 5, c = e & 31;
        Lk[r][c] = (r < nb && c <= r) ? A[(size_t) (k0 + r) * ld + (k0 + c)] : ((r == c) ? 1.0 : 0.0);
        Xr[r][c] = (r < rrows && c < nb) ? A[(size_t) (rbase + r) * ld + (k0 + c)] : 0.0;
        if (need_c) Xc[r][c] = (r < crows && c < nb) ? A[(size_t) (crow0 + r) * ld + (k0 + c)] : 0.0;
    }
    __syncthreads();
    DBG_T(1);
    // potf2 of L_kk fused with the row solves X <- X L_kk^-T, right-looking, all 8 warps:
    //   thread (lane r, warp g) owns rows r and columns c = g, g+8, g+16, g+24.
    //   Every warp recomputes the pivot reciprocal and the multipliers l_r = a_rj / sqrt(d) of column j
    //   in registers (no communication), updates its own columns of L_kk, Xr and Xc, one barrier per pivot.
    //   Column j itself is left un-scaled in shared memory (it is read-only from now on) and rescaled by
    //   dinv[j] after the loop.  The pivot chain is the latency floor of the whole solve: rsqrt + multiply
    //   instead of sqrt + divide.
    for (int j = 0; j < nb; j++) {
        const double d = Lk[j][j];
        const double arj = Lk[lane][j];
        const double xr = Xr[lane][j];
        const double xc = need_c ? Xc[lane][j] : 0.0;
        const bool bad = !(d > 0.0) || !isfinite(d);
        const double rinv = bad ? 1.0 : rsqrt(d);
        const double l = (lane > j) ? arj * rinv : 0.0;
        const double xjr = xr * rinv, xjc = xc * rinv;
        if (tid == 0) { dinv[j] = rinv; if (bad) fail_s = 1; }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = warp + 8 * q;                      // warp-uniform
            if (c > j && c < nb) {
                const double lc = __shfl_sync(0xffffffffu, l, c);   // L[c][j]
                if (lane >= c) Lk[lane][c] = fma(-l, lc, Lk[lane][c]);
                Xr[lane][c] = fma(-xjr, lc, Xr[lane][c]);
                if (need_c) Xc[lane][c] = fma(-xjc, lc, Xc[lane][c]);
            }
        }
        __syncthreads();
    }
    if (tid < NB && tid >= nb) dinv[tid] = 1.0;
    __syncthreads();
    DBG_T(2);
    // rescale: L[r][c] = a_rc dinv[c] (c < r), L[c][c] = d_c dinv[c], X[r][c] *= dinv[c]
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e >> 5, c = e & 31;
        const double sc_c = dinv[c];
        if (c <= r && r < nb) Lk[r][c] *= sc_c;
        Xr[r][c] *= sc_c;
        if (need_c) Xc[r][c] *= sc_c;
    }
    __syncthreads();
    if (warp == 3 && blockIdx.x == 0) {
        // Z = L^-1 (lane c solves L z = e_c) for the blocked back substitution
        const int c = lane;
        for (int r = 0; r < NB; r++) {
            double sacc = (r == c) ? 1.0 : 0.0;
            for (int t = c; t < r; t++) sacc = fma(-Lk[r][t], Z[t][c], sacc);
            Z[r][c] = (r < c) ? 0.0 : sacc * dinv[r];
        }
    } else if (warp == 4 && blockIdx.x == 0) {
        for (int c = 0; c < nb; c++) if (lane < nb && c <= lane) A[(size_t) (k0 + lane) * ld + (k0 + c)] = Lk[lane][c];
        if (lane == 0 && fail_s) sc->chol_fail = 1;
    }
    __syncthreads();
    DBG_T(3);
    if (blockIdx.x == 0) {
        for (int e = tid; e < NB * NB; e += 256) Linv_all[(size_t) k * NB * NB + e] = Z[e >> 5][e & 31];
    }
    // panel write-back: the owner writes the RHS segment, first-column tiles write their row block
    if (blockIdx.x == 0 || cb == k + 1) {
        for (int e = tid; e < NB * NB; e += 256) {
            const int r = e >> 5, c = e & 31;
            if (r < rrows && c < nb) A[(size_t) (rbase + r) * ld + (k0 + c)] = Xr[r][c];
        }
    }
    if (cb < 0) return;
    // trailing tile (rb, cb):  A[r][c] -= sum_t Xr[r][t] Xc[c][t]   (c <= r on the diagonal tile)
    const double (*XC)[NB + 1] = need_c ? Xc : Xr;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int e = tid + q * 256;
        const int r = e >> 5, c = e & 31;
        if (r < rrows && c < crows && (rb != cb || c <= r)) {
            double acc = 0.0;
#pragma unroll 8
            for (int t = 0; t < NB; t++) acc = fma(Xr[r][t], XC[c][t], acc);
            A[(size_t) (rbase + r) * ld + (crow0 + c)] -= acc;
        }
    }
    DBG_T(4);
}

// ---- back substitution L^T x = y (y = row n of A) with inverted diagonal blocks, single CTA ---------
// per 32-row block (last to first): x_k = L_kk^-T y_k, then y[0:k0] -= L[k0:k0+nb, 0:k0]^T x_k.
// Two barriers per block instead of one per row.
__global__ void __launch_bounds__(512) chol_backsolve_blocked_kernel(const double *A, int ld, int n, const double *Linv_all, double *x)
{
    __shared__ double xk[NB];
    __shared__ double ys[NB];
    __shared__ double Li[NB][NB + 1];
    const int tid = threadIdx.x;
    const double *yrow = A + (size_t) n * ld;
    constexpr int COLS = 3;                     // n <= 1536
    double y[COLS];
#pragma unroll
    for (int q = 0; q < COLS; q++) { const int c = tid + q * 512; y[q] = (c < n) ? yrow[c] : 0.0; }
    const int nbk = (n + NB - 1) / NB;
    for (int kb = nbk - 1; kb >= 0; kb--) {
        const int k0 = kb * NB, nb = min(NB, n - k0);
        // everything that does not depend on x_k is issued first: L_kk^-1 block and this thread's
        // column of the 32 L rows (first 512 columns; the rare wider case is fetched later)
        const double z0 = Linv_all[(size_t) kb * NB * NB + tid], z1 = Linv_all[(size_t) kb * NB * NB + 512 + tid];
        double l[NB];
#pragma unroll
        for (int r = 0; r < NB; r++) l[r] = (tid < k0 && r < nb) ? A[(size_t) (k0 + r) * ld + tid] : 0.0;
#pragma unroll
        for (int q = 0; q < COLS; q++) { const int c = tid + q * 512; if (c >= k0 && c < k0 + nb) ys[c - k0] = y[q]; }
        Li[tid >> 5][tid & 31] = z0;
        Li[16 + (tid >> 5)][tid & 31] = z1;
        __syncthreads();
        // x_k = L_kk^-T y_k : 8 lanes per output, 4 terms each, shuffle-reduced
        if (tid < 256) {
            const int o = tid >> 3, part = tid & 7;
            double sacc = 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) { const int r = part * 4 + u; if (r >= o && r < nb) sacc = fma(Li[r][o], ys[r], sacc); }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
            if (part == 0) { xk[o] = sacc; if (o < nb) x[k0 + o] = sacc; }
        }
        __syncthreads();
        {
            double sacc = 0.0;
#pragma unroll
            for (int r = 0; r < NB; r++) sacc = fma(l[r], xk[r], sacc);
            y[0] -= sacc;
        }
#pragma unroll
        for (int q = 1; q < COLS; q++) {
            const int c = tid + q * 512;
            if (c < k0) {
                double sacc = 0.0;
#pragma unroll 8
                for (int r = 0; r < nb; r++) sacc = fma(A[(size_t) (k0 + r) * ld + c], xk[r], sacc);
                y[q] -= sacc;
            }
        }
    }
}

// ---- back substitution L^T x = y (y = row n of A), single CTA, row-oriented (large systems) --------
// thread c owns y_c; rows are consumed from the last to the first in batches whose L entries are
// prefetched (double-buffered) so the per-row critical path is one broadcast + one FMA:
// x_i = y_i / L_ii, then y_c -= L_ic x_i for c < i (row i of L is contiguous => coalesced).
template <int COLS, int BS_BATCH>
__global__ void __launch_bounds__(1024) chol_backsolve_kernel(const double *A, int ld, int n, double *x)
{
    __shared__ double xi_s[2];
    const int tid = threadIdx.x;
    const double *yrow = A + (size_t) n * ld;
    double y[COLS];
#pragma unroll
    for (int q = 0; q < COLS; q++) { const int c = tid + q * 1024; y[q] = (c < n) ? yrow[c] : 0.0; }
    double cur[COLS][BS_BATCH], nxt[COLS][BS_BATCH];
    auto fetch = [&](double (&dst)[COLS][BS_BATCH], int itop) {
#pragma unroll
        for (int b = 0; b < BS_BATCH; b++) {
            const int i = itop - b;
#pragma unroll
            for (int q = 0; q < COLS; q++) { const int c = tid + q * 1024; dst[q][b] = (i >= 0 && c <= i) ? A[(size_t) i * ld + c] : 0.0; }
        }
    };
    fetch(cur, n - 1);
    for (int itop = n - 1; itop >= 0; itop -= BS_BATCH) {
        fetch(nxt, itop - BS_BATCH);
#pragma unroll
        for (int b = 0; b < BS_BATCH; b++) {
            const int i = itop - b;
            if (i < 0) break;
            const int owner = i & 1023, oq = i >> 10;
            if (tid == owner) {
                double yi = 0.0, lii = 1.0;
#pragma unroll
                for (int q = 0; q < COLS; q++) if (q == oq) { yi = y[q]; lii = cur[q][b]; }
                const double xi = yi / lii;
                xi_s[i & 1] = xi;
                x[i] = xi;
            }
            __syncthreads();
            const double xi = xi_s[i & 1];
#pragma unroll
            for (int q = 0; q < COLS; q++) { const int c = tid + q * 1024; if (c < i) y[q] = fma(-cur[q][b], xi, y[q]); }
        }
#pragma unroll
        for (int q = 0; q < COLS; q++)
#pragma unroll
            for (int b = 0; b < BS_BATCH; b++) cur[q][b] = nxt[q][b];
    }
}

// symmetric S (both triangles filled by the Schur kernel) + E in row n.  x receives the solution.
// linv_ws: ceil(n/32) * 1024 doubles.
int chol_solve(cudaStream_t st, double *A, int n, double *linv_ws, double *x, Scalars *sc)
{
    const int ld = n, nrows = n + 1;
    if (n > 1024 * 9) { set_error("reduced camera system of dimension %d exceeds the supported %d", n, 1024 * 9); return BSFM_ERR_UNSUPPORTED; }
    auto backsolve = [&]() -> int {
        if (n <= 4096) chol_backsolve_kernel<4, 2><<<1, 1024, 0, st>>>(A, ld, n, x);
        else chol_backsolve_kernel<9, 1><<<1, 1024, 0, st>>>(A, ld, n, x);
        BSFM_KERNEL_CHECK();
        return BSFM_OK;
    };
    if (n <= 1536) {
        // small reduced systems (<= ~170 cameras): one fused launch per 32-column step
        const int nbk = (n + NB - 1) / NB;
        for (int k = 0; k < nbk; k++) {
            const int R = nbk - 1 - k;                      // remaining column blocks
            const int tiles = R * (R + 1) / 2 + R;          // (rb, cb) incl. the RHS row block
            chol_step_kernel<<<1 + tiles, 256, 0, st>>>(A, ld, n, k, linv_ws, sc);
            BSFM_KERNEL_CHECK();
        }
        chol_backsolve_blocked_kernel<<<1, 512, 0, st>>>(A, ld, n, linv_ws, x);
        BSFM_KERNEL_CHECK();
        return BSFM_OK;
    }
    const bool small = false;
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
    return backsolve();
}

}  // namespace ba
}  // namespace bsfm

#ifdef BSFM_DEBUG_CLOCKS
extern "C" int bsfm_debug_read(long long *out64)
{
    return cudaMemcpyFromSymbol(out64, bsfm::ba::g_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif
