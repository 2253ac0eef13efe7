// ba_chol_large.cu -- dense SPD solve of LARGE reduced camera systems (n > 1536; config 3: 1000 cameras -> 9000^2).
// Reference: sba_Axb_Chol = LAPACK dpotrf + dpotrs, lib/sba-1.5/sba_lapack.c:374-485.
//
// Right-looking blocked Cholesky (lower, row-major, factor OUT OF PLACE in Lmat, right-hand side carried as matrix
// row n so the forward substitution falls out of the panel solves), outer panels of 256 columns, FOUR launches per
// panel:
//   chol_diag_kernel      1 CTA : fp64 factorisation of the 256 x 256 diagonal block (32-column sub-steps: one warp
//                                 factors the 32 x 32 block in REGISTERS with shuffles, thread-per-row solves, rank-32
//                                 in-block update) + the inverses of its eight 32 x 32 diagonal sub-blocks
//   chol_trsm_kernel      64 rows x 256 columns per CTA : L_panel = A_panel L_kk^-T as a blocked forward substitution
//                                 (small GEMMs against the sub-block inverses); optionally emits the int8 slices the
//                                 tcgen05 trailing update consumes (ba_chol_tc.cu)
//   trailing update       A -= L_panel L_panel^T, K = 256: tcgen05 int8-slice (Ozaki) kernel or fp64 DMMA kernel
//   (after the last panel) chol_backsolve_big_kernel per panel, last to first: L^T x = y
// The small-system path (ba_chol.cu: one fused launch per 32 columns) stays as it is: it is latency-bound, this one is
// throughput-bound.
#include "ba_kernels.cuh"
#include "ba_chol_large.cuh"
#include "ba_chol_potf2.cuh"
#include "common.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <vector>

namespace bsfm {
namespace ba {

// fp64 tensor-core tile product (DMMA): D(8x8) += A(8x4) B(4x8); lane = 4 g + tg holds A[g][tg], B[tg][g] (= B^T[g][tg]),
// D[g][2 tg], D[g][2 tg + 1]
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- 256 x 256 diagonal block ----------------------------------------------------------------------------------
// One CTA of 256 threads works on the block in global memory (it is L2-resident: 512 KB); per 32-column sub-step:
//   warp 0           : potf2 of the 32 x 32 block, lane = row, the row lives in registers, pivots / multipliers travel
//                      by shuffle (no barrier inside the 32-pivot chain)
//   threads 0..223   : one panel row each (rows below the sub-block inside the diagonal block): x <- x L_ss^-T
//   last warp        : Z = L_ss^-1 (lane = column), needed by chol_trsm_kernel
//   all warps        : rank-32 update of the rest of the diagonal block on the fp64 tensor cores, one 32 x 32 tile
//                      per warp and turn (DMMA: 8 shared-memory loads per 16 MMAs; a scalar FMA loop was bound by
//                      its shared-memory loads)
// Shared-memory row pitch 36 doubles: conflict-free for the DMMA fragment pattern [row g][column tg].
constexpr int DG_LD = LNB + 4;
constexpr int DG_XROWS = LNBO - LNB;          // 224 panel rows at most inside the diagonal block
constexpr int DG_SMEM_DOUBLES = (3 * LNB + DG_XROWS) * DG_LD + LNB;
static_assert(DG_LD == TP, "tile primitives of ba_chol_potf2.cuh use a pitch of 36 doubles");
constexpr int DG_THREADS = 256;    // 255 registers per thread: the register-resident 32-wide rows need them
__global__ void __launch_bounds__(DG_THREADS, 1) chol_diag_kernel(double *A, double *Lout, int ld, int n, int k0, double *linv_all, Scalars *sc, long long *dbg)
{
#define DG_T(i) do { if (dbg && tid == 0) { const long long now = clock64(); dbg[i] += now - tlast; tlast = now; } } while (0)
    long long tlast = dbg ? clock64() : 0;
    extern __shared__ __align__(16) double dsm[];
    double (*Ls)[DG_LD] = reinterpret_cast<double (*)[DG_LD]>(dsm);
    double (*Zs)[DG_LD] = reinterpret_cast<double (*)[DG_LD]>(dsm + LNB * DG_LD);
    double (*Ws)[DG_LD] = reinterpret_cast<double (*)[DG_LD]>(dsm + 2 * LNB * DG_LD);
    double (*Xs)[DG_LD] = reinterpret_cast<double (*)[DG_LD]>(dsm + 3 * LNB * DG_LD);
    double *dinv = dsm + (3 * LNB + DG_XROWS) * DG_LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tg = lane & 3;
    const int nb = min(LNBO, n - k0);
    const int nsub = (nb + LNB - 1) / LNB;
    for (int e = tid; e < LNB * DG_LD; e += DG_THREADS) Zs[0][e] = 0.0;       // the strictly upper blocks of Z stay zero
    for (int s = 0; s < nsub; s++) {
        const int c0 = k0 + s * LNB;                      // first matrix column of the sub-block
        const int w = min(LNB, k0 + nb - c0);             // its width
        const int below = max(0, k0 + nb - (c0 + LNB));   // rows below it inside the diagonal block
        const int nblk = (below + LNB - 1) / LNB;
        for (int e = tid; e < LNB * LNB; e += DG_THREADS) {
            const int r = e >> 5, c = e & 31;
            Ls[r][c] = (r < w && c <= r) ? A[(size_t) (c0 + r) * ld + (c0 + c)] : ((r == c) ? 1.0 : 0.0);
        }
        for (int e0 = 0; e0 < nblk * LNB * LNB; e0 += DG_THREADS * 8) {      // 8 independent loads per thread in flight
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int e = e0 + tid + i * DG_THREADS, r = e >> 5, c = e & 31;
                v[i] = (r < below && c < w) ? A[(size_t) (c0 + LNB + r) * ld + (c0 + c)] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int e = e0 + tid + i * DG_THREADS;
                if (e < nblk * LNB * LNB) Xs[e >> 5][e & 31] = v[i];
            }
        }
        __syncthreads();
        DG_T(0);
        if (warp == 0) {
            // potf2 of the 32 x 32 block (ba_chol_potf2.cuh): lane = row, 8-column register panels, DMMA rank-8 updates
            const bool bad = warp_potf2_32_tc(Ls, dinv, lane, true);
            if (bad && lane == 0) sc->chol_fail = 1;
            DG_T(5);                                            // dev accounting: potf2 alone; slot 1 is then the wait for the inverse
        } else if (warp == 1) {
            const long long t1 = dbg ? clock64() : 0;
            warp_tile_inverse(Ls, dinv, Zs, Ws, lane);          // Z = L_ss^-1 assembled one panel behind warp 0
            if (dbg && lane == 0) dbg[6] += clock64() - t1;
        }
        __syncthreads();
        DG_T(1);
        // panel rows below the sub-block: X <- X L_ss^-T = X Z^T, 8-row strips on the fp64 tensor cores
        for (int strip = warp; strip * 8 < below; strip += DG_THREADS / 32) warp_rows_times_ZT(Xs, strip * 8, Zs, lane);
        __syncthreads();
        DG_T(2);
        for (int e = tid; e < LNB * LNB; e += DG_THREADS) {
            const int r = e >> 5, c = e & 31;
            if (r < w && c <= r) Lout[(size_t) (c0 + r) * ld + (c0 + c)] = Ls[r][c];
            linv_all[(size_t) (c0 / LNB) * LNB * LNB + e] = (c <= r) ? Zs[r][c] : 0.0;
        }
        for (int e = tid; e < below * LNB; e += DG_THREADS) {
            const int r = e >> 5, c = e & 31;
            if (c < w) Lout[(size_t) (c0 + LNB + r) * ld + (c0 + c)] = Xs[r][c];
        }
        DG_T(3);
        // rank-32 update of the remaining block: tile (bi, bj), bj <= bi, on the fp64 tensor cores
        for (int blk = warp; blk < nblk * (nblk + 1) / 2; blk += DG_THREADS / 32) {
            int bi = 0, rem = blk;
            while (rem > bi) { rem -= bi + 1; bi++; }
            const int bj = rem;
            // the tile's current values are fetched first: the independent loads stay in flight under the MMAs
            double acc[4][4][2], cur[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int gr = bi * LNB + i * 8 + g;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int gc = bj * LNB + j * 8 + 2 * tg;
                    const double *src = A + (size_t) (c0 + LNB + gr) * ld + (c0 + LNB + gc);
                    cur[i][j][0] = (gr < below && gc <= gr) ? src[0] : 0.0;
                    cur[i][j][1] = (gr < below && gc + 1 <= gr) ? src[1] : 0.0;
                    acc[i][j][0] = 0.0; acc[i][j][1] = 0.0;
                }
            }
#pragma unroll
            for (int ks = 0; ks < LNB; ks += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; i++) af[i] = Xs[bi * LNB + i * 8 + g][ks + tg];
#pragma unroll
                for (int j = 0; j < 4; j++) bf[j] = Xs[bj * LNB + j * 8 + g][ks + tg];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int gr = bi * LNB + i * 8 + g;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int gc = bj * LNB + j * 8 + 2 * tg;
                    double *dst = A + (size_t) (c0 + LNB + gr) * ld + (c0 + LNB + gc);
                    if (gr < below && gc <= gr) dst[0] = cur[i][j][0] - acc[i][j][0];
                    if (gr < below && gc + 1 <= gr) dst[1] = cur[i][j][1] - acc[i][j][1];
                }
            }
        }
        __syncthreads();
        DG_T(4);
    }
#undef DG_T
}

// ---- panel solve: rows below the diagonal block (and the right-hand side row n) ----------------------------------
// X_s = (B_s - sum_{t<s} X_t L_st^T) Z_ss^T for the eight 32-column sub-blocks s, 64 rows per CTA, X kept in shared
// memory, every product on the fp64 tensor cores (DMMA): warp (wr = warp & 3, wc = warp >> 2) owns the 16 x 16 piece
// [16 wr, +16) x [16 wc, +16) of the current 64 x 32 sub-block as 2 x 2 accumulator tiles.
// Source rows: matrix rows of A below the diagonal block (ident == 0), or rows of the 256 x 256 identity (ident == 1:
// the result is L_kk^-T, the inverse the back substitution multiplies with; blockIdx.y = panel).
constexpr int TR_ROWS = 64;
constexpr int TR_LDX = LNBO + 4;       // pitch = 4 (mod 16) doubles: conflict-free DMMA fragment loads
constexpr int TR_LDB = LNB + 4;
constexpr int TR_NLB = LNBO / LNB;      // blocks staged per sub-step: L_s0 .. L_s,s-1 and Z_ss
constexpr int TR_SMEM_DOUBLES = TR_ROWS * TR_LDX + TR_NLB * LNB * TR_LDB + TR_ROWS;
struct TrsmArgs {
    const double *A;        // ident == 0: source rows (matrix, pitch ld)
    double *dst;            // destination rows: dst[(dst_row0 + r) * dst_ld + dst_col0 + c]
    const double *Lfac;     // factor matrix (the panel's diagonal block is read from it)
    const double *linv_all;
    int ld, dst_ld, nrows, ident;
    int k0, nb;             // ident == 0: the panel.  ident == 1: panel = blockIdx.y * 256
    int n;
};
__global__ void __launch_bounds__(256) chol_trsm_kernel(const TrsmArgs T, SliceOut so)
{
    extern __shared__ double dsm[];
    double *Xs = dsm;                                                                     // [64][260]
    double (*Lb)[LNB][TR_LDB] = reinterpret_cast<double (*)[LNB][TR_LDB]>(dsm + TR_ROWS * TR_LDX);               // [8][32][36]
    unsigned long long *rmax = reinterpret_cast<unsigned long long *>(dsm + TR_ROWS * TR_LDX + TR_NLB * LNB * TR_LDB);   // [64] bits of the row maxima
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tg = lane & 3;
    const int k0 = T.ident ? (int) blockIdx.y * LNBO : T.k0;
    const int nb = T.ident ? min(LNBO, T.n - k0) : T.nb;
    const int row0 = T.ident ? (int) blockIdx.x * TR_ROWS : k0 + nb + (int) blockIdx.x * TR_ROWS;   // ident: row of the identity block
    const int rcount = T.ident ? max(0, min(TR_ROWS, nb - row0)) : min(TR_ROWS, T.nrows - row0);
    if (rcount <= 0) return;
    const double *Lfac = T.Lfac;
    const int ld = T.ld;
    for (int r0 = 0; r0 < TR_ROWS; r0 += 32) {       // thread = column, 32 rows per batch: 32 independent loads in flight
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int r = r0 + i, c = tid;
            v[i] = 0.0;
            if (r < rcount && c < nb) v[i] = T.ident ? ((row0 + r == c) ? 1.0 : 0.0) : T.A[(size_t) (row0 + r) * ld + (k0 + c)];
        }
#pragma unroll
        for (int i = 0; i < 32; i++) Xs[(r0 + i) * TR_LDX + tid] = v[i];
    }
    if (tid < TR_ROWS) rmax[tid] = 0ull;
    const int nsub = (nb + LNB - 1) / LNB;
    const int wr = warp & 3, wc = warp >> 2;
    const double *xrow0 = Xs + (wr * 16 + g) * TR_LDX, *xrow1 = xrow0 + 8 * TR_LDX;
    double pre[4 * TR_NLB];      // the blocks of the NEXT sub-step, 4 elements of each per thread
    auto trsm_prefetch = [&](int sn) {
        const int wn = min(LNB, nb - sn * LNB);
#pragma unroll
        for (int t = 0; t < TR_NLB; t++) {
            if (t > sn) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = tid + 256 * j, c = e >> 5, u = e & 31;
                double v = 0.0;
                if (t < sn) { if (c < wn) v = Lfac[(size_t) (k0 + sn * LNB + c) * ld + (k0 + t * LNB + u)]; }
                else v = T.linv_all[(size_t) (k0 / LNB + sn) * LNB * LNB + e];
                pre[4 * t + j] = v;
            }
        }
    };
    trsm_prefetch(0);
    __syncthreads();
    for (int s = 0; s < nsub; s++) {
        const int w = min(LNB, nb - s * LNB);
        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const double *src = Xs + (wr * 16 + i * 8 + g) * TR_LDX + s * LNB + wc * 16 + j * 8 + 2 * tg;
                acc[i][j][0] = src[0]; acc[i][j][1] = src[1];
            }
        // staging: the blocks L_s0 .. L_s,s-1 and Z_ss were requested one sub-step ahead (they depend on nothing computed here),
        // so their L2 latency hides under the previous sub-step's MMAs instead of costing up to four round trips per sub-step
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TR_NLB; t++)
            if (t <= s)
#pragma unroll
                for (int j = 0; j < 4; j++) { const int e = tid + 256 * j; Lb[t][e >> 5][e & 31] = pre[4 * t + j]; }
        __syncthreads();
        if (s + 1 < nsub) trsm_prefetch(s + 1);
        for (int t = 0; t < s; t++) {
#pragma unroll
            for (int ks = 0; ks < LNB; ks += 4) {
                const double a0 = -xrow0[t * LNB + ks + tg], a1 = -xrow1[t * LNB + ks + tg];
                const double b0 = Lb[t][wc * 16 + g][ks + tg], b1 = Lb[t][wc * 16 + 8 + g][ks + tg];
                dmma_m8n8k4(acc[0][0][0], acc[0][0][1], a0, b0); dmma_m8n8k4(acc[0][1][0], acc[0][1][1], a0, b1);
                dmma_m8n8k4(acc[1][0][0], acc[1][0][1], a1, b0); dmma_m8n8k4(acc[1][1][0], acc[1][1][1], a1, b1);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double *dstp = Xs + (wr * 16 + i * 8 + g) * TR_LDX + s * LNB + wc * 16 + j * 8 + 2 * tg;
                dstp[0] = acc[i][j][0]; dstp[1] = acc[i][j][1];
            }
        __syncthreads();
        double out[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) { out[i][j][0] = 0.0; out[i][j][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < LNB; ks += 4) {      // X_s = Y Z^T (Z lower triangular: Z[c][u] = 0 for u > c)
            const double a0 = xrow0[s * LNB + ks + tg], a1 = xrow1[s * LNB + ks + tg];
            const double b0 = Lb[s][wc * 16 + g][ks + tg], b1 = Lb[s][wc * 16 + 8 + g][ks + tg];
            dmma_m8n8k4(out[0][0][0], out[0][0][1], a0, b0); dmma_m8n8k4(out[0][1][0], out[0][1][1], a0, b1);
            dmma_m8n8k4(out[1][0][0], out[1][0][1], a1, b0); dmma_m8n8k4(out[1][1][0], out[1][1][1], a1, b1);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; i++) {
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = wc * 16 + j * 8 + 2 * tg;
                double *dstp = Xs + (wr * 16 + i * 8 + g) * TR_LDX + s * LNB + c;
                const double v0 = (c < w) ? out[i][j][0] : 0.0, v1 = (c + 1 < w) ? out[i][j][1] : 0.0;
                dstp[0] = v0; dstp[1] = v1;
                m = fmax(m, fmax(fabs(v0), fabs(v1)));
            }
            if (so.slices) {      // row maxima for the int8 slicing: non-negative doubles order like their bit patterns
                m = fmax(m, shfl_d(m, lane ^ 1));
                m = fmax(m, shfl_d(m, lane ^ 2));
                if (tg == 0) atomicMax(&rmax[wr * 16 + i * 8 + g], (unsigned long long) __double_as_longlong(m));
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < TR_ROWS * LNBO; e += 256) {
        const int rr = e >> 8, c = e & 255;
        if (rr < rcount && c < nb) {
            if (T.ident) T.dst[((size_t) blockIdx.y * LNBO + row0 + rr) * LNBO + c] = Xs[rr * TR_LDX + c];
            else T.dst[(size_t) (row0 + rr) * T.dst_ld + (k0 + c)] = Xs[rr * TR_LDX + c];
        }
    }
    if (so.slices) emit_slices(so, Xs, TR_LDX, row0, rcount, rmax);
}

// ---- back substitution L^T x = y, one launch per outer panel (last to first) --------------------------------------
// y lives in row n of the factor matrix and is updated in place.  Every CTA first forms the panel's own 256 unknowns
// x_p = L_pp^-T y_p with the inverse computed after the factorisation (Xinv = L_pp^-T, row-major 256 x 256, upper
// triangular; one warp per row, coalesced), then CTA b removes their contribution from 64 of the remaining right-hand
// side entries:  y[c] -= sum_r L[k0 + r][c] x[r].
__global__ void __launch_bounds__(1024) chol_backsolve_big_kernel(double *L, int ld, int n, int k0, int nb, const double *Xinv_all, double *x)
{
    __shared__ double ys[LNBO];
    __shared__ double xs[LNBO];
    __shared__ double part[16][64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *y = L + (size_t) n * ld;
    if (tid < LNBO) ys[tid] = (tid < nb) ? y[k0 + tid] : 0.0;
    __syncthreads();
    const double *Xi = Xinv_all + (size_t) (k0 / LNBO) * LNBO * LNBO;
    {   // warp = 8 rows of x_p; all 64 loads of a lane are independent (one memory round trip)
        const int i0 = warp * 8;
        double a[8];
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int c = lane + 32 * q;
            const double yv = ys[c];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u;
                const double zv = (c >= i && c < nb && i < nb) ? Xi[(size_t) i * LNBO + c] : 0.0;
                a[u] = fma(zv, yv, a[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[u] += shfl_d(a[u], lane ^ o);
            if (lane == 0) { xs[i0 + u] = a[u]; if (blockIdx.x == 0 && i0 + u < nb) x[k0 + i0 + u] = a[u]; }
        }
    }
    __syncthreads();
    const int col = blockIdx.x * 64 + (tid & 63), grp = tid >> 6;      // 16 row groups of 16
    double acc = 0.0;
    if (col < k0) {
        const int r_lo = grp * 16;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int rr = r_lo + q;
            if (rr < nb) acc = fma(L[(size_t) (k0 + rr) * ld + col], xs[rr], acc);
        }
    }
    part[grp][tid & 63] = acc;
    __syncthreads();
    if (tid < 64 && col < k0) {
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < 16; q++) sacc += part[q][tid];
        y[col] -= sacc;
    }
}

__global__ void __launch_bounds__(256) chol_syrk_dmma_kernel(double *A, const double *L, int ld, int nrows, int cb, int ce, int kb, int ke);

// ---- optional per-kernel-class device timing (bench.py's roofline of the dominant kernel; off by default) --------
namespace {
struct CholProfile {
    bool on = false;
    float ms[4] = {0, 0, 0, 0};          // diag, trsm, trailing update, back substitution
    int launches[4] = {0, 0, 0, 0};
    double int8_ops = 0.0, fp64_flops = 0.0;      // trailing update only
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t get() { if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); } return pool[used++]; }
    void begin(int cls, cudaStream_t st) { if (!on) return; Span sp{cls, get(), get()}; cudaEventRecord(sp.a, st); spans.push_back(sp); }
    void end(cudaStream_t st) { if (!on) return; cudaEventRecord(spans.back().b, st); launches[spans.back().cls]++; }
    void collect()
    {
        for (auto &sp : spans) { cudaEventSynchronize(sp.b); float t = 0; cudaEventElapsedTime(&t, sp.a, sp.b); ms[sp.cls] += t; }
        spans.clear(); used = 0;
    }
};
thread_local CholProfile g_prof;
}  // namespace

static long long *g_diag_dbg = nullptr;     // dev-only (BSFM_DIAG_PROF=1): cycles of the first panel's diagonal-block phases
static int set_smem_attr_once(int dev)
{
    static std::atomic<int> done[64];
    if (dev < 0 || dev >= 64) return BSFM_OK;
    if (done[dev].load(std::memory_order_acquire)) return BSFM_OK;
    BSFM_CUDA_TRY(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (DG_SMEM_DOUBLES * sizeof(double))));
    BSFM_CUDA_TRY(cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (TR_SMEM_DOUBLES * sizeof(double))));
    done[dev].store(1, std::memory_order_release);
    return BSFM_OK;
}

// Same contract as chol_solve (ba_chol.cu): symmetric A (both triangles) + right-hand side in row n; Lmat (n+1) x n;
// linv_ws ceil(n/32) * 1024 doubles; xinv_ws ceil(n/256) * 65536 doubles (L_pp^-T of every panel); x receives the solution.  `ws` = int8-slice workspace of the tensor-core
// trailing update (null: fp64 DMMA update).
int chol_solve_large(cudaStream_t st, double *A, double *Lmat, int n, double *linv_ws, double *xinv_ws, double *x, Scalars *sc, const TcWorkspace *ws)
{
    const int ld = n, nrows = n + 1;
    int dev = 0;
    BSFM_CUDA_TRY(cudaGetDevice(&dev));
    { int rc = set_smem_attr_once(dev); if (rc != BSFM_OK) return rc; }
    const bool use_tc = ws && ws->slices && tc_syrk_available();
    static const bool diag_prof = getenv("BSFM_DIAG_PROF") != nullptr;
    if (diag_prof && !g_diag_dbg) { cudaMalloc(&g_diag_dbg, 8 * sizeof(long long)); cudaMemset(g_diag_dbg, 0, 8 * sizeof(long long)); }
    // Look-ahead (tensor-core path): the diagonal block of panel p+1 is a ~0.15 ms serial chain on ONE SM.  The trailing update
    // of panel p is therefore split into the two tile columns panel p+1 lives in (launched first) and the rest, and
    // diag(p+1) runs on a side stream next to the rest, which leaves it one SM (grid = SMs - 1):
    //   main : trsm(p)  col(p) -E1->  rest(p) ............ -wait E2->  trsm(p+1) ...
    //   side :                 wait E1  diag(p+1) -E2->
    // BSFM_BA_CHOL_LOOKAHEAD=0 serialises everything on the main stream (also the fp64 DMMA fallback).
    static const bool lookahead_env = []() { const char *e = getenv("BSFM_BA_CHOL_LOOKAHEAD"); return !(e && e[0] == '0'); }();
    const bool lookahead = use_tc && lookahead_env;
    struct SideStream { int device = -1; cudaStream_t st = nullptr; cudaEvent_t e_col = nullptr, e_diag = nullptr; };
    static thread_local SideStream side;
    if (lookahead && side.device != dev) {
        if (side.st) { cudaStreamDestroy(side.st); cudaEventDestroy(side.e_col); cudaEventDestroy(side.e_diag); side = SideStream(); }
        BSFM_CUDA_TRY(cudaStreamCreateWithFlags(&side.st, cudaStreamNonBlocking));
        BSFM_CUDA_TRY(cudaEventCreateWithFlags(&side.e_col, cudaEventDisableTiming));
        BSFM_CUDA_TRY(cudaEventCreateWithFlags(&side.e_diag, cudaEventDisableTiming));
        side.device = dev;
    }
    const int sms = tc_sm_count();
    bool diag_done = false;       // diag(p) already issued by the look-ahead of panel p-1
    // The last panels are latency-bound (a 110 us diagonal chain per panel that no trailing update hides any more): once the
    // remaining matrix fits the one-launch dataflow factorisation (ba_chol_dataflow.cu, <= 640), it takes over in place -- same
    // factor blocks, same 32 x 32 inverses, so the batched panel inverses and the back substitution below do not notice.
    static const int tail_max = []() { const char *e = getenv("BSFM_BA_CHOL_TAIL"); return e ? atoi(e) : 640; }();
    auto tail_fits = [&](int kk) { return kk > 0 && kk < n && n - kk <= tail_max && n - kk > 2 * LNB; };
    for (int k0 = 0; k0 < n; k0 += LNBO) {
        const int nb = std::min(LNBO, n - k0);
        const int k1 = k0 + nb;
        if (!diag_done && tail_fits(k0)) {
            bool used = false;
            g_prof.begin(0, st);
            int rc = chol_dataflow_factor(st, A + (size_t) k0 * ld + k0, Lmat + (size_t) k0 * ld + k0, ld, n - k0, linv_ws + (size_t) (k0 / LNB) * LNB * LNB,
                                          linv_ws + chol_large_pub_offset(n), sc, &used);
            g_prof.end(st);
            if (rc != BSFM_OK) return rc;
            if (used) break;
        }
        if (!diag_done) {
            g_prof.begin(0, st);
            chol_diag_kernel<<<1, DG_THREADS, DG_SMEM_DOUBLES * sizeof(double), st>>>(A, Lmat, ld, n, k0, linv_ws, sc, k0 == 0 ? g_diag_dbg : nullptr);
            BSFM_KERNEL_CHECK();
            g_prof.end(st);
        } else {
            BSFM_CUDA_TRY(cudaStreamWaitEvent(st, side.e_diag, 0));
        }
        diag_done = false;
        const int rows_below = nrows - k1;     // >= 1: the right-hand side row
        SliceOut so = {};
        if (use_tc && k1 < n) so = tc_slice_out(*ws, k1, n);
        g_prof.begin(1, st);
        TrsmArgs T = {};
        T.A = A; T.dst = Lmat; T.Lfac = Lmat; T.linv_all = linv_ws; T.ld = ld; T.dst_ld = ld; T.nrows = nrows; T.ident = 0; T.k0 = k0; T.nb = nb; T.n = n;
        chol_trsm_kernel<<<(rows_below + TR_ROWS - 1) / TR_ROWS, 256, TR_SMEM_DOUBLES * sizeof(double), st>>>(T, so);
        BSFM_KERNEL_CHECK();
        g_prof.end(st);
        if (k1 < n) {
            g_prof.begin(2, st);
            if (g_prof.on) {
                const double tr = (double) (n - k1), pairs = tr * (tr + 1.0) * 0.5 + tr;     // lower triangle + right-hand side row
                g_prof.fp64_flops += 2.0 * pairs * nb;
                if (use_tc) g_prof.int8_ops += 2.0 * pairs * nb * (ws->ns * (ws->ns + 1) / 2);
            }
            if (use_tc && lookahead && !tail_fits(k1)) {
                const int col_tiles = LNBO / TC_TILE;      // the tile columns of the next panel
                int rc = tc_syrk_update(st, *ws, A, Lmat, ld, nrows, k1, n, k0, k1, 0, col_tiles, 0);
                if (rc != BSFM_OK) return rc;
                BSFM_CUDA_TRY(cudaEventRecord(side.e_col, st));
                BSFM_CUDA_TRY(cudaStreamWaitEvent(side.st, side.e_col, 0));
                chol_diag_kernel<<<1, DG_THREADS, DG_SMEM_DOUBLES * sizeof(double), side.st>>>(A, Lmat, ld, n, k1, linv_ws, sc, nullptr);
                BSFM_KERNEL_CHECK();
                BSFM_CUDA_TRY(cudaEventRecord(side.e_diag, side.st));
                diag_done = true;
                if (g_prof.on) g_prof.launches[0]++;
                rc = tc_syrk_update(st, *ws, A, Lmat, ld, nrows, k1, n, k0, k1, col_tiles, -1, std::max(1, sms - 1));
                if (rc != BSFM_OK) return rc;
            } else if (use_tc) {
                int rc = tc_syrk_update(st, *ws, A, Lmat, ld, nrows, k1, n, k0, k1);
                if (rc != BSFM_OK) return rc;
            } else {
                const int BT = 128;
                dim3 grid((n - k1 + BT - 1) / BT, (nrows - k1 + BT - 1) / BT);
                chol_syrk_dmma_kernel<<<grid, 256, 0, st>>>(A, Lmat, ld, nrows, k1, n, k0, k1);
                BSFM_KERNEL_CHECK();
            }
            g_prof.end(st);
        }
    }
    if (diag_done) BSFM_CUDA_TRY(cudaStreamWaitEvent(st, side.e_diag, 0));
    g_prof.begin(3, st);
    // L_pp^-T of every panel at once (the panel solve applied to the rows of the identity), then the back substitution
    const int npan = (n + LNBO - 1) / LNBO;
    double *xinv = xinv_ws;
    {
        TrsmArgs T = {};
        T.A = nullptr; T.dst = xinv; T.Lfac = Lmat; T.linv_all = linv_ws; T.ld = ld; T.dst_ld = LNBO; T.nrows = 0; T.ident = 1; T.k0 = 0; T.nb = 0; T.n = n;
        SliceOut none = {};
        chol_trsm_kernel<<<dim3(LNBO / TR_ROWS, npan), 256, TR_SMEM_DOUBLES * sizeof(double), st>>>(T, none);
        BSFM_KERNEL_CHECK();
    }
    for (int k0 = (npan - 1) * LNBO; k0 >= 0; k0 -= LNBO) {
        const int nb = std::min(LNBO, n - k0);
        chol_backsolve_big_kernel<<<std::max(1, (k0 + 63) / 64), 1024, 0, st>>>(Lmat, ld, n, k0, nb, xinv, x);
        BSFM_KERNEL_CHECK();
    }
    g_prof.end(st);
    if (g_prof.on) g_prof.launches[3] += npan;     // end() counted one
    return BSFM_OK;
}

}  // namespace ba
}  // namespace bsfm

extern "C" int bsfm_ba_chol_profile(int enable)
{
    using namespace bsfm::ba;
    g_prof.collect();
    for (int q = 0; q < 4; q++) { g_prof.ms[q] = 0; g_prof.launches[q] = 0; }
    g_prof.int8_ops = g_prof.fp64_flops = 0.0;
    g_prof.on = enable != 0;
    return BSFM_OK;
}
extern "C" int bsfm_ba_chol_profile_read(float ms[4], int launches[4], double *tensor_int8_ops, double *fp64_equiv_flops)
{
    using namespace bsfm::ba;
    g_prof.collect();
    for (int q = 0; q < 4; q++) { if (ms) ms[q] = g_prof.ms[q]; if (launches) launches[q] = g_prof.launches[q]; }
    if (tensor_int8_ops) *tensor_int8_ops = g_prof.int8_ops;
    if (fp64_equiv_flops) *fp64_equiv_flops = g_prof.fp64_flops;
    return BSFM_OK;
}

extern "C" int bsfm_debug_diag_prof(long long *out8)
{
    if (!bsfm::ba::g_diag_dbg) return -1;
    if (cudaMemcpy(out8, bsfm::ba::g_diag_dbg, 8 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    cudaMemset(bsfm::ba::g_diag_dbg, 0, 8 * sizeof(long long));
    return 0;
}
