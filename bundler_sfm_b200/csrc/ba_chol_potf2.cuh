// ba_chol_potf2.cuh -- the 32 x 32 tile primitives of every dense solve of the BA path (reference: dpotf2 / dtrsm inside dpotrf,
// lib/sba-1.5/sba_lapack.c:429).  Measured on B200 (profiles/r2_fp64_issue_rates.log): a vector DFMA issues every 2 cycles per SM
// sub-partition, a DMMA (mma.sync.m8n8k4.f64, 256 FMAs) every 16 -- the same peak FLOP rate, but one instruction instead of eight
// plus the shuffles / shared-memory loads that feed them.  The 32-pivot chain is LATENCY bound (shuffle -> rsqrt -> multiply ->
// update, ~240 cycles per pivot), so everything that is not the chain itself is taken off the factoring warp or batched into DMMAs
// (profiles/r2_chol_dataflow_v2_cta0_cycles.log -> r2_chol_dataflow_v3_regtiles.log: 13.2k -> 7.7k cycles per tile):
//   warp_potf2_32_tc   one warp: four 8-column register panels (pivots / multipliers by shuffle, reciprocal square root as a
//                      MUFU seed + one cubic correction, validity test on the integer pipe), the rank-8 update of the columns to
//                      the right as DMMA tiles; signals a named barrier after every panel
//   warp_tile_inverse  a second warp (another sub-partition) follows panel by panel and assembles Z = L^-1: 8 x 8 diagonal
//                      inverses by substitution, the off-diagonal blocks as DMMA products (Z_10 = -Z_11 L_10 Z_00, ...)
//   warp_rows_times_ZT X <- X Z^T for an 8-row strip: the triangular solve of panel rows as 20 DMMAs
// All tiles live in shared memory with a row pitch of 36 doubles (conflict-free for the DMMA fragment pattern [row g][column tg]).
#pragma once
#include <cuda_runtime.h>

namespace bsfm {
namespace ba {

constexpr int TP = 36;        // shared-memory row pitch (doubles) of every 32-wide tile handled here

__device__ __forceinline__ double potf2_shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// D(8x8) += A(8x4) B(4x8); lane = 4 g + tg holds A[g][tg], B[tg][g], D[g][2 tg], D[g][2 tg + 1]
__device__ __forceinline__ void tile_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double neg_d(double x) { return __hiloint2double(__double2hiint(x) ^ (int) 0x80000000, __double2loint(x)); }

// 1 / sqrt(d) for a positive normal d: the hardware seed (MUFU.RSQ64H: reads the high word only, ~2^-20.5 relative) and one cubic
// (Halley) correction y (1 + e/2 + 3 e^2 / 8), e = 1 - d y^2: relative error 2.5 eps0^3 ~ 2^-60 before rounding, 5 fp64 instructions
__device__ __forceinline__ double rsqrt_pos_normal(double d)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double t = d * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double u = y0 * e;
    return fma(u, p, y0);
}
__device__ __forceinline__ bool not_pos_normal(double d)
{
    return (unsigned) (__double2hiint(d) - 0x00100000) >= 0x7FE00000u;      // zero, negative, denormal, inf or nan
}

// ---- scalar variant (kept for the fused-step path of ba_chol.cu) ----
// lane = row.  Four 8-column panels: the panel lives in registers (8 pivots unrolled, multipliers by shuffle), the columns right
// of it are updated in shared memory in a run-time loop (a fully unrolled 32-pivot register version is ~3000 straight-line
// instructions and ran at the instruction-fetch rate).  The lower triangle of Ls is overwritten by L (upper part zeroed),
// dinv[j] = 1 / L[j][j] (as rsqrt of the pivot).  Returns true if a pivot was not positive (the factor is then meaningless).
template <int LDP>
__device__ __forceinline__ bool warp_potf2_32(double (*Ls)[LDP], double *dinv, int lane)
{
    bool bad = false;
    for (int jb = 0; jb < 32; jb += 8) {
        double p[8];
#pragma unroll
        for (int q = 0; q < 8; q++) p[q] = Ls[lane][jb + q];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = jb + q;
            const double d = potf2_shfl_d(p[q], j);
            const bool isbad = !(d > 0.0) || !isfinite(d);
            bad |= isbad;
            const double rinv = isbad ? 1.0 : rsqrt(d);
            const double l = p[q] * rinv;               // lane == j: sqrt(d); lane > j: L[lane][j]; lane < j: 0
            p[q] = l;
            if (lane == j) dinv[j] = rinv;
#pragma unroll
            for (int q2 = 1; q2 < 8; q2++) {
                if (q2 > q) {
                    const double lc = potf2_shfl_d(l, jb + q2);
                    if (lane >= jb + q2) p[q2] = fma(-l, lc, p[q2]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) Ls[lane][jb + q] = (jb + q <= lane) ? p[q] : 0.0;
#pragma unroll 2
        for (int c = jb + 8; c < 32; c++) {
            double v = Ls[lane][c];
#pragma unroll
            for (int q = 0; q < 8; q++) v = fma(-p[q], potf2_shfl_d(p[q], c), v);
            if (lane >= c) Ls[lane][c] = v;
        }
    }
    return bad;
}

// named barriers 2..5: "panel b of the tile is final in shared memory" (the factoring warp arrives, the inverse warp waits)
constexpr int POTF2_BAR0 = 2;
__device__ __forceinline__ void potf2_bar_arrive(int b) { asm volatile("bar.arrive %0, 64;" ::"r"(POTF2_BAR0 + b) : "memory"); }
__device__ __forceinline__ void potf2_bar_wait(int b) { asm volatile("bar.sync %0, 64;" ::"r"(POTF2_BAR0 + b) : "memory"); }

// ---- tensor-core variant: one warp, Ls[32][TP] (16-byte aligned), dinv[32] ----
// signal: arrive on named barrier POTF2_BAR0 + b once columns [8 b, 8 b + 8) of L and dinv are final (warp_tile_inverse waits there)
__device__ __forceinline__ bool warp_potf2_32_tc(double (*Ls)[TP], double *dinv, int lane, bool signal)
{
    const int g = lane >> 2, tg = lane & 3;
    bool bad = false;
#pragma unroll 1
    for (int jb = 0; jb < 32; jb += 8) {
        double p[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double2 t = *reinterpret_cast<const double2 *>(&Ls[lane][jb + 2 * q]);
            p[2 * q] = t.x; p[2 * q + 1] = t.y;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = jb + q;
            const double d = potf2_shfl_d(p[q], j);
            const bool isbad = not_pos_normal(d);
            bad |= isbad;
            const double rinv = isbad ? 1.0 : rsqrt_pos_normal(d);
            const double l = p[q] * rinv;               // lane == j: sqrt(d); lane > j: L[lane][j]
            p[q] = l;
            if (lane == j) dinv[j] = rinv;
#pragma unroll
            for (int q2 = 1; q2 < 8; q2++) {
                if (q2 > q) {
                    const double lc = potf2_shfl_d(l, jb + q2);
                    if (lane >= jb + q2) p[q2] = fma(-l, lc, p[q2]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double2 t;
            t.x = (jb + 2 * q <= lane) ? p[2 * q] : 0.0;
            t.y = (jb + 2 * q + 1 <= lane) ? p[2 * q + 1] : 0.0;
            *reinterpret_cast<double2 *>(&Ls[lane][jb + 2 * q]) = t;
        }
        __syncwarp();
        if (signal) potf2_bar_arrive(jb >> 3);
        // rank-8 update of the 8 x 8 tiles (ri, cj), jb / 8 < cj <= ri <= 3, on the fp64 tensor cores
        const int b0 = jb >> 3;
        if (b0 < 3) {
            double af[4][2];                                 // panel rows of row tile rt: [g][4 ks + tg]
#pragma unroll
            for (int rt = 1; rt < 4; rt++)
#pragma unroll
                for (int ks = 0; ks < 2; ks++) af[rt][ks] = (rt > b0) ? Ls[8 * rt + g][jb + 4 * ks + tg] : 0.0;
#pragma unroll
            for (int cj = 1; cj < 4; cj++) {
                if (cj > b0) {
                    const double nb0 = neg_d(af[cj][0]), nb1 = neg_d(af[cj][1]);
#pragma unroll
                    for (int ri = 1; ri < 4; ri++) {
                        if (ri >= cj) {
                            double2 c = *reinterpret_cast<const double2 *>(&Ls[8 * ri + g][8 * cj + 2 * tg]);
                            tile_dmma(c.x, c.y, af[ri][0], nb0);
                            tile_dmma(c.x, c.y, af[ri][1], nb1);
                            *reinterpret_cast<double2 *>(&Ls[8 * ri + g][8 * cj + 2 * tg]) = c;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    return bad;
}

// 8 x 8 inverse of the diagonal block b of L (lanes 0..7, lane = column): Zs[8b + r][8b + c]
__device__ __forceinline__ void inv8_block(const double (*Ls)[TP], const double *dinv, double (*Zs)[TP], int b, int lane)
{
    const int o = 8 * b;
    if (lane < 8) {
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const double z = v[t] * dinv[o + t];
            v[t] = z;
#pragma unroll
            for (int r = 1; r < 8; r++)
                if (r > t) v[r] = fma(-Ls[o + r][o + t], z, v[r]);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) Zs[o + r][o + lane] = (r >= lane) ? v[r] : 0.0;
    }
    __syncwarp();
}

// C(8x8 tile at Cs[cr..][cc..]) = sign * sum over the listed 8-wide k blocks of A[ar..][ak..] * B[bk..][bc..]   (B k-major)
struct TileRef { const double (*m)[TP]; int r, c; };
__device__ __forceinline__ void tile_acc_kmajor(double &c0, double &c1, const double (*As)[TP], int ar, int ak, const double (*Bs)[TP], int bk, int bc, int g, int tg)
{
#pragma unroll
    for (int ks = 0; ks < 8; ks += 4) tile_dmma(c0, c1, As[ar + g][ak + ks + tg], Bs[bk + ks + tg][bc + g]);
}
__device__ __forceinline__ void tile_store(double (*Cs)[TP], int cr, int cc, double c0, double c1, int g, int tg)
{
    double2 t; t.x = c0; t.y = c1;
    *reinterpret_cast<double2 *>(&Cs[cr + g][cc + 2 * tg]) = t;
}

// Z = L^-1 of the 32 x 32 tile being factored by warp_potf2_32_tc(signal = true), assembled by ONE other warp behind it.
// Zs: the strictly upper 8 x 8 blocks must be zero on entry (they are never written); Ws: 32 x 32 scratch.
// Block inverse of a lower triangular [[T, 0], [B, D]]: [[T^-1, 0], [-D^-1 B T^-1, D^-1]], applied at 8 and at 16.
__device__ __forceinline__ void warp_tile_inverse(const double (*Ls)[TP], const double *dinv, double (*Zs)[TP], double (*Ws)[TP], int lane)
{
    const int g = lane >> 2, tg = lane & 3;
    double c0, c1;
    // ---- panel 0 ----
    potf2_bar_wait(0);
    inv8_block(Ls, dinv, Zs, 0, lane);
    // ---- panel 1 ----
    potf2_bar_wait(1);
    inv8_block(Ls, dinv, Zs, 1, lane);
    c0 = c1 = 0.0; tile_acc_kmajor(c0, c1, Ls, 8, 0, Zs, 0, 0, g, tg);              // W_10 = L_10 Z_00
    tile_store(Ws, 8, 0, c0, c1, g, tg);
    __syncwarp();
    c0 = c1 = 0.0; tile_acc_kmajor(c0, c1, Zs, 8, 8, Ws, 8, 0, g, tg);              // Z_10 = -Z_11 W_10
    tile_store(Zs, 8, 0, neg_d(c0), neg_d(c1), g, tg);
    __syncwarp();
    // W_B0 = L[16:32, 0:16] Z[0:16, 0:16]   (rows 16..31 of columns 0..15 are final after panel 1)
#pragma unroll
    for (int rt = 2; rt < 4; rt++) {
        c0 = c1 = 0.0;
        tile_acc_kmajor(c0, c1, Ls, 8 * rt, 0, Zs, 0, 0, g, tg);
        tile_acc_kmajor(c0, c1, Ls, 8 * rt, 8, Zs, 8, 0, g, tg);
        tile_store(Ws, 8 * rt, 0, c0, c1, g, tg);
        c0 = c1 = 0.0;
        tile_acc_kmajor(c0, c1, Ls, 8 * rt, 8, Zs, 8, 8, g, tg);
        tile_store(Ws, 8 * rt, 8, c0, c1, g, tg);
    }
    __syncwarp();
    // ---- panel 2 ----
    potf2_bar_wait(2);
    inv8_block(Ls, dinv, Zs, 2, lane);
    c0 = c1 = 0.0; tile_acc_kmajor(c0, c1, Ls, 24, 16, Zs, 16, 16, g, tg);          // W_32 = L_32 Z_22
    tile_store(Ws, 24, 16, c0, c1, g, tg);
#pragma unroll
    for (int ct = 0; ct < 2; ct++) {                                               // Z[16:24, 0:16] = -Z_22 W[16:24, 0:16]
        c0 = c1 = 0.0; tile_acc_kmajor(c0, c1, Zs, 16, 16, Ws, 16, 8 * ct, g, tg);
        tile_store(Zs, 16, 8 * ct, neg_d(c0), neg_d(c1), g, tg);
    }
    __syncwarp();
    // ---- panel 3 ----
    potf2_bar_wait(3);
    inv8_block(Ls, dinv, Zs, 3, lane);
    c0 = c1 = 0.0; tile_acc_kmajor(c0, c1, Zs, 24, 24, Ws, 24, 16, g, tg);          // Z_32 = -Z_33 W_32
    tile_store(Zs, 24, 16, neg_d(c0), neg_d(c1), g, tg);
    __syncwarp();
#pragma unroll
    for (int ct = 0; ct < 2; ct++) {                                               // Z[24:32, 0:16] = -(Z_32 W[16:24] + Z_33 W[24:32])
        c0 = c1 = 0.0;
        tile_acc_kmajor(c0, c1, Zs, 24, 16, Ws, 16, 8 * ct, g, tg);
        tile_acc_kmajor(c0, c1, Zs, 24, 24, Ws, 24, 8 * ct, g, tg);
        tile_store(Zs, 24, 8 * ct, neg_d(c0), neg_d(c1), g, tg);
    }
    __syncwarp();
}

// X[r0 .. r0+8)[0..32) <- X Z^T (in place; Z lower triangular: block b of the result = sum_{a <= b} X_a Z_ba^T), one warp per strip
__device__ __forceinline__ void warp_rows_times_ZT(double (*Xs)[TP], int r0, const double (*Zs)[TP], int lane)
{
    const int g = lane >> 2, tg = lane & 3;
    double a[8];
#pragma unroll
    for (int ks = 0; ks < 8; ks++) a[ks] = Xs[r0 + g][4 * ks + tg];
    double c[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
        c[nb][0] = 0.0; c[nb][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < 4; kb++) {
            if (kb <= nb) {
                tile_dmma(c[nb][0], c[nb][1], a[2 * kb], Zs[8 * nb + g][8 * kb + tg]);
                tile_dmma(c[nb][0], c[nb][1], a[2 * kb + 1], Zs[8 * nb + g][8 * kb + 4 + tg]);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int nb = 0; nb < 4; nb++) tile_store(Xs, r0, 8 * nb, c[nb][0], c[nb][1], g, tg);
}

}  // namespace ba
}  // namespace bsfm
