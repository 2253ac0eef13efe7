// ba_chol_potf2.cuh -- Cholesky factorisation of ONE 32 x 32 tile in shared memory by ONE warp (the pivot chain of every dense
// solve of the BA path; reference: dpotf2 inside dpotrf, lib/sba-1.5/sba_lapack.c:429).
#pragma once
#include <cuda_runtime.h>

namespace bsfm {
namespace ba {

__device__ __forceinline__ double potf2_shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

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

}  // namespace ba
}  // namespace bsfm
