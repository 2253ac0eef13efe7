// ba_chol_dataflow.cu -- dense SPD factorisation of SMALL reduced camera systems (n <= 1536; config 2: 50 cameras -> 450)
// as ONE co-resident launch per factorisation.  Reference: dpotrf inside sba_Axb_Chol, lib/sba-1.5/sba_lapack.c:429.
//
// Small systems are bound by the chain of n sequential pivots, not by FLOPs.  The fused-step path (ba_chol.cu) pays one
// launch per 32 columns (15 for config 2) and a 16-warp, barrier-per-pivot factorisation of the diagonal tile in every CTA
// (~600 cycles per pivot).  Here the whole factorisation is a dataflow over 32 x 32 tiles inside one cooperative launch:
//   CTA 0        owns the critical path: for every 32-column step k it factors the diagonal tile (k,k) with ONE warp
//                (8-column register panels + shuffles, ~230 cycles per pivot: warp_potf2_32), publishes L_kk, then itself
//                applies step k to the NEXT diagonal tile (k+1,k+1) (solve of the sub-diagonal block + rank-32 update) and
//                goes on to step k+1 -- look-ahead by construction.  A spare warp forms L_kk^-1 for the back substitution.
//   CTAs 1..G-1  own the other tiles statically; for every step and owned tile they wait for L_kk and for the two panel
//                blocks of their rows (per-tile version flags in global memory, acquire / release), solve them
//                (thread per row), apply the rank-32 update, write the factor block if theirs is the first column, and
//                publish the tile's version.  No grid-wide barrier anywhere.
// Same data contract as chol_solve (factor OUT OF PLACE in Lmat, right-hand side as matrix row n).  Waits are bounded
// (a lost flag sets sc->chol_fail = 2 instead of hanging the device).
#include "ba_kernels.cuh"
#include "ba_chol_large.cuh"
#include "ba_chol_potf2.cuh"
#include "common.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>

namespace bsfm {
namespace ba {

constexpr int DF_THREADS = 256;
constexpr long long DF_SPIN_LIMIT = 400000;      // polls of ~0.5-1 us: a lost flag costs well under a second, never a hang

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 polls until *flag >= need (bounded); the CTA continues after the barrier
__device__ __forceinline__ void df_wait(const int *flag, int need, Scalars *sc)
{
    if (threadIdx.x == 0) {
        long long spins = 0;
        while (ld_acquire(flag) < need) {
            if (++spins > DF_SPIN_LIMIT) { sc->chol_fail = 2; break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
}

// tile index of (rb, cb), 1 <= cb <= rb <= nbk (block row nbk = right-hand side), column-major over cb
__device__ __forceinline__ int df_tile_index(int rb, int cb, int nbk)
{
    // tiles of columns 1 .. cb-1: sum_{c=1}^{cb-1} (nbk - c + 1)
    const int before = (cb - 1) * (nbk + 1) - (cb - 1) * cb / 2;
    return before + (rb - cb);
}

// dev-only accounting of CTA 0 (read and reset by bsfm_debug_df_prof): [0] kernel ns, [1] potf2, [2] wait for the helpers' loads,
// [3] sub-diagonal solve, [4] wait for the publish, [5] phase 3, [6] steps, [7] launches   ([1..5] in SM cycles)
__device__ unsigned long long g_df_prof[8];
__device__ __forceinline__ unsigned long long df_clock() { unsigned long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__device__ __forceinline__ unsigned long long df_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }

__global__ void __launch_bounds__(DF_THREADS, 1) chol_dataflow_kernel(double *A, double *Lout, int ld, int n, double *linv_all, Scalars *sc,
                                                                     int *ver /* (nbk+1) x (nbk+1) tile versions, init -1 */, int *diag_ready /* nbk, init -1 */, int variant /* dev experiments */)
{
    __shared__ __align__(16) double tiles[5][LNB][TP];      // CTA 0: L_kk, next diagonal tile, X, Z, W;  workers: Z, X_r, X_c
    __shared__ double dinv[LNB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tg = lane & 3;
    const int nbk = (n + LNB - 1) / LNB;
    const int VW = nbk + 1;                  // row pitch of `ver`
    const int workers = (int) gridDim.x - 1;

    auto rows_of = [&](int rb) { return (rb == nbk) ? 1 : min(LNB, n - rb * LNB); };
    auto rbase_of = [&](int rb) { return (rb == nbk) ? n : rb * LNB; };

    if (blockIdx.x == 0) {
        // ================= the diagonal chain =================
        // warp 0 factors the diagonal tile (warp_potf2_32_tc), warp 1 assembles its inverse one panel behind, warps 2..7 poll the
        // flags and fetch the inputs of the look-ahead meanwhile.  Then: X <- X Z^T for the sub-diagonal block (warps 0..3, DMMA)
        // while warps 4..7 publish L_kk and Z_kk; then the rank-32 update of the next diagonal tile, which never leaves shared memory.
        double (*Lc)[TP] = tiles[0], (*Ln)[TP] = tiles[1];
        double (*Xr)[TP] = tiles[2], (*Zs)[TP] = tiles[3], (*Ws)[TP] = tiles[4];
        const unsigned long long ns0 = df_ns();
        unsigned long long pc[5] = {0, 0, 0, 0, 0}, tc = 0;
        {
            const int nb0 = min(LNB, n);
            for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                const int r = e >> 5, c = e & 31;
                Lc[r][c] = (r < nb0 && c <= r) ? A[(size_t) r * ld + c] : ((r == c) ? 1.0 : 0.0);
                Zs[r][c] = 0.0;                              // the strictly upper blocks of Z stay zero
            }
        }
        __syncthreads();
        for (int k = 0; k < nbk; k++) {
            const int k0 = k * LNB, nb = min(LNB, n - k0);
            const bool last = (k + 1 == nbk);
            const int rb = last ? nbk : k + 1;                 // last step: the right-hand side segment instead of a diagonal tile
            const int rrows = rows_of(rb), rbase = rbase_of(rb);
            // ---- phase 1: potf2(k) and its inverse  ||  inputs of the look-ahead ----
            tc = df_clock();
            if (warp == 0) {
                const bool bad = warp_potf2_32_tc(Lc, dinv, lane, true);
                if (bad && lane == 0) sc->chol_fail = 1;
                { const unsigned long long t = df_clock(); pc[0] += t - tc; tc = t; }
            } else if (warp == 1) {
                warp_tile_inverse(Lc, dinv, Zs, Ws, lane);
            } else {
                if (tid == 64) {
                    long long spins = 0;
                    while (ld_acquire(&ver[rb * VW + k]) < k - 1 || (!last && ld_acquire(&ver[rb * VW + rb]) < k - 1)) {
                        if (++spins > DF_SPIN_LIMIT) { sc->chol_fail = 2; break; }
                        __nanosleep(20);
                    }
                }
                asm volatile("bar.sync 1, 192;" ::: "memory");
                for (int e = tid - 64; e < LNB * LNB; e += DF_THREADS - 64) {
                    const int r = e >> 5, c = e & 31;
                    Xr[r][c] = (r < rrows && c < nb) ? __ldcg(&A[(size_t) (rbase + r) * ld + (k0 + c)]) : 0.0;
                    if (!last) Ln[r][c] = (r < rrows && c <= r) ? __ldcg(&A[(size_t) (rbase + r) * ld + (rbase + c)]) : ((r == c) ? 1.0 : 0.0);
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[1] += t - tc; tc = t; }
            // ---- phase 2: sub-diagonal block X <- X Z^T  ||  publish L_kk and Z_kk ----
            const int role = (variant & 4) ? (warp ^ 4) : warp;          // dev: swap which half solves and which half publishes
            if (role < 4) {
                warp_rows_times_ZT(Xr, 8 * role, Zs, lane);
                { const unsigned long long t = df_clock(); pc[2] += t - tc; tc = t; }
            } else {
                for (int e = (tid ^ ((variant & 4) ? 128 : 0)) - 128; e < LNB * LNB; e += 128) {
                    const int r = e >> 5, c = e & 31;
                    if (r < nb && c <= r) Lout[(size_t) (k0 + r) * ld + (k0 + c)] = Lc[r][c];
                    linv_all[(size_t) k * LNB * LNB + e] = (c <= r) ? Zs[r][c] : 0.0;
                }
                asm volatile("bar.sync 6, 128;" ::: "memory");
                if (tid == ((variant & 4) ? 0 : ((variant & 1) ? 224 : 128))) {      // one cumulative fence after the barrier, not 128
                    if (variant & 2) *reinterpret_cast<volatile int *>(&diag_ready[k]) = 1;        // dev: timing without the fence (NOT correct)
                    else { __threadfence(); st_release(&diag_ready[k], 1); }
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[3] += t - tc; tc = t; }
            // ---- phase 3: factor block out, rank-32 update of the next diagonal tile (stays in shared memory) ----
            for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                const int r = e >> 5, c = e & 31;
                if (r < rrows && c < nb) Lout[(size_t) (rbase + r) * ld + (k0 + c)] = Xr[r][c];
            }
            if (!last) {
                for (int t = warp; t < 10; t += DF_THREADS / 32) {       // lower 8 x 8 tiles (ri, cj)
                    int ri = 0, rem = t;
                    while (rem > ri) { rem -= ri + 1; ri++; }
                    const int cj = rem;
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < LNB; ks += 4) tile_dmma(c0, c1, Xr[8 * ri + g][ks + tg], Xr[8 * cj + g][ks + tg]);
                    double2 cur = *reinterpret_cast<const double2 *>(&Ln[8 * ri + g][8 * cj + 2 * tg]);
                    cur.x -= c0; cur.y -= c1;
                    *reinterpret_cast<double2 *>(&Ln[8 * ri + g][8 * cj + 2 * tg]) = cur;
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[4] += t - tc; tc = t; }
            double (*tmp)[TP] = Lc; Lc = Ln; Ln = tmp;
        }
        if (tid == 0) {
            g_df_prof[0] += df_ns() - ns0;
            for (int i = 0; i < 5; i++) g_df_prof[1 + i] += pc[i];
            g_df_prof[6] += nbk;
            g_df_prof[7] += 1;
        }
        return;
    }

    // ================= workers: static tile ownership =================
    double (*Zs)[TP] = tiles[0], (*Xr)[TP] = tiles[1], (*Xc)[TP] = tiles[2];
    // load block (rb, cb) of A (rows x width, zero padded) into smem
    auto load_block = [&](double (*dst)[TP], int rb, int cb, int width) {
        const int rows = rows_of(rb), rbase = rbase_of(rb), c0 = cb * LNB;
        for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
            const int r = e >> 5, c = e & 31;
            dst[r][c] = (r < rows && c < width) ? __ldcg(&A[(size_t) (rbase + r) * ld + (c0 + c)]) : 0.0;     // L2: written by another SM
        }
    };
    const int w = (int) blockIdx.x - 1;
    for (int k = 0; k + 1 <= nbk; k++) {
        const int k0 = k * LNB, nb = min(LNB, n - k0);
        bool have_z = false;
        // tiles (rb, cb), k < cb <= rb <= nbk, cb < nbk, owned by this CTA; (k+1, k+1) belongs to CTA 0 at this step
        for (int cb = k + 1; cb < nbk; cb++) {
            for (int rb = cb; rb <= nbk; rb++) {
                if (df_tile_index(rb, cb, nbk) % workers != w) continue;
                if (rb == cb && cb == k + 1) continue;
                if (!have_z) {
                    df_wait(&diag_ready[k], 1, sc);
                    for (int e = tid; e < LNB * LNB; e += DF_THREADS) Zs[e >> 5][e & 31] = __ldcg(&linv_all[(size_t) k * LNB * LNB + e]);
                    have_z = true;
                }
                df_wait(&ver[rb * VW + k], k - 1, sc);
                if (cb != rb) df_wait(&ver[cb * VW + k], k - 1, sc);
                if (rb == cb && k > 0) df_wait(&ver[rb * VW + cb], k - 1, sc);   // diagonal tiles change hands (CTA 0 did step cb-1 only, but be safe)
                load_block(Xr, rb, k, nb);
                if (cb != rb) load_block(Xc, cb, k, nb);
                __syncthreads();
                const int rrows = rows_of(rb), crows = rows_of(cb);
                // both panel blocks: X <- X Z^T, one 8-row strip per warp
                if (warp < 4) warp_rows_times_ZT(Xr, 8 * warp, Zs, lane);
                else if (cb != rb) warp_rows_times_ZT(Xc, 8 * (warp - 4), Zs, lane);
                __syncthreads();
                const double (*XC)[TP] = (cb != rb) ? Xc : Xr;
                const int rbase = rbase_of(rb), c0 = cb * LNB;
                if (cb == k + 1) {          // first column: this tile's row block of the factor
                    for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                        const int r = e >> 5, c = e & 31;
                        if (r < rrows && c < nb) Lout[(size_t) (rbase + r) * ld + (k0 + c)] = Xr[r][c];
                    }
                }
                // tile -= X_r X_c^T : sixteen 8 x 8 DMMA tiles, two per warp
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int t = warp * 2 + i, ri = t >> 2, cj = t & 3;
                    if (rb == cb && cj > ri) continue;
                    if (8 * ri >= rrows || 8 * cj >= crows) continue;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < LNB; ks += 4) tile_dmma(a0, a1, Xr[8 * ri + g][ks + tg], XC[8 * cj + g][ks + tg]);
                    const int r = 8 * ri + g, c = 8 * cj + 2 * tg;
                    double *dst = &A[(size_t) (rbase + r) * ld + (c0 + c)];
                    if (r < rrows && c < crows && (rb != cb || c <= r)) dst[0] = __ldcg(dst) - a0;
                    if (r < rrows && c + 1 < crows && (rb != cb || c + 1 <= r)) dst[1] = __ldcg(dst + 1) - a1;
                }
                __syncthreads();
                if (tid == 0) { __threadfence(); st_release(&ver[rb * VW + cb], k); }
            }
        }
    }
}

__global__ void __launch_bounds__(512) chol_backsolve_blocked_kernel(const double *A, int ld, int n, const double *Linv_all, double *x, double *ywork);

static int df_max_blocks()
{
    static const int v = []() { const char *e = getenv("BSFM_BA_CHOL_DATAFLOW_MAXBLK"); return e ? atoi(e) : 20; }();
    return v;
}

// *used = false: not usable here (caller falls back to the fused-step path); else launched
int chol_solve_dataflow(cudaStream_t st, double *A, double *Lmat, int n, double *linv_ws, double *x, Scalars *sc, bool *used)
{
    *used = false;
    static const bool enabled = []() { const char *e = getenv("BSFM_BA_CHOL_DATAFLOW"); return !(e && e[0] == '0'); }();
    if (!enabled) return BSFM_OK;
    int dev = 0;
    BSFM_CUDA_TRY(cudaGetDevice(&dev));
    static std::atomic<int> coop_ok[64];      // 0 unknown, 1 yes, -1 no
    static std::atomic<int> sm_count[64];
    if (dev < 0 || dev >= 64) return BSFM_OK;
    if (coop_ok[dev].load() == 0) {
        int coop = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev].store(sms);
        coop_ok[dev].store(coop ? 1 : -1);
    }
    if (coop_ok[dev].load() != 1) return BSFM_OK;
    const int nbk = (n + LNB - 1) / LNB;
    if (nbk < 2 || nbk > df_max_blocks()) return BSFM_OK;      // larger systems: the workers' serial tile loop loses to the fused-step path
    int ntile = 0;
    for (int c = 1; c < nbk; c++) ntile += nbk - c + 1;
    static const int grid_cap = []() { const char *e = getenv("BSFM_DF_GRID"); return e ? atoi(e) : 1 << 20; }();      // dev experiments
    const int grid = std::min(grid_cap, std::min(sm_count[dev].load(), 1 + std::max(1, ntile)));
    if (grid < 2) return BSFM_OK;
    // flags live behind the back-substitution scratch of the workspace (chol_extra_ws_doubles leaves > 64k doubles there)
    double *ywork = linv_ws + (size_t) nbk * LNB * LNB;
    int *flags = reinterpret_cast<int *>(ywork + n + 64);
    int *ver = flags, *diag_ready = flags + (nbk + 1) * (nbk + 1);
    BSFM_CUDA_TRY(cudaMemsetAsync(flags, 0xFF, (size_t) ((nbk + 1) * (nbk + 1) + nbk + 8) * sizeof(int), st));
    int ld = n;
    static const int variant = []() { const char *e = getenv("BSFM_DF_VARIANT"); return e ? atoi(e) : 0; }();
    int var = variant;
    void *args[] = {&A, &Lmat, &ld, &n, &linv_ws, &sc, &ver, &diag_ready, &var};
    BSFM_CUDA_TRY(cudaLaunchCooperativeKernel((const void *) chol_dataflow_kernel, dim3(grid), dim3(DF_THREADS), args, 0, st));
    count_launch();
    chol_backsolve_blocked_kernel<<<1, 512, 0, st>>>(Lmat, ld, n, linv_ws, x, ywork);
    BSFM_KERNEL_CHECK();
    *used = true;
    return BSFM_OK;
}

}  // namespace ba
}  // namespace bsfm

// dev-only: read and reset chol_dataflow_kernel's CTA-0 counters (see g_df_prof)
extern "C" int bsfm_debug_df_prof(unsigned long long *out8)
{
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out8, bsfm::ba::g_df_prof, 8 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return cudaMemcpyToSymbol(bsfm::ba::g_df_prof, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
