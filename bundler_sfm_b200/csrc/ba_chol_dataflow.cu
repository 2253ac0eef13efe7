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
constexpr int DF_LD = LNB + 1;
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

// x <- x L^-T for one 32-wide row per thread (row in registers), L lower triangular in shared memory, dinv = 1 / diag
__device__ __forceinline__ void row_solve_32(double *xrow /* DF_LD pitch row in smem */, const double (*Ls)[DF_LD], const double *dinv)
{
    double v[LNB];
#pragma unroll
    for (int c = 0; c < LNB; c++) v[c] = xrow[c];
#pragma unroll
    for (int t = 0; t < LNB; t++) {
        const double x = v[t] * dinv[t];
        v[t] = x;
#pragma unroll
        for (int c = 1; c < LNB; c++)
            if (c > t) v[c] = fma(-x, Ls[c][t], v[c]);
    }
#pragma unroll
    for (int c = 0; c < LNB; c++) xrow[c] = v[c];
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
                                                                     int *ver /* (nbk+1) x (nbk+1) tile versions, init -1 */, int *diag_ready /* nbk, init -1 */)
{
    __shared__ double Ls[LNB][DF_LD];        // L_kk
    __shared__ double Xr[LNB][DF_LD];
    __shared__ double Xc[LNB][DF_LD];
    __shared__ double Zs[LNB][DF_LD];
    __shared__ double dinv[LNB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nbk = (n + LNB - 1) / LNB;
    const int VW = nbk + 1;                  // row pitch of `ver`
    const int workers = (int) gridDim.x - 1;

    auto rows_of = [&](int rb) { return (rb == nbk) ? 1 : min(LNB, n - rb * LNB); };
    auto rbase_of = [&](int rb) { return (rb == nbk) ? n : rb * LNB; };
    // load block (rb, cb) of A (rows x width, zero padded) into smem
    auto load_block = [&](double (*dst)[DF_LD], int rb, int cb, int width) {
        const int rows = rows_of(rb), rbase = rbase_of(rb), c0 = cb * LNB;
        for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
            const int r = e >> 5, c = e & 31;
            dst[r][c] = (r < rows && c < width) ? __ldcg(&A[(size_t) (rbase + r) * ld + (c0 + c)]) : 0.0;     // L2: written by another SM
        }
    };

    if (blockIdx.x == 0) {
        // ================= the diagonal chain =================
        // warp 0 runs the chain (potf2 -> solve of the sub-diagonal block -> rank-32 update of the next diagonal tile, which never
        // leaves shared memory); warps 1..7 hide everything else under it: they poll the flags and fetch the blocks of the NEXT
        // look-ahead while warp 0 factors, and publish L_kk / form L_kk^-1 while warp 0 solves.
        __shared__ double Dn[LNB][DF_LD];        // the next diagonal tile
        double (*Lc)[DF_LD] = Ls, (*Ln)[DF_LD] = Dn;
        const unsigned long long ns0 = df_ns();
        unsigned long long pc[5] = {0, 0, 0, 0, 0}, tc = 0;
        {
            const int nb0 = min(LNB, n);
            for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                const int r = e >> 5, c = e & 31;
                Lc[r][c] = (r < nb0 && c <= r) ? A[(size_t) r * ld + c] : ((r == c) ? 1.0 : 0.0);
            }
        }
        __syncthreads();
        for (int k = 0; k < nbk; k++) {
            const int k0 = k * LNB, nb = min(LNB, n - k0);
            const bool last = (k + 1 == nbk);
            const int rb = last ? nbk : k + 1;                 // last step: the right-hand side segment instead of a diagonal tile
            const int rrows = rows_of(rb), rbase = rbase_of(rb);
            // ---- phase 1: potf2(k)  ||  inputs of the look-ahead ----
            tc = df_clock();
            if (warp == 0) {
                const bool bad = warp_potf2_32<DF_LD>(Lc, dinv, lane);
                if (bad && lane == 0) sc->chol_fail = 1;
                __syncwarp();
                dinv[lane] = 1.0 / Lc[lane][lane];               // the divisors every CTA uses (the workers form them the same way)
                { const unsigned long long t = df_clock(); pc[0] += t - tc; tc = t; }
            } else {
                if (tid == 32) {
                    long long spins = 0;
                    while (ld_acquire(&ver[rb * VW + k]) < k - 1 || (!last && ld_acquire(&ver[rb * VW + rb]) < k - 1)) {
                        if (++spins > DF_SPIN_LIMIT) { sc->chol_fail = 2; break; }
                        __nanosleep(20);
                    }
                }
                asm volatile("bar.sync 1, 224;" ::: "memory");
                for (int e = tid - 32; e < LNB * LNB; e += DF_THREADS - 32) {
                    const int r = e >> 5, c = e & 31;
                    Xr[r][c] = (r < rrows && c < nb) ? __ldcg(&A[(size_t) (rbase + r) * ld + (k0 + c)]) : 0.0;
                    if (!last) Ln[r][c] = (r < rrows && c <= r) ? __ldcg(&A[(size_t) (rbase + r) * ld + (rbase + c)]) : ((r == c) ? 1.0 : 0.0);
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[1] += t - tc; tc = t; }
            // ---- phase 2: solve of the sub-diagonal block  ||  publish L_kk, form L_kk^-1 ----
            if (warp == 0) {
                if (lane < rrows) row_solve_32(Xr[lane], Lc, dinv);
                { const unsigned long long t = df_clock(); pc[2] += t - tc; tc = t; }
            } else {
                for (int e = tid - 32; e < LNB * LNB; e += DF_THREADS - 32) {
                    const int r = e >> 5, c = e & 31;
                    if (r < nb && c <= r) Lout[(size_t) (k0 + r) * ld + (k0 + c)] = Lc[r][c];
                }
                __threadfence();
                asm volatile("bar.sync 1, 224;" ::: "memory");
                if (tid == 32) st_release(&diag_ready[k], 1);
                if (warp == DF_THREADS / 32 - 1) {
                    double v[LNB];
#pragma unroll
                    for (int r = 0; r < LNB; r++) v[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int t = 0; t < LNB; t++) {
                        const double z = v[t] * dinv[t];
                        v[t] = z;
#pragma unroll
                        for (int r = 1; r < LNB; r++)
                            if (r > t) v[r] = fma(-Lc[r][t], z, v[r]);
                    }
#pragma unroll
                    for (int r = 0; r < LNB; r++) Zs[r][lane] = v[r];
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[3] += t - tc; tc = t; }
            // ---- phase 3: factor block, inverse, rank-32 update of the next diagonal tile (stays in shared memory) ----
            for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                const int r = e >> 5, c = e & 31;
                if (r < rrows && c < nb) Lout[(size_t) (rbase + r) * ld + (k0 + c)] = Xr[r][c];
                linv_all[(size_t) k * LNB * LNB + e] = (c <= r) ? Zs[r][c] : 0.0;
                if (!last && r < rrows && c <= r) {
                    double acc = 0.0;
#pragma unroll 8
                    for (int t = 0; t < LNB; t++) acc = fma(Xr[r][t], Xr[c][t], acc);
                    Ln[r][c] -= acc;
                }
            }
            __syncthreads();
            { const unsigned long long t = df_clock(); pc[4] += t - tc; tc = t; }
            double (*tmp)[DF_LD] = Lc; Lc = Ln; Ln = tmp;
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
    const int w = (int) blockIdx.x - 1;
    for (int k = 0; k + 1 <= nbk; k++) {
        const int k0 = k * LNB, nb = min(LNB, n - k0);
        bool have_l = false;
        // tiles (rb, cb), k < cb <= rb <= nbk, cb < nbk, owned by this CTA; (k+1, k+1) belongs to CTA 0 at this step
        for (int cb = k + 1; cb < nbk; cb++) {
            for (int rb = cb; rb <= nbk; rb++) {
                if (df_tile_index(rb, cb, nbk) % workers != w) continue;
                if (rb == cb && cb == k + 1) continue;
                if (!have_l) {
                    df_wait(&diag_ready[k], 1, sc);
                    for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                        const int r = e >> 5, c = e & 31;
                        Ls[r][c] = (r < nb && c <= r) ? __ldcg(&Lout[(size_t) (k0 + r) * ld + (k0 + c)]) : ((r == c) ? 1.0 : 0.0);
                    }
                    __syncthreads();
                    if (tid < LNB) dinv[tid] = 1.0 / Ls[tid][tid];
                    have_l = true;
                }
                df_wait(&ver[rb * VW + k], k - 1, sc);
                if (cb != rb) df_wait(&ver[cb * VW + k], k - 1, sc);
                if (rb == cb && k > 0) df_wait(&ver[rb * VW + cb], k - 1, sc);   // diagonal tiles change hands (CTA 0 did step cb-1 only, but be safe)
                load_block(Xr, rb, k, nb);
                if (cb != rb) load_block(Xc, cb, k, nb);
                __syncthreads();
                const int rrows = rows_of(rb), crows = rows_of(cb);
                if (tid < rrows) row_solve_32(Xr[tid], Ls, dinv);
                else if (cb != rb && tid >= 32 && tid < 32 + crows) row_solve_32(Xc[tid - 32], Ls, dinv);
                __syncthreads();
                const double (*XC)[DF_LD] = (cb != rb) ? Xc : Xr;
                const int rbase = rbase_of(rb), c0 = cb * LNB;
                if (cb == k + 1) {          // first column: this tile's row block of the factor
                    for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                        const int r = e >> 5, c = e & 31;
                        if (r < rrows && c < nb) Lout[(size_t) (rbase + r) * ld + (k0 + c)] = Xr[r][c];
                    }
                }
                for (int e = tid; e < LNB * LNB; e += DF_THREADS) {
                    const int r = e >> 5, c = e & 31;
                    if (r < rrows && c < crows && (rb != cb || c <= r)) {
                        double acc = 0.0;
#pragma unroll 8
                        for (int t = 0; t < LNB; t++) acc = fma(Xr[r][t], XC[c][t], acc);
                        double *dst = &A[(size_t) (rbase + r) * ld + (c0 + c)];
                        *dst = __ldcg(dst) - acc;
                    }
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) st_release(&ver[rb * VW + cb], k);
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
    const int grid = std::min(sm_count[dev].load(), 1 + std::max(1, ntile));
    if (grid < 2) return BSFM_OK;
    // flags live behind the back-substitution scratch of the workspace (chol_extra_ws_doubles leaves > 64k doubles there)
    double *ywork = linv_ws + (size_t) nbk * LNB * LNB;
    int *flags = reinterpret_cast<int *>(ywork + n + 64);
    int *ver = flags, *diag_ready = flags + (nbk + 1) * (nbk + 1);
    BSFM_CUDA_TRY(cudaMemsetAsync(flags, 0xFF, (size_t) ((nbk + 1) * (nbk + 1) + nbk + 8) * sizeof(int), st));
    int ld = n;
    void *args[] = {&A, &Lmat, &ld, &n, &linv_ws, &sc, &ver, &diag_ready};
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
