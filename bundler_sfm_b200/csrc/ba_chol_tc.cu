// ba_chol_tc.cu -- trailing update of the large-system Cholesky on the 5th-generation tensor cores.
// Reference step: the rank-k update inside LAPACK dpotrf called by sba_Axb_Chol, lib/sba-1.5/sba_lapack.c:429.
//
//   A[r][c] -= sum_{t in panel} L[r][t] L[c][t]        (K = 256 per outer panel, fp64 in, fp64 out)
//
// tcgen05 has no fp64 kind, so the contraction runs on `tcgen05.mma kind::i8` over the int8 slices of the panel
// (ba_chol_large.cuh): NS (NS+1)/2 exact int8 x int8 -> int32 products per K-half, grouped by level d = k + l, each
// level folded into fp64 registers by the epilogue with weight 2^(-8 (d+2)) and finally scaled by 2^(e_r + e_c).
//
// Persistent, one CTA per SM, warp-specialised (the structure of the MATCH kernel, match_kernels.cu):
//   warp 0     : TMA producer -- cp.async.bulk of 16 KB slice tiles (pre-swizzled in HBM) into 2 NS dedicated slots
//                (A_k = slice k of the row tile, B_l = slice l of the column tile); a slot is re-filled for the next
//                (tile, K-half) as soon as the LAST level has used it: the last level runs outside-in
//                ((0,NS-1), (NS-1,0), (1,NS-2), ...) so slices 0 -- needed first -- are released first
//   warp 1     : TMEM allocator + single-thread MMA issuer: M = N = 128, K = 4 x 32, a/b signedness per slice
//                (slice 0 signed, the others unsigned), one TMEM accumulator stage (128 columns) per level, 4 stages
//   warps 2..17: epilogue: thread = accumulator lane = matrix COLUMN (the column tile is the MMA's M side), 32 matrix
//                rows each; per level  acc += (double) C_d * 2^(-8 (d+2));  after the 2 NS levels of a tile the fp64
//                read-modify-write of A, coalesced along matrix rows
// Algorithmic work per 128 x 128 x 256 tile update: 8.4 MFLOP (fp64-equivalent) = NS (NS+1) x 4.2 M int8 MACs.
#include "ba_chol_large.cuh"
#include "ba_kernels.cuh"
#include "tc_ptx.cuh"
#ifndef BSFM_TCS_DBG
#define BSFM_TCS_DBG 0
#endif
#include "common.h"
#include <atomic>
#include <cstdlib>

namespace bsfm {
namespace ba {
using namespace bsfm::ptx;

constexpr int TCS_NS_MAX = 7;
constexpr int TCS_EPI_WARPS = 16;            // 4 per TMEM lane quadrant: 32 accumulator columns each
constexpr int TCS_THREADS = 64 + TCS_EPI_WARPS * 32;
constexpr int TCS_STAGES = 4;               // TMEM accumulator stages of 128 columns

struct TcSyrkParams {
    const uint8_t *slices;
    const double *rscale;
    double *A;
    int ns, ld, nrows, cb, ce;
    int tile0, nct, ntiles;
    int nrt, tc0, tc1;          // row tiles; tile columns [tc0, tc1) of this launch (column-major enumeration inside the range)
    unsigned long long *prof;   // dev-only cycle accounting of CTA 0 (BSFM_TCS_PROF=1), else null
};
// prof[0..3] MMA thread: wait slices, wait accumulator stage, issue, levels;  [4..5] producer: wait slot, loads;
// [8..11] epilogue warp 2 lane 0: wait accumulator, fold, write-back, tiles
#define TCS_T(var) const long long var = P.prof ? clock64() : 0
#define TCS_ADD(i, v) do { if (P.prof && blockIdx.x == 0) atomicAdd(&P.prof[i], (unsigned long long) (v)); } while (0)

// tile idx of a launch -> (row tile ti, column tile tj): columns tc0 .. tc1-1 one after the other, rows tj .. nrt-1 inside
// a column (lower triangle).  Column-major order keeps the column tile's slices hot in L2 for consecutive CTAs.
__device__ __forceinline__ void tcs_decode_tile(int idx, const TcSyrkParams &P, int &ti, int &tj)
{
    int c = P.tc0;
    while (idx >= P.nrt - c) { idx -= P.nrt - c; c++; }
    tj = c; ti = c + idx;
}

// UMMA instruction descriptor, kind::i8: c S32 (2) [4,6), a/b format [7,10)/[10,13) 0 = u8, 1 = s8, K-major,
// N >> 3 [17,23), M >> 4 [24,29)
__device__ __forceinline__ uint32_t tcs_idesc(bool a_signed, bool b_signed)
{
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t) (TC_TILE >> 3) << 17) | ((uint32_t) (TC_TILE >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_regs8(uint32_t (&v)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait_regs16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}

// exact int32 -> fp64 without the conversion unit (I2F.F64 runs on the quarter-rate XU pipe and bounded the epilogue):
// the bits 0x43300000:(x ^ 0x80000000) are the double 2^52 + 2^31 + x
__device__ __forceinline__ double i2d(uint32_t x)
{
    return __hiloint2double(0x43300000, (int) (x ^ 0x80000000u)) - 4503601774854144.0;
}

// exact int64 -> fp64 for |t| < 2^51: the bits 0x4330000000000000 + (t + 2^51) are the double 2^52 + 2^51 + t
__device__ __forceinline__ double ll2d(long long t)
{
    return __longlong_as_double(0x4330000000000000ll + (t + (1ll << 51))) - 6755399441055744.0;
}

__global__ void __launch_bounds__(TCS_THREADS, 1) tc_syrk_kernel(const TcSyrkParams P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw_addr);
    const int ns = P.ns;
    const uint32_t sA = base;                                   // A_k at sA + k * 16 KB
    const uint32_t sB = base + (uint32_t) ns * TC_SLICE_BYTES;   // B_l
    const uint32_t bar0 = base + 2u * (uint32_t) ns * TC_SLICE_BYTES;
    const uint32_t bar_full = bar0;                              // [2 * NS_MAX] : A_0.., then B_0.. at + 8 * NS_MAX
    const uint32_t bar_empty = bar0 + 8 * 2 * TCS_NS_MAX;
    const uint32_t bar_tfull = bar_empty + 8 * 2 * TCS_NS_MAX;   // [4]
    const uint32_t bar_tempty = bar_tfull + 8 * TCS_STAGES;      // [4]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * (size_t) ns * TC_SLICE_BYTES + 8 * (4 * TCS_NS_MAX + 2 * TCS_STAGES));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * TCS_NS_MAX; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < TCS_STAGES; s++) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, TCS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;      // running (tile, K-half) counter: slot phase = it & 1
            for (int idx = blockIdx.x; idx < P.ntiles; idx += gridDim.x) {
                int ti, tj;
                tcs_decode_tile(idx, P, ti, tj);
                if ((P.ld & 1) == 0) {
                    // the fp64 tile this CTA will read-modify-write ~25 us from now: pull its 128 rows (1 KB each) into L2 so
                    // that the write-back of all CTAs (they run in lockstep) does not hit HBM in one burst
                    const int r0 = (P.tile0 + ti) * TC_TILE, c0 = (P.tile0 + tj) * TC_TILE;
                    const int bytes = min(TC_TILE, P.ce - c0) * 8 & ~15;
                    for (int r = 0; r < TC_TILE && r0 + r < P.nrows; r++)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.A + (size_t) (r0 + r) * P.ld + c0), "r"(bytes) : "memory");
                }
                for (int half = 0; half < 2; half++, it++) {
                    const uint32_t ph = it & 1u;
                    for (int k = 0; k < ns; k++) {
                        TCS_T(p0);
                        mbar_wait_backoff(bar_empty + 8 * k, ph ^ 1u);
                        TCS_T(p1);
                        TCS_ADD(4, p1 - p0); TCS_ADD(5, 1);
                        mbar_expect_tx(bar_full + 8 * k, TC_SLICE_BYTES);
                        tma_bulk_g2s(sA + k * TC_SLICE_BYTES, P.slices + tc_slice_offset(P.tile0 + tj, half, k, ns), TC_SLICE_BYTES, bar_full + 8 * k);
                        mbar_wait_backoff(bar_empty + 8 * (TCS_NS_MAX + k), ph ^ 1u);
                        mbar_expect_tx(bar_full + 8 * (TCS_NS_MAX + k), TC_SLICE_BYTES);
                        tma_bulk_g2s(sB + k * TC_SLICE_BYTES, P.slices + tc_slice_offset(P.tile0 + ti, half, k, ns), TC_SLICE_BYTES, bar_full + 8 * (TCS_NS_MAX + k));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, q = 0;    // q: running level counter -> TMEM stage q & 3, phase (q >> 2) & 1
            for (int idx = blockIdx.x; idx < P.ntiles; idx += gridDim.x) {
                for (int half = 0; half < 2; half++, it++) {
                    const uint32_t ph = it & 1u;
                    for (int d = 0; d < ns; d++, q++) {
                        TCS_T(m0);
                        mbar_wait(bar_full + 8 * d, ph);
                        mbar_wait(bar_full + 8 * (TCS_NS_MAX + d), ph);
                        TCS_T(m1);
                        const uint32_t stage = q & (TCS_STAGES - 1);
                        mbar_wait(bar_tempty + 8 * stage, ((q >> 2) & 1u) ^ 1u);
                        TCS_T(m2);
                        tc_fence_after();
                        const uint32_t tmem_d = tmem_base + stage * TC_TILE;
                        const bool last = (d == ns - 1);
                        for (int pi = 0; pi <= d; pi++) {
                            // last level: outside-in, so that slices 0 (needed first by the next K-half) are released first
                            int k = pi;
                            if (last) k = (pi & 1) ? (d - (pi >> 1)) : (pi >> 1);
                            const int l = d - k;
                            const uint64_t adesc = make_sw128_desc(sA + k * TC_SLICE_BYTES);
                            const uint64_t bdesc = make_sw128_desc(sB + l * TC_SLICE_BYTES);
                            const uint32_t idesc = tcs_idesc(k == 0, l == 0);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)
                                tc_mma_i8(tmem_d, adesc + (uint64_t) (kk * 2), bdesc + (uint64_t) (kk * 2), idesc, (pi > 0 || kk > 0) ? 1u : 0u);
                            if (last) { tc_commit(bar_empty + 8 * k); tc_commit(bar_empty + 8 * (TCS_NS_MAX + l)); }
                        }
                        tc_commit(bar_tfull + 8 * stage);
                        TCS_T(m3);
                        TCS_ADD(0, m1 - m0); TCS_ADD(1, m2 - m1); TCS_ADD(2, m3 - m2); TCS_ADD(3, 1);
                    }
                }
            }
        }
    } else {
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
        const int ch = (warp - 2) >> 2;            // column quarter (32 accumulator columns = 32 matrix rows)
        uint32_t q = 0;
        for (int idx = blockIdx.x; idx < P.ntiles; idx += gridDim.x) {
            int ti, tj;
            tcs_decode_tile(idx, P, ti, tj);
            double acc[32];
#pragma unroll
            for (int j = 0; j < 32; j++) acc[j] = 0.0;
            for (int lev = 0; lev < 2 * ns; lev++, q++) {
                const int d = (lev >= ns) ? lev - ns : lev;
                const double sd = __longlong_as_double((long long) (1023 - 8 * (d + 2)) << 52);
                const uint32_t stage = q & (TCS_STAGES - 1);
                TCS_T(e0);
                mbar_wait(bar_tfull + 8 * stage, (q >> 2) & 1u);
                TCS_T(e1);
                if (warp == 2 && lane == 0) TCS_ADD(8, e1 - e0);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t) (quad * 32) << 16) + stage * TC_TILE + ch * 32;
                uint32_t va[16], vb[16];
                tmem_ld16(taddr, va);
                tmem_ld16(taddr + 16, vb);
                tmem_ld_wait_regs16(va);
                tmem_ld_wait_regs16(vb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * stage);     // accumulator stage free again
#if BSFM_TCS_DBG == 1                      // dev build: no arithmetic (ceiling of the MMA / TMEM side)
                acc[0] += (double) (va[0] ^ vb[15]);
#elif BSFM_TCS_DBG == 2                    // dev build: half of the columns only (is the fold arithmetic the bound?)
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fma(i2d(va[j]), sd, acc[j]);
                acc[16] += (double) vb[3];
#else
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fma(i2d(va[j]), sd, acc[j]);
#pragma unroll
                for (int j = 0; j < 16; j++) acc[16 + j] = fma(i2d(vb[j]), sd, acc[16 + j]);
#endif
                TCS_T(e2);
                if (warp == 2 && lane == 0) TCS_ADD(9, e2 - e1);
            }
            TCS_T(w0);
            // The MMA's M side is the COLUMN tile: thread = matrix column, register j = matrix row, so that for every j
            // the 32 lanes of a warp touch 32 consecutive doubles of one matrix row (coalesced 256-byte segments).
            //   A[row_j][col] -= acc[j] * 2^e_col * 2^e_row_j
            const int col = (P.tile0 + tj) * TC_TILE + quad * 32 + lane;
            const int row0 = (P.tile0 + ti) * TC_TILE + ch * 32;
            const double sc = (col < P.ce) ? P.rscale[col] : 0.0;
            double *acol = P.A + col;
            const bool fast = (ti > tj) && (row0 + 31 < P.nrows) && ((P.tile0 + tj) * TC_TILE + quad * 32 + 31 < P.ce);   // warp-uniform
            if (fast) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double t[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) t[i] = acol[(size_t) (row0 + 8 * b + i) * P.ld];
#pragma unroll
                    for (int i = 0; i < 8; i++) t[i] = fma(-(acc[8 * b + i] * sc), P.rscale[row0 + 8 * b + i], t[i]);
#pragma unroll
                    for (int i = 0; i < 8; i++) acol[(size_t) (row0 + 8 * b + i) * P.ld] = t[i];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int row = row0 + j;
                    if (row < P.nrows && col < P.ce && col <= row) {
                        double *dst = acol + (size_t) row * P.ld;
                        *dst = fma(-(acc[j] * sc), P.rscale[row], *dst);
                    }
                }
            }
            TCS_T(w1);
            if (warp == 2 && lane == 0) { TCS_ADD(10, w1 - w0); TCS_ADD(11, 1); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static unsigned long long *g_tcs_prof = nullptr;
int tc_sm_count()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    return sms;
}
bool tc_syrk_available()
{
    static const bool on = []() { const char *e = getenv("BSFM_BA_TC"); return !(e && e[0] == '0'); }();
    return on;
}
int tc_slices_wanted()
{
    static const int ns = []() {
        const char *e = getenv("BSFM_BA_TC_SLICES");
        int v = e ? atoi(e) : TC_NS_DEFAULT;
        return (v < 3 || v > TCS_NS_MAX) ? TC_NS_DEFAULT : v;
    }();
    return ns;
}

int tc_syrk_update(cudaStream_t st, const TcWorkspace &ws, double *A, const double *, int ld, int nrows, int cb, int ce, int kb, int ke,
                   int col_tile_begin, int col_tile_end, int max_ctas)
{
    if (ke - kb != LNBO || (cb % TC_TILE) != 0 || ws.ns < 1 || ws.ns > TCS_NS_MAX) { set_error("tc_syrk_update: unsupported panel geometry"); return BSFM_ERR_ARG; }
    int dev = 0;
    BSFM_CUDA_TRY(cudaGetDevice(&dev));
    static std::atomic<int> attr_done[64];
    static std::atomic<int> sm_count[64];
    const size_t smem = 2 * (size_t) ws.ns * TC_SLICE_BYTES + 8 * (4 * TCS_NS_MAX + 2 * TCS_STAGES) + 16 + 1024;
    if (dev >= 0 && dev < 64 && !attr_done[dev].load(std::memory_order_acquire)) {
        BSFM_CUDA_TRY(cudaFuncSetAttribute(tc_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (2 * (size_t) TCS_NS_MAX * TC_SLICE_BYTES + 8 * (4 * TCS_NS_MAX + 2 * TCS_STAGES) + 16 + 1024)));
        int sms = 0;
        BSFM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        sm_count[dev].store(sms);
        attr_done[dev].store(1, std::memory_order_release);
    }
    TcSyrkParams P;
    P.slices = ws.slices; P.rscale = ws.rscale; P.A = A; P.ns = ws.ns; P.ld = ld; P.nrows = nrows; P.cb = cb; P.ce = ce;
    P.tile0 = cb / TC_TILE;
    P.nrt = (nrows - cb + TC_TILE - 1) / TC_TILE;
    P.nct = (ce - cb + TC_TILE - 1) / TC_TILE;
    P.tc0 = std::max(0, col_tile_begin);
    P.tc1 = (col_tile_end < 0) ? P.nct : std::min(P.nct, col_tile_end);
    if (P.tc0 >= P.tc1) return BSFM_OK;
    P.ntiles = 0;
    for (int c = P.tc0; c < P.tc1; c++) P.ntiles += P.nrt - c;
    const int sms = (dev >= 0 && dev < 64) ? sm_count[dev].load() : 148;
    int grid = std::min(P.ntiles, sms > 0 ? sms : 148);
    if (max_ctas > 0) grid = std::min(grid, max_ctas);
    static unsigned long long *d_prof = []() -> unsigned long long * {
        if (!getenv("BSFM_TCS_PROF")) return nullptr;
        unsigned long long *q = nullptr;
        if (cudaMalloc(&q, 16 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
        cudaMemset(q, 0, 16 * sizeof(unsigned long long));
        return q;
    }();
    P.prof = d_prof;
    g_tcs_prof = d_prof;
    tc_syrk_kernel<<<grid, TCS_THREADS, smem, st>>>(P);
    BSFM_KERNEL_CHECK();
    return BSFM_OK;
}

}  // namespace ba
}  // namespace bsfm

// dev-only: cycle counters of tc_syrk_kernel's CTA 0 (BSFM_TCS_PROF=1); reading resets them
extern "C" int bsfm_debug_tcs_prof(unsigned long long *out16)
{
    if (!bsfm::ba::g_tcs_prof) return -1;
    if (cudaMemcpy(out16, bsfm::ba::g_tcs_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    cudaMemset(bsfm::ba::g_tcs_prof, 0, 16 * sizeof(unsigned long long));
    return 0;
}
