// ba_chol_large.cuh -- shared definitions of the large-system Cholesky path (ba_chol_large.cu, ba_chol_tc.cu).
//
// int8-slice ("Ozaki") representation consumed by the tcgen05 trailing update:
//   row r of the current 256-column panel of L is scaled by 2^-e_r so that every entry lies in (-0.5, 0.5) and is
//   written as an 8 NS-bit two's-complement fixed-point number  v 2^-e_r = sum_k b_k(r,t) 2^(-8 (k+1)),  k = 0..NS-1,
//   b_0 SIGNED (int8), b_1.. UNSIGNED (uint8).  Then
//     (L L^T)[r][c] = 2^(e_r + e_c) sum_d 2^(-8 (d+2)) C_d[r][c],    C_d = sum_{k+l=d} B_k B_l^T   (exact in int32)
//   and levels d >= NS are dropped (relative size <= 2^-(8 NS - 11): 3e-14 for NS = 7, the size of fp64 rounding in a
//   K = 256 dot product).
// HBM layout of the slices: 128-row tiles (tile = matrix row / 128) x 2 K-halves (128 panel columns each) x NS
// slices, each a 16 KB block in the UMMA "K-major SWIZZLE_128B" shared-memory image (8-row x 128-byte atoms, 16-byte
// chunk c of row r stored at chunk c ^ (r & 7)), so one cp.async.bulk brings a slice tile into shared memory ready
// for tcgen05.mma -- the layout the MATCH descriptors use (match_kernels.cuh).
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

namespace bsfm {
namespace ba {

struct Scalars;

constexpr int LNB = 32;          // sub-block
constexpr int LNBO = 256;        // outer panel
constexpr int TC_TILE = 128;     // rows per slice tile = UMMA M = UMMA N
constexpr int TC_KHALF = 128;    // panel columns per K-half (= bytes per slice row)
constexpr int TC_SLICE_BYTES = TC_TILE * TC_KHALF;   // 16 KB
constexpr int TC_NS_MAX = 8;
constexpr int TC_NS_DEFAULT = 7;

struct SliceOut {
    uint8_t *slices;     // null: do not emit
    double *rscale;      // 2^e_r per matrix row
    int ns;
};

struct TcWorkspace {
    uint8_t *slices;     // ceil((n+1)/128) tiles x 2 x ns x 16 KB
    double *rscale;      // n+1
    int ns;
    int ntiles;
};

inline size_t tc_slices_bytes(int n, int ns) { return (size_t) ((n + 1 + TC_TILE - 1) / TC_TILE) * 2 * ns * TC_SLICE_BYTES; }
bool tc_syrk_available();
int tc_slices_wanted();
inline SliceOut tc_slice_out(const TcWorkspace &ws, int, int) { SliceOut so; so.slices = ws.slices; so.rscale = ws.rscale; so.ns = ws.ns; return so; }
// A[r][c] -= sum_{t in [kb,ke)} L[r][t] L[c][t]  for c in [cb,ce), r in [c, nrows)  from the slices of panel [kb,ke)
// restricted to the 128-column tile columns [col_tile_begin, col_tile_end) of the trailing matrix (end < 0: all) on at most
// `max_ctas` CTAs (0: one per SM)
int tc_syrk_update(cudaStream_t st, const TcWorkspace &ws, double *A, const double *L, int ld, int nrows, int cb, int ce, int kb, int ke,
                   int col_tile_begin = 0, int col_tile_end = -1, int max_ctas = 0);
int tc_sm_count();
int chol_solve_large(cudaStream_t st, double *A, double *Lmat, int n, double *linv_ws, double *xinv_ws, double *x, Scalars *sc, const TcWorkspace *ws);
// doubles of linear-solver workspace behind the 32 x 32 inverses: back-substitution scratch (small path) or the 256 x 256 panel inverses (large path)
constexpr int DF_MAX_N = 24 * LNB;       // largest system the dataflow path (ba_chol_dataflow.cu) takes: it needs (n + 1) x n doubles more
inline size_t chol_extra_ws_doubles(int n)
{
    // small systems: (n + 1) x n behind the scratch; large ones finish their last panels with the dataflow kernel: room for its largest case
    return (size_t) ((n + LNBO - 1) / LNBO) * LNBO * LNBO + (size_t) n + 64 + (size_t) (std::min(n, DF_MAX_N) + 1) * std::min(n, DF_MAX_N) + 64;
}
// where the large path keeps that scratch inside linv_ws
inline size_t chol_large_pub_offset(int n) { return (size_t) ((n + LNB - 1) / LNB) * LNB * LNB + (size_t) ((n + LNBO - 1) / LNBO) * LNBO * LNBO + (size_t) n + 64; }
int chol_dataflow_factor(cudaStream_t st, const double *A, double *Lmat, int ld, int n, double *linv_blocks, double *pub, Scalars *sc, bool *used);

__host__ __device__ __forceinline__ size_t tc_slice_offset(int tile, int half, int k, int ns)
{
    return ((size_t) (tile * 2 + half) * ns + k) * TC_SLICE_BYTES;
}

// Called by all 256 threads of chol_trsm_kernel after the panel rows [row0, row0 + 64) are final in shared memory
// (Xs[r * ldx + t], t = 0..255, zero beyond the panel); rmax[r] = bit pattern of max_t |Xs[r][t]|.
// Rows >= row0 + rcount get zero slices.
__device__ __forceinline__ void emit_slices(const SliceOut &so, const double *Xs, int ldx, int row0, int rcount, const unsigned long long *rmax)
{
    __shared__ double s_mul[64];
    const int tid = threadIdx.x;
    if (tid < 64) {
        const double m = __longlong_as_double((long long) rmax[tid]);
        int e = 0;
        if (m > 0.0 && isfinite(m)) {
            int q;
            frexp(m, &q);          // m = f 2^q, f in [0.5, 1)
            e = max(q + 1, -900);  // |v| 2^-e < 0.5
        }
        s_mul[tid] = (tid < rcount) ? ldexp(1.0, 8 * so.ns - e) : 0.0;
        if (tid < rcount) so.rscale[row0 + tid] = ldexp(1.0, e);
    }
    __syncthreads();
    const int ns = so.ns;
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        const int item = tid + 256 * it;          // (row, 16-column chunk)
        const int r = item & 63, ch = item >> 6;  // ch 0..15: columns ch*16 .. +15
        const int row = row0 + r;
        const int tile = row >> 7, rt = row & 127;
        const int half = ch >> 3, c16 = ch & 7;
        const double mul = s_mul[r];
        long long v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = __double2ll_rd(Xs[r * ldx + ch * 16 + j] * mul);
        uint8_t *base = so.slices + tc_slice_offset(tile, half, 0, ns) + (size_t) (rt >> 3) * 1024 + (size_t) (rt & 7) * 128 + (size_t) ((c16 ^ (rt & 7)) << 4);
        for (int k = 0; k < ns; k++) {
            const int sh = 8 * (ns - 1 - k);
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] = (uint32_t) ((v[4 * q] >> sh) & 0xff) | ((uint32_t) ((v[4 * q + 1] >> sh) & 0xff) << 8) |
                       ((uint32_t) ((v[4 * q + 2] >> sh) & 0xff) << 16) | ((uint32_t) ((v[4 * q + 3] >> sh) & 0xff) << 24);
            }
            *reinterpret_cast<uint4 *>(base + (size_t) k * TC_SLICE_BYTES) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

}  // namespace ba
}  // namespace bsfm
