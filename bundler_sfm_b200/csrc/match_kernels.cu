// match_kernels.cu -- sm_100a kernels of the MATCH hot path.  See match_kernels.cuh for layout.
#include "match_kernels.cuh"
#include "tc_ptx.cuh"
#include <climits>

namespace bsfm {
namespace match {

// ---------------------------------------------------------------------------------------------
// prep: raw host-order descriptors (n_i x 128 per image, concatenated) -> squared norms (for the per-image
// sort), then the padded, norm-sorted, swizzled layout.  One warp per row; lane l handles bytes [4l, 4l+4).
// ---------------------------------------------------------------------------------------------
__global__ void raw_norm_kernel(const uint8_t *__restrict__ raw, int64_t total_keys, int32_t *__restrict__ norms_raw, int32_t *__restrict__ iota)
{
    const int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= total_keys) return;
    const uint8_t *src = raw + row * DESC_BYTES + lane * 4;
    const uint32_t w = (uint32_t) src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16) | ((uint32_t) src[3] << 24);
    uint32_t s = __dp4a(w, w, 0u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { norms_raw[row] = (int32_t) s; iota[row] = (int32_t) row; }
}

// sorted_src[key_off[img] + pos] = raw row (global) that is the pos-th smallest-norm key of image img
__global__ void prep_kernel(const uint8_t *__restrict__ raw, const int64_t *__restrict__ key_off,
                            const int32_t *__restrict__ img_doff, const int32_t *__restrict__ tile_img,
                            const int32_t *__restrict__ sorted_src, const int32_t *__restrict__ sorted_norms,
                            uint8_t *__restrict__ keys_sw, int32_t *__restrict__ norms, int32_t *__restrict__ perm,
                            int64_t row_begin, int64_t row_end, int64_t key_base)
{
    // rows [row_begin, row_end) of the device layout; `raw`, `sorted_src`, `sorted_norms` hold the keys [key_base, ...) only
    // (a rank of a sharded build prepares its own images; key_base = 0 and the whole row range otherwise)
    int64_t row = row_begin + (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= row_end) return;
    int img = tile_img[row >> 7];
    uint32_t w = 0;
    int32_t nrm = NORM_PAD, orig = -1;
    if (img >= 0) {
        int64_t k = row - img_doff[img];
        int64_t n = key_off[img + 1] - key_off[img];
        if (k < n) {
            const int64_t srow = sorted_src[key_off[img] - key_base + k];      // row inside `raw`
            const uint8_t *src = raw + srow * DESC_BYTES + lane * 4;
            w = (uint32_t) src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16) | ((uint32_t) src[3] << 24);
            nrm = sorted_norms[key_off[img] - key_base + k];
            orig = (int32_t) (srow + key_base - key_off[img]);
        }
    }
    int c = lane >> 2;  // 16-byte chunk
    uint32_t *dst = reinterpret_cast<uint32_t *>(keys_sw + sw_chunk_offset(row, c)) + (lane & 3);
    *dst = w;
    if (lane == 0) { norms[row] = nrm; perm[row] = orig; }
}

// ---------------------------------------------------------------------------------------------
// DP4A kernel: one CTA per work unit (128 query rows x whole database image).  CUDA-core second
// implementation used to cross-check the tensor-core kernel at full size and as the selectable
// BSFM_MATCH_KERNEL_DP4A path.  Exact top-2 with index per row; emits final matches directly.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void top2_insert(int &m1, int &m2, int &mi, int t, int idx)
{
    // keeps (m1 <= m2); first-seen wins ties (pr_queue_k.h:102-117 strict '>')
    if (t < m1) { m2 = m1; m1 = t; mi = idx; }
    else if (t < m2) { m2 = t; }
}

__global__ void __launch_bounds__(256) match_dp4a_kernel(MatchParams P)
{
    __shared__ uint32_t sQ[32][TILE_Q];
    __shared__ uint32_t sD[32][128];
    __shared__ int32_t sN[128];

    const int u = P.unit_begin + blockIdx.x;
    if (u >= P.unit_end) return;
    const int k = find_run_image(P.run_imgs, P.num_run_imgs, u);
    const RunImage R = P.run_imgs[k];
    const int64_t a_row0 = (int64_t) (R.atile0 + (u - R.unit0)) * TILE_Q;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    // query tile -> sQ[k4][row]
    for (int idx = tid; idx < TILE_Q * 8; idx += 256) {
        int c = idx >> 7, row = idx & 127;
        uint4 v = *reinterpret_cast<const uint4 *>(P.keys_sw + sw_chunk_offset(a_row0 + row, c));
        sQ[c * 4 + 0][row] = v.x; sQ[c * 4 + 1][row] = v.y; sQ[c * 4 + 2][row] = v.z; sQ[c * 4 + 3][row] = v.w;
    }

    int m1[8], m2[8], mi[8];
#pragma unroll
    for (int a = 0; a < 8; a++) { m1[a] = INT_MAX; m2[a] = INT_MAX; mi[a] = -1; }

    const int ntiles = R.ntiles_db * 2;  // 128-row database tiles
    for (int tile = 0; tile < ntiles; tile++) {
        const int64_t d_row0 = (int64_t) R.db_row0 + (int64_t) tile * 128;
        __syncthreads();
        for (int idx = tid; idx < 128 * 8; idx += 256) {
            int c = idx >> 7, row = idx & 127;
            uint4 v = *reinterpret_cast<const uint4 *>(P.keys_sw + sw_chunk_offset(d_row0 + row, c));
            sD[c * 4 + 0][row] = v.x; sD[c * 4 + 1][row] = v.y; sD[c * 4 + 2][row] = v.z; sD[c * 4 + 3][row] = v.w;
        }
        if (tid < 128) sN[tid] = P.norms[d_row0 + tid];
        __syncthreads();

        uint32_t acc[8][8];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++) acc[a][b] = 0;
#pragma unroll 4
        for (int k4 = 0; k4 < 32; k4++) {
            uint32_t q[8], d[8];
#pragma unroll
            for (int a = 0; a < 8; a++) q[a] = sQ[k4][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; b++) d[b] = sD[k4][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 8; b++) acc[a][b] = __dp4a(q[a], d[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int col = tile * 128 + tx + 16 * b;
            const int nb = sN[tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; a++) {
                int t = nb - 2 * (int) acc[a][b];
                top2_insert(m1[a], m2[a], mi[a], t, col);
            }
        }
    }

    // merge the 16 column-threads of every row (lanes differing in the low 4 bits)
#pragma unroll
    for (int a = 0; a < 8; a++) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            int o1 = __shfl_xor_sync(0xffffffffu, m1[a], o);
            int o2 = __shfl_xor_sync(0xffffffffu, m2[a], o);
            int oi = __shfl_xor_sync(0xffffffffu, mi[a], o);
            int hi = max(m1[a], o1);
            int lo2 = min(m2[a], o2);
            if (o1 < m1[a]) mi[a] = oi;
            m1[a] = min(m1[a], o1);
            m2[a] = min(hi, lo2);
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int r = ty + 16 * a;
            const int na = P.norms[a_row0 + r];
            if (na >= NORM_PAD_HALF) continue;  // padding query row
            const int d0 = na + m1[a];
            const int d1 = (m2[a] >= NORM_PAD_HALF) ? INT_MAX : na + m2[a];
            if (ratio_pass(P, d0, d1)) {
                int pos = atomicAdd(&P.counters[1], 1);
                if (pos < P.match_cap) {
                    P.match_slot[pos] = match_sort_key(P, R, a_row0 + r);
                    P.match_idx2[pos] = mi[a];
                } else {
                    P.counters[2] = 1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel: persistent, warp-specialised.
//   warp 0      : TMA producer (cp.async.bulk, one 16 KB query tile per unit, 32 KB database tiles)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::i8, M128 N256 K32 x4)
//   warps 2..9  : epilogue; warp w reads TMEM lane quadrant w%4 (thread == query row) and column
//                 half (w-2)/4 of every 256-column tile; the two halves of a row are merged at unit end
// Accumulators double-buffered in TMEM (2 x 256 columns); shared-memory database ring of 4 stages.
// Epilogue per element: t = |p|^2 - 2*dot (IMAD), chunk minimum (VIMNMX3 tree).  Per 32-column
// chunk the row keeps (m1 = smallest chunk-min, s2 = second smallest chunk-min, bchunk).  d0 is
// exact; d1 is bounded above by s2, so only rows passing the ratio test with that bound can match:
// they are pushed as candidates and resolved exactly by match_verify_kernel.
// ---------------------------------------------------------------------------------------------

using namespace bsfm::ptx;   // tc_ptx.cuh
constexpr uint32_t TC_IDESC = (2u << 4) | ((uint32_t) (TILE_DB >> 3) << 17) | ((uint32_t) (TILE_Q >> 4) << 24);
constexpr uint32_t TC_IDESC_N128 = (2u << 4) | ((uint32_t) (128 >> 3) << 17) | ((uint32_t) (TILE_Q >> 4) << 24);
constexpr uint32_t TC_IDESC_PAIR = (2u << 4) | ((uint32_t) (TILE_DB >> 3) << 17) | ((uint32_t) ((2 * TILE_Q) >> 4) << 24);
// 3-input MAX (VIMNMX3) over the 32 values of a chunk and `init`: four independent chains
__device__ __forceinline__ int chunk_max(const uint32_t (&v)[32], int init)
{
    int a0 = max(max((int) v[0], (int) v[1]), init);
    int a1 = max(max((int) v[2], (int) v[3]), (int) v[4]);
    int a2 = max(max((int) v[5], (int) v[6]), (int) v[7]);
    int a3 = max(max((int) v[8], (int) v[9]), (int) v[10]);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        a0 = max(max((int) v[11 + 8 * q], (int) v[12 + 8 * q]), a0);
        a1 = max(max((int) v[13 + 8 * q], (int) v[14 + 8 * q]), a1);
        a2 = max(max((int) v[15 + 8 * q], (int) v[16 + 8 * q]), a2);
        a3 = max(max((int) v[17 + 8 * q], (int) v[18 + 8 * q]), a3);
    }
    a0 = max(max((int) v[27], (int) v[28]), a0);
    a1 = max(max((int) v[29], (int) v[30]), a1);
    a2 = max((int) v[31], a2);
    return max(max(a0, a1), max(a2, a3));
}

struct UnitInfo {
    int64_t a_row0;
    int32_t db_row0, n, ntiles_db;
};
// Work units of a CTA ascend, and most consecutive units stay inside the same run image: every role keeps a cursor
// (the image's unit range and fields in registers) and only touches global memory when a unit leaves the image --
// a binary search per unit would put a chain of dependent global loads on the MMA thread's critical path.
struct UnitCursor {
    int32_t k = -1, unit0 = 0, unit_end = 0;      // current run image and its unit range [unit0, unit_end)
    int32_t atile0 = 0, db_row0 = 0, n = 0, ntiles_db = 0;
};
__device__ __forceinline__ void cursor_load(const MatchParams &P, UnitCursor &c)
{
    const RunImage *R = P.run_imgs + c.k;
    c.unit0 = R->unit0; c.unit_end = R->unit0 + R->nunits;
    c.atile0 = R->atile0; c.db_row0 = R->db_row0; c.n = R->n; c.ntiles_db = R->ntiles_db;
}
__device__ __forceinline__ UnitInfo decode_unit(const MatchParams &P, UnitCursor &c, int u)
{
    if (c.k < 0 || u >= c.unit_end) {
        if (c.k >= 0 && c.k + 1 < P.num_run_imgs) { c.k++; cursor_load(P, c); }       // usually the next image
        if (c.k < 0 || u < c.unit0 || u >= c.unit_end) { c.k = find_run_image(P.run_imgs, P.num_run_imgs, u); cursor_load(P, c); }
    }
    UnitInfo U;
    U.a_row0 = (int64_t) (c.atile0 + (u - c.unit0)) * TILE_Q;
    U.db_row0 = c.db_row0;
    U.n = c.n;
    U.ntiles_db = c.ntiles_db;
    return U;
}

// BOUND = true : bound epilogue (every database image of the launch fits the norm staging buffer)
// BOUND = false: exact chunk-minimum epilogue (any image size)
// Dev-only cycle accounting of the pipeline roles (scripts/dev_match_prof.py builds a library with -DBSFM_TC_PROFILE):
// g_prof[0..3] MMA thread: wait b_full, wait t_empty, issue, tiles;  [4..6] producer: wait b_empty, issue, tiles;
// [8..11] epilogue warp 2 lane 0: wait t_full, loads+reduce until release, after release, tiles.  CTA 0 only.
#ifdef BSFM_TC_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, v) do { if (blockIdx.x == 0) atomicAdd(&g_prof[i], (unsigned long long) (v)); } while (0)
extern "C" int bsfm_debug_prof(unsigned long long *out)
{
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out, g_prof, sizeof z) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(g_prof, z, sizeof z) == cudaSuccess ? 0 : -1;
}
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(i, v) do {} while (0)
#endif
#ifndef BSFM_TC_MMA_SPIN
#define BSFM_TC_MMA_SPIN 1
#endif
// PAIR  = true : two CTAs of a cluster (an SM pair) issue ONE cta_group::2 MMA per database tile: M = 256 (128
//                query rows per CTA), every CTA stages only HALF of the 256-row database tile, which halves the
//                L2 -> SM traffic per MAC (the single-CTA kernel is L2-bandwidth bound, profiles/r1_match_tc_v6).
// QUAD = true : (bound epilogue, single CTA) FOUR accumulator stages of 128 columns instead of two of 256: every 256-row
//                database tile is issued as two N = 128 MMAs groups into consecutive stages and the 16 epilogue warps form four
//                groups, one per stage.  The stage hand-over (MMA -> commit -> epilogue wake-up -> TMEM loads -> release -> MMA)
//                costs ~650 cycles on top of the loads; with four stages in flight twice as many hand-overs overlap.
template <bool BOUND, bool PAIR, bool QUAD = false>
__device__ __forceinline__ void match_tc_body(const MatchParams &P)
{
    static_assert(!QUAD || (BOUND && !PAIR), "the four-stage variant exists for the bound epilogue on single CTAs");
    constexpr int NTS = QUAD ? 4 : 2;                 // accumulator stages
    constexpr int TS_COLS = QUAD ? 128 : TILE_DB;     // TMEM columns per stage
    extern __shared__ uint8_t smem_raw[];
    // manual 1024-byte alignment (SWIZZLE_128B atoms)
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw_addr);

    const uint32_t sA = base + TC_SMEM_A;
    const uint32_t sB = base + TC_SMEM_B;
    int32_t *sN = reinterpret_cast<int32_t *>(smem + TC_SMEM_N);
    const uint32_t bar0 = base + TC_SMEM_BAR;
    // barrier map (8 bytes each)
    constexpr int BST = PAIR ? 2 * TC_B_STAGES : TC_B_STAGES;       // database tile stages
    constexpr uint32_t BSLOT = PAIR ? TC_B_BYTES / 2 : TC_B_BYTES;   // bytes this CTA stages per database tile
    const uint32_t bar_b_full = bar0;                          // [BST]
    const uint32_t bar_b_empty = bar0 + 8 * BST;               // [BST]
    const uint32_t bar_a_full = bar0 + 16 * BST;               // [2]
    const uint32_t bar_a_empty = bar_a_full + 16;              // [2]
    const uint32_t bar_t_full = bar_a_empty + 16;              // [NTS]
    const uint32_t bar_t_empty = bar_t_full + 8 * NTS;         // [NTS]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TC_SMEM_BAR + 8 * (2 * BST + 4 + 2 * NTS));
    static_assert(8 * (2 * BST + 4 + 2 * NTS) + 4 <= 256, "barrier region");
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs), 1 = peer

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // pair mode: the leader's "full" barriers also count the peer's relay arrival (its half is staged), and
        // the leader's accumulator-empty barriers count one arrival per epilogue warp of the group in BOTH CTAs
        const uint32_t full_cnt = (PAIR && crank == 0) ? 2 : 1;
        for (int s = 0; s < BST; s++) { mbar_init(bar_b_full + 8 * s, full_cnt); mbar_init(bar_b_empty + 8 * s, 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_a_full + 8 * s, full_cnt);
            mbar_init(bar_a_empty + 8 * s, 1);
        }
        for (int s = 0; s < NTS; s++) {
            mbar_init(bar_t_full + 8 * s, 1);
            // bound mode: one warp group per stage; pair: one arrival per warp of the group in both CTAs
            mbar_init(bar_t_empty + 8 * s, QUAD ? TC_EPI_WARPS / 4 : (!BOUND ? TC_EPI_THREADS : (PAIR ? 2 : 1) * TC_EPI_WARPS / 2));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer's barriers must be initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // pair mode: cluster c takes the unit pairs (2c, 2c+1), (2c + 2C, ...): both units share the database image
    // (images are padded to 256 rows, so every image contributes an even number of 128-row query units)
    const int u_first = PAIR ? P.unit_begin + 2 * (int) (blockIdx.x >> 1) + (int) crank : P.unit_begin + (int) blockIdx.x;
    const int u_step = PAIR ? (int) (gridDim.x & ~1u) : (int) gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t bs = 0, bph = 0, as = 0, aph = 0;
            UnitCursor cur;
            for (int u = u_first; u < P.unit_end; u += u_step) {
                const UnitInfo U = decode_unit(P, cur, u);
                mbar_wait_backoff(bar_a_empty + 8 * as, aph ^ 1);
                mbar_expect_tx(bar_a_full + 8 * as, TC_A_BYTES);
                tma_bulk_g2s(sA + as * TC_A_BYTES, P.keys_sw + (size_t) U.a_row0 * DESC_BYTES, TC_A_BYTES, bar_a_full + 8 * as);
                as ^= 1; if (as == 0) aph ^= 1;
                // pair mode: this CTA stages rows [128 crank, 128 crank + 128) of every 256-row database tile
                const uint8_t *src = P.keys_sw + (size_t) U.db_row0 * DESC_BYTES + (PAIR ? crank * BSLOT : 0u);
                for (int t = 0; t < U.ntiles_db; t++) {
                    PROF_T(p0);
                    mbar_wait_backoff(bar_b_empty + 8 * bs, bph ^ 1);
                    PROF_T(p1);
                    mbar_expect_tx(bar_b_full + 8 * bs, BSLOT);
                    tma_bulk_g2s(sB + bs * BSLOT, src + (size_t) t * TC_B_BYTES, BSLOT, bar_b_full + 8 * bs);
                    if (++bs == BST) { bs = 0; bph ^= 1; }
                    PROF_T(p2);
                    PROF_ADD(4, p1 - p0); PROF_ADD(5, p2 - p1); PROF_ADD(6, 1);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t bs = 0, bph = 0, as = 0, aph = 0, ts = 0, tph = 0;
            if (PAIR && crank != 0) {
                // peer CTA: no MMA to issue.  This warp relays "my half is staged" to the leader's barriers.
                UnitCursor cur;
            for (int u = u_first; u < P.unit_end; u += u_step) {
                    const UnitInfo U = decode_unit(P, cur, u);
                    mbar_wait_backoff(bar_a_full + 8 * as, aph);
                    mbar_arrive_cluster(bar_a_full + 8 * as, 0);
                    as ^= 1; if (as == 0) aph ^= 1;
                    for (int t = 0; t < U.ntiles_db; t++) {
                        mbar_wait_backoff(bar_b_full + 8 * bs, bph);
                        mbar_arrive_cluster(bar_b_full + 8 * bs, 0);
                        if (++bs == BST) { bs = 0; bph ^= 1; }
                    }
                }
            } else {
            UnitCursor cur;
            for (int u = u_first; u < P.unit_end; u += u_step) {
                const UnitInfo U = decode_unit(P, cur, u);
                PROF_T(qa0);
                if (PAIR) mbar_wait_backoff_cluster(bar_a_full + 8 * as, aph); else mbar_wait_backoff(bar_a_full + 8 * as, aph);
                PROF_T(qa1);
                PROF_ADD(12, qa1 - qa0); PROF_ADD(13, 1);
                const uint64_t adesc = make_sw128_desc(sA + as * TC_A_BYTES);
                for (int t = 0; t < U.ntiles_db; t++) {
                    PROF_T(q0);
                    if (PAIR) {
                        mbar_wait_backoff_cluster(bar_b_full + 8 * bs, bph);
                        mbar_wait_backoff_cluster(bar_t_empty + 8 * ts, tph ^ 1);
                    } else {
                        mbar_wait_backoff(bar_b_full + 8 * bs, bph);
                        PROF_T(q1);
                        PROF_ADD(0, q1 - q0);
                        // the accumulator-stage hand-over is on the critical path: poll without sleeping
                        if (BSFM_TC_MMA_SPIN) mbar_wait(bar_t_empty + 8 * ts, tph ^ 1); else mbar_wait_backoff(bar_t_empty + 8 * ts, tph ^ 1);
                    }
                    PROF_T(q2);
                    tc_fence_after();
                    const uint64_t bdesc = make_sw128_desc(sB + bs * BSLOT);
                    if constexpr (QUAD) {
                        // rows 0..127 and 128..255 of the database tile into two consecutive 128-column stages
#pragma unroll
                        for (int half = 0; half < 2; half++) {
                            if (half) { mbar_wait(bar_t_empty + 8 * ts, tph ^ 1); tc_fence_after(); }
                            const uint32_t tmem_d = tmem_base + ts * TS_COLS;
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)      // second half: +128 rows x 128 B = +1024 in the descriptor start field
                                tc_mma_i8(tmem_d, adesc + (uint64_t) (kk * 2), bdesc + (uint64_t) (half * 1024 + kk * 2), TC_IDESC_N128, kk > 0);
                            if (half) tc_commit(bar_b_empty + 8 * bs);
                            tc_commit(bar_t_full + 8 * ts);
                            ts = (ts + 1) & 3; if (ts == 0) tph ^= 1;
                        }
                        if (++bs == BST) { bs = 0; bph ^= 1; }
                    } else {
                    const uint32_t tmem_d = tmem_base + ts * TILE_DB;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {  // K = 4 x 32 bytes; +32 B = +2 in the descriptor start field
                        if (PAIR) tc_mma_i8_pair(tmem_d, adesc + (uint64_t) (kk * 2), bdesc + (uint64_t) (kk * 2), TC_IDESC_PAIR, kk > 0);
                        else tc_mma_i8(tmem_d, adesc + (uint64_t) (kk * 2), bdesc + (uint64_t) (kk * 2), TC_IDESC, kk > 0);
                    }
                    if (PAIR) { tc_commit_pair(bar_b_empty + 8 * bs); tc_commit_pair(bar_t_full + 8 * ts); }
                    else { tc_commit(bar_b_empty + 8 * bs); tc_commit(bar_t_full + 8 * ts); }
                    if (++bs == BST) { bs = 0; bph ^= 1; }
                    ts ^= 1; if (ts == 0) tph ^= 1;
                    }
                    PROF_T(q3);
                    PROF_ADD(1, q2 - q0); PROF_ADD(2, q3 - q2); PROF_ADD(3, 1);   // [1] includes [0]
                }
                if (PAIR) tc_commit_pair(bar_a_empty + 8 * as); else tc_commit(bar_a_empty + 8 * as);
                as ^= 1; if (as == 0) aph ^= 1;
            }
            }
        }
    } else {
        // ===================== epilogue (TC_EPI_WARPS warps) =====================
        constexpr int NPART = TC_EPI_WARPS / 4;            // column parts per 256-column tile
        constexpr int NCH = TILE_DB / CHUNK / NPART;       // 32-column chunks per warp per tile
        const int quad = warp & 3;               // TMEM lane quadrant this warp may read
        const int part = (warp - 2) >> 2;        // which 256/NPART columns of each tile
        const int row = quad * 32 + lane;        // query row inside the unit
        const int etid = threadIdx.x - 64;       // 0..TC_EPI_THREADS-1
        const int neg2 = P.neg2;                 // runtime -2: keeps the multiply-add on the FMA pipe (IMAD)
        int *xch = reinterpret_cast<int *>(smem + TC_SMEM_XCH);   // [NPART-1][5][128] exchange between column parts
        int *sNall = reinterpret_cast<int *>(smem + TC_SMEM_NALL);
        int staged_row0 = -1;
        uint32_t ts = 0, tph = 0;
        // bound mode: two warp groups, each owning one accumulator stage (see the tile loop)
        // (four-stage variant: four groups, one per 128-column stage, every warp reduces a whole stage of its lane quadrant)
        constexpr int GNPART = QUAD ? 1 : NPART / 2;      // column parts per stage inside a group
        constexpr int GNCH = TS_COLS / CHUNK / (GNPART > 0 ? GNPART : 1);
        const uint32_t grp = QUAD ? (uint32_t) part : ((uint32_t) part & 1u);        // accumulator stage this warp serves
        const int gpart = QUAD ? 0 : (part >> 1);
        uint32_t gtile = 0, gph = 0;                      // running tile number at unit start, phase of the group's stage
        UnitCursor cur;
            for (int u = u_first; u < P.unit_end; u += u_step) {
            const UnitInfo U = decode_unit(P, cur, u);
            PROF_T(eu0);
            const int na = P.norms[U.a_row0 + row];
            // row state.  exact mode : m1 = smallest chunk-min of t, s2 = second smallest chunk-min, bchunk.
            //             bound mode : m1 = smallest chunk LOWER bound, s2 = second smallest lower bound,
            //                          ubb = UPPER bound of the best chunk, uo = smallest upper bound of the others.
            int m1 = INT_MAX, s2 = INT_MAX, bchunk = 0, ubb = INT_MAX, uo = INT_MAX;
            // Database norms: when the whole image fits (<= TC_NORM_CAP rows) they are staged ONCE per
            // (CTA, database image) -- consecutive units of a CTA share the image -- so the tile loop has
            // no CTA-level barrier; larger images fall back to staging 256 norms per tile.
            const bool whole = BOUND ? true : (U.ntiles_db * TILE_DB <= TC_NORM_CAP);   // host guarantees it when BOUND
            // bound mode (default): keys are norm-sorted inside an image, so a chunk's norms span
            // [norm(first), norm(last)] and  min_c t  lies in  [nbmin - 2 vmax, nbmax - 2 vmax]  with vmax the
            // chunk's largest dot product: the hot loop is a bare 3-input MAX tree (0.5 instruction / element).
            constexpr bool bound = BOUND;   // host guarantees `whole` for every image when BOUND
            if (whole && staged_row0 != U.db_row0) {
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");   // everyone done with the old image
                for (int q = etid; q < U.ntiles_db * TILE_DB; q += TC_EPI_THREADS) sNall[q] = P.norms[(size_t) U.db_row0 + q];
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
                staged_row0 = U.db_row0;
            }
            if constexpr (BOUND) {
                // The epilogue warps form TWO groups; group g reduces the tiles whose running number is g (mod 2),
                // i.e. it owns accumulator stage g.  The groups run half a tile period apart, so while one waits
                // for its accumulator / its first TMEM load the other keeps the ALU pipe busy.  Inside a tile the
                // TMEM loads are software pipelined (chunk c+1 in flight while chunk c is reduced) and the stage
                // is handed back to the MMA warp as soon as its last chunk sits in registers.
                static_assert(BCHUNK % CHUNK == 0 && (GNCH * CHUNK) % BCHUNK == 0, "bound chunk must tile a column part");
                const uint32_t tlane = tmem_base + ((uint32_t) (quad * 32) << 16) + grp * TS_COLS + gpart * (TS_COLS / GNPART);
                // hs = running stage-sized step inside the unit: a 256-column tile (two stages) or one of its 128-column halves (four)
                constexpr int NSTEP = QUAD ? 2 : 1;       // steps per database tile
                for (int hs = (int) ((grp - gtile) & (uint32_t) (NTS - 1)); hs < NSTEP * U.ntiles_db; hs += NTS) {
                    const int t = QUAD ? (hs >> 1) : hs;
                    const int gp = QUAD ? (hs & 1) : gpart;      // which 128-column half of tile t this warp reduces
                    const int *nbs = sNall + t * TILE_DB + gp * 128;
                    PROF_T(e0);
                    mbar_wait(bar_t_full + 8 * grp, gph);
                    PROF_T(e1);
                    tc_fence_after();
                    uint32_t va[32], vb[32];
                    auto bookkeep = [&](int vmax, int cb) {
                        const int chunk_id = t * (TILE_DB / BCHUNK) + gp * (GNCH * CHUNK / BCHUNK) + cb;
                        const int lb = vmax * neg2 + nbs[cb * BCHUNK];
                        const int ub = vmax * neg2 + nbs[cb * BCHUNK + BCHUNK - 1];
                        const bool nb_best = lb < m1;
                        s2 = nb_best ? m1 : min(s2, lb);
                        uo = min(uo, nb_best ? ubb : ub);
                        ubb = nb_best ? ub : ubb;
                        bchunk = nb_best ? chunk_id : bchunk;
                        m1 = min(m1, lb);
                    };
                    auto release_stage = [&]() {
                        tc_fence_before();
                        __syncwarp();       // one arrival per warp: 32 same-address arrivals per warp would serialise
                        if (lane == 0) {
                            if (PAIR) mbar_arrive_cluster(bar_t_empty + 8 * grp, 0);   // the LEADER's MMA overwrites both CTAs' stage
                            else mbar_arrive(bar_t_empty + 8 * grp);                   // accumulator stage free again
                        }
                        gph ^= 1;
#ifdef BSFM_TC_PROFILE
                        if (warp == 2 && lane == 0) { PROF_ADD(8, e1 - e0); PROF_ADD(9, clock64() - e1); PROF_ADD(11, 1); }
#endif
                    };
                    tmem_ld32(tlane, va);
                    int vmax = 0;
#pragma unroll
                    for (int c = 0; c < GNCH; c++) {
                        uint32_t (&v)[32] = (c & 1) ? vb : va;
                        uint32_t (&vn)[32] = (c & 1) ? va : vb;
                        tmem_ld_wait_regs(v);
                        if (c + 1 < GNCH) tmem_ld32(tlane + (c + 1) * CHUNK, vn);
                        else release_stage();
                        vmax = chunk_max(v, vmax);
                        if (((c + 1) * CHUNK) % BCHUNK == 0) {
                            bookkeep(vmax, ((c + 1) * CHUNK) / BCHUNK - 1);
                            vmax = 0;
                        }
                    }
#ifdef BSFM_TC_PROFILE
                    if (warp == 2 && lane == 0) PROF_ADD(10, clock64() - e1);   // whole tile after the wait
#endif
                }
                gtile += (uint32_t) (NSTEP * U.ntiles_db);
            } else {
            int nrm_next = (!whole && etid < TILE_DB) ? P.norms[(size_t) U.db_row0 + etid] : 0;
            for (int t = 0; t < U.ntiles_db; t++) {
                const int4 *nb4;
                if (whole) {
                    nb4 = reinterpret_cast<const int4 *>(sNall + t * TILE_DB + part * (TILE_DB / NPART));
                } else {
                    // stage this tile's 256 database norms in shared memory (buffer = accumulator stage)
                    if (etid < TILE_DB) {
                        sN[ts * TILE_DB + etid] = nrm_next;
                        if (t + 1 < U.ntiles_db) nrm_next = P.norms[(size_t) U.db_row0 + (size_t) (t + 1) * TILE_DB + etid];
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
                    nb4 = reinterpret_cast<const int4 *>(sN + ts * TILE_DB + part * (TILE_DB / NPART));
                }
                mbar_wait(bar_t_full + 8 * ts, tph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t) (quad * 32) << 16) + ts * TILE_DB + part * (TILE_DB / NPART);
                // software pipeline over the TMEM loads: chunk c+1 is in flight while chunk c is reduced
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    uint32_t (&v)[32] = (c & 1) ? vb : va;
                    uint32_t (&vn)[32] = (c & 1) ? va : vb;
                    tmem_ld_wait_regs(v);
                    if (c + 1 < NCH) tmem_ld32(taddr + (c + 1) * CHUNK, vn);
                    const int chunk_id = t * (TILE_DB / CHUNK) + part * NCH + c;
                    int cm = INT_MAX;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int4 nb = nb4[c * 8 + q];
                        // t = |p|^2 - 2 q.p as IMAD with a runtime multiplier: keeps the arithmetic on the FMA
                        // pipe and leaves the (slower) ALU pipe to the 3-input minima
                        const int t0 = (int) v[4 * q + 0] * neg2 + nb.x;
                        const int t1 = (int) v[4 * q + 1] * neg2 + nb.y;
                        const int t2 = (int) v[4 * q + 2] * neg2 + nb.z;
                        const int t3 = (int) v[4 * q + 3] * neg2 + nb.w;
                        // two 3-input minima per four elements (VIMNMX3)
                        const int x3 = min(min(t0, t1), t2);
                        cm = min(min(x3, t3), cm);
                    }
                    // (m1, s2) <- two smallest of {m1, s2, cm}
                    const int hi = max(m1, cm);
                    if (cm < m1) bchunk = chunk_id;
                    m1 = min(m1, cm);
                    s2 = min(s2, hi);
                }
                tc_fence_before();
                mbar_arrive(bar_t_empty + 8 * ts);
                ts ^= 1; if (ts == 0) tph ^= 1;
            }
            }
            PROF_T(eu1);
            // merge the column parts of every row (parts 1.. -> shared memory -> part 0)
            if (part > 0) { int *x = xch + (part - 1) * 640; x[row] = m1; x[128 + row] = s2; x[256 + row] = bchunk; x[384 + row] = ubb; x[512 + row] = uo; }
            asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI_THREADS) : "memory");
            bool cand = false;
            int f5 = INT_MAX, f6 = INT_MAX;
            if (part == 0) {
#pragma unroll
                for (int pp = 0; pp < NPART - 1; pp++) {
                    const int *x = xch + pp * 640;
                    const int o1 = x[row], o2 = x[128 + row], ob = x[256 + row], oub = x[384 + row], ouo = x[512 + row];
                    const int hi = max(m1, o1);
                    const bool other_wins = o1 < m1;
                    uo = min(min(uo, ouo), other_wins ? ubb : oub);     // the loser's best chunk joins the "others"
                    ubb = other_wins ? oub : ubb;
                    if (other_wins) bchunk = ob;
                    m1 = min(m1, o1);
                    s2 = min(min(s2, o2), hi);
                }
                if (na < NORM_PAD_HALF) {
                    if (bound) {
                        // two distinct keys with t <= max(ubb, uo) exist, so d1 <= na + max(ubb, uo); d0 >= na + m1
                        const int ubs = max(ubb, uo);
                        const int d1u = (ubs >= NORM_PAD_HALF) ? INT_MAX : na + ubs;
                        cand = (double) (na + m1) < P.ratio_sq * (double) d1u;
                        f5 = s2; f6 = uo;
                    } else {
                        // exact d0, upper bound on d1
                        const int d0 = na + m1;
                        const int d1u = (s2 >= NORM_PAD_HALF) ? INT_MAX : na + s2;
                        cand = (double) d0 < P.ratio_sq * (double) d1u;
                        f5 = d1u;
                    }
                }
            }
            asm volatile("bar.sync 3, %0;" ::"n"(TC_EPI_THREADS) : "memory");   // xch may be overwritten by the next unit
            const unsigned ball = __ballot_sync(0xffffffffu, cand);
            if (ball) {
                int basepos = 0;
                if (lane == 0) basepos = atomicAdd(&P.counters[0], __popc(ball));
                basepos = __shfl_sync(0xffffffffu, basepos, 0);
                if (cand) {
                    const int pos = basepos + __popc(ball & ((1u << lane) - 1u));
                    if (pos < P.cand_cap) {
                        constexpr int CW = BOUND ? BCHUNK : CHUNK;   // width of the chunk the verify kernel recomputes
                        const int col0 = bchunk * CW;
                        const size_t cap = (size_t) P.cand_cap;
                        P.cand[0 * cap + pos] = (u - P.unit_begin) * TILE_Q + row;
                        P.cand[1 * cap + pos] = (int32_t) (U.a_row0 + row);
                        P.cand[2 * cap + pos] = U.db_row0 + col0;
                        P.cand[3 * cap + pos] = col0;
                        P.cand[4 * cap + pos] = min(CW, U.n - col0);
                        P.cand[5 * cap + pos] = f5;
                        P.cand[6 * cap + pos] = bound ? f6 : INT_MIN;   // INT_MIN marks an exact-mode candidate
                    } else {
                        P.counters[2] = 1;
                    }
                }
            }
#ifdef BSFM_TC_PROFILE
            if (warp == 2 && lane == 0) { PROF_ADD(14, clock64() - eu1); PROF_ADD(15, eu1 - eu0); }
#endif
        }
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // neither CTA may leave (or free tensor memory) while the pair's MMAs / remote arrives are in flight
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_kernel(MatchParams P) { match_tc_body<false, false>(P); }
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_bound_kernel(MatchParams P) { match_tc_body<true, false>(P); }
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_quad_kernel(MatchParams P) { match_tc_body<true, false, true>(P); }
// launched with cluster dimension (2, 1, 1)
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_pair_kernel(MatchParams P) { match_tc_body<true, true>(P); }

// ---------------------------------------------------------------------------------------------
// verify: one warp per candidate.  Recomputes the 32 distances of the winning chunk with the
// DEFINITION (sum of squared differences, kd_pr_search.cpp:200-209), finds the exact nearest
// column and the chunk's own second minimum, applies the final ratio test.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sqdist_rows(const uint8_t *keys_sw, int64_t ra, int64_t rb)
{
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(keys_sw + sw_chunk_offset(ra, c));
        const uint4 b = *reinterpret_cast<const uint4 *>(keys_sw + sw_chunk_offset(rb, c));
        uint32_t dx = __vabsdiffu4(a.x, b.x), dy = __vabsdiffu4(a.y, b.y), dz = __vabsdiffu4(a.z, b.z), dw = __vabsdiffu4(a.w, b.w);
        acc = __dp4a(dx, dx, acc); acc = __dp4a(dy, dy, acc); acc = __dp4a(dz, dz, acc); acc = __dp4a(dw, dw, acc);
    }
    return (int) acc;
}

__global__ void __launch_bounds__(256) match_verify_kernel(MatchParams P, int ncand)
{
    const int w = (int) (((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= ncand) return;
    const size_t cap = (size_t) P.cand_cap;
    const int slot = P.cand[0 * cap + w];
    const int qrow = P.cand[1 * cap + w];
    const int db0 = P.cand[2 * cap + w];
    const int col0 = P.cand[3 * cap + w];
    const int nvalid = P.cand[4 * cap + w];
    const int f5 = P.cand[5 * cap + w];
    const int f6 = P.cand[6 * cap + w];

    // up to 64 rows per candidate (bound chunks may be 64 wide): two distances per lane, lane-local top-2 first
    int m1 = INT_MAX, m2 = INT_MAX, mi = lane;
    if (lane < nvalid) m1 = sqdist_rows(P.keys_sw, qrow, (int64_t) db0 + lane);
    if (lane + 32 < nvalid) {
        const int d = sqdist_rows(P.keys_sw, qrow, (int64_t) db0 + lane + 32);
        if (d < m1) { m2 = m1; m1 = d; mi = lane + 32; } else m2 = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int o1 = __shfl_xor_sync(0xffffffffu, m1, o);
        int o2 = __shfl_xor_sync(0xffffffffu, m2, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        int hi = max(m1, o1);
        int lo2 = min(m2, o2);
        if (o1 < m1 || (o1 == m1 && oi < mi)) mi = oi;
        m1 = min(m1, o1);
        m2 = min(hi, lo2);
    }
    if (lane == 0) {
        bool match = false, hard = false;
        if (f6 == INT_MIN) {
            // exact mode: d0 = m1 is the global minimum, f5 bounds d1 from above through the other chunks
            const int d1 = min(f5, m2);
            match = ratio_pass(P, m1, d1);
        } else {
            // bound mode: the chunk with the smallest lower bound was recomputed exactly (m1, m2).  It holds the
            // global minimum iff m1 <= every other chunk's lower bound; d1 lies in [min(m2, L2), min(m2, Uo)].
            const int na = P.norms[qrow];
            const int l2d = (f5 >= NORM_PAD_HALF) ? INT_MAX : na + f5;
            const int uod = (f6 >= NORM_PAD_HALF) ? INT_MAX : na + f6;
            if (m1 > l2d) hard = true;
            else {
                const int d1_low = min(m2, l2d), d1_up = min(m2, uod);
                if (ratio_pass(P, m1, d1_low)) match = true;
                else if (ratio_pass(P, m1, d1_up)) hard = true;   // undecided inside the bound slack
            }
        }
        if (hard) {
            int pos = atomicAdd(&P.counters[3], 1);
            if (pos < P.cand_cap) { P.hard[pos] = slot; P.hard[(size_t) P.cand_cap + pos] = qrow; }
            else P.counters[2] = 1;
        }
        if (match) {
            int pos = atomicAdd(&P.counters[1], 1);
            if (pos < P.match_cap) {
                const int u = P.unit_begin + (slot >> 7);
                const RunImage R = P.run_imgs[find_run_image(P.run_imgs, P.num_run_imgs, u)];
                P.match_slot[pos] = match_sort_key(P, R, qrow);
                P.match_idx2[pos] = col0 + mi;
            } else {
                P.counters[2] = 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// full scan: one block per "hard" row (bounds undecided): exact two nearest keys over the whole database
// image with the definition, final ratio test.  Rare by construction.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) match_fullscan_kernel(MatchParams P, int nhard)
{
    // one 256-thread block per hard row: every warp scans a strided eighth of the image, then the warps' top-2 merge
    __shared__ int s_m1[8], s_m2[8], s_mi[8];
    const int w = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (w >= nhard) return;
    const int slot = P.hard[w];
    const int qrow = P.hard[(size_t) P.cand_cap + w];
    const int u = P.unit_begin + (slot >> 7);
    const RunImage R = P.run_imgs[find_run_image(P.run_imgs, P.num_run_imgs, u)];
    int m1 = INT_MAX, m2 = INT_MAX, mi = INT_MAX;
    for (int c = threadIdx.x; c < R.n; c += 256) {
        const int d = sqdist_rows(P.keys_sw, qrow, (int64_t) R.db_row0 + c);
        if (d < m1) { m2 = m1; m1 = d; mi = c; } else if (d < m2) m2 = d;
    }
    // (value, index) top-2 merge; ties keep the smaller index (any tie of the two best fails the ratio test anyway)
    auto merge = [&](int o1, int o2, int oi) {
        const int hi = max(m1, o1);
        const int lo2 = min(m2, o2);
        if (o1 < m1 || (o1 == m1 && oi < mi)) mi = oi;
        m1 = min(m1, o1);
        m2 = min(hi, lo2);
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int o1 = __shfl_xor_sync(0xffffffffu, m1, o);
        const int o2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        merge(o1, o2, oi);
    }
    if (lane == 0) { s_m1[warp] = m1; s_m2[warp] = m2; s_mi[warp] = mi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; q++) merge(s_m1[q], s_m2[q], s_mi[q]);
        if (ratio_pass(P, m1, m2)) {
            int pos = atomicAdd(&P.counters[1], 1);
            if (pos < P.match_cap) { P.match_slot[pos] = match_sort_key(P, R, qrow); P.match_idx2[pos] = mi; }
            else P.counters[2] = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// finalize: sorted (slot, idx2) -> (idx1, idx2) records + per-pair counts.
// ---------------------------------------------------------------------------------------------
__global__ void match_finalize_kernel(const uint32_t *__restrict__ slots, const int32_t *__restrict__ idx2,
                                      int nmatch, const RunImage *__restrict__ run_imgs, int K, int unit_begin,
                                      const int32_t *__restrict__ tile_img, const int32_t *__restrict__ img_doff,
                                      const int32_t *__restrict__ perm,
                                      int32_t *__restrict__ out_pairs /* [nmatch][2] */,
                                      int32_t *__restrict__ pair_counts)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nmatch) return;
    const uint32_t slot = slots[e];
    const int u = unit_begin + (int) (slot >> 7);
    const int r = (int) (slot & 127u);
    const int k = find_run_image(run_imgs, K, u);
    const RunImage R = run_imgs[k];
    const int atile = R.atile0 + (u - R.unit0);
    const int j = tile_img[atile];
    const int idx1 = atile * TILE_Q + r - img_doff[j];
    out_pairs[2 * (size_t) e + 0] = idx1;
    out_pairs[2 * (size_t) e + 1] = perm[(size_t) R.db_row0 + idx2[e]];   // norm-sorted position -> caller's key index
    atomicAdd(&pair_counts[R.pair0 + (j - R.start_img)], 1);
}

}  // namespace match
}  // namespace bsfm
