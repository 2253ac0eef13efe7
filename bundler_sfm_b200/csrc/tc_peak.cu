// tc_peak.cu -- measured int8 tensor-pipe ceiling of this GPU: a plain tcgen05.mma kind::i8 loop ("measure
// peak with a plain tcgen05 i8 GEMM on the same box").  One CTA per SM; one thread issues back-to-back MMAs
// (M = 128, N = 128 or 256, K = 4 x 32 per tile) on operand tiles resident in shared memory -- a ring of four (A, B) tile
// pairs of pseudo-random bytes, K-major SWIZZLE_128B like the real kernels, so consecutive tiles read different operands --
// alternating between two TMEM accumulator stages; no loads, no epilogue.  What it reports is the rate the MATCH search kernel (N = 256) and the BA trailing update (N = 128) are
// measured against (bench.py roofline.peak); it is an upper bound for any kernel built from the same instruction shape.
#include "common.h"
#include "tc_ptx.cuh"
#include <atomic>

namespace bsfm {
using namespace bsfm::ptx;

__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, int n_tile)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw_addr);
    constexpr int STAGES = 4, STAGE_BYTES = 16384 + 32768;      // ring of (A 128 x 128 B, B 256 x 128 B) tiles, as a pipelined GEMM keeps
    const uint32_t bar = base + STAGES * STAGE_BYTES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + STAGES * STAGE_BYTES + 16);
    for (int q = threadIdx.x; q < STAGES * STAGE_BYTES / 4; q += blockDim.x) {      // pseudo-random operand bytes
        uint32_t v = (uint32_t) q * 2654435761u + blockIdx.x * 40503u;
        v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
        reinterpret_cast<uint32_t *>(smem)[q] = v;
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes of the tiles -> async proxy (MMA) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        uint64_t adesc[STAGES], bdesc[STAGES];
#pragma unroll
        for (int st = 0; st < STAGES; st++) { adesc[st] = make_sw128_desc(base + st * STAGE_BYTES); bdesc[st] = make_sw128_desc(base + st * STAGE_BYTES + 16384); }
        const uint32_t idesc = (2u << 4) | ((uint32_t) (n_tile >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);      // u8 x u8 -> s32
        uint32_t phase = 0;
        for (int it = 0; it < iters; it++) {
            const uint32_t tmem_d = tmem_base + (uint32_t) (it & 1) * 256;
            const uint64_t ad = adesc[it & (STAGES - 1)], bd = bdesc[it & (STAGES - 1)];      // every tile reads different operands
#pragma unroll
            for (int kk = 0; kk < 4; kk++) tc_mma_i8(tmem_d, ad + (uint64_t) (kk * 2), bd + (uint64_t) (kk * 2), idesc, kk > 0);
            if ((it & 15) == 15 || it == iters - 1) {      // bound the number of MMAs in flight
                tc_commit(bar);
                mbar_wait(bar, phase);
                phase ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace bsfm

// Runs the loop and returns the measured rate in TOP/s (2 ops per multiply-add), < 0 on error.  n_tile = 128 or 256.
extern "C" double bsfm_measure_int8_peak(int n_tile, int iters)
{
    using namespace bsfm;
    clear_error();
    if (require_device() != BSFM_OK) return -1.0;
    if ((n_tile != 128 && n_tile != 256) || iters < 16) { set_error("bsfm_measure_int8_peak: n_tile must be 128 or 256, iters >= 16"); return -1.0; }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { set_error("device query failed"); return -1.0; }
    const size_t smem = 4 * (16384 + 32768) + 64 + 1024;
    if (cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) { set_error("cudaFuncSetAttribute failed"); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    i8_peak_kernel<<<sms, 128, smem>>>(iters / 8 + 16, n_tile);      // warm-up
    cudaEventRecord(e0);
    i8_peak_kernel<<<sms, 128, smem>>>(iters, n_tile);
    cudaEventRecord(e1);
    count_launch(2);
    float ms = 0.f;
    const cudaError_t e = cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess || ms <= 0.f) { set_error("bsfm_measure_int8_peak: %s", cudaGetErrorString(e)); return -1.0; }
    const double ops = (double) sms * iters * 4.0 * (128.0 * n_tile * 32.0) * 2.0;
    return ops / (ms * 1e-3) / 1e12;
}

// ---- fp64 issue rates (dev measurement behind DESIGN.md's notes on the Cholesky pivot chain) ----
// mode 0: vector DFMA, 8 independent chains per thread;  mode 1: DMMA m8n8k4, 4 independent accumulator pairs per thread.
// One CTA per SM, `warps` warps; returns SM cycles per warp instruction per SM sub-partition (warps spread over the 4 of them).
namespace bsfm {
__global__ void fp64_rate_kernel(int iters, int mode, double *sink, unsigned long long *cycles)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0 + 1e-12 * threadIdx.x, c = 1e-13;
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(a[i]), "+d"(a[i + 1]) : "d"(b), "d"(c));
        }
    }
    const unsigned long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
}  // namespace bsfm

extern "C" double bsfm_measure_fp64_issue_cycles(int mode, int warps, int iters)
{
    using namespace bsfm;
    clear_error();
    if (require_device() != BSFM_OK) return -1.0;
    if (warps < 1 || warps > 32 || iters < 16 || (mode != 0 && mode != 1)) { set_error("bsfm_measure_fp64_issue_cycles: bad arguments"); return -1.0; }
    double *sink = nullptr; unsigned long long *cyc = nullptr, h = 0;
    if (cudaMalloc(&sink, 8) != cudaSuccess || cudaMalloc(&cyc, 8) != cudaSuccess) { set_error("cudaMalloc failed"); return -1.0; }
    fp64_rate_kernel<<<1, 32 * warps>>>(iters, mode, sink, cyc);
    fp64_rate_kernel<<<1, 32 * warps>>>(iters, mode, sink, cyc);
    count_launch(2);
    const cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(sink); cudaFree(cyc);
    if (e != cudaSuccess) { set_error("bsfm_measure_fp64_issue_cycles: %s", cudaGetErrorString(e)); return -1.0; }
    const double instr_per_warp = (double) iters * (mode == 0 ? 8.0 : 4.0);
    const double warps_per_smsp = (warps + 3) / 4;
    return (double) h / (instr_per_warp * warps_per_smsp);
}
