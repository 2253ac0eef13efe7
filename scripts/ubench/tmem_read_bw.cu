// Micro-benchmark (dev tool, not part of the library): tcgen05.ld throughput per SM for the shapes the MATCH
// epilogue uses.  One CTA per SM, W warps issue x32 loads (32 lanes x 32 columns x 4 B = 4 KB each) with D loads in
// flight per warp.  Prints cycles per load and bytes/cycle/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int DEPTH>
__global__ void __launch_bounds__(544) bw_kernel(int iters, int nwarps, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t taddr = base + ((uint32_t) ((warp & 3) * 32) << 16) + (uint32_t) ((warp >> 2) * 64);
        uint32_t va[32], vb[32];
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            ld32(taddr, va);
            if (DEPTH == 2) ld32(taddr + 32, vb);
            ldwait();
            acc += va[0] ^ va[13] ^ va[31];
            if (DEPTH == 2) acc += vb[0] ^ vb[13] ^ vb[31];
        }
        t1 = clock64();
    }
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
    }
}

int main()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long *d_cyc; uint32_t *d_sink;
    cudaMalloc(&d_cyc, sms * sizeof(long long)); cudaMalloc(&d_sink, 4);
    const int iters = 20000;
    for (int depth = 1; depth <= 2; depth++)
        for (int nw : {4, 8, 16}) {
            for (int rep = 0; rep < 2; rep++) {
                if (depth == 1) bw_kernel<1><<<sms, 544>>>(iters, nw, d_cyc, d_sink);
                else bw_kernel<2><<<sms, 544>>>(iters, nw, d_cyc, d_sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long c; cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
            const double loads = (double) iters * depth * nw;           // x32 loads issued by the CTA (thread 0's warp timing)
            printf("depth %d warps %2d: %.1f cycles per iteration of warp 0, %.1f cycles per x32 load per SM, %.1f B/cycle/SM\n",
                   depth, nw, (double) c / iters, (double) c / loads, loads * 4096.0 / (double) c);
        }
    return 0;
}
