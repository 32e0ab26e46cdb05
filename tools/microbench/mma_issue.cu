// Issue cost of tcgen05 operations from one thread (sm_100a): per iteration [fence::after_thread_sync, 2 x tcgen05.mma M128 N64 K16 into one
// accumulator tile, 1 or 2 tcgen05.commit], nothing else running on the SM; then the same from 1, 2, 3 warps concurrently (each its own
// accumulator columns and barriers).  Operands: zero-filled shared memory, unswizzled K-major 32-byte rows (the layout of the RANSAC filter).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void issue(uint32_t d, uint32_t a_lo, uint32_t b1_lo, uint32_t b2_lo, uint32_t bar1, uint32_t bar2, int commits, int N)
{
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = 16u | (1u << 14);
    asm volatile("{\n\t.reg .b64 da, db1, db2;\n\t.reg .pred pt, pf;\n\t"
                 "mov.b64 da, {%1, %4};\n\tmov.b64 db1, {%2, %4};\n\tmov.b64 db2, {%3, %4};\n\t"
                 "setp.eq.u32 pt, %0, %0;\n\tsetp.ne.u32 pf, %0, %0;\n\t"
                 "tcgen05.fence::after_thread_sync;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db1, %5, pf;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db2, %5, pt;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(b1_lo), "r"(b2_lo), "r"(hi), "r"(idesc) : "memory");
    if (commits >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
    if (commits >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
}
__global__ void __launch_bounds__(128) k(int iters, int nwarps, int commits, int N, long long* cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tb; __shared__ uint64_t bar[8];
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int b = 0; b < 8; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
                            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nwarps && lane == 0) {
        const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | (8u << 16), b_lo = ((smem_u32(smem + 4096) >> 4) & 0x3FFFu) | (8u << 16);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            issue(tb + (uint32_t)(warp * 128 + (it & 1) * 64), a_lo, b_lo, b_lo + 128, smem_u32(&bar[2 * warp]), smem_u32(&bar[2 * warp + 1]), commits, N);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cyc[warp] = t1 - t0;
    }
    __syncthreads();
    // drain: wait until the tensor pipe is idle before freeing TMEM
    if (threadIdx.x == 0) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[7])) : "memory"); }
    __nanosleep(20000);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
int main()
{
    long long* cyc; cudaMalloc(&cyc, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024);
    const int iters = 4000;
    for (int N : {64, 128, 256}) for (int commits : {0, 1, 2}) for (int nw : {1, 2, 3}) {
        if (nw * 128 > 512) continue;
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 16384 + 1024>>>(iters, nw, commits, N, cyc); cudaDeviceSynchronize(); }
        long long c[8]; cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
        printf("N %3d, %d commit(s) per issue, %d issuing warp(s): %.1f cycles per issue (warp 0)%s  %s\n", N, commits, nw, (double)c[0] / iters,
               nw > 1 ? "" : "", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
