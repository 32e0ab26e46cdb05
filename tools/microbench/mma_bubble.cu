// tcgen05.mma issue/retire cost on sm_100a as a function of N, the number of K steps per accumulator tile and how many distinct
// accumulator tiles the issuing thread rotates through.  One CTA per SM, one issuing thread, operands = zero-filled shared memory
// (SWIZZLE_64B K-major bf16, 64-byte rows, the K1-TC layout), kind::f16, M = 128, cta_group::1.
//   cycles per group = [ (first MMA, scale_d = 0) + (ksteps - 1) accumulating MMAs ] to one accumulator tile, then the next tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bubble mma_bubble.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw64(const void* smem)
{
    return (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// groups: accumulator-tile visits; ksteps: MMAs per visit; ndt: distinct tiles rotated (tile t at TMEM column (t % ndt) * N); commit_every: commit after
// every c-th group (0 = only at the end); arows: 0 = every group uses the same A rows, 1 = alternate between two A row blocks
__global__ void __launch_bounds__(128) k(int N, int groups, int ksteps, int ndt, int commit_every, int arows, int always_acc, long long* cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tb; __shared__ uint64_t bar;
    uint16_t* A = reinterpret_cast<uint16_t*>(smem);            // 256 rows x 32 bf16 (16 KB)
    uint16_t* B = A + 256 * 32;                                  // 256 rows x 32 bf16 (16 KB)
    for (int i = threadIdx.x; i < 2 * 256 * 32 / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
                            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a0 = desc_sw64(A), a1 = desc_sw64(A + 128 * 32), b0 = desc_sw64(B);
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tb + (uint32_t)((g % ndt) * N);
            const uint64_t ad = (arows && (g & 1)) ? a1 : a0;
            for (int ks = 0; ks < ksteps; ++ks) mma(d, ad + (uint64_t)(2 * (ks & 1)), b0 + (uint64_t)(2 * (ks & 1)), idesc, (ks > 0 || always_acc) ? 1u : 0u);
            if (commit_every > 0 && (g + 1) % commit_every == 0 && g + 1 < groups) { commit(&bar); mbar_wait(&bar, ph); ph ^= 1u; }
            if (commit_every < 0 && g + 1 < groups) { commit(&bar); ph ^= 1u; }      // commit without waiting (as the kernel does)
        }
        commit(&bar); mbar_wait(&bar, ph);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
static void run(int N, int ksteps, int ndt, int commit_every, int arows, long long* cyc, int always_acc = 0)
{
    const int groups = 2000;
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 32768 + 1024>>>(N, groups, ksteps, ndt, commit_every, arows, always_acc, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("acc-always %d N %3d ksteps %2d tiles %d commit+wait every %d arows %d: %7.1f cycles/group  (%6.1f per MMA, %6.1f per 128x128 outputs)  %s\n", always_acc, N, ksteps, ndt, commit_every, arows,
           (double)c / groups, (double)c / groups / ksteps, (double)c / groups * 128.0 / N, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    long long* cyc; cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024);
    for (int N : {64, 128, 256})
        for (int ks : {1, 2, 4, 16}) { run(N, ks, 1, 0, 0, cyc); if (N * 2 <= 512) run(N, ks, 2, 0, 0, cyc); }
    for (int N : {64, 128, 256})
        for (int ks : {1, 2, 4}) { run(N, ks, 2, 0, 1, cyc, 1); }
    run(128, 2, 4, 0, 1, cyc, 1); run(128, 2, 4, -1, 1, cyc, 1); run(128, 2, 4, -1, 1, cyc, 0); run(128, 2, 4, 1, 1, cyc, 1);
    run(128, 2, 4, 0, 0, cyc); run(128, 2, 2, 0, 1, cyc); run(128, 2, 4, 0, 1, cyc); run(64, 2, 8, 0, 0, cyc);
    run(128, 2, 2, 1, 0, cyc); run(128, 2, 2, 2, 0, cyc); run(256, 2, 2, 1, 0, cyc);
    return 0;
}
