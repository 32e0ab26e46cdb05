// Warp-to-warp hand-off latency inside one CTA on sm_100a: ping-pong between lane 0 of warp 0 and lane 0 of warp 1, N round trips.
//   mode 0: mbarrier arrive -> mbarrier.try_wait spin                (what K1-TC uses between the epilogue warps and the MMA thread)
//   mode 1: mbarrier arrive -> mbarrier.test_wait spin
//   mode 2: st.release.cta flag -> ld.acquire.cta spin (shared memory)
//   mode 3: st.volatile flag -> ld.volatile spin (no ordering)
//   mode 4: as 0 with tcgen05.fence::before_thread_sync before the arrive and ::after_thread_sync after the wait
//   mode 5: as 0, but 4 arriving lanes (4 warps) per hop in one direction, as in the kernel (count = 4)
// Prints cycles per ONE-WAY hop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handshake handshake.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_try(uint64_t* b, uint32_t par)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void mbar_test(uint64_t* b, uint32_t par)
{
    asm volatile("{\n\t.reg .pred p;\n\tS_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra S_%=;\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void st_rel(uint32_t* p, uint32_t v) { asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acq(uint32_t* p) { uint32_t v; asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_vol(uint32_t* p) { uint32_t v; asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }

__global__ void __launch_bounds__(192) k(int mode, int iters, long long* cyc)
{
    __shared__ uint64_t bar[2];
    __shared__ uint32_t flag[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], mode == 5 ? 4 : 1); mbar_init(&bar[1], 1); flag[0] = flag[1] = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0 = 0;
    if (warp == 0 && lane == 0) t0 = clock64();
    if (mode == 5) {
        // warps 0..3 (lane 0 each) arrive on bar[0]; warp 4 waits for it and answers on bar[1]; warps 0..3 all wait for bar[1]
        if (lane == 0 && warp < 4) for (int i = 0; i < iters; ++i) { mbar_arrive(&bar[0]); mbar_try(&bar[1], i & 1); }
        if (lane == 0 && warp == 4) for (int i = 0; i < iters; ++i) { mbar_try(&bar[0], i & 1); mbar_arrive(&bar[1]); }
    } else if (lane == 0 && warp < 2) {
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) {
                if (mode == 4) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (mode == 0 || mode == 4) { mbar_arrive(&bar[0]); mbar_try(&bar[1], i & 1); }
                else if (mode == 1) { mbar_arrive(&bar[0]); mbar_test(&bar[1], i & 1); }
                else if (mode == 2) { st_rel(&flag[0], i + 1); while (ld_acq(&flag[1]) != (uint32_t)(i + 1)) {} }
                else { st_vol(&flag[0], i + 1); while (ld_vol(&flag[1]) != (uint32_t)(i + 1)) {} }
                if (mode == 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            } else {
                if (mode == 0 || mode == 4) { mbar_try(&bar[0], i & 1); if (mode == 4) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); } mbar_arrive(&bar[1]); }
                else if (mode == 1) { mbar_test(&bar[0], i & 1); mbar_arrive(&bar[1]); }
                else if (mode == 2) { while (ld_acq(&flag[0]) != (uint32_t)(i + 1)) {} st_rel(&flag[1], i + 1); }
                else { while (ld_vol(&flag[0]) != (uint32_t)(i + 1)) {} st_vol(&flag[1], i + 1); }
            }
        }
    }
    if (warp == 0 && lane == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}
int main()
{
    long long* cyc; cudaMalloc(&cyc, 8);
    const int iters = 20000;
    const char* names[] = { "mbarrier arrive -> try_wait", "mbarrier arrive -> test_wait", "st.release -> ld.acquire (smem flag)", "st.volatile -> ld.volatile (smem flag)",
                            "mbarrier + tcgen05 fences", "mbarrier, 4 arriving warps -> 1 waiter -> 4 waiters" };
    for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 192>>>(mode, iters, cyc); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d %-52s: %6.1f cycles per one-way hop   %s\n", mode, names[mode], (double)c / iters / 2.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
