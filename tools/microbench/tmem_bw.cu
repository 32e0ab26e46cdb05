// TMEM -> register (tcgen05.ld) bandwidth probe on sm_100a: W warps per CTA (1 CTA/SM) each issue tcgen05.ld.32x32b.xN in a loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]);
template <> __device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr));
}
template <> __device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(taddr));
}
template <int X>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cyc) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
                   asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32]; uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    #pragma unroll
    for (int c = 0; c < 512; c += X) { ld<X>(base + c, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[0] ^ r[X - 1]; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int X> void run(int warps, float* out, long long* cyc) {
  const int iters = 200;
  k<X><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<X><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double bytes = (double)warps * iters * (512.0 / X) * X * 32 * 4;   // per SM
  printf("x%-3d warps %2d: %lld cycles, %.1f B/clk/SM (%.1f per warp-quarter)  err=%s\n", X, warps, c, bytes / c, bytes / c / 4, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) { run<32>(w, out, cyc); run<8>(w, out, cyc); }
  return 0;
}
