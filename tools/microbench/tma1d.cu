// 1-D TMA bulk copy (cp.async.bulk) latency/throughput probe: each CTA streams `bytes`-sized tiles through a ring of
// STAGES smem buffers with mbarriers, no compute.  Reports GB/s aggregate and us per tile per CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../buffer_b200/csrc/bfr_common.cuh"
using namespace bfr;
template<int STAGES>
__global__ void __launch_bounds__(128) k(const float* __restrict__ src, size_t tile_floats, int ntiles, size_t cta_stride_floats, float* out, int bytes)
{
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * bytes);
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); mbar_fence_init(); }
  __syncthreads();
  const float* base = src + (size_t)blockIdx.x * cta_stride_floats;
  if (threadIdx.x == 0) for (int s = 0; s < STAGES && s < ntiles; ++s) { mbar_expect_tx(&full[s], bytes); tma_load_1d(buf + (size_t)s * bytes / 4, base + (size_t)s * tile_floats, bytes, &full[s]); }
  float acc = 0.f;
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % STAGES; const uint32_t ph = (t / STAGES) & 1;
    mbar_wait(&full[s], ph);
    acc += buf[(size_t)s * bytes / 4 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && t + STAGES < ntiles) { mbar_expect_tx(&full[s], bytes); tma_load_1d(buf + (size_t)s * bytes / 4, base + (size_t)(t + STAGES) * tile_floats, bytes, &full[s]); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main(){
  const size_t total = (size_t)1 << 30;  // 1 GiB of floats region
  float* src; cudaMalloc(&src, total); cudaMemset(src, 0, total);
  float* out; cudaMalloc(&out, 4096 * 128 * 4);
  for (int bytes : {2048, 8192, 16384}) for (int stages : {1, 4}) for (int grid : {148, 296, 592}) {
    int ntiles = 2000;
    size_t tile_floats = bytes / 4, stride = ((total / 4) / grid) / 1024 * 1024; if ((size_t)ntiles * tile_floats > stride) ntiles = stride / tile_floats;
    size_t smem = (size_t)stages * bytes + 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (stages == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<grid, 128, smem>>>(src, tile_floats, ntiles, stride, out, bytes); }
      else { cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<4><<<grid, 128, smem>>>(src, tile_floats, ntiles, stride, out, bytes); }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("bytes %5d stages %d grid %3d: %.3f ms  %.1f GB/s  %.2f us/tile/CTA  err=%s\n", bytes, stages, grid, ms, (double)grid * ntiles * bytes / ms * 1e-6, ms * 1e3 / ntiles, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
