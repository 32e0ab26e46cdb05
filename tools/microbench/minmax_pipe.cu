// Throughput of the min/max flavours an argmax epilogue can use (sm_100a), per scheduler: FMNMX (max.f32), FMNMX3 (3-input max.f32),
// HMNMX2 (max.f16x2 / max.bf16x2), alone and mixed 1:1 with FMNMX3 (do they share the ALU pipe?), and mixed with FFMA (FMA pipe).
// 16 independent chains per thread, W warps per scheduler (CTA = 4 W warps, one CTA per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_pipe minmax_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, unsigned* out, long long* cyc, unsigned seed)
{
    unsigned r[16]; float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = seed * (threadIdx.x + 1) + i * 0x3c003c01u; f[i] = __uint_as_float(0x3f800000u + ((seed + i * threadIdx.x) & 0xffff)); }
    const unsigned b = seed ^ 0x3a003b00u, c = seed ^ 0x39003800u; const float fb = __uint_as_float(0x3f810000u ^ (seed & 0xff)), fc = __uint_as_float(0x3f820000u ^ (seed & 0xf0));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (MODE == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc));
            if (MODE == 2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (MODE == 3) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (MODE == 4) { if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); else asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc)); }
            if (MODE == 5) { if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc)); }
            if (MODE == 6) { if (i & 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc)); }
            if (MODE == 7) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
            if (MODE == 8) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[i]) : "r"(b));
            if (MODE == 9) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, unsigned* out, long long* cyc)
{
    const int iters = 2000;
    for (int W : {1, 2, 4}) {
        for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, 128 * W>>>(iters, out, cyc, 12345u + rep); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %d warp(s)/scheduler: %.2f cycles per warp instruction and scheduler  %s\n", name, W, (double)c / (iters * 16.0 * W), cudaGetErrorString(cudaGetLastError()));
    }
}
int main()
{
    unsigned* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    run<0>("max.f32 (FMNMX)", out, cyc); run<1>("max.f32 3-input (FMNMX3)", out, cyc); run<2>("max.f16x2 (HMNMX2)", out, cyc); run<3>("max.bf16x2", out, cyc);
    run<4>("max.f16x2 : FMNMX3 = 1:1", out, cyc); run<5>("max.f16x2 : FFMA = 1:1", out, cyc); run<6>("FMNMX3 : FFMA = 1:1", out, cyc); run<7>("fma.f16x2 (HFMA2)", out, cyc);
    run<8>("prmt", out, cyc); run<9>("max.u16x2 (VIMNMX?)", out, cyc);
    return 0;
}
