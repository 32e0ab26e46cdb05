// Probe for a later round: what does tcgen05.mma kind::f16 with an F16 accumulator (idesc d-format 0) do on sm_100a?
//   1. TMEM footprint: does a 128 x N F16 tile occupy N or N/2 32-bit columns?  (N/2 would let K1-TC keep twice as many tiles in flight)
//   2. accuracy: max |D_f16 - D_exact| over random unit-norm 32-d bf16 rows (two K = 16 MMAs, the K1-TC shape), against the FP32 accumulator
// One CTA.  A and B are written to shared memory by the threads in the SWIZZLE_NONE K-major core-matrix layout (8 rows x 16 bytes per core
// matrix, LBO = distance between the two K halves of a 16-element slice, SBO = distance between 8-row groups), so no TMA is needed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16_accum f16_accum.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, no swizzle: start address, LBO (bytes between core matrices along K), SBO (bytes between 8-row groups), version 1, layout 0
__device__ __forceinline__ uint64_t desc_noswz(const void* smem, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                   "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
constexpr int N = 128;      // streamed rows (MMA N)
constexpr int KMAX = 64;    // operand depth allocated (mode 3 uses all of it: 4 MMAs; the other modes use 32)
// element (row, k) of a [rows x 32] bf16 operand in the no-swizzle K-major layout: core matrix = 8 rows x 8 elements (16 bytes per row)
__device__ __forceinline__ int opnd_index(int row, int k, int rows) { return ((k >> 3) * (rows >> 3) + (row >> 3)) * 64 + (row & 7) * 8 + (k & 7); }

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* out_f32, float* out_f16, uint32_t* raw_f16, int mode)
{
    __shared__ __align__(1024) __nv_bfloat16 sa[128 * KMAX], sb[N * KMAX];
    const int KD = mode == 3 ? KMAX : 32;
    __shared__ uint32_t tb; __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 128 * KD; i += 128) sa[opnd_index(i / KD, i % KD, 128)] = A[i];
    for (int i = threadIdx.x; i < N * KD; i += 128) sb[opnd_index(i / KD, i % KD, N)] = B[i];
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
                            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tb + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    uint32_t r[32];
    // mode 0: bf16 operands, F32 accumulator | mode 1: the operand bits are f16, F16 accumulator | mode 2: bf16 operands, F16 accumulator
    for (int pass = ((mode == 0 || mode == 3) ? 0 : 1); pass < ((mode == 0 || mode == 3) ? 1 : 2); ++pass) {
        if (threadIdx.x == 0) {
            const uint32_t dfmt = pass == 0 ? 1u : 0u;
            const uint32_t abfmt = (mode == 1 || mode == 3) ? 0u : 1u; // 0 = F16, 1 = BF16
            const uint32_t idesc = (dfmt << 4) | (abfmt << 7) | (abfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int ks = 0; ks < KD / 16; ++ks) {                    // K = 16 per instruction = two core matrices along K
                const uint64_t ad = desc_noswz(sa + ks * 2 * (128 / 8) * 64, (128 / 8) * 128, 128);
                const uint64_t bd = desc_noswz(sb + ks * 2 * (N / 8) * 64, (N / 8) * 128, 128);
                mma(tb + (pass ? 256u : 0u), ad, bd, idesc, ks);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < N; c += 32) {
            ld32(lane_addr + (pass ? 256u : 0u) + c, r);
            for (int i = 0; i < 32; ++i) {
                if (pass == 0) out_f32[threadIdx.x * N + c + i] = __uint_as_float(r[i]);
                else raw_f16[threadIdx.x * N + c + i] = r[i];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    }
    (void)out_f16;
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int M = 128, KD = mode == 3 ? KMAX : 32;
    uint16_t *hA = new uint16_t[M * KD], *hB = new uint16_t[N * KD];
    float *fA = new float[M * KD], *fB = new float[N * KD];
    srand(7);
    auto fill = [&](uint16_t* h, float* f, int rows) {
        for (int i = 0; i < rows; ++i) {
            float v[KMAX], nn = 0; for (int k = 0; k < KD; ++k) { v[k] = (rand() / (float)RAND_MAX) * 2 - 1; nn += v[k] * v[k]; }
            for (int k = 0; k < KD; ++k) {
                const float x = mode == 3 ? 8.0f * v[k] : v[k] / sqrtf(nn);          // mode 3: magnitudes up to 8, products up to 64, heavy cancellation
                if (mode == 1 || mode == 3) { __half hx = __float2half(x); memcpy(&h[i * KD + k], &hx, 2); f[i * KD + k] = __half2float(hx); }
                else { __nv_bfloat16 bx = __float2bfloat16(x); memcpy(&h[i * KD + k], &bx, 2); f[i * KD + k] = __bfloat162float(bx); }
            }
        }
    };
    fill(hA, fA, M); fill(hB, fB, N);
    __nv_bfloat16 *dA, *dB; float *d32, *d16; uint32_t* draw;
    cudaMalloc(&dA, M * KD * 2); cudaMalloc(&dB, N * KD * 2); cudaMalloc(&d32, M * N * 4); cudaMalloc(&d16, M * N * 4); cudaMalloc(&draw, M * N * 4);
    cudaMemcpy(dA, hA, M * KD * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * KD * 2, cudaMemcpyHostToDevice);
    cudaMemset(draw, 0, M * N * 4); cudaMemset(d32, 0, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 0);
    probe<<<1, 128>>>(dA, dB, d32, d16, draw, mode);
    cudaError_t e = cudaDeviceSynchronize();
    const char* names[] = { "bf16 operands, F32 accumulator", "f16 operands, F16 accumulator", "bf16 operands, F16 accumulator", "f16 operands up to 8, K = 64, F32 accumulator" };
    printf("mode %d (%s): kernel: %s\n", mode, names[mode], cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    float* o32 = new float[M * N]; uint32_t* raw = new uint32_t[M * N];
    cudaMemcpy(o32, d32, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(raw, draw, M * N * 4, cudaMemcpyDeviceToHost);
    double e32 = 0, e16_lo = 0, e16_packed = 0, rel_abs = 0, rel_seq = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0, sabs = 0; float seq = 0.0f;
            for (int k = 0; k < KD; ++k) { const double pr = (double)fA[i * KD + k] * fB[j * KD + k]; s += pr; sabs += fabs(pr); seq = fmaf(fA[i * KD + k], fB[j * KD + k], seq); }
            e32 = fmax(e32, fabs(o32[i * N + j] - s));
            rel_abs = fmax(rel_abs, fabs(o32[i * N + j] - s) / sabs);
            rel_seq = fmax(rel_seq, fabs((double)seq - s) / sabs);
            __half lo; uint16_t b = (uint16_t)(raw[i * N + j] & 0xFFFF); memcpy(&lo, &b, 2);
            e16_lo = fmax(e16_lo, fabs(__half2float(lo) - s));
            uint32_t w = raw[i * N + j / 2]; uint16_t hb = (uint16_t)((j & 1) ? (w >> 16) : (w & 0xFFFF)); __half hv; memcpy(&hv, &hb, 2);
            e16_packed = fmax(e16_packed, fabs(__half2float(hv) - s));
        }
    if (mode == 0) printf("F32 accumulator: max |D - exact| = %.3e (operand layout check: must be ~1e-7)\n", e32);
    else if (mode == 3) printf("F32 accumulator over %d products: max |D - exact| = %.3e; relative to sum|a_k b_k|: %.3e = 2^%.1f (a sequential FP32 fma chain on the same data: 2^%.1f)\n",
                               KD, e32, rel_abs, log2(rel_abs), log2(rel_seq));
    else {
        printf("F16 accumulator, one value per column (low half):   max |D - exact| = %.3e\n", e16_lo);
        printf("F16 accumulator, two values packed per column:      max |D - exact| = %.3e\n", e16_packed);
        printf("raw row 0, columns 0..3: %08x %08x %08x %08x   columns %d..%d: %08x %08x\n", raw[0], raw[1], raw[2], raw[3], N / 2, N / 2 + 1, raw[N / 2], raw[N / 2 + 1]);
        printf("(2^-11 = 4.9e-4 is the F16 rounding step near 1; 2^-8 |a||b| = 3.9e-3 is the bf16 operand bound)\n");
    }
    return 0;
}
