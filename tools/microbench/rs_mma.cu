// Probe for the tensor-core RANSAC scoring filter (csrc/ransac_tc.cuh): residual components x_i(h,c) = sum_j R_ij s_j + t_i - q_i of
// 128 correspondences x 85 hypotheses as ONE accumulator tile, computed by two K = 16 tcgen05.mma kind::f16 that share the A operand:
//   A row (correspondence c)   = [s_hi(3) s_lo(3) 1 q_hi(3) q_lo(3) 0 0 0]          (f16, hi + lo = 2-level split of the FP32 value)
//   B1 row (hypothesis h, i)   = [R_i,hi(3) R_i,hi(3) t_i,hi -e_i(3) -e_i(3) 0 0 0]
//   B2 row                     = [R_i,lo(3) 0 0 0     t_i,lo 0 ...]
// Checks (1) which shared-memory descriptor describes a K-major tile with 32-byte rows (SWIZZLE_32B with / without the XOR, or the
// unswizzled core-matrix layout with either meaning of LBO / SBO), (2) the accuracy of the split against the FP32 chain of the oracle and
// against float64, (3) the cost of one accumulator-tile visit of this shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rs_mma rs_mma.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo16, uint32_t sbo16, uint32_t layout)
{
    return (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint64_t)lbo16 << 16) | ((uint64_t)sbo16 << 32) | (1ull << 46) | ((uint64_t)layout << 61);
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

// byte offset of the 16-byte chunk kc (0/1) of row r inside a K-major tile with 32-byte rows
__host__ __device__ inline uint32_t chunk_off(int variant, int r, int kc)
{
    switch (variant) {
    case 0: return (uint32_t)(r * 32 + ((kc ^ ((r >> 2) & 1)) * 16));      // SWIZZLE_32B, address bit 4 ^= bit 7
    case 1: return (uint32_t)(r * 32 + kc * 16);                            // SWIZZLE_32B descriptor, linear rows (expected wrong)
    default: return (uint32_t)((r >> 3) * 256 + kc * 128 + (r & 7) * 16);   // no swizzle: 8 x 16 B core matrices, K chunks 128 B apart, row groups 256 B apart
    }
}

// rows are given linearly ([rows][16] halves); the kernel places them according to `variant`, runs D = A B1^T (+ A B2^T), writes D [128][256]
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B1, const __half* B2, int variant, int two, float* D, int visits, long long* cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tb; __shared__ uint64_t bar;
    unsigned char* sA = smem, *sB1 = smem + 4096, *sB2 = smem + 4096 + 8192;
    for (int i = threadIdx.x; i < 128 * 2; i += blockDim.x) *reinterpret_cast<uint4*>(sA + chunk_off(variant, i >> 1, i & 1)) = reinterpret_cast<const uint4*>(A)[i];
    for (int i = threadIdx.x; i < 256 * 2; i += blockDim.x) {
        *reinterpret_cast<uint4*>(sB1 + chunk_off(variant, i >> 1, i & 1)) = reinterpret_cast<const uint4*>(B1)[i];
        *reinterpret_cast<uint4*>(sB2 + chunk_off(variant, i >> 1, i & 1)) = reinterpret_cast<const uint4*>(B2)[i];
    }
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tb)) : "memory");
                            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t a, b1, b2;
        if (variant <= 1)      { a = make_desc(sA, 1, 16, 6); b1 = make_desc(sB1, 1, 16, 6); b2 = make_desc(sB2, 1, 16, 6); }     // SWIZZLE_32B: 8-row groups 256 B apart
        else if (variant == 2) { a = make_desc(sA, 8, 16, 0); b1 = make_desc(sB1, 8, 16, 0); b2 = make_desc(sB2, 8, 16, 0); }     // LBO = K stride (128 B), SBO = row-group stride (256 B)
        else                   { a = make_desc(sA, 16, 8, 0); b1 = make_desc(sB1, 16, 8, 0); b2 = make_desc(sB2, 16, 8, 0); }     // the other reading of LBO / SBO
        const long long t0 = clock64();
        for (int v = 0; v < visits; ++v) {
            const uint32_t d = tmem + (uint32_t)((v & 1) * 256);
            mma(d, a, b1, idesc, 0u);
            if (two) mma(d, a, b2, idesc, 1u);
        }
        commit(&bar); mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && cyc) cyc[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (blockIdx.x == 0 && D) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                           "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                           "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr) : "memory");
            for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static double urand() { return (double)rand() / RAND_MAX; }
static void split2(float v, __half& hi, __half& lo) { hi = __float2half_rn(v); lo = __float2half_rn(v - __half2float(hi)); }

int main()
{
    srand(7);
    const int NC = 128, NH = 85;
    std::vector<float> s(NC * 3), q(NC * 3), R(NH * 9), t(NH * 3);
    for (int c = 0; c < NC; ++c) for (int k = 0; k < 3; ++k) { s[3 * c + k] = (float)(3.0 * urand() - 1.5); q[3 * c + k] = (float)(3.0 * urand() - 1.5); }
    for (int h = 0; h < NH; ++h) {                                      // random rotation (axis-angle) + translation
        double ax = urand() - .5, ay = urand() - .5, az = urand() - .5, n = sqrt(ax * ax + ay * ay + az * az); ax /= n; ay /= n; az /= n;
        const double th = 6.28 * urand(), cs = cos(th), sn = sin(th), C = 1 - cs;
        const double Rm[9] = { cs + ax * ax * C, ax * ay * C - az * sn, ax * az * C + ay * sn, ay * ax * C + az * sn, cs + ay * ay * C, ay * az * C - ax * sn,
                               az * ax * C - ay * sn, az * ay * C + ax * sn, cs + az * az * C };
        for (int k = 0; k < 9; ++k) R[9 * h + k] = (float)Rm[k];
        for (int k = 0; k < 3; ++k) t[3 * h + k] = (float)(4.0 * urand() - 2.0);
    }
    std::vector<__half> A(NC * 16), B1(256 * 16), B2(256 * 16);
    for (auto& x : A) x = __float2half(0.f); for (auto& x : B1) x = __float2half(0.f); for (auto& x : B2) x = __float2half(0.f);
    for (int c = 0; c < NC; ++c) {
        __half hi, lo;
        for (int k = 0; k < 3; ++k) { split2(s[3 * c + k], hi, lo); A[16 * c + k] = hi; A[16 * c + 3 + k] = lo; split2(q[3 * c + k], hi, lo); A[16 * c + 7 + k] = hi; A[16 * c + 10 + k] = lo; }
        A[16 * c + 6] = __float2half(1.0f);
    }
    for (int h = 0; h < NH; ++h) for (int i = 0; i < 3; ++i) {
        const int row = 3 * h + i; __half hi, lo;
        for (int j = 0; j < 3; ++j) { split2(R[9 * h + 3 * i + j], hi, lo); B1[16 * row + j] = hi; B1[16 * row + 3 + j] = hi; B2[16 * row + j] = lo; }
        split2(t[3 * h + i], hi, lo); B1[16 * row + 6] = hi; B2[16 * row + 6] = lo;
        B1[16 * row + 7 + i] = __float2half(-1.0f); B1[16 * row + 10 + i] = __float2half(-1.0f);
    }
    __half *dA, *dB1, *dB2; float* dD; long long* dcyc;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB1, B1.size() * 2); cudaMalloc(&dB2, B2.size() * 2); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dcyc, 8);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB1, B1.data(), B1.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB2, B2.data(), B2.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 4096 + 8192 + 8192 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> D(128 * 256);
    for (int variant = 0; variant < 4; ++variant) {
        cudaMemset(dD, 0, 128 * 256 * 4);
        probe<<<1, 128, smem>>>(dA, dB1, dB2, variant, 1, dD, 1, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double e_true = 0, e_chain = 0, chain_true = 0, e_rel = 0;
        for (int c = 0; c < NC; ++c) for (int h = 0; h < NH; ++h) for (int i = 0; i < 3; ++i) {
            const float* Rr = &R[9 * h + 3 * i];
            const double xt = (double)Rr[0] * s[3 * c] + (double)Rr[1] * s[3 * c + 1] + (double)Rr[2] * s[3 * c + 2] + (double)t[3 * h + i] - (double)q[3 * c + i];
            const float xc = fmaf(Rr[0], s[3 * c], fmaf(Rr[1], s[3 * c + 1], fmaf(Rr[2], s[3 * c + 2], t[3 * h + i]))) - q[3 * c + i];
            const double Bp = fabs(Rr[0] * s[3 * c]) + fabs(Rr[1] * s[3 * c + 1]) + fabs(Rr[2] * s[3 * c + 2]) + fabs(t[3 * h + i]) + fabs(q[3 * c + i]);
            const double xd = D[c * 256 + 3 * h + i];
            e_true = fmax(e_true, fabs(xd - xt)); e_chain = fmax(e_chain, fabs(xd - (double)xc)); chain_true = fmax(chain_true, fabs((double)xc - xt));
            e_rel = fmax(e_rel, fabs(xd - xt) / Bp);
        }
        printf("variant %d (%s): %s  max|x_tc - x_f64| = %.3e (= 2^%.1f of sum|terms|)  max|x_tc - x_fp32chain| = %.3e  max|x_fp32chain - x_f64| = %.3e\n", variant,
               variant == 0 ? "SWIZZLE_32B, bit4^=bit7" : variant == 1 ? "SWIZZLE_32B desc, linear rows" : variant == 2 ? "no swizzle, LBO=128B SBO=256B" : "no swizzle, LBO=256B SBO=128B",
               cudaGetErrorString(e), e_true, log2(e_rel), e_chain, chain_true);
    }
    // timing: visits alternate between two accumulator tiles; one or two MMAs (K = 16 each) per visit, all 148 SMs busy
    for (int variant : {0, 2}) for (int two = 0; two < 2; ++two) {
        const int visits = 4000;
        for (int rep = 0; rep < 2; ++rep) { probe<<<148, 128, smem>>>(dA, dB1, dB2, variant, two, nullptr, visits, dcyc); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("variant %d, %d MMA(s) per visit (M128 N256 K16): %.1f cycles per accumulator-tile visit  %s\n", variant, two + 1, (double)c / visits, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
