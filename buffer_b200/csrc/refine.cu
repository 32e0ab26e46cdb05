// refine.cu — K4: weighted Kabsch (rigid_transform_3d) and the <=20-round post_refinement loop, entirely on device.
//
// Replaces rigid_transform_3d (reference models/BUFFER.py:424-464: diag_embed(weights) [n x n], Am^T W Bm and a
// GPU->CPU->GPU torch.svd per call) and buffer.post_refinement (models/BUFFER.py:382-418: one host sync per round for
// int(inlier_num)).  One CTA of 256 threads per pair / batch element; the weighted sums use a FIXED reduction tree
// (lane l accumulates elements l, l+256, ... sequentially; xor-butterfly 16,8,4,2,1 inside each warp; the 8 warp
// totals added in warp order) that oracle/bfr_oracle.c mirrors, so the whole loop — inlier counts, number of rounds and
// the final transform — is bit-reproducible against the CPU oracle.  The 3x3 SVD is the same closed form as K2.
#include "bfr_common.cuh"
#include "bfr_kernels.h"

namespace bfr {

constexpr int RF_THREADS = 256;

template <int NV>
BFR_DEVINL void block_sum(float (&v)[NV], float (*red)[RF_THREADS / 32])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = __fadd_rn(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    __syncthreads();                            // red[] free to overwrite
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k][warp] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float tot = red[k][0];
#pragma unroll
        for (int w = 1; w < RF_THREADS / 32; ++w) tot = __fadd_rn(tot, red[k][w]);
        v[k] = tot;
    }
}

BFR_DEVINL int block_sum_int(int v, int* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < RF_THREADS / 32; ++w) tot += red[w];
    return tot;
}

// Weighted Kabsch over n points given by functor f(i, a[3], b[3]) -> weight.  Every thread returns the same T
// (row-major 3x4: R | t).  Matches orc_rigid_transform_3d.
template <typename F>
BFR_DEVINL void weighted_kabsch_block(int n, F f, float R[9], float t[3], float (*red)[RF_THREADS / 32])
{
    float s7[7] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for (int i = threadIdx.x; i < n; i += RF_THREADS) {
        float a[3], b[3];
        const float w = f(i, a, b);
        s7[0] = __fadd_rn(s7[0], w);
#pragma unroll
        for (int r = 0; r < 3; ++r) { s7[1 + r] = __fmaf_rn(w, a[r], s7[1 + r]); s7[4 + r] = __fmaf_rn(w, b[r], s7[4 + r]); }
    }
    block_sum<7>(s7, red);
    const float den = __fadd_rn(s7[0], 1e-6f);      // reference: / (sum(weights) + 1e-6), models/BUFFER.py:441-444
    float ca[3], cb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { ca[r] = __fdiv_rn(s7[1 + r], den); cb[r] = __fdiv_rn(s7[4 + r], den); }
    float h9[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for (int i = threadIdx.x; i < n; i += RF_THREADS) {
        float a[3], b[3];
        const float w = f(i, a, b);
#pragma unroll
        for (int r = 0; r < 3; ++r) { a[r] = __fsub_rn(a[r], ca[r]); b[r] = __fsub_rn(b[r], cb[r]); }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float wa = __fmul_rn(w, a[r]);
#pragma unroll
            for (int c = 0; c < 3; ++c) h9[3 * r + c] = __fmaf_rn(wa, b[c], h9[3 * r + c]);
        }
    }
    block_sum<9>(h9, red);
    kabsch_rotation(h9, R);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = __fsub_rn(cb[r], __fmaf_rn(R[3 * r + 2], ca[2], __fmaf_rn(R[3 * r + 1], ca[1], __fmul_rn(R[3 * r + 0], ca[0]))));
}

// rigid_transform_3d(A[bs,n,3], B[bs,n,3], weights[bs,n] or NULL, weight_threshold) -> T[bs,4,4]
__global__ void __launch_bounds__(RF_THREADS) rigid_transform_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ w,
                                                                     int n, float weight_threshold, float* __restrict__ T)
{
    __shared__ float red[9][RF_THREADS / 32];
    const int b = blockIdx.x;
    const float* Ab = A + (size_t)b * n * 3; const float* Bb = B + (size_t)b * n * 3; const float* wb = w ? w + (size_t)b * n : nullptr;
    float R[9], t[3];
    weighted_kabsch_block(n, [&](int i, float a[3], float q[3]) {
        a[0] = __ldg(Ab + 3 * (size_t)i); a[1] = __ldg(Ab + 3 * (size_t)i + 1); a[2] = __ldg(Ab + 3 * (size_t)i + 2);
        q[0] = __ldg(Bb + 3 * (size_t)i); q[1] = __ldg(Bb + 3 * (size_t)i + 1); q[2] = __ldg(Bb + 3 * (size_t)i + 2);
        float wi = wb ? __ldg(wb + i) : 1.0f;
        if (wi < weight_threshold) wi = 0.0f;        // models/BUFFER.py:437
        return wi;
    }, R, t, red);
    if (threadIdx.x == 0) {
        float* o = T + 16 * (size_t)b;
#pragma unroll
        for (int r = 0; r < 3; ++r) { o[4 * r] = R[3 * r]; o[4 * r + 1] = R[3 * r + 1]; o[4 * r + 2] = R[3 * r + 2]; o[4 * r + 3] = t[r]; }
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
    }
}

// post_refinement for P pairs; corr = 8-float records at corr_off[p], corr_cnt[p] of them (ALL mutual correspondences,
// as the reference passes ss_kpts/tt_kpts, models/BUFFER.py:328).
__global__ void __launch_bounds__(RF_THREADS) post_refinement_kernel(const float* __restrict__ T0, const float4* __restrict__ corr, const int32_t* __restrict__ corr_off,
                                                                     const int32_t* __restrict__ corr_cnt, float thr, int max_iter,
                                                                     float* __restrict__ Tout, int32_t* __restrict__ iters_out, int32_t* __restrict__ inliers_out)
{
    __shared__ float red[9][RF_THREADS / 32];
    __shared__ int redi[RF_THREADS / 32];
    const int p = blockIdx.x;
    const int n = corr_cnt[p];
    const float4* c = corr + 2 * (size_t)corr_off[p];
    float R[9], t[3];
    {
        const float* ti = T0 + 16 * (size_t)p;
#pragma unroll
        for (int r = 0; r < 3; ++r) { R[3 * r] = ti[4 * r]; R[3 * r + 1] = ti[4 * r + 1]; R[3 * r + 2] = ti[4 * r + 2]; t[r] = ti[4 * r + 3]; }
    }
    int prev = 0, it = 0;
    for (; it < max_iter; ++it) {
        int cnt = 0;
        for (int i = threadIdx.x; i < n; i += RF_THREADS) {
            const float4 a = __ldg(&c[2 * (size_t)i]), q = __ldg(&c[2 * (size_t)i + 1]);
            cnt += (__fsqrt_rn(resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z)) < thr) ? 1 : 0;
        }
        cnt = block_sum_int(cnt, redi);
        if (cnt == prev) break;                         // models/BUFFER.py:405-407
        prev = cnt;
        float Rn[9], tn[3];
        weighted_kabsch_block(n, [&](int i, float a3[3], float q3[3]) {
            const float4 a = __ldg(&c[2 * (size_t)i]), q = __ldg(&c[2 * (size_t)i + 1]);
            a3[0] = a.x; a3[1] = a.y; a3[2] = a.z; q3[0] = q.x; q3[1] = q.y; q3[2] = q.z;
            const float L2 = __fsqrt_rn(resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z));
            if (!(L2 < thr)) return 0.0f;
            const float u = __fdiv_rn(L2, thr);
            return __fdiv_rn(1.0f, __fmaf_rn(u, u, 1.0f));   // models/BUFFER.py:415
        }, Rn, tn, red);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = tn[k];
    }
    if (threadIdx.x == 0) {
        float* o = Tout + 16 * (size_t)p;
#pragma unroll
        for (int r = 0; r < 3; ++r) { o[4 * r] = R[3 * r]; o[4 * r + 1] = R[3 * r + 1]; o[4 * r + 2] = R[3 * r + 2]; o[4 * r + 3] = t[r]; }
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
        if (iters_out) iters_out[p] = it;
        if (inliers_out) inliers_out[p] = prev;
    }
}

cudaError_t rigid_transform_launch(const float* A, const float* B, const float* w, int bs, int n, float weight_threshold, float* T, cudaStream_t stream)
{
    if (bs > 0) rigid_transform_kernel<<<bs, RF_THREADS, 0, stream>>>(A, B, w, n, weight_threshold, T);
    return cudaGetLastError();
}

cudaError_t post_refinement_launch(const float* T0, const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, float thr, int max_iter,
                                   float* Tout, int32_t* iters_out, int32_t* inliers_out, cudaStream_t stream)
{
    if (P > 0) post_refinement_kernel<<<P, RF_THREADS, 0, stream>>>(T0, reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, thr, max_iter, Tout, iters_out, inliers_out);
    return cudaGetLastError();
}

}  // namespace bfr
