// extras.cu — the "next" rows of SURVEY.md §8(f) that reuse the hot path's building blocks:
//   f1  get_matching_indices (reference models/BUFFER.py:361-380; twins ThreeDMatch/dataset.py:14-22, trainer.py:38-54):
//       SE(3)-transform the source points, brute-force 3-D nearest neighbour, keep pairs closer than the voxel size.
//       Replaces the last knn_cuda.KNN call of models/BUFFER.py (:374) and an O(N*M*3) broadcast in the trainer.
//   f3  batched 3x3 SVD with the torch_batch_svd contract (utils/common.py:10, call site :715 in cal_Z_axis).
// Arithmetic mirrors oracle/bfr_oracle.c (orc_get_matching_indices, orc_svd3) operation for operation.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include <cmath>

namespace bfr {

constexpr int KNN3_THREADS = 256;
constexpr int KNN3_TILE = 2048;

// one thread per source point; targets stream through shared memory in tiles (all lanes read the same target: broadcast)
__global__ void __launch_bounds__(KNN3_THREADS) knn3_kernel(const float* __restrict__ source, int N, const float* __restrict__ target, int M,
                                                            const float* __restrict__ T, int32_t* __restrict__ nn, float* __restrict__ dist)
{
    __shared__ float4 tile[KNN3_TILE];
    const int i = blockIdx.x * KNN3_THREADS + threadIdx.x;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < N) {
        const float x = source[3 * (size_t)i], y = source[3 * (size_t)i + 1], z = source[3 * (size_t)i + 2];
        px = __fadd_rn(__fmaf_rn(T[2], z, __fmaf_rn(T[1], y, __fmul_rn(T[0], x))), T[3]);
        py = __fadd_rn(__fmaf_rn(T[6], z, __fmaf_rn(T[5], y, __fmul_rn(T[4], x))), T[7]);
        pz = __fadd_rn(__fmaf_rn(T[10], z, __fmaf_rn(T[9], y, __fmul_rn(T[8], x))), T[11]);
    }
    float best = INFINITY; int arg = -1;
    for (int j0 = 0; j0 < M; j0 += KNN3_TILE) {
        const int n = min(KNN3_TILE, M - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += KNN3_THREADS)
            tile[j] = make_float4(target[3 * (size_t)(j0 + j)], target[3 * (size_t)(j0 + j) + 1], target[3 * (size_t)(j0 + j) + 2], 0.f);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 q = tile[j];
            const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            if (d2 < best || arg < 0) { best = d2; arg = j0 + j; }
        }
    }
    if (i < N) { nn[i] = arg; dist[i] = __fsqrt_rn(__fadd_rn(best, 1e-12f)); }
}

// ordered compaction of the pairs with dist < voxel (one CTA, ascending source index like the reference's boolean mask)
__global__ void __launch_bounds__(256) knn3_select_kernel(const int32_t* __restrict__ nn, const float* __restrict__ dist, int N, float voxel,
                                                          int64_t* __restrict__ pairs, int32_t* __restrict__ count, int64_t* __restrict__ nn_out)
{
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool flag = false; int j = -1;
        if (i < N) { j = nn[i]; flag = (j >= 0) && (dist[i] < voxel); if (nn_out) nn_out[i] = (int64_t)j; }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int pre = base_s, tot = 0;
        for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; if (w < warp) pre += c; tot += c; }
        if (flag) { const int pos = pre + __popc(bal & ((1u << lane) - 1u)); pairs[2 * (size_t)pos] = (int64_t)i; pairs[2 * (size_t)pos + 1] = (int64_t)j; }
        __syncthreads();
        if (threadIdx.x == 0) base_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_s;
}

// ---- f3: 3x3 SVD, one matrix per thread ------------------------------------------------------------------------
BFR_DEVINL void svd3(const float x[9], float u[9], float s[3], float v[9])
{
    float W[3][3], V[3][3] = { { 1.f, 0.f, 0.f }, { 0.f, 1.f, 0.f }, { 0.f, 0.f, 1.f } };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) W[r][c] = x[3 * r + c];
#pragma unroll 1
    for (int sweep = 0; sweep < 4; ++sweep) { jacobi_pair<0, 1>(W, V); jacobi_pair<0, 2>(W, V); jacobi_pair<1, 2>(W, V); }
    float n2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) n2[k] = __fmaf_rn(W[2][k], W[2][k], __fmaf_rn(W[1][k], W[1][k], __fmul_rn(W[0][k], W[0][k])));
    int o0 = 0, o1 = 1, o2 = 2;
    auto N2 = [&](int k) { return k == 0 ? n2[0] : (k == 1 ? n2[1] : n2[2]); };
    if (N2(o1) > N2(o0)) { const int t = o0; o0 = o1; o1 = t; }
    if (N2(o2) > N2(o1)) { const int t = o1; o1 = o2; o2 = t; }
    if (N2(o1) > N2(o0)) { const int t = o0; o0 = o1; o1 = t; }
    const int o[3] = { o0, o1, o2 };
    float uu[3][3], vv[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[k] = __fsqrt_rn(N2(o[k]));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            vv[k][r] = o[k] == 0 ? V[r][0] : (o[k] == 1 ? V[r][1] : V[r][2]);
            uu[k][r] = o[k] == 0 ? W[r][0] : (o[k] == 1 ? W[r][1] : W[r][2]);
        }
    }
    const float tiny = __fmul_rn(1e-6f, s[0]);
    if (s[0] > 0.0f) {
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[0][r] = __fdiv_rn(uu[0][r], s[0]);
    } else { uu[0][0] = 1.0f; uu[0][1] = 0.0f; uu[0][2] = 0.0f; }
    if (s[1] > tiny && s[1] > 0.0f) {
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[1][r] = __fdiv_rn(uu[1][r], s[1]);
        gram_schmidt2(uu[0], uu[1]);
    } else {
        int m = 0;
        if (fabsf(uu[0][1]) < fabsf(uu[0][m])) m = 1;
        if (fabsf(uu[0][2]) < fabsf(m == 0 ? uu[0][0] : uu[0][1])) m = 2;
        const float e[3] = { m == 0 ? 1.0f : 0.0f, m == 1 ? 1.0f : 0.0f, m == 2 ? 1.0f : 0.0f };
        cross3(uu[0], e, uu[1]);
        const float nn_ = __fsqrt_rn(dot3(uu[1], uu[1]));
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[1][r] = __fdiv_rn(uu[1][r], nn_);
    }
    if (s[2] > tiny && s[2] > 0.0f) {
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[2][r] = __fdiv_rn(uu[2][r], s[2]);
        const float d0 = dot3(uu[0], uu[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[2][r] = __fmaf_rn(-d0, uu[0][r], uu[2][r]);
        gram_schmidt2(uu[1], uu[2]);
    } else {
        float c[3], cv[3];
        cross3(uu[0], uu[1], c);
        cross3(vv[0], vv[1], cv);
        const float sg = dot3(cv, vv[2]) < 0.0f ? -1.0f : 1.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[2][r] = __fmul_rn(sg, c[r]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r) { u[3 * r + k] = uu[k][r]; v[3 * r + k] = vv[k][r]; }
}

__global__ void svd3_kernel(const float* __restrict__ x, int B, float* __restrict__ u, float* __restrict__ s, float* __restrict__ v)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float xi[9], ui[9], si[3], vi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) xi[k] = x[9 * (size_t)b + k];
    svd3(xi, ui, si, vi);
#pragma unroll
    for (int k = 0; k < 9; ++k) { u[9 * (size_t)b + k] = ui[k]; v[9 * (size_t)b + k] = vi[k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) s[3 * (size_t)b + k] = si[k];
}

// ---- f4: farthest point sampling (pointnet2_ops.furthest_point_sample, reference models/BUFFER.py:266-267) ---------------
// One CTA per cloud; the running min-distance array lives in the caller's workspace; each round every thread updates its
// strided points and the CTA reduces (value, ~index) with a packed 64-bit max (ties -> lowest index).
constexpr int FPS_THREADS = 1024;
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float* __restrict__ xyz, int N, int npoint, int32_t* __restrict__ idx, float* __restrict__ temp)
{
    __shared__ unsigned long long red[FPS_THREADS / 32];
    __shared__ int s_old;
    const int b = blockIdx.x;
    const float* p = xyz + (size_t)b * N * 3;
    float* tmp = temp + (size_t)b * N;
    int32_t* out = idx + (size_t)b * npoint;
    for (int k = threadIdx.x; k < N; k += FPS_THREADS) tmp[k] = 1e10f;
    if (threadIdx.x == 0) { out[0] = 0; s_old = 0; }
    __syncthreads();
    for (int j = 1; j < npoint; ++j) {
        const int old = s_old;
        const float x1 = p[3 * (size_t)old], y1 = p[3 * (size_t)old + 1], z1 = p[3 * (size_t)old + 2];
        unsigned long long best = 0ull;                                  // (key(-1) would be smaller than any d2 >= 0 key; 0 = "none")
        for (int k = threadIdx.x; k < N; k += FPS_THREADS) {
            const float x2 = p[3 * (size_t)k], y2 = p[3 * (size_t)k + 1], z2 = p[3 * (size_t)k + 2];
            const float mag = __fmaf_rn(z2, z2, __fmaf_rn(y2, y2, __fmul_rn(x2, x2)));
            if (mag <= 1e-3f) continue;
            const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float d2 = d < tmp[k] ? d : tmp[k];
            tmp[k] = d2;
            const unsigned long long cand = pack_best(float_key(d2), (uint32_t)k);
            best = cand > best ? cand : best;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long v = threadIdx.x < FPS_THREADS / 32 ? red[threadIdx.x] : 0ull;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o); v = other > v ? other : v; }
            if (threadIdx.x == 0) { const int nxt = v ? (int)packed_index(v) : 0; s_old = nxt; out[j] = nxt; }
        }
        __syncthreads();
    }
}

cudaError_t fps_launch(const float* xyz, int B, int N, int npoint, int32_t* idx, float* temp, cudaStream_t stream)
{
    if (B > 0 && N > 0 && npoint > 0) fps_kernel<<<B, FPS_THREADS, 0, stream>>>(xyz, N, npoint, idx, temp);
    return cudaGetLastError();
}

size_t knn3_workspace_bytes(int N) { return (size_t)(N > 0 ? N : 1) * 8 + 64; }

cudaError_t get_matching_indices_launch(const float* source, int N, const float* target, int M, const float* T, float voxel,
                                        int64_t* pairs, int32_t* count, int64_t* nn_out, float* dist_out, void* ws, cudaStream_t stream)
{
    int32_t* nn = reinterpret_cast<int32_t*>(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    float* dist = dist_out ? dist_out : reinterpret_cast<float*>(nn + (N > 0 ? N : 1));
    if (N > 0) knn3_kernel<<<(N + KNN3_THREADS - 1) / KNN3_THREADS, KNN3_THREADS, 0, stream>>>(source, N, target, M, T, nn, dist);
    knn3_select_kernel<<<1, 256, 0, stream>>>(nn, dist, N, voxel, pairs, count, nn_out);
    return cudaGetLastError();
}

cudaError_t svd3_launch(const float* x, int B, float* u, float* s, float* v, cudaStream_t stream)
{
    if (B > 0) svd3_kernel<<<(B + 127) / 128, 128, 0, stream>>>(x, B, u, s, v);
    return cudaGetLastError();
}

}  // namespace bfr
