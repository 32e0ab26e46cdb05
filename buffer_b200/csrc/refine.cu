// refine.cu — K4: weighted Kabsch (rigid_transform_3d) and the <=20-round post_refinement loop, entirely on device.
//
// Replaces rigid_transform_3d (reference models/BUFFER.py:424-464: diag_embed(weights) [n x n], Am^T W Bm and a
// GPU->CPU->GPU torch.svd per call) and buffer.post_refinement (models/BUFFER.py:382-418: one host sync per round for
// int(inlier_num)).  The weighted sums use a FIXED reduction tree that oracle/bfr_oracle.c mirrors, so the whole loop —
// inlier counts, number of rounds and the final transform — is bit-reproducible against the CPU oracle:
//     C = 1 (n <= 16384) or 8 (larger n) blocks of 256 lanes; lane l of block c accumulates elements c*256 + l, + 256 C, ...
//     sequentially; xor-butterfly 16,8,4,2,1 inside each warp; the 8 warp totals of a block added in warp order; the C block totals
//     added in block order.
// One CTA of 256 threads per pair / batch element walks its C blocks one after the other; for huge pairs (BASELINE config 5: 100k
// correspondences) post_refinement_cluster_kernel gives each block of the tree to one CTA of an 8-CTA thread-block cluster and
// exchanges the block totals through distributed shared memory — same tree, same bits, 8 SMs instead of one.
// The 3x3 SVD is the same closed form as K2.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace bfr {

constexpr int RF_THREADS = 256;
constexpr int RF_BIG_N = 16384;                 // more elements than this: the tree has RF_BIG_C blocks of 256 lanes
constexpr int RF_BIG_C = 8;                     // = the (portable) cluster size of post_refinement_cluster_kernel

BFR_DEVINL int rf_blocks(int n) { return n > RF_BIG_N ? RF_BIG_C : 1; }

struct RfSmem {
    float red[9][RF_THREADS / 32];
    int redi[RF_THREADS / 32];
    float xch[9];                               // this CTA's block total, read by the other CTAs of the cluster
    int xchi;
};

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

// Sum of NV per-element quantities over n elements with the fixed tree above.  acc(i, v) adds element i's terms to v.
// CLUSTER: this CTA is block `crank` of an RF_BIG_C-CTA cluster and n > RF_BIG_N (every CTA of the cluster calls this in lock step);
// otherwise one CTA walks the C blocks itself.  Every thread of every participating CTA returns the same totals.
template <bool CLUSTER, int NV, typename F>
BFR_DEVINL void tree_sum(int n, F acc, float (&tot)[NV], RfSmem& sm)
{
    const int C = rf_blocks(n);
    if (CLUSTER && C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        const int crank = (int)cluster.block_rank();
        float v[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = 0.0f;
        for (int i = crank * RF_THREADS + threadIdx.x; i < n; i += RF_THREADS * RF_BIG_C) acc(i, v);
        block_sum<NV>(v, sm.red);
        if (threadIdx.x == 0)
#pragma unroll
            for (int k = 0; k < NV; ++k) sm.xch[k] = v[k];
        cluster.sync();
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] = 0.0f;
        for (int c = 0; c < RF_BIG_C; ++c) {
            const float* remote = cluster.map_shared_rank(sm.xch, c);
#pragma unroll
            for (int k = 0; k < NV; ++k) tot[k] = (c == 0) ? remote[k] : __fadd_rn(tot[k], remote[k]);
        }
        cluster.sync();                         // everybody has read xch before it is overwritten
    } else {
        for (int c = 0; c < C; ++c) {
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.0f;
            for (int i = c * RF_THREADS + threadIdx.x; i < n; i += RF_THREADS * C) acc(i, v);
            block_sum<NV>(v, sm.red);
#pragma unroll
            for (int k = 0; k < NV; ++k) tot[k] = (c == 0) ? v[k] : __fadd_rn(tot[k], v[k]);
        }
    }
}

template <bool CLUSTER, typename F>
BFR_DEVINL int tree_count(int n, F pred, RfSmem& sm)
{
    if (CLUSTER && rf_blocks(n) > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        const int crank = (int)cluster.block_rank();
        int cnt = 0;
        for (int i = crank * RF_THREADS + threadIdx.x; i < n; i += RF_THREADS * RF_BIG_C) cnt += pred(i) ? 1 : 0;
        cnt = block_sum_int(cnt, sm.redi);
        if (threadIdx.x == 0) sm.xchi = cnt;
        cluster.sync();
        int tot = 0;
        for (int c = 0; c < RF_BIG_C; ++c) tot += *cluster.map_shared_rank(&sm.xchi, c);
        cluster.sync();
        return tot;
    }
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += RF_THREADS) cnt += pred(i) ? 1 : 0;
    return block_sum_int(cnt, sm.redi);
}

// Weighted Kabsch over n points given by functor f(i, a[3], b[3]) -> weight.  Every thread returns the same T
// (row-major 3x4: R | t).  Matches orc_rigid_transform_3d.
template <bool CLUSTER, typename F>
BFR_DEVINL void weighted_kabsch_block(int n, F f, float R[9], float t[3], RfSmem& sm)
{
    float s7[7];
    tree_sum<CLUSTER, 7>(n, [&](int i, float (&v)[7]) {
        float a[3], b[3];
        const float w = f(i, a, b);
        v[0] = __fadd_rn(v[0], w);
#pragma unroll
        for (int r = 0; r < 3; ++r) { v[1 + r] = __fmaf_rn(w, a[r], v[1 + r]); v[4 + r] = __fmaf_rn(w, b[r], v[4 + r]); }
    }, s7, sm);
    const float den = __fadd_rn(s7[0], 1e-6f);      // reference: / (sum(weights) + 1e-6), models/BUFFER.py:441-444
    float ca[3], cb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { ca[r] = __fdiv_rn(s7[1 + r], den); cb[r] = __fdiv_rn(s7[4 + r], den); }
    float h9[9];
    tree_sum<CLUSTER, 9>(n, [&](int i, float (&v)[9]) {
        float a[3], b[3];
        const float w = f(i, a, b);
#pragma unroll
        for (int r = 0; r < 3; ++r) { a[r] = __fsub_rn(a[r], ca[r]); b[r] = __fsub_rn(b[r], cb[r]); }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float wa = __fmul_rn(w, a[r]);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[3 * r + c] = __fmaf_rn(wa, b[c], v[3 * r + c]);
        }
    }, h9, sm);
    kabsch_rotation(h9, R);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = __fsub_rn(cb[r], __fmaf_rn(R[3 * r + 2], ca[2], __fmaf_rn(R[3 * r + 1], ca[1], __fmul_rn(R[3 * r + 0], ca[0]))));
}

// rigid_transform_3d(A[bs,n,3], B[bs,n,3], weights[bs,n] or NULL, weight_threshold) -> T[bs,4,4]
__global__ void __launch_bounds__(RF_THREADS) rigid_transform_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ w,
                                                                     int n, float weight_threshold, float* __restrict__ T)
{
    __shared__ RfSmem sm;
    const int b = blockIdx.x;
    const float* Ab = A + (size_t)b * n * 3; const float* Bb = B + (size_t)b * n * 3; const float* wb = w ? w + (size_t)b * n : nullptr;
    float R[9], t[3];
    weighted_kabsch_block<false>(n, [&](int i, float a[3], float q[3]) {
        a[0] = __ldg(Ab + 3 * (size_t)i); a[1] = __ldg(Ab + 3 * (size_t)i + 1); a[2] = __ldg(Ab + 3 * (size_t)i + 2);
        q[0] = __ldg(Bb + 3 * (size_t)i); q[1] = __ldg(Bb + 3 * (size_t)i + 1); q[2] = __ldg(Bb + 3 * (size_t)i + 2);
        float wi = wb ? __ldg(wb + i) : 1.0f;
        if (wi < weight_threshold) wi = 0.0f;        // models/BUFFER.py:437
        return wi;
    }, R, t, sm);
    if (threadIdx.x == 0) {
        float* o = T + 16 * (size_t)b;
#pragma unroll
        for (int r = 0; r < 3; ++r) { o[4 * r] = R[3 * r]; o[4 * r + 1] = R[3 * r + 1]; o[4 * r + 2] = R[3 * r + 2]; o[4 * r + 3] = t[r]; }
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
    }
}

// post_refinement for one pair; corr = 8-float records at corr_off[p], corr_cnt[p] of them (ALL mutual correspondences,
// as the reference passes ss_kpts/tt_kpts, models/BUFFER.py:328).
template <bool CLUSTER>
BFR_DEVINL void post_refinement_pair(int p, bool writer, const float* __restrict__ T0, const float4* __restrict__ corr, const int32_t* __restrict__ corr_off,
                                     const int32_t* __restrict__ corr_cnt, float thr, int max_iter,
                                     float* __restrict__ Tout, int32_t* __restrict__ iters_out, int32_t* __restrict__ inliers_out, RfSmem& sm)
{
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
        const int cnt = tree_count<CLUSTER>(n, [&](int i) {
            const float4 a = __ldg(&c[2 * (size_t)i]), q = __ldg(&c[2 * (size_t)i + 1]);
            return __fsqrt_rn(resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z)) < thr;
        }, sm);
        if (cnt == prev) break;                         // models/BUFFER.py:405-407
        prev = cnt;
        float Rn[9], tn[3];
        weighted_kabsch_block<CLUSTER>(n, [&](int i, float a3[3], float q3[3]) {
            const float4 a = __ldg(&c[2 * (size_t)i]), q = __ldg(&c[2 * (size_t)i + 1]);
            a3[0] = a.x; a3[1] = a.y; a3[2] = a.z; q3[0] = q.x; q3[1] = q.y; q3[2] = q.z;
            const float L2 = __fsqrt_rn(resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z));
            if (!(L2 < thr)) return 0.0f;
            const float u = __fdiv_rn(L2, thr);
            return __fdiv_rn(1.0f, __fmaf_rn(u, u, 1.0f));   // models/BUFFER.py:415
        }, Rn, tn, sm);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = tn[k];
    }
    if (writer && threadIdx.x == 0) {
        float* o = Tout + 16 * (size_t)p;
#pragma unroll
        for (int r = 0; r < 3; ++r) { o[4 * r] = R[3 * r]; o[4 * r + 1] = R[3 * r + 1]; o[4 * r + 2] = R[3 * r + 2]; o[4 * r + 3] = t[r]; }
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
        if (iters_out) iters_out[p] = it;
        if (inliers_out) inliers_out[p] = prev;
    }
}

__global__ void __launch_bounds__(RF_THREADS) post_refinement_kernel(const float* __restrict__ T0, const float4* __restrict__ corr, const int32_t* __restrict__ corr_off,
                                                                     const int32_t* __restrict__ corr_cnt, float thr, int max_iter,
                                                                     float* __restrict__ Tout, int32_t* __restrict__ iters_out, int32_t* __restrict__ inliers_out)
{
    __shared__ RfSmem sm;
    post_refinement_pair<false>((int)blockIdx.x, true, T0, corr, corr_off, corr_cnt, thr, max_iter, Tout, iters_out, inliers_out, sm);
}

// one 8-CTA cluster per pair (launched with cluster dimension RF_BIG_C): pairs with more than RF_BIG_N correspondences split the
// reduction tree over the cluster; smaller pairs are done by every CTA on its own (identical results), CTA 0 writes
__global__ void __launch_bounds__(RF_THREADS) post_refinement_cluster_kernel(const float* __restrict__ T0, const float4* __restrict__ corr, const int32_t* __restrict__ corr_off,
                                                                             const int32_t* __restrict__ corr_cnt, float thr, int max_iter,
                                                                             float* __restrict__ Tout, int32_t* __restrict__ iters_out, int32_t* __restrict__ inliers_out)
{
    __shared__ RfSmem sm;
    const int p = (int)blockIdx.x / RF_BIG_C;
    const bool big = rf_blocks(corr_cnt[p]) > 1;
    if (!big && cg::this_cluster().block_rank() != 0) return;           // small pair: one CTA is enough (no cluster barrier is ever reached)
    post_refinement_pair<true>(p, cg::this_cluster().block_rank() == 0, T0, corr, corr_off, corr_cnt, thr, max_iter, Tout, iters_out, inliers_out, sm);
}

cudaError_t rigid_transform_launch(const float* A, const float* B, const float* w, int bs, int n, float weight_threshold, float* T, cudaStream_t stream)
{
    if (bs > 0) rigid_transform_kernel<<<bs, RF_THREADS, 0, stream>>>(A, B, w, n, weight_threshold, T);
    return cudaGetLastError();
}

// max_count: host-side upper bound of corr_cnt (0 = unknown).  Above RF_BIG_N the cluster kernel is used; the result does not depend
// on the choice (the single-CTA kernel walks the same tree).
cudaError_t post_refinement_launch(const float* T0, const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, float thr, int max_iter,
                                   float* Tout, int32_t* iters_out, int32_t* inliers_out, int max_count, cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    if (max_count > RF_BIG_N) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)P * RF_BIG_C); cfg.blockDim = dim3(RF_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = RF_BIG_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const float4* c4 = reinterpret_cast<const float4*>(corr);
        return cudaLaunchKernelEx(&cfg, post_refinement_cluster_kernel, T0, c4, corr_off, corr_cnt, thr, max_iter, Tout, iters_out, inliers_out);
    }
    post_refinement_kernel<<<P, RF_THREADS, 0, stream>>>(T0, reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, thr, max_iter, Tout, iters_out, inliers_out);
    return cudaGetLastError();
}

}  // namespace bfr
