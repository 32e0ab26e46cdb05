// ransac.cu — K2 + K3: Philox hypothesis sampling, batched closed-form Kabsch, cheap checks, SE(3) inlier scoring
// with a packed max-reduction; plus the a3/a4 stage of BUFFER (LRF hypotheses + scoring) on the same scoring core.
//
// Replaces the Open3D 0.13 CPU RANSAC call of the reference (models/BUFFER.py:313-326) and the [A,A,3] broadcast
// scoring block (models/BUFFER.py:303-311).  Semantics and exact arithmetic: oracle/bfr_oracle.c (orc_hypothesis,
// orc_count_inliers, orc_ransac, orc_score_hypotheses); DESIGN.md §K2/§K3.
//
// Structure: one CTA owns a contiguous range of hypothesis indices of one pair.  Every thread draws one hypothesis per
// round (Philox counter = (h, pair_id, 0, 0)) and runs the cheap checks (repeated index, edge lengths); survivors (~10 %)
// are compacted into queue 1 so that the 3-point Kabsch + distance check runs on dense warps; what passes (a few %) goes
// to queue 2.  Whenever queue 2 holds a full block's worth, each thread takes one hypothesis and scores it against ALL
// correspondences, which stream
// through shared memory in 2048-correspondence chunks laid out pair-interleaved so that the transform is FFMA2
// (two correspondences per instruction) with warp-uniform LDS.128 broadcasts.  The inlier count never leaves the
// thread; the CTA's best (count << 32 | ~h) goes out with one 64-bit atomicMax.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include <cmath>

namespace bfr {

constexpr int RS_THREADS = 256;
constexpr int RS_CHUNK = 2048;                  // correspondences per shared-memory chunk
#ifndef RS_S1_N
#define RS_S1_N 4
#endif
constexpr int RS_S1 = RS_S1_N;                     // stage-1 hypotheses per thread and round
constexpr int RS_QCAP = 2 * RS_THREADS;            // queue 2 holds < RS_THREADS leftovers plus the survivors of one fit block
constexpr int RS_Q1CAP = (1 + RS_S1) * RS_THREADS; // queue 1 holds < RS_THREADS leftovers plus one round's survivors

struct __align__(16) RsSmem {
    float4 chunk[RS_CHUNK / 2][4];              // per pair of correspondences: (sx sx' sy sy')(sz sz' qx qx')(qy qy' qz qz')(w w' - -)
    float q[12][RS_QCAP];                       // queue 2: hypotheses that passed every check: R (9) + t (3), SoA
    uint32_t qh[RS_QCAP];
    uint4 q1[RS_Q1CAP];                        // queue 1: survivors of the cheap checks: {h, i0, i1, i2}
    int q1count;
    unsigned long long red[RS_THREADS / 32];
    int qcount;
};

// cooperative load of correspondences [c0, c0 + RS_CHUNK) of one pair into the pair-interleaved layout
BFR_DEVINL void load_chunk(RsSmem& sm, const float4* __restrict__ corr, int K, int c0)
{
    for (int g = threadIdx.x; g < RS_CHUNK / 2; g += RS_THREADS) {
        const int c = c0 + 2 * g;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), b0 = make_float4(1e18f, 1e18f, 1e18f, 0.f), a1 = a0, b1 = b0;   // padding never counts
        if (c < K)     { a0 = __ldg(&corr[2 * (size_t)c]);     b0 = __ldg(&corr[2 * (size_t)c + 1]); }
        if (c + 1 < K) { a1 = __ldg(&corr[2 * (size_t)c + 2]); b1 = __ldg(&corr[2 * (size_t)c + 3]); }
        sm.chunk[g][0] = make_float4(a0.x, a1.x, a0.y, a1.y);
        sm.chunk[g][1] = make_float4(a0.z, a1.z, b0.x, b1.x);
        sm.chunk[g][2] = make_float4(b0.y, b1.y, b0.z, b1.z);
        sm.chunk[g][3] = make_float4(a0.w, a1.w, 0.f, 0.f);
    }
}

// ++count iff d < thr, as one FSETP + one predicated IADD (the compiler's select form costs an extra add per test)
BFR_DEVINL void count_if_lt(int& count, float d, float thr)
{
    asm("{\n\t.reg .pred q;\n\tsetp.lt.f32 q, %1, %2;\n\t@q add.s32 %0, %0, 1;\n\t}" : "+r"(count) : "f"(d), "f"(thr));
}

// inlier count of one hypothesis over pairs [g_begin, g_end) of the chunk currently in shared memory
template <bool PER_CORR_THR>
BFR_DEVINL int score_chunk(const RsSmem& sm, int g_begin, int g_end, const float R[9], const float t[3], float d2max)
{
    f32x2 Rb[9], tb[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = pack2(R[k], R[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) tb[k] = pack2(t[k], t[k]);
    int count = 0;
#pragma unroll 4
    for (int g = g_begin; g < g_end; ++g) {
        const float4 L0 = sm.chunk[g][0], L1 = sm.chunk[g][1], L2 = sm.chunk[g][2];     // warp-uniform -> broadcast
        const f32x2 sx = pack2(L0.x, L0.y), sy = pack2(L0.z, L0.w), sz = pack2(L1.x, L1.y);
        const f32x2 qx = pack2(L1.z, L1.w), qy = pack2(L2.x, L2.y), qz = pack2(L2.z, L2.w);
        const f32x2 x = sub2(fma2(Rb[0], sx, fma2(Rb[1], sy, fma2(Rb[2], sz, tb[0]))), qx);
        const f32x2 y = sub2(fma2(Rb[3], sx, fma2(Rb[4], sy, fma2(Rb[5], sz, tb[1]))), qy);
        const f32x2 z = sub2(fma2(Rb[6], sx, fma2(Rb[7], sy, fma2(Rb[8], sz, tb[2]))), qz);
        const f32x2 d2 = fma2(x, x, fma2(y, y, mul2(z, z)));
        float da, db;
        unpack2(d2, da, db);
        if (PER_CORR_THR) {
            const float2 w = *reinterpret_cast<const float2*>(&sm.chunk[g][3]);
            count_if_lt(count, da, w.x);
            count_if_lt(count, db, w.y);
        } else {
            count_if_lt(count, da, d2max);
            count_if_lt(count, db, d2max);
        }
    }
    return count;
}

// score `n` queued hypotheses against all K correspondences; fold into `best`.  A full queue (n = RS_THREADS) gives every thread one
// hypothesis.  A partial flush (n < RS_THREADS) would leave most warps idle while the chunks still stream through shared memory, so the
// nw = ceil(n / 32) warps' worth of hypotheses are replicated over the 8 / nw groups of warps and every group scores its own slice of each
// chunk (the loads stay warp-uniform broadcasts); the partial counts are integer sums, so the total is exact whatever the split.
BFR_DEVINL void score_queue(RsSmem& sm, const float4* __restrict__ corr, int K, int n, float d2max, unsigned long long& best)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (n + 31) >> 5;                                     // warps that hold hypotheses
    const int nparts = (RS_THREADS / 32) / nw;                        // correspondence slices (1 for a full queue)
    const int part = warp / nw, hi = (warp % nw) * 32 + lane;         // this thread: hypothesis hi of the queue, slice `part`
    const bool warp_has_work = part < nparts;
    const bool mine = warp_has_work && hi < n;
    float R[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, t[3] = { 0.f, 0.f, 0.f }; uint32_t h = 0;
    if (mine) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = sm.q[k][hi];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = sm.q[9 + k][hi];
        h = sm.qh[hi];
    }
    int* partial = reinterpret_cast<int*>(sm.q1);                     // queue 1 is empty whenever a partial flush runs
    if (nparts > 1) partial[threadIdx.x] = 0;                         // (made visible by the barriers of the chunk loop)
    int count = 0;
    for (int c0 = 0; c0 < K; c0 += RS_CHUNK) {
        __syncthreads();                       // previous chunk fully consumed
        load_chunk(sm, corr, K, c0);
        __syncthreads();
        const int npairs = (min(RS_CHUNK, K - c0) + 1) >> 1;
        if (warp_has_work) count += score_chunk<false>(sm, (part * npairs) / nparts, ((part + 1) * npairs) / nparts, R, t, d2max);
    }
    if (nparts > 1) {
        if (mine) atomicAdd(&partial[hi], count);
        __syncthreads();
        count = partial[hi < RS_THREADS ? hi : 0];
    }
    if (mine && part == 0) {
        const unsigned long long packed = ((unsigned long long)(uint32_t)count << 32) | (unsigned long long)(0xFFFFFFFFu - h);
        best = packed > best ? packed : best;
    }
}

__global__ void __launch_bounds__(RS_THREADS, 2)
ransac_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt,
              uint64_t seed, uint32_t pair_id_base, uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th,
              unsigned long long* __restrict__ best_packed, int32_t* __restrict__ valid_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(smem_raw);
    const int p = blockIdx.y;
    const int K = corr_cnt[p];
    if (K < 3 || h_end <= h_begin) return;
    const float4* corr_p = corr + 2 * (size_t)corr_off[p];
    const uint32_t nh = h_end - h_begin;
    const uint32_t hb = h_begin + (uint32_t)(((unsigned long long)blockIdx.x * nh) / gridDim.x);
    const uint32_t he = h_begin + (uint32_t)(((unsigned long long)(blockIdx.x + 1) * nh) / gridDim.x);
    const float d2max = __fmul_rn(dist_th, dist_th), sim2 = __fmul_rn(similar_th, similar_th);
    const uint32_t pair_id = pair_id_base + (uint32_t)p;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) { sm.qcount = 0; sm.q1count = 0; }
    __syncthreads();
    unsigned long long best = 0ull;
    int n_scored = 0;                                                 // hypotheses that passed every check (thread 0's tally)

    // stage 2 on n queue-1 entries (thread i takes entry i): Kabsch + distance check on dense warps, survivors -> queue 2;
    // scores a full block of queue 2 whenever one is available
    auto fit_queue1 = [&](int base, int n) {
        float R[9], t[3];
        bool ok = false; uint32_t h = 0;
        if ((int)threadIdx.x < n) {
            const uint4 e = sm.q1[base + threadIdx.x];
            h = e.x;
            const uint32_t id[3] = { e.y, e.z, e.w };
            float s[3][3], q[3][3];
            load_sample(corr_p, id, s, q);
            ok = hypothesis_fit(s, q, d2max, R, t);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (bal) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&sm.qcount, __popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
            if (ok) {
#pragma unroll
                for (int k = 0; k < 9; ++k) sm.q[k][pos] = R[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) sm.q[9 + k][pos] = t[k];
                sm.qh[pos] = h;
            }
        }
        __syncthreads();
        const int qn = sm.qcount;
        __syncthreads();
        if (qn >= RS_THREADS) {
            score_queue(sm, corr_p, K, RS_THREADS, d2max, best);
            n_scored += RS_THREADS;
            const int rem = qn - RS_THREADS;                           // move the overflow [RS_THREADS, qn) down to the front
            float mv[12]; uint32_t mh = 0;
            if ((int)threadIdx.x < rem) {
#pragma unroll
                for (int k = 0; k < 12; ++k) mv[k] = sm.q[k][RS_THREADS + threadIdx.x];
                mh = sm.qh[RS_THREADS + threadIdx.x];
            }
            __syncthreads();
            if ((int)threadIdx.x < rem) {
#pragma unroll
                for (int k = 0; k < 12; ++k) sm.q[k][threadIdx.x] = mv[k];
                sm.qh[threadIdx.x] = mh;
            }
            if (threadIdx.x == 0) sm.qcount = rem;
            __syncthreads();
        }
    };

    for (uint32_t base = hb; base < he; base += RS_S1 * RS_THREADS) {
        // stage 1: RS_S1 independent hypotheses per thread (their sample gathers overlap), cheap checks only (~10 % survive at 70 % outliers)
        uint32_t hh[RS_S1], id[RS_S1][3];
        bool ok[RS_S1];
#pragma unroll
        for (int u = 0; u < RS_S1; ++u) {
            hh[u] = base + (uint32_t)u * RS_THREADS + threadIdx.x;
            id[u][0] = id[u][1] = id[u][2] = 0u;
            ok[u] = false;
            if (hh[u] < he) { float s[3][3], q[3][3]; ok[u] = hypothesis_precheck(corr_p, (uint32_t)K, seed, pair_id, hh[u], sim2, id[u], s, q); }
        }
#pragma unroll
        for (int u = 0; u < RS_S1; ++u) {                              // queue order = hypothesis order within the round (any order gives the same best)
            const unsigned bal = __ballot_sync(0xffffffffu, ok[u]);
            if (bal) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&sm.q1count, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
                if (ok[u]) sm.q1[pos] = make_uint4(hh[u], id[u][0], id[u][1], id[u][2]);
            }
        }
        __syncthreads();
        const int n1 = sm.q1count;
        __syncthreads();                       // everyone has read q1count before it changes
        if (n1 >= RS_THREADS) {
            int done = 0;
            for (; n1 - done >= RS_THREADS; done += RS_THREADS) fit_queue1(done, RS_THREADS);
            const int rem = n1 - done;                                 // < RS_THREADS leftovers move to the front
            uint4 mv = make_uint4(0u, 0u, 0u, 0u);
            if ((int)threadIdx.x < rem) mv = sm.q1[done + threadIdx.x];
            __syncthreads();
            if ((int)threadIdx.x < rem) sm.q1[threadIdx.x] = mv;
            if (threadIdx.x == 0) sm.q1count = rem;
            __syncthreads();
        }
    }
    {
        const int n1 = sm.q1count;             // flush queue 1, then queue 2
        __syncthreads();
        if (n1 > 0) fit_queue1(0, n1);
    }
    const int qn = sm.qcount;
    if (qn > 0) { score_queue(sm, corr_p, K, qn, d2max, best); n_scored += qn; }
    if (valid_count && threadIdx.x == 0 && n_scored) atomicAdd(valid_count + p, n_scored);

    // block max -> one atomic
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    __syncthreads();
    if (lane == 0) sm.red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = 0ull;
        for (int w = 0; w < RS_THREADS / 32; ++w) b = sm.red[w] > b ? sm.red[w] : b;
        if (b) atomicMax(best_packed + p, b);
    }
}

// decode the packed best of each pair and regenerate the winning minimal-sample fit (counter-based RNG: no broadcast
// of R,t needed, also across GPUs).  Identity when nothing was valid / K < 3 (Open3D's default result; reference
// ThreeDMatch/test.py:242-245 maps failure to eye(4)).
__global__ void ransac_finalize_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt, int P,
                                       uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th,
                                       const unsigned long long* __restrict__ best_packed, float* __restrict__ T, int32_t* __restrict__ inliers, int64_t* __restrict__ best_h)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) out[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    int32_t cnt = 0; int64_t bh = -1;
    const unsigned long long b = best_packed[p];
    const int K = corr_cnt[p];
    if (b != 0ull && K >= 3) {
        const uint32_t h = 0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull);
        float R[9], t[3];
        if (make_hypothesis(corr + 2 * (size_t)corr_off[p], (uint32_t)K, seed, pair_id_base + (uint32_t)p, h,
                            __fmul_rn(dist_th, dist_th), __fmul_rn(similar_th, similar_th), R, t)) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { out[4 * r] = R[3 * r]; out[4 * r + 1] = R[3 * r + 1]; out[4 * r + 2] = R[3 * r + 2]; out[4 * r + 3] = t[r]; }
            cnt = (int32_t)(b >> 32); bh = (int64_t)h;
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) T[16 * (size_t)p + k] = out[k];
    if (inliers) inliers[p] = cnt;
    if (best_h) best_h[p] = bh;
}

// ---- a3: per-correspondence pose hypotheses from local reference frames (models/BUFFER.py:294-301) ----------------
__global__ void lrf_hypotheses_kernel(const float* __restrict__ cs, const float* __restrict__ ss_R, const float* __restrict__ tt_R,
                                      const float* __restrict__ ss_kpts, const float* __restrict__ tt_kpts, int A, float* __restrict__ R_out, float* __restrict__ t_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    const float c = cs[2 * i], s = cs[2 * i + 1];
    float Rt[9], Rs[9], Mx[9], R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rt[k] = tt_R[9 * (size_t)i + k]; Rs[k] = ss_R[9 * (size_t)i + k]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Mx[3 * r + 0] = __fmaf_rn(Rt[3 * r + 1], s, __fmul_rn(Rt[3 * r + 0], c));
        Mx[3 * r + 1] = __fmaf_rn(Rt[3 * r + 1], c, -__fmul_rn(Rt[3 * r + 0], s));
        Mx[3 * r + 2] = Rt[3 * r + 2];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            R[3 * r + cc] = __fmaf_rn(Mx[3 * r + 2], Rs[3 * cc + 2], __fmaf_rn(Mx[3 * r + 1], Rs[3 * cc + 1], __fmul_rn(Mx[3 * r + 0], Rs[3 * cc + 0])));
    const float px = ss_kpts[3 * (size_t)i], py = ss_kpts[3 * (size_t)i + 1], pz = ss_kpts[3 * (size_t)i + 2];
#pragma unroll
    for (int k = 0; k < 9; ++k) R_out[9 * (size_t)i + k] = R[k];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t_out[3 * (size_t)i + r] = __fsub_rn(tt_kpts[3 * (size_t)i + r], __fmaf_rn(R[3 * r + 2], pz, __fmaf_rn(R[3 * r + 1], py, __fmul_rn(R[3 * r + 0], px))));
}

// ---- a4: score explicit hypotheses (models/BUFFER.py:303-311) ----------------------------------------------------
// records for the scoring core: {sx sy sz thr^2 | qx qy qz 0}
__global__ void make_records_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const float* __restrict__ thr, float thr_scalar, int C, float4* __restrict__ rec)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float th = thr ? thr[c] : thr_scalar;
    rec[2 * (size_t)c] = make_float4(src[3 * (size_t)c], src[3 * (size_t)c + 1], src[3 * (size_t)c + 2], __fmul_rn(th, th));
    rec[2 * (size_t)c + 1] = make_float4(tgt[3 * (size_t)c], tgt[3 * (size_t)c + 1], tgt[3 * (size_t)c + 2], 0.0f);
}

__global__ void __launch_bounds__(RS_THREADS, 2)
score_hypotheses_kernel(const float* __restrict__ Rh, const float* __restrict__ th, int H, const float4* __restrict__ rec, int C,
                        int32_t* __restrict__ counts, unsigned long long* __restrict__ best_packed)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(smem_raw);
    const int h = blockIdx.x * RS_THREADS + threadIdx.x;
    const bool mine = h < H;
    float R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = mine ? Rh[9 * (size_t)h + k] : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = mine ? th[3 * (size_t)h + k] : 0.0f;
    const bool warp_has_work = (int)(blockIdx.x * RS_THREADS + (threadIdx.x & ~31u)) < H;
    int count = 0;
    for (int c0 = 0; c0 < C; c0 += RS_CHUNK) {
        __syncthreads();
        load_chunk(sm, rec, C, c0);
        __syncthreads();
        const int npairs = (min(RS_CHUNK, C - c0) + 1) >> 1;
        if (warp_has_work) count += score_chunk<true>(sm, 0, npairs, R, t, 0.0f);
    }
    unsigned long long best = 0ull;
    if (mine) {
        if (counts) counts[h] = count;
        // +1 so that an all-zero-count winner is still distinguishable from "empty"; torch.argmax -> first maximum
        best = ((unsigned long long)(uint32_t)(count + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)h);
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    if (lane == 0) sm.red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = 0ull;
        for (int w = 0; w < RS_THREADS / 32; ++w) b = sm.red[w] > b ? sm.red[w] : b;
        if (b) atomicMax(best_packed, b);
    }
}

// inlier mask of the winning hypothesis (models/BUFFER.py:311) + its index
__global__ void score_mask_kernel(const float* __restrict__ Rh, const float* __restrict__ th, const float4* __restrict__ rec, int C,
                                  const unsigned long long* __restrict__ best_packed, uint8_t* __restrict__ mask, int64_t* __restrict__ best_idx)
{
    const unsigned long long b = *best_packed;
    if (b == 0ull) { if (blockIdx.x == 0 && threadIdx.x == 0 && best_idx) *best_idx = -1; return; }
    const uint32_t h = 0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0 && best_idx) *best_idx = (int64_t)h;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C || !mask) return;
    float R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Rh[9 * (size_t)h + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = th[3 * (size_t)h + k];
    const float4 a = rec[2 * (size_t)c], q = rec[2 * (size_t)c + 1];
    mask[c] = (resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z) < a.w) ? 1 : 0;
}

// ---- host launchers -------------------------------------------------------------------------------------------
static cudaError_t ensure_smem(const void* fn)
{
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem));
}

cudaError_t ransac_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                          uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, int splits, unsigned long long* best_packed, int32_t* valid_count,
                          cudaStream_t stream)
{
    static bool once = false;
    if (!once) { cudaError_t e = ensure_smem((const void*)ransac_kernel); if (e != cudaSuccess) return e; once = true; }
    if (P <= 0 || h_end <= h_begin) return cudaSuccess;
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)splits, (unsigned)P);
    ransac_kernel<<<grid, RS_THREADS, sizeof(RsSmem), stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, seed, pair_id_base,
                                                                h_begin, h_end, dist_th, similar_th, best_packed, valid_count);
    return cudaGetLastError();
}

cudaError_t ransac_finalize_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                                   float dist_th, float similar_th, const unsigned long long* best_packed, float* T, int32_t* inliers, int64_t* best_h, cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    ransac_finalize_kernel<<<(P + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, P, seed, pair_id_base,
                                                                dist_th, similar_th, best_packed, T, inliers, best_h);
    return cudaGetLastError();
}

cudaError_t lrf_hypotheses_launch(const float* cs, const float* ss_R, const float* tt_R, const float* ss_kpts, const float* tt_kpts, int A,
                                  float* R_out, float* t_out, cudaStream_t stream)
{
    if (A > 0) lrf_hypotheses_kernel<<<(A + 127) / 128, 128, 0, stream>>>(cs, ss_R, tt_R, ss_kpts, tt_kpts, A, R_out, t_out);
    return cudaGetLastError();
}

size_t score_workspace_bytes(int C) { return (size_t)(C > 0 ? C : 1) * 32 + 64; }

cudaError_t score_hypotheses_launch(const float* R, const float* t, int H, const float* src, const float* tgt, int C, const float* thr, float thr_scalar,
                                    int32_t* counts, unsigned long long* best_packed, int64_t* best_idx, uint8_t* mask, void* ws, cudaStream_t stream)
{
    static bool once = false;
    if (!once) { cudaError_t e = ensure_smem((const void*)score_hypotheses_kernel); if (e != cudaSuccess) return e; once = true; }
    float4* rec = reinterpret_cast<float4*>(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    cudaError_t e = cudaMemsetAsync(best_packed, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (C > 0) make_records_kernel<<<(C + 255) / 256, 256, 0, stream>>>(src, tgt, thr, thr_scalar, C, rec);
    if (H > 0) score_hypotheses_kernel<<<(H + RS_THREADS - 1) / RS_THREADS, RS_THREADS, sizeof(RsSmem), stream>>>(R, t, H, rec, C, counts, best_packed);
    score_mask_kernel<<<(max(C, 1) + 255) / 256, 256, 0, stream>>>(R, t, rec, C, best_packed, mask, best_idx);
    return cudaGetLastError();
}

}  // namespace bfr
