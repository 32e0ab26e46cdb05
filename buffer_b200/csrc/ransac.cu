// ransac.cu — K2 + K3: Philox hypothesis sampling, batched closed-form Kabsch, cheap checks, SE(3) inlier scoring
// with a packed max-reduction; plus the a3/a4 stage of BUFFER (LRF hypotheses + scoring) on the same scoring core.
//
// Replaces the Open3D 0.13 CPU RANSAC call of the reference (models/BUFFER.py:313-326) and the [A,A,3] broadcast
// scoring block (models/BUFFER.py:303-311).  Semantics and exact arithmetic: oracle/bfr_oracle.c (orc_hypothesis,
// orc_count_inliers, orc_ransac, orc_score_hypotheses, orc_lrf_vote); DESIGN.md §K2/§K3.
//
// ransac_kernel: persistent CTAs (512 threads, one per SM) walk work items = (pair, slice of the hypothesis range) round-robin.
// A pair's correspondences (up to RS_CHUNK = 5120, 24 bytes each: 120 KB) are loaded ONCE per item into shared memory in a
// pair-interleaved layout and serve both the random sample gathers of stage 1 and the scoring loop; larger pairs stream through the
// same buffer in chunks and gather their samples from global memory.  Per round every thread draws RS_S1 hypotheses (Philox counter
// = (h, pair_id, 0, 0)) and runs the cheap checks (repeated index, edge lengths); survivors (~10 %) are compacted into queue 1 so that
// the 3-point Kabsch + distance check runs on dense warps; what passes (a few %) goes to queue 2.  Whenever queue 2 holds a full block's
// worth, each thread takes one hypothesis and scores it against ALL correspondences: FFMA2 over two correspondences per instruction with
// warp-uniform LDS.128 broadcasts.  The inlier count never leaves the thread; a CTA's best (count << 32 | ~h) goes out with one 64-bit
// atomicMax per item.
// confidence < 1 (Open3D's RANSACConvergenceCriteria, models/BUFFER.py:323-324): one item per pair, rounds of 512 hypotheses, and after
// every round the sequential rule of Open3D (stop once iteration >= ceil(log(1-c)/log(1-fitness^3)) of the best so far) is replayed
// in hypothesis order, so the result equals a one-thread sequential run (oracle orc_ransac_confidence) exactly.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include <cmath>

namespace bfr {

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CHUNK = 5120;                  // correspondences per shared-memory chunk; pairs with K <= RS_CHUNK stay resident
#ifndef RS_S1_N
#define RS_S1_N 4
#endif
constexpr int RS_S1 = RS_S1_N;                     // stage-1 hypotheses per thread and round
constexpr int RS_QCAP = 2 * RS_THREADS;            // queue 2 holds < RS_THREADS leftovers plus the survivors of one fit block
constexpr int RS_Q1CAP = (1 + RS_S1) * RS_THREADS; // queue 1 holds < RS_THREADS leftovers plus one round's survivors

#ifdef RS_TIMING
__device__ unsigned long long g_rs_dbg[8];
#define RST(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#else
#define RST(acc, stmt) { stmt; }
#endif

struct __align__(16) RsSmem {
    float4 chunk[RS_CHUNK / 2][3];              // per pair of correspondences: (sx sx' sy sy')(sz sz' qx qx')(qy qy' qz qz')
    float q[12][RS_QCAP];                       // queue 2: hypotheses that passed every check: R (9) + t (3), SoA
    uint32_t qh[RS_QCAP];
    uint4 q1[RS_Q1CAP];                         // queue 1: survivors of the cheap checks: {h, i0, i1, i2}; idle: partial counts / round results
    unsigned long long red[RS_WARPS];
    unsigned long long seq_best;                // confidence mode: state of the sequential replay
    uint32_t seq_bound;
    int seq_stop;
    int q1count;
    int qcount;
};
static_assert(sizeof(RsSmem) <= 227 * 1024, "RsSmem must fit one CTA's shared memory");
static_assert(RS_Q1CAP * 4 >= 5 * RS_THREADS, "queue 1 doubles as partial counts + round results of the confidence mode");

// a4 / LRF-vote scoring keeps the smaller 256-thread shape (per-correspondence thresholds in a 4th float4)
constexpr int SC_THREADS = 256;
constexpr int SC_CHUNK = 2048;
struct __align__(16) ScSmem {
    float4 chunk[SC_CHUNK / 2][4];              // ... + (w w' - -): per-correspondence squared-distance thresholds
    unsigned long long red[SC_THREADS / 32];
};

// cooperative load of correspondences [c0, c0 + CHUNK) of one pair (8-float records) into the pair-interleaved layout
template <int NF, int CHUNK, int THREADS>
BFR_DEVINL void load_chunk(float4 (*chunk)[NF], const float4* __restrict__ corr, int K, int c0)
{
    for (int g = threadIdx.x; g < CHUNK / 2; g += THREADS) {
        const int c = c0 + 2 * g;
        if (c >= K) break;                                              // pairs beyond the end are never read
        const float4 a0 = __ldg(&corr[2 * (size_t)c]), b0 = __ldg(&corr[2 * (size_t)c + 1]);
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = make_float4(1e18f, 1e18f, 1e18f, 0.f);   // odd tail: a far-away dummy never counts
        if (c + 1 < K) { a1 = __ldg(&corr[2 * (size_t)c + 2]); b1 = __ldg(&corr[2 * (size_t)c + 3]); }
        chunk[g][0] = make_float4(a0.x, a1.x, a0.y, a1.y);
        chunk[g][1] = make_float4(a0.z, a1.z, b0.x, b1.x);
        chunk[g][2] = make_float4(b0.y, b1.y, b0.z, b1.z);
        if (NF == 4) chunk[g][NF - 1] = make_float4(a0.w, a1.w, 0.f, 0.f);
    }
}

// ++count iff d < thr, as one FSETP + one predicated IADD (the compiler's select form costs an extra add per test)
BFR_DEVINL void count_if_lt(int& count, float d, float thr)
{
    asm("{\n\t.reg .pred q;\n\tsetp.lt.f32 q, %1, %2;\n\t@q add.s32 %0, %0, 1;\n\t}" : "+r"(count) : "f"(d), "f"(thr));
}

// inlier count of one hypothesis over correspondence pairs [g_begin, g_end) of the chunk currently in shared memory
template <int NF, bool PER_CORR_THR>
BFR_DEVINL int score_pairs(const float4 (*chunk)[NF], int g_begin, int g_end, const float R[9], const float t[3], float d2max)
{
    f32x2 Rb[9], tb[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rb[k] = pack2(R[k], R[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) tb[k] = pack2(t[k], t[k]);
    int count = 0;
#pragma unroll 4
    for (int g = g_begin; g < g_end; ++g) {
        const float4 L0 = chunk[g][0], L1 = chunk[g][1], L2 = chunk[g][2];     // warp-uniform -> broadcast
        const f32x2 sx = pack2(L0.x, L0.y), sy = pack2(L0.z, L0.w), sz = pack2(L1.x, L1.y);
        const f32x2 qx = pack2(L1.z, L1.w), qy = pack2(L2.x, L2.y), qz = pack2(L2.z, L2.w);
        const f32x2 x = sub2(fma2(Rb[0], sx, fma2(Rb[1], sy, fma2(Rb[2], sz, tb[0]))), qx);
        const f32x2 y = sub2(fma2(Rb[3], sx, fma2(Rb[4], sy, fma2(Rb[5], sz, tb[1]))), qy);
        const f32x2 z = sub2(fma2(Rb[6], sx, fma2(Rb[7], sy, fma2(Rb[8], sz, tb[2]))), qz);
        const f32x2 d2 = fma2(x, x, fma2(y, y, mul2(z, z)));
        float da, db;
        unpack2(d2, da, db);
        if (PER_CORR_THR) {
            const float2 w = *reinterpret_cast<const float2*>(&chunk[g][NF - 1]);
            count_if_lt(count, da, w.x);
            count_if_lt(count, db, w.y);
        } else {
            count_if_lt(count, da, d2max);
            count_if_lt(count, db, d2max);
        }
    }
    return count;
}

// minimal-sample gather from the resident chunk: record c lives in pair g = c / 2, slot c % 2
BFR_DEVINL void load_sample_smem(const RsSmem& sm, const uint32_t id[3], float s[3][3], float q[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* f = reinterpret_cast<const float*>(&sm.chunk[id[i] >> 1][0]) + (id[i] & 1u);
        s[i][0] = f[0]; s[i][1] = f[2]; s[i][2] = f[4]; q[i][0] = f[6]; q[i][1] = f[8]; q[i][2] = f[10];
    }
}

template <bool RES>
BFR_DEVINL bool precheck(const RsSmem& sm, const float4* __restrict__ corr_p, uint32_t K, uint64_t seed, uint32_t pair_id, uint32_t h, float sim2, uint32_t id[3])
{
    sample3(seed, pair_id, h, K, id);
    if (id[0] == id[1] || id[0] == id[2] || id[1] == id[2]) return false;
    float s[3][3], q[3][3];
    if (RES) load_sample_smem(sm, id, s, q); else load_sample(corr_p, id, s, q);
    return edge_lengths_ok(s, q, sim2);
}

// score `n` queued hypotheses against all K correspondences.  A full queue (n = RS_THREADS) gives every thread one hypothesis.  A partial
// flush (n < RS_THREADS) would leave most warps idle, so the nw = ceil(n / 32) warps' worth of hypotheses are replicated over the
// RS_WARPS / nw groups of warps and every group scores its own slice of each chunk (the loads stay warp-uniform broadcasts); the partial
// counts are integer sums, so the total is exact whatever the split.  Returns true in the thread that owns queue entry hi (index h, count).
template <bool RES>
BFR_DEVINL bool score_queue(RsSmem& sm, const float4* __restrict__ corr, int K, int n, float d2max, uint32_t& h, int& count, int& hi)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (n + 31) >> 5;                                     // warps that hold hypotheses
    const int nparts = RS_WARPS / nw;                                 // correspondence slices (1 for a full queue)
    const int part = warp / nw;
    hi = (warp % nw) * 32 + lane;                                     // this thread: hypothesis hi of the queue, slice `part`
    const bool warp_has_work = part < nparts;
    const bool mine = warp_has_work && hi < n;
    float R[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, t[3] = { 0.f, 0.f, 0.f };
    h = 0;
    if (mine) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = sm.q[k][hi];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = sm.q[9 + k][hi];
        h = sm.qh[hi];
    }
    int* partial = reinterpret_cast<int*>(sm.q1);                     // queue 1 is empty whenever a partial flush runs
    if (nparts > 1) { partial[threadIdx.x] = 0; __syncthreads(); }
    count = 0;
    for (int c0 = 0; c0 < K; c0 += RS_CHUNK) {
        if (!RES) {
            __syncthreads();                   // previous chunk fully consumed
            load_chunk<3, RS_CHUNK, RS_THREADS>(sm.chunk, corr, K, c0);
            __syncthreads();
        }
        const int npairs = (min(RS_CHUNK, K - c0) + 1) >> 1;
        if (warp_has_work) count += score_pairs<3, false>(sm.chunk, (part * npairs) / nparts, ((part + 1) * npairs) / nparts, R, t, d2max);
    }
    if (nparts > 1) {
        if (mine) atomicAdd(&partial[hi], count);
        __syncthreads();
        count = partial[hi < RS_THREADS ? hi : 0];
        __syncthreads();                       // partial[] (= queue 1) may be refilled after this
    }
    return mine && part == 0;
}

// stage 2 on n queue-1 entries starting at `base` (thread i takes entry base + i): Kabsch + distance check on dense warps, survivors -> queue 2
template <bool RES>
BFR_DEVINL void fit_block(RsSmem& sm, const float4* __restrict__ corr_p, int base, int n, float d2max)
{
    const int lane = threadIdx.x & 31;
    float R[9], t[3];
    bool ok = false; uint32_t h = 0;
    if ((int)threadIdx.x < n) {
        const uint4 e = sm.q1[base + threadIdx.x];
        h = e.x;
        const uint32_t id[3] = { e.y, e.z, e.w };
        float s[3][3], q[3][3];
        if (RES) load_sample_smem(sm, id, s, q); else load_sample(corr_p, id, s, q);
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
}

BFR_DEVINL unsigned long long pack_count(int count, uint32_t h) { return ((unsigned long long)(uint32_t)count << 32) | (unsigned long long)(0xFFFFFFFFu - h); }

// one work item: hypotheses [hb, he) of pair p.  CONF: Open3D confidence rule (hb = the pair's first hypothesis, one item per pair).
template <bool RES, bool CONF>
BFR_DEVINL void ransac_item(RsSmem& sm, const float4* __restrict__ corr_p, int K, uint64_t seed, uint32_t pair_id, uint32_t hb, uint32_t he,
                            float d2max, float sim2, float confidence, unsigned long long& best, int& n_scored, long long (&tm)[3])
{
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { sm.qcount = 0; sm.q1count = 0; sm.seq_best = 0ull; sm.seq_bound = he - hb; sm.seq_stop = 0; }
    if (RES) load_chunk<3, RS_CHUNK, RS_THREADS>(sm.chunk, corr_p, K, 0);
    __syncthreads();

    if (!CONF) {
        // fit one block of queue 1, then score a full block of queue 2 if there is one
        auto fit_and_score = [&](int base, int n) {
            RST(tm[1], fit_block<RES>(sm, corr_p, base, n, d2max));
            __syncthreads();
            const int qn = sm.qcount;
            __syncthreads();
            if (qn >= RS_THREADS) {
                uint32_t h; int count, hi;
                bool mine;
                RST(tm[2], mine = score_queue<RES>(sm, corr_p, K, RS_THREADS, d2max, h, count, hi));
                if (mine) { const unsigned long long pk = pack_count(count, h); best = pk > best ? pk : best; }
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
#ifdef RS_TIMING
            const long long ts_ = clock64();
#endif
#pragma unroll
            for (int u = 0; u < RS_S1; ++u) {
                hh[u] = base + (uint32_t)u * RS_THREADS + threadIdx.x;
                id[u][0] = id[u][1] = id[u][2] = 0u;
                ok[u] = hh[u] < he && precheck<RES>(sm, corr_p, (uint32_t)K, seed, pair_id, hh[u], sim2, id[u]);
            }
#ifdef RS_TIMING
            tm[0] += clock64() - ts_;
#endif
#pragma unroll
            for (int u = 0; u < RS_S1; ++u) {                              // queue order is irrelevant: any order gives the same best
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
                for (; n1 - done >= RS_THREADS; done += RS_THREADS) fit_and_score(done, RS_THREADS);
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
            if (n1 > 0) fit_and_score(0, n1);
            if (threadIdx.x == 0) sm.q1count = 0;
            __syncthreads();
        }
        const int qn = sm.qcount;
        if (qn > 0) {
            uint32_t h; int count, hi;
            bool mine;
            RST(tm[2], mine = score_queue<RES>(sm, corr_p, K, qn, d2max, h, count, hi));
            if (mine) { const unsigned long long pk = pack_count(count, h); best = pk > best ? pk : best; }
            n_scored += qn;
        }
    } else {
        // Open3D's convergence criterion, replayed exactly: rounds of RS_THREADS hypotheses in index order; every valid hypothesis of a round
        // is scored, then the round's (iteration, count) list is sorted by iteration and walked sequentially: a hypothesis only counts if its
        // iteration number is still below the bound set by the best hypothesis BEFORE it (RANSACConvergenceCriteria, models/BUFFER.py:323-324)
        const double log_1m_conf = det_log(__dsub_rn(1.0, (double)confidence));
        const uint32_t max_iter = he - hb;
        uint32_t* res_h = reinterpret_cast<uint32_t*>(sm.q1) + 1 * RS_THREADS;       // ints [0, RS_THREADS) of q1 = partial counts of score_queue
        int* res_c = reinterpret_cast<int*>(sm.q1) + 2 * RS_THREADS;
        uint32_t* srt_h = reinterpret_cast<uint32_t*>(sm.q1) + 3 * RS_THREADS;
        int* srt_c = reinterpret_cast<int*>(sm.q1) + 4 * RS_THREADS;
        for (uint32_t base = hb; base < he; base += RS_THREADS) {
            {
                const uint32_t hh = base + threadIdx.x;
                uint32_t id[3] = { 0u, 0u, 0u };
                const bool ok = hh < he && precheck<RES>(sm, corr_p, (uint32_t)K, seed, pair_id, hh, sim2, id);
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (bal) {
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(&sm.q1count, __popc(bal));
                    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
                    if (ok) sm.q1[pos] = make_uint4(hh, id[0], id[1], id[2]);
                }
            }
            __syncthreads();
            const int n1 = sm.q1count;
            __syncthreads();
            if (n1 > 0) fit_block<RES>(sm, corr_p, 0, n1, d2max);
            __syncthreads();
            const int qn = sm.qcount;
            __syncthreads();
            if (threadIdx.x == 0) { sm.q1count = 0; sm.qcount = 0; }
            if (qn > 0) {
                uint32_t h; int count, hi;
                const bool mine = score_queue<RES>(sm, corr_p, K, qn, d2max, h, count, hi);
                n_scored += qn;
                if (mine) { res_h[hi] = h - hb; res_c[hi] = count; }          // iteration number relative to the pair's first hypothesis
                __syncthreads();
                if ((int)threadIdx.x < qn) {                                   // rank by iteration number (all distinct)
                    const uint32_t my = res_h[threadIdx.x];
                    int rank = 0;
                    for (int k = 0; k < qn; ++k) rank += (res_h[k] < my) ? 1 : 0;
                    srt_h[rank] = my; srt_c[rank] = res_c[threadIdx.x];
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    unsigned long long sb = sm.seq_best; uint32_t bound = sm.seq_bound;
                    for (int k = 0; k < qn; ++k) {
                        const uint32_t it = srt_h[k];
                        if (it >= bound) { sm.seq_stop = 1; break; }           // the sequential loop ended before this iteration
                        const unsigned long long pk = pack_count(srt_c[k], it + hb);
                        if (pk > sb) {
                            sb = pk;
                            const uint32_t nb = ransac_exit_bound((uint32_t)srt_c[k], (uint32_t)K, log_1m_conf, max_iter);
                            bound = nb < bound ? nb : bound;
                        }
                    }
                    sm.seq_best = sb; sm.seq_bound = bound;
                }
            }
            __syncthreads();
            if (sm.seq_stop || (base - hb) + RS_THREADS >= sm.seq_bound) break;   // block-uniform
        }
        __syncthreads();
        if (threadIdx.x == 0) best = sm.seq_best;
    }
}

__global__ void __launch_bounds__(RS_THREADS, 1)
ransac_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt, int P, int splits,
              uint64_t seed, uint32_t pair_id_base, uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, float confidence,
              unsigned long long* __restrict__ best_packed, int32_t* __restrict__ valid_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(smem_raw);
    const float d2max = __fmul_rn(dist_th, dist_th), sim2 = __fmul_rn(similar_th, similar_th);
    const bool conf = confidence > 0.0f && confidence < 1.0f;
    const uint32_t nh = h_end - h_begin;
    const int lane = threadIdx.x & 31;
    long long tm[3] = { 0, 0, 0 };                                    // RS_TIMING: stage 1 / fit / score cycles (dead code otherwise)
#ifdef RS_TIMING
    const long long t_begin = clock64();
    int scored_total = 0;
#endif
    for (int item = blockIdx.x; item < P * splits; item += gridDim.x) {
        const int p = item / splits, sp = item % splits;
        const int K = corr_cnt[p];
        if (K < 3) continue;
        const float4* corr_p = corr + 2 * (size_t)corr_off[p];
        const uint32_t hb = h_begin + (uint32_t)(((unsigned long long)sp * nh) / (unsigned)splits);
        const uint32_t he = h_begin + (uint32_t)(((unsigned long long)(sp + 1) * nh) / (unsigned)splits);
        const uint32_t pair_id = pair_id_base + (uint32_t)p;
        unsigned long long best = 0ull;
        int n_scored = 0;                                             // hypotheses that passed every check (thread 0's tally)
        __syncthreads();                                              // the previous item is done with shared memory
        if (K <= RS_CHUNK) {
            if (conf) ransac_item<true, true>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, best, n_scored, tm);
            else      ransac_item<true, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, best, n_scored, tm);
        } else {
            if (conf) ransac_item<false, true>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, best, n_scored, tm);
            else      ransac_item<false, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, best, n_scored, tm);
        }
        if (valid_count && threadIdx.x == 0 && n_scored) atomicAdd(valid_count + p, n_scored);
#ifdef RS_TIMING
        scored_total += n_scored;
#endif
        // block max -> one atomic
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
        __syncthreads();
        if (lane == 0) sm.red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0ull;
            for (int w = 0; w < RS_WARPS; ++w) b = sm.red[w] > b ? sm.red[w] : b;
            if (b) atomicMax(best_packed + p, b);
        }
    }
#ifdef RS_TIMING
    if (threadIdx.x == 0) {
        atomicAdd(&g_rs_dbg[0], (unsigned long long)(clock64() - t_begin)); atomicAdd(&g_rs_dbg[1], (unsigned long long)tm[2]);
        atomicAdd(&g_rs_dbg[2], (unsigned long long)(tm[1] + tm[2])); atomicAdd(&g_rs_dbg[3], 1ull); atomicAdd(&g_rs_dbg[4], (unsigned long long)scored_total);
        atomicAdd(&g_rs_dbg[5], (unsigned long long)tm[0]);
    }
#endif
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
// R = tt_R Rz(c, s) ss_R^T, t = q - R p
BFR_DEVINL void lrf_pose(float c, float s, const float Rt[9], const float Rs[9], const float p[3], const float q[3], float R[9], float t[3])
{
    float Mx[9];
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
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = __fsub_rn(q[r], __fmaf_rn(R[3 * r + 2], p[2], __fmaf_rn(R[3 * r + 1], p[1], __fmul_rn(R[3 * r + 0], p[0]))));
}

__global__ void lrf_hypotheses_kernel(const float* __restrict__ cs, const float* __restrict__ ss_R, const float* __restrict__ tt_R,
                                      const float* __restrict__ ss_kpts, const float* __restrict__ tt_kpts, int A, float* __restrict__ R_out, float* __restrict__ t_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    float Rt[9], Rs[9], R[9], t[3], p[3], q[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rt[k] = tt_R[9 * (size_t)i + k]; Rs[k] = ss_R[9 * (size_t)i + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = ss_kpts[3 * (size_t)i + k]; q[k] = tt_kpts[3 * (size_t)i + k]; }
    lrf_pose(cs[2 * i], cs[2 * i + 1], Rt, Rs, p, q, R, t);
#pragma unroll
    for (int k = 0; k < 9; ++k) R_out[9 * (size_t)i + k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t_out[3 * (size_t)i + k] = t[k];
}

// ---- a4: score explicit hypotheses (models/BUFFER.py:303-311) ----------------------------------------------------
// records for the scoring core: {sx sy sz T | qx qy qz 0}, T = sqrt_threshold(thr): d2 < T  <=>  sqrt(d2) < thr (the reference's test)
__global__ void make_records_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const float* __restrict__ thr, float thr_scalar, int C, float4* __restrict__ rec)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float th = thr ? thr[c] : thr_scalar;
    rec[2 * (size_t)c] = make_float4(src[3 * (size_t)c], src[3 * (size_t)c + 1], src[3 * (size_t)c + 2], sqrt_threshold(th));
    rec[2 * (size_t)c + 1] = make_float4(tgt[3 * (size_t)c], tgt[3 * (size_t)c + 1], tgt[3 * (size_t)c + 2], 0.0f);
}

// block max of (count + 1) << 32 | ~h  ->  one atomicMax (+1 so that an all-zero-count winner differs from "empty"; torch.argmax -> first maximum)
BFR_DEVINL void publish_best(ScSmem& sm, bool mine, int count, uint32_t h, unsigned long long* best_packed)
{
    unsigned long long best = mine ? pack_count(count + 1, h) : 0ull;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    if (lane == 0) sm.red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = 0ull;
        for (int w = 0; w < SC_THREADS / 32; ++w) b = sm.red[w] > b ? sm.red[w] : b;
        if (b) atomicMax(best_packed, b);
    }
}

__global__ void __launch_bounds__(SC_THREADS, 2)
score_hypotheses_kernel(const float* __restrict__ Rh, const float* __restrict__ th, int H, const float4* __restrict__ rec, int C,
                        int32_t* __restrict__ counts, unsigned long long* __restrict__ best_packed)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScSmem& sm = *reinterpret_cast<ScSmem*>(smem_raw);
    const int h = blockIdx.x * SC_THREADS + threadIdx.x;
    const bool mine = h < H;
    float R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = mine ? Rh[9 * (size_t)h + k] : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = mine ? th[3 * (size_t)h + k] : 0.0f;
    const bool warp_has_work = (int)(blockIdx.x * SC_THREADS + (threadIdx.x & ~31u)) < H;
    int count = 0;
    for (int c0 = 0; c0 < C; c0 += SC_CHUNK) {
        __syncthreads();
        load_chunk<4, SC_CHUNK, SC_THREADS>(sm.chunk, rec, C, c0);
        __syncthreads();
        const int npairs = (min(SC_CHUNK, C - c0) + 1) >> 1;
        if (warp_has_work) count += score_pairs<4, true>(sm.chunk, 0, npairs, R, t, 0.0f);
    }
    if (mine && counts) counts[h] = count;
    publish_best(sm, mine, count, (uint32_t)h, best_packed);
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

// ---- a3 + a4 fused, batched over pairs (models/BUFFER.py:294-311): the LRF vote ---------------------------------------------
// Correspondence c of a pair proposes R_c = tt_R[c] Rz(angle_c) ss_R[c]^T, t_c = q_c - R_c s_c with angle_c = ind_c 2 pi / azi_n + 1e-6;
// every proposal is scored on all correspondences of the pair against the per-correspondence threshold |s_c| pi / azi_n inlier_th; the
// first proposal with the most inliers wins and its inliers, compacted in order, are the `corr` that the reference hands to RANSAC
// (:311-316).  R and t only ever exist in registers; nothing goes to the host.
BFR_DEVINL float lrf_angle(float ind, float azi_n)
{   // torch float32 evaluation order of `ind * 2 * np.pi / azi_n + 1e-6` (:295)
    return __fadd_rn(__fdiv_rn(__fmul_rn(__fmul_rn(ind, 2.0f), 3.14159274101257324f), azi_n), 1e-6f);
}
BFR_DEVINL float vote_threshold(float sx, float sy, float sz, float azi_n, float inlier_th)
{   // sqrt(sum(ss^2)) * np.pi / azi_n * inlier_th in float32, left to right (:306-307)
    const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz)));
    return __fmul_rn(__fdiv_rn(__fmul_rn(n, 3.14159274101257324f), azi_n), inlier_th);
}

// thresholds into the 4th float of every record (in place, once per call)
__global__ void vote_thresholds_kernel(float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt, float azi_n, float inlier_th)
{
    const int p = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= corr_cnt[p]) return;
    float4* r = corr + 2 * ((size_t)corr_off[p] + c);
    float4 a = r[0];
    a.w = sqrt_threshold(vote_threshold(a.x, a.y, a.z, azi_n, inlier_th));
    r[0] = a;
}

BFR_DEVINL void vote_hypothesis(const float4* __restrict__ corr_p, const float* __restrict__ ind, const float* __restrict__ ss_R, const float* __restrict__ tt_R,
                                size_t row, int h, float azi_n, float R[9], float t[3])
{
    float Rt[9], Rs[9], sn, cs;
#pragma unroll
    for (int k = 0; k < 9; ++k) { Rt[k] = tt_R[9 * (row + h) + k]; Rs[k] = ss_R[9 * (row + h) + k]; }
    const float4 a = corr_p[2 * (size_t)h], b = corr_p[2 * (size_t)h + 1];
    const float p[3] = { a.x, a.y, a.z }, q[3] = { b.x, b.y, b.z };
    det_sincos(lrf_angle(ind[row + h], azi_n), sn, cs);
    lrf_pose(cs, sn, Rt, Rs, p, q, R, t);
}

__global__ void __launch_bounds__(SC_THREADS, 2)
lrf_vote_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt,
                const float* __restrict__ ind, const float* __restrict__ ss_R, const float* __restrict__ tt_R, float azi_n,
                int32_t* __restrict__ counts, unsigned long long* __restrict__ best_packed)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScSmem& sm = *reinterpret_cast<ScSmem*>(smem_raw);
    const int p = blockIdx.y;
    const int A = corr_cnt[p];
    if ((int)(blockIdx.x * SC_THREADS) >= A) return;
    const size_t row = (size_t)corr_off[p];
    const float4* corr_p = corr + 2 * row;
    const int h = blockIdx.x * SC_THREADS + threadIdx.x;
    const bool mine = h < A;
    float R[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, t[3] = { 0.f, 0.f, 0.f };
    if (mine) vote_hypothesis(corr_p, ind, ss_R, tt_R, row, h, azi_n, R, t);
    const bool warp_has_work = (int)(blockIdx.x * SC_THREADS + (threadIdx.x & ~31u)) < A;
    int count = 0;
    for (int c0 = 0; c0 < A; c0 += SC_CHUNK) {
        __syncthreads();
        load_chunk<4, SC_CHUNK, SC_THREADS>(sm.chunk, corr_p, A, c0);
        __syncthreads();
        const int npairs = (min(SC_CHUNK, A - c0) + 1) >> 1;
        if (warp_has_work) count += score_pairs<4, true>(sm.chunk, 0, npairs, R, t, 0.0f);
    }
    if (mine && counts) counts[row + h] = count;
    publish_best(sm, mine, count, (uint32_t)h, best_packed + p);
}

// inliers of the winning proposal, compacted in ascending order into sub_corr (records at the pair's offset) -> the RANSAC input
__global__ void __launch_bounds__(SC_THREADS)
lrf_vote_select_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt,
                       const float* __restrict__ ind, const float* __restrict__ ss_R, const float* __restrict__ tt_R, float azi_n,
                       const unsigned long long* __restrict__ best_packed, float4* __restrict__ sub_corr, int32_t* __restrict__ sub_cnt,
                       int64_t* __restrict__ best_idx, int64_t* __restrict__ inlier_ind)
{
    __shared__ int warp_cnt[SC_THREADS / 32];
    __shared__ int base_s;
    const int p = blockIdx.x;
    const int A = corr_cnt[p];
    const size_t row = (size_t)corr_off[p];
    const float4* corr_p = corr + 2 * row;
    const unsigned long long b = best_packed[p];
    if (b == 0ull || A <= 0) {
        if (threadIdx.x == 0) { sub_cnt[p] = 0; if (best_idx) best_idx[p] = -1; }
        return;
    }
    const int hbest = (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull));
    float R[9], t[3];
    vote_hypothesis(corr_p, ind, ss_R, tt_R, row, hbest, azi_n, R, t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < A; c0 += SC_THREADS) {
        const int c = c0 + threadIdx.x;
        bool flag = false; float4 a = make_float4(0.f, 0.f, 0.f, 0.f), q = a;
        if (c < A) { a = corr_p[2 * (size_t)c]; q = corr_p[2 * (size_t)c + 1]; flag = resid2(R, t, a.x, a.y, a.z, q.x, q.y, q.z) < a.w; }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int pre = base_s, tot = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) { const int n = warp_cnt[w]; if (w < warp) pre += n; tot += n; }
        if (flag) {
            const int pos = pre + __popc(bal & ((1u << lane) - 1u));
            sub_corr[2 * (row + pos)] = make_float4(a.x, a.y, a.z, 0.0f);
            sub_corr[2 * (row + pos) + 1] = q;
            if (inlier_ind) inlier_ind[row + pos] = (int64_t)c;
        }
        __syncthreads();
        if (threadIdx.x == 0) base_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { sub_cnt[p] = base_s; if (best_idx) best_idx[p] = (int64_t)hbest; }
}

#ifdef RS_TIMING
}
extern "C" __attribute__((visibility("default"))) void bfr_dbg_ransac_counters(unsigned long long* out) { cudaMemcpyFromSymbol(out, bfr::g_rs_dbg, 64); unsigned long long z[8] = {0}; cudaMemcpyToSymbol(bfr::g_rs_dbg, z, 64); }
namespace bfr {
#endif
// ---- host launchers -------------------------------------------------------------------------------------------
static int sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int v = cached[dev & 63].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

cudaError_t ransac_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                          uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, float confidence, int splits,
                          unsigned long long* best_packed, int32_t* valid_count, cudaStream_t stream)
{
    static std::atomic<unsigned long long> attr_done{0};
    cudaError_t e = ensure_dyn_smem((const void*)ransac_kernel, (int)sizeof(RsSmem), attr_done);
    if (e != cudaSuccess) return e;
    if (P <= 0 || h_end <= h_begin) return cudaSuccess;
    const bool conf = confidence > 0.0f && confidence < 1.0f;
    if (splits < 1 || conf) splits = 1;                               // the convergence rule is sequential in the hypothesis index: one CTA per pair
    const long long items = (long long)P * splits;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(items < sms ? items : sms);      // persistent: one 512-thread CTA per SM, items dealt round-robin
    ransac_kernel<<<grid, RS_THREADS, sizeof(RsSmem), stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, P, splits, seed, pair_id_base,
                                                                h_begin, h_end, dist_th, similar_th, confidence, best_packed, valid_count);
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
    static std::atomic<unsigned long long> attr_done{0};
    cudaError_t e = ensure_dyn_smem((const void*)score_hypotheses_kernel, (int)sizeof(ScSmem), attr_done);
    if (e != cudaSuccess) return e;
    float4* rec = reinterpret_cast<float4*>(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    e = cudaMemsetAsync(best_packed, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (C > 0) make_records_kernel<<<(C + 255) / 256, 256, 0, stream>>>(src, tgt, thr, thr_scalar, C, rec);
    if (H > 0) score_hypotheses_kernel<<<(H + SC_THREADS - 1) / SC_THREADS, SC_THREADS, sizeof(ScSmem), stream>>>(R, t, H, rec, C, counts, best_packed);
    score_mask_kernel<<<(max(C, 1) + 255) / 256, 256, 0, stream>>>(R, t, rec, C, best_packed, mask, best_idx);
    return cudaGetLastError();
}

// the LRF vote for P pairs: corr = all mutual matches (records; their 4th float is overwritten with the vote threshold), ind / ss_R / tt_R
// row-aligned with corr.  vote_best [P] is scratch (zeroed here); sub_corr gets each pair's inlier subset at the pair's offset.
cudaError_t lrf_vote_launch(float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, const float* ind, const float* ss_R, const float* tt_R,
                            float azi_n, float inlier_th, int32_t* counts, unsigned long long* vote_best, float* sub_corr, int32_t* sub_cnt,
                            int64_t* best_idx, int64_t* inlier_ind, cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    static std::atomic<unsigned long long> attr_done{0};
    cudaError_t e = ensure_dyn_smem((const void*)lrf_vote_kernel, (int)sizeof(ScSmem), attr_done);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(vote_best, 0, (size_t)P * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    const int nblk = ((max_count > 0 ? max_count : 1) + SC_THREADS - 1) / SC_THREADS;
    dim3 grid((unsigned)nblk, (unsigned)P);
    vote_thresholds_kernel<<<grid, SC_THREADS, 0, stream>>>(reinterpret_cast<float4*>(corr), corr_off, corr_cnt, azi_n, inlier_th);
    lrf_vote_kernel<<<grid, SC_THREADS, sizeof(ScSmem), stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, ind, ss_R, tt_R, azi_n, counts, vote_best);
    lrf_vote_select_kernel<<<P, SC_THREADS, 0, stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, ind, ss_R, tt_R, azi_n, vote_best,
                                                         reinterpret_cast<float4*>(sub_corr), sub_cnt, best_idx, inlier_ind);
    return cudaGetLastError();
}

}  // namespace bfr
