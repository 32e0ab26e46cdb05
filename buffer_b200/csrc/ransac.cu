// ransac.cu — K2 + K3: Philox hypothesis sampling, batched closed-form Kabsch, cheap checks, SE(3) inlier scoring
// with a packed max-reduction; plus the a3/a4 stage of BUFFER (LRF hypotheses + scoring) on the same scoring core.
//
// Replaces the Open3D 0.13 CPU RANSAC call of the reference (models/BUFFER.py:313-326) and the [A,A,3] broadcast
// scoring block (models/BUFFER.py:303-311).  Semantics and exact arithmetic: oracle/bfr_oracle.c (orc_hypothesis,
// orc_count_inliers, orc_ransac, orc_score_hypotheses, orc_lrf_vote); DESIGN.md §K2/§K3.
//
// ransac_kernel: persistent CTAs (512 worker threads + four tensor-core warps, one CTA per SM) walk work items = (pair, slice of the
// hypothesis range) round-robin.  A pair's correspondences (up to RS_CHUNK = 5120, 24 bytes each: 120 KB) are loaded ONCE per item into
// shared memory in a pair-interleaved layout and serve the random sample gathers of stage 1, the fits and the exact scoring loop; larger
// pairs stream through the same buffer in chunks and gather their samples from global memory.  Per round every thread draws RS_S1
// hypotheses (Philox counter = (h, pair_id, 0, 0)) and runs the cheap checks (repeated index, edge lengths); survivors (~10 %) are
// compacted into queue 1 so that the 3-point Kabsch + distance check runs on dense warps; what passes (a few %) goes to queue 2.
//
// Scoring (inlier count of every queued hypothesis over ALL correspondences of the pair) has two implementations with identical results:
//  * exact FP32 (score_queue): each thread takes one hypothesis; FFMA2 over two correspondences per instruction with warp-uniform
//    LDS.128 broadcasts; ~11 instructions per (hypothesis, correspondence).
//  * tensor-core filter (tc_flush, the default for shared-memory-resident pairs when the caller passes scratch memory): the residual
//    components x_i(h,c) = sum_j R_ij s_j + t_i - q_i are bilinear in (R_i, t_i | 1) and (s, 1, q), so a 128-correspondence x
//    20-hypothesis block of all three components is ONE 128 x 64 accumulator tile of two K = 16 tcgen05.mma kind::f16 on 2-level f16
//    splits of the FP32 operands (relative operand error 2^-24); four warp groups run four such pipelines side by side.  The epilogue (thread = correspondence, TMEM -> registers) needs three
//    FMAs, a sign-bit add and a band test per (h,c); only pairs whose approximate d^2 lies within a rigorous error band of the threshold
//    (none to a handful per pair) are re-evaluated with the oracle's FP32 chain, so the counts stay bit-exact.
// The inlier count never leaves the CTA; a CTA's best (count << 32 | ~h) goes out with one 64-bit atomicMax per item.
// confidence < 1 (Open3D's RANSACConvergenceCriteria, models/BUFFER.py:323-324): one item per pair, rounds of 512 hypotheses, and after
// every round the sequential rule of Open3D (stop once iteration >= ceil(log(1-c)/log(1-fitness^3)) of the best so far) is replayed
// in hypothesis order, so the result equals a one-thread sequential run (oracle orc_ransac_confidence) exactly.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include "bfr_tcgen05.cuh"
#include <cuda_fp16.h>
#include <cmath>

namespace bfr {

constexpr int RS_THREADS = 512;                 // worker threads: stage 1, fits, scoring / tensor-core epilogue
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_LAUNCH = RS_THREADS + 128;     // + four warps that feed the tensor core, one per epilogue warp group (20 warps: 96 registers, like 17 would be)
constexpr int RS_CHUNK = 5120;                  // correspondences per shared-memory chunk; pairs with K <= RS_CHUNK stay resident
#ifndef RS_S1_N
#define RS_S1_N 8
#endif
constexpr int RS_S1 = RS_S1_N;                     // stage-1 hypotheses per thread and round
constexpr int RS_QCAP = 2 * RS_THREADS;            // queue 2 (exact scoring): < RS_THREADS leftovers plus the survivors of one fit block
constexpr int RS_Q1CAP = ((1 + RS_S1) > 5 ? (1 + RS_S1) : 5) * RS_THREADS; // queue 1 holds < RS_THREADS leftovers plus one round's survivors (hypothesis index only)
constexpr int RS_MAX_CTAS = 160;                   // upper bound of the persistent grid (sizes the tensor-core scratch)

// tensor-core scoring filter
constexpr int RT_GROUPS = RS_WARPS / 4;         // warp groups of the epilogue: 4 warps = the 128 TMEM lanes (correspondences) of a tile
constexpr int RT_EPI_WARPS = 4 * RT_GROUPS;     // epilogue warps (any further worker warps sit a flush out)
constexpr int RT_HT = 20;                       // hypotheses per accumulator tile: 2 chunks of 32 TMEM columns = 5 hypothesis pairs x (x x' y y' z z') + 2 unused
constexpr int RT_FLUSH = RT_GROUPS * RT_HT;     // hypotheses per flush: every group ping-pongs between two 128 x 64 accumulator tiles (4 x 2 x 64 = 512 TMEM columns)
constexpr int RT_N = RT_GROUPS * 64;            // B rows of a flush (a group's MMAs read its 64-row slice)
constexpr int RT_QCAP = RT_FLUSH + RS_THREADS;  // queue 2 (tensor-core scoring): < RT_FLUSH leftovers plus the survivors of one fit block
constexpr int RT_TILE = 128;                    // correspondences per A tile (MMA M)
constexpr int RT_TILE_BYTES = RT_TILE * 32;     // K = 16 f16 per row
constexpr int RT_STAGES = 8;                    // A tiles in shared memory: two super-stages of RT_SUPER tiles
constexpr int RT_SUPER = 4;                     // A tiles per TMA copy and per full / empty barrier pair
constexpr int RT_MAX_TILES = RS_CHUNK / RT_TILE;
constexpr int RT_MIN_TC = 48;                    // smaller flushes (the tail of an item) are cheaper on the exact FP32 loop than 40 tile hand-offs
constexpr float RT_RANGE = 8192.0f;             // largest |s|_1, |q_i|, |t_i| the f16 splits are used for (beyond: exact FP32 scoring)
constexpr float RT_PAD_Q = 32768.0f;            // target coordinate of the padding rows of the last A tile: never an inlier, never in the band

#ifdef RS_TRACE
constexpr int RTR_EV = 8, RTR_TILES = 12;
__device__ unsigned int g_trace[20 * RTR_EV * RTR_TILES];      // clocks of lane 0 of every warp of CTA 0 during one flush: [warp][tile][event]
#define RTR(ev, tile) { if (trace_on && lane == 0 && (tile) < RTR_TILES) sm.trace[warp][(tile) * RTR_EV + (ev)] = (unsigned)clock64(); }
#else
#define RTR(ev, tile) { }
#endif
#ifdef RS_TIMING
__device__ unsigned long long g_rs_dbg[8];
__device__ unsigned long long g_rt_dbg[8];      // tc_flush, summed over thread 0 of every CTA: [0] B operands + bounds, [1] the tile loop, [4] count reduction, [5] flushes, [6] A tiles
#define RST(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#else
#define RST(acc, stmt) { stmt; }
#endif

struct __align__(1024) RsSmem {
    float4 chunk[RS_CHUNK / 2][3];              // per pair of correspondences: (sx sx' sy sy')(sz sz' qx qx')(qy qy' qz qz')
    union {
        struct {                                // exact scoring
            float q[12][RS_QCAP];               // queue 2: hypotheses that passed every check: R (9) + t (3), SoA
            uint32_t qh[RS_QCAP];
        } ex;
        struct {                                // tensor-core scoring
            unsigned char a_ring[RT_STAGES][RT_TILE_BYTES];   // A tiles: 128 correspondences x 16 f16, unswizzled K-major core matrices
            unsigned char b_op[2][RT_N * 32];                 // [B1 | B2]: 256 (hypothesis, component) rows x 16 f16, group g = rows 64 g ..
            float q[12][RT_QCAP];
            uint32_t qh[RT_QCAP];
            int wcnt[RS_WARPS][RT_HT];                        // per-warp inlier counts of the group's hypotheses (current flush)
        } tc;
    } u;
    uint32_t q1[RS_Q1CAP];                      // queue 1: survivors of the cheap checks (hypothesis index); idle: partial counts / round results
    unsigned long long red[RS_WARPS];
    uint64_t a_full[RT_STAGES], a_empty[RT_STAGES], acc_empty[RT_GROUPS][2], acc_full[RT_GROUPS][2];     // acc_*[warp group][accumulator buffer]
    uint64_t flush_go;                          // one arrival per flush (or at the end of the kernel): the tensor-core warps may read tc_ntiles and the B operands
    int tc_ntiles;                              // A tiles of the current item; < 0: the tensor-core warps leave
    unsigned long long seq_best;                // confidence mode: state of the sequential replay
    uint32_t seq_bound;
    int seq_stop;
    int q1count;
    int qcount;
    uint32_t tmem_base;
#ifdef RS_TRACE
    unsigned int trace[20][8 * 12];
#endif
    uint32_t stat[4];                           // float bits: max |s|_1, max |q_i| of the pair, max |t_i| of the flush; [3] != 0: out of range / non-finite
};
static_assert(sizeof(RsSmem) <= 227 * 1024, "RsSmem must fit one CTA's shared memory");
static_assert(RS_Q1CAP >= 5 * RS_THREADS, "queue 1 doubles as partial counts + round results of the confidence mode");
static_assert(RT_TILE_BYTES >= 2 * RS_THREADS * (int)sizeof(int), "an A-ring stage doubles as scratch of the exact fallback of a flush");
static_assert(offsetof(RsSmem, u) % 1024 == 0, "operand tiles start on a 1 KB boundary");

// a4 / LRF-vote scoring keeps the smaller 256-thread shape (per-correspondence thresholds in a 4th float4)
constexpr int SC_THREADS = 256;
constexpr int SC_CHUNK = 2048;
struct __align__(16) ScSmem {
    float4 chunk[SC_CHUNK / 2][4];              // ... + (w w' - -): per-correspondence squared-distance thresholds
    unsigned long long red[SC_THREADS / 32];
};

// the kernel's shared memory, re-derived from the symbol so that out-of-line functions keep shared-state-space loads and stores too
extern __shared__ __align__(1024) unsigned char rs_smem_raw[];
BFR_DEVINL RsSmem& rs_smem() { return *reinterpret_cast<RsSmem*>(rs_smem_raw); }

// barrier of the worker threads (the tensor-core warps are not part of it; a flush is handed to them through the flush_go mbarrier)
BFR_DEVINL void rs_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RS_THREADS) : "memory"); }

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

// one record of the resident chunk: record c lives in pair g = c / 2, slot c % 2
BFR_DEVINL void load_record_smem(const RsSmem& sm, uint32_t c, float s[3], float q[3])
{
    const float* f = reinterpret_cast<const float*>(&sm.chunk[c >> 1][0]) + (c & 1u);
    s[0] = f[0]; s[1] = f[2]; s[2] = f[4]; q[0] = f[6]; q[1] = f[8]; q[2] = f[10];
}
BFR_DEVINL void load_sample_smem(const RsSmem& sm, const uint32_t id[3], float s[3][3], float q[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) load_record_smem(sm, id[i], s[i], q[i]);
}

// Stage 1 of one hypothesis.  RES (samples in shared memory): branch-free - the repeated-index test, all three gathers and all three edge
// tests are always evaluated and combined at the end (same comparisons as edge_lengths_ok, so the same verdict), which lets the RS_S1
// hypotheses of a thread interleave: with early exits they ran one after another, each waiting for its own Philox chain and gathers.
template <bool RES>
BFR_DEVINL bool precheck(const RsSmem& sm, const float4* __restrict__ corr_p, uint32_t K, uint64_t seed, uint32_t pair_id, uint32_t h, float sim2)
{
    uint32_t id[3];
    sample3(seed, pair_id, h, K, id);
#ifndef RS_S1_BRANCHY
    if (RES) {
        float s[3][3], q[3][3];
        load_sample_smem(sm, id, s, q);
        bool ok = (id[0] != id[1]) & (id[0] != id[2]) & (id[1] != id[2]);
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int a = (e == 2) ? 1 : 0, b = (e == 0) ? 1 : 2;
            float dx = __fsub_rn(s[a][0], s[b][0]), dy = __fsub_rn(s[a][1], s[b][1]), dz = __fsub_rn(s[a][2], s[b][2]);
            const float ds2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
            dx = __fsub_rn(q[a][0], q[b][0]); dy = __fsub_rn(q[a][1], q[b][1]); dz = __fsub_rn(q[a][2], q[b][2]);
            const float dt2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
            ok &= !((ds2 < __fmul_rn(dt2, sim2)) | (dt2 < __fmul_rn(ds2, sim2)));
        }
        return ok;
    }
#endif
    if (id[0] == id[1] || id[0] == id[2] || id[1] == id[2]) return false;
    float s[3][3], q[3][3];
    if (RES) load_sample_smem(sm, id, s, q); else load_sample(corr_p, id, s, q);
    return edge_lengths_ok(s, q, sim2);
}

template <bool TC> BFR_DEVINL float& q2(RsSmem& sm, int k, int i) { if (TC) return sm.u.tc.q[k][i]; else return sm.u.ex.q[k][i]; }
template <bool TC> BFR_DEVINL uint32_t& q2h(RsSmem& sm, int i) { if (TC) return sm.u.tc.qh[i]; else return sm.u.ex.qh[i]; }

// exact FP32 scoring of the `n` queued hypotheses [base, base + n) against all K correspondences.  A full block (n = RS_THREADS) gives
// every thread one hypothesis.  With n < RS_THREADS most warps would idle, so the nw = ceil(n / 32) warps' worth of hypotheses are
// replicated over the RS_WARPS / nw groups of warps and every group scores its own slice of each chunk (the loads stay warp-uniform
// broadcasts); the partial counts are integer sums, so the total is exact whatever the split.  Returns true in the thread that owns queue
// entry base + hi (index h, count).  `partial`: RS_THREADS ints of scratch (only touched when n < RS_THREADS).
template <bool RES, bool TC>
BFR_DEVINL bool score_queue(RsSmem& sm, const float4* __restrict__ corr, int K, int base, int n, float d2max, int* partial, uint32_t& h, int& count, int& hi)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (n + 31) >> 5;                                     // warps that hold hypotheses
    const int nparts = RS_WARPS / nw;                                 // correspondence slices (1 for a full queue)
    const int part = warp / nw;
    hi = (warp % nw) * 32 + lane;                                     // this thread: hypothesis hi of the block, slice `part`
    const bool warp_has_work = part < nparts;
    const bool mine = warp_has_work && hi < n;
    float R[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, t[3] = { 0.f, 0.f, 0.f };
    h = 0;
    if (mine) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = q2<TC>(sm, k, base + hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = q2<TC>(sm, 9 + k, base + hi);
        h = q2h<TC>(sm, base + hi);
    }
    if (nparts > 1) { partial[threadIdx.x] = 0; rs_sync(); }
    count = 0;
    for (int c0 = 0; c0 < K; c0 += RS_CHUNK) {
        if (!RES) {
            rs_sync();                         // previous chunk fully consumed
            load_chunk<3, RS_CHUNK, RS_THREADS>(sm.chunk, corr, K, c0);
            rs_sync();
        }
        const int npairs = (min(RS_CHUNK, K - c0) + 1) >> 1;
        if (warp_has_work) count += score_pairs<3, false>(sm.chunk, (part * npairs) / nparts, ((part + 1) * npairs) / nparts, R, t, d2max);
    }
    if (nparts > 1) {
        if (mine) atomicAdd(&partial[hi], count);
        rs_sync();
        count = partial[hi < RS_THREADS ? hi : 0];
        rs_sync();                             // partial[] (= queue 1) may be refilled after this
    }
    return mine && part == 0;
}

BFR_DEVINL unsigned long long pack_count(int count, uint32_t h) { return ((unsigned long long)(uint32_t)count << 32) | (unsigned long long)(0xFFFFFFFFu - h); }

// ---- tensor-core scoring filter ----------------------------------------------------------------------------------------------------
// Operand rows (16 f16 = 32 bytes, K-major, unswizzled: 8 x 16-byte core matrices, the two K halves 128 B apart, 8-row groups 256 B apart):
//   A row of correspondence c:        [s_hi(3) s_lo(3) 1 q_hi(3) q_lo(3) 0 0 0]                    hi = rn_f16(v), lo = rn_f16(v - hi)
//   B1 row of (hypothesis h, comp i): [R_i,hi(3) R_i,hi(3) t_i,hi -e_i(3) -e_i(3) 0 0 0]
//   B2 row:                           [R_i,lo(3) 0 0 0 t_i,lo 0 ...]
// A B1^T + A B2^T = x_i(h,c) up to |error| <= 3 * 2^-24 * B' from the splits (B' = sum_j |R_ij s_j| + |t_i| + |q_i|) plus the
// accumulation error of the tensor core (measured: total 2^-21.6 B', tools/microbench/rs_mma.cu; budgeted here: 2^-19 B').
BFR_DEVINL uint32_t rt_row_offset(int r) { return (uint32_t)((r >> 3) * 256 + (r & 7) * 16); }     // first 16-byte half; the second is + 128
BFR_DEVINL void split_f16(float v, __half& hi, __half& lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(__fsub_rn(v, __half2float(hi)));
}
BFR_DEVINL uint32_t pack_h2(__half lo16, __half hi16) { return (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16); }
BFR_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
BFR_DEVINL float min3abs(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c))); return d; }

// Once per item: the pair's A tiles (f16 splits of every resident correspondence, padded to whole tiles) go to this CTA's scratch in
// global memory (it stays in L2; the tensor-core warps stream it back tile by tile with 1-D TMA copies, once per flush), and the
// pair's magnitude bounds are reduced.  Returns false if the pair must be scored exactly (coordinates out of range / non-finite).
BFR_DEVINL bool tc_prepare_pair(RsSmem& sm, int K, unsigned char* __restrict__ scratch, float& s1max, float& qmax)
{
    if (threadIdx.x < 4) sm.stat[threadIdx.x] = 0u;
    rs_sync();
    const int ntiles = (K + RT_TILE - 1) / RT_TILE;
    float s1 = 0.0f, qm = 0.0f; bool bad = false;
    const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
    for (int c = threadIdx.x; c < ntiles * RT_TILE; c += RS_THREADS) {
        float s[3] = { 0.f, 0.f, 0.f }, q[3] = { RT_PAD_Q, RT_PAD_Q, RT_PAD_Q };
        if (c < K) {
            load_record_smem(sm, (uint32_t)c, s, q);
            const float a = fabsf(s[0]) + fabsf(s[1]) + fabsf(s[2]), b = fmaxf(fmaxf(fabsf(q[0]), fabsf(q[1])), fabsf(q[2]));
            bad |= !(a <= RT_RANGE) || !(fabsf(q[0]) <= RT_RANGE) || !(fabsf(q[1]) <= RT_RANGE) || !(fabsf(q[2]) <= RT_RANGE);
            s1 = fmaxf(s1, a); qm = fmaxf(qm, b);
        }
        __half sh[3], sl[3], qh[3], ql[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { split_f16(s[k], sh[k], sl[k]); split_f16(q[k], qh[k], ql[k]); }
        const uint4 c0 = make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sl[0]), pack_h2(sl[1], sl[2]), pack_h2(one, qh[0]));
        const uint4 c1 = make_uint4(pack_h2(qh[1], qh[2]), pack_h2(ql[0], ql[1]), pack_h2(ql[2], zero), 0u);
        unsigned char* dst = scratch + (size_t)(c >> 7) * RT_TILE_BYTES + rt_row_offset(c & (RT_TILE - 1));
        *reinterpret_cast<uint4*>(dst) = c0;
        *reinterpret_cast<uint4*>(dst + 128) = c1;
    }
    s1 = warp_max(s1); qm = warp_max(qm);
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&sm.stat[0], __float_as_uint(s1)); atomicMax(&sm.stat[1], __float_as_uint(qm));      // non-negative floats order like their bits
        if (bad) sm.stat[3] = 1u;
    }
    __threadfence();
    fence_proxy_async_all();                    // the scratch is read by the async proxy (TMA)
    rs_sync();
    s1max = __uint_as_float(sm.stat[0]); qmax = __uint_as_float(sm.stat[1]);
    if (threadIdx.x == 0) sm.tc_ntiles = ntiles;
    return sm.stat[3] == 0u;
}

// Feeding the tensor core: one thread of the dedicated 13th warp.  Tile numbers `g` run on across flushes and items, so every mbarrier
// just keeps flipping phases.  Per flush it polls, round-robin over the warp groups, for "this group's next accumulator buffer has been
// handed back and the A tile has landed" and issues that tile's two MMAs (B1 then B2 accumulate; commit -> acc_full of the buffer and
// a_empty of the A stage), and keeps the TMA ring of A tiles full.  Everything on that path is precomputed: low words of the
// shared-memory descriptors (the high word is the same for all) and barrier addresses.
constexpr uint32_t RT_DESC_HI = 16u | (1u << 14);                    // SBO = 256 B (8-row groups), descriptor version 1, no swizzle
BFR_DEVINL uint32_t rt_desc_lo(const void* smem) { return ((smem_u32(smem) >> 4) & 0x3FFFu) | (8u << 16); }     // LBO = 128 B (K direction)
BFR_DEVINL void rt_issue_tile(uint32_t tmem_d, uint32_t a_lo, uint32_t b1_lo, uint32_t b2_lo, uint32_t bar_acc_full, uint32_t bar_a_empty)
{
    // instruction descriptor (kind::f16): D = F32, A = B = F16, both K-major, N = 64, M = 128
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n\t.reg .b64 da, db1, db2;\n\t.reg .pred pt, pf;\n\t"
                 "mov.b64 da, {%1, %4};\n\tmov.b64 db1, {%2, %4};\n\tmov.b64 db2, {%3, %4};\n\t"
                 "setp.eq.u32 pt, %0, %0;\n\tsetp.ne.u32 pf, %0, %0;\n\t"
                 "tcgen05.fence::after_thread_sync;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db1, %5, pf;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db2, %5, pt;\n\t"
                 "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b1_lo), "r"(b2_lo), "r"(RT_DESC_HI), "r"(idesc), "r"(bar_acc_full) : "memory");
    if (bar_a_empty)                                                  // last tile of an A super-stage: it is free once these (and all earlier) MMAs have completed
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a_empty) : "memory");
}
static_assert(RT_STAGES == 2 * RT_SUPER, "two super-stages");
BFR_DEVINL void tc_warp_loop(RsSmem& sm, const unsigned char* __restrict__ scratch, int grp)
{
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_grp = sm.tmem_base + (uint32_t)(grp * 128);
    const uint32_t a_lo0 = rt_desc_lo(sm.u.tc.a_ring[0]);
    const uint32_t b1_lo = rt_desc_lo(sm.u.tc.b_op[0]) + (uint32_t)(grp * (64 * 32 >> 4)), b2_lo = rt_desc_lo(sm.u.tc.b_op[1]) + (uint32_t)(grp * (64 * 32 >> 4));
    const uint32_t bar_af = smem_u32(&sm.a_full[0]), bar_ae = smem_u32(&sm.a_empty[0]), bar_cf = smem_u32(&sm.acc_full[grp][0]);
    const uint32_t ring = smem_u32(sm.u.tc.a_ring[0]);
    // The A tiles travel in super-tiles of RT_SUPER = 4 (one 16 KB TMA copy, one full / empty barrier pair per super-stage, two
    // super-stages): per tile the issuing thread then waits for nothing but the hand-back of its accumulator buffer and commits once -
    // every asynchronous operation on that path costs it ~100 cycles.
    uint32_t g0 = 0, G0 = 0;                                          // A tiles / super-tiles consumed before this flush
    for (uint32_t flush = 0;; ++flush) {
        mbar_wait(&sm.flush_go, flush & 1u);
        const int ntiles = *reinterpret_cast<volatile int*>(&sm.tc_ntiles);
        if (ntiles < 0) break;
        const int nst = (ntiles + RT_SUPER - 1) / RT_SUPER;
        if (lane == 0) {
            // super-tile j of this flush -> its super-stage, once every group's MMAs on the previous occupant have completed; the copies are
            // dealt out over the four warps (super-tile G by warp G % 4)
            auto load = [&](int j) {
                const uint32_t G = G0 + (uint32_t)j, ss = G & 1u;
                if (G >= 2u) mbar_wait(&sm.a_empty[ss], ((G >> 1) & 1u) ^ 1u);
                const uint32_t bytes = (uint32_t)min(RT_SUPER, ntiles - RT_SUPER * j) * RT_TILE_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_af + ss * 8u), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ring + ss * (RT_SUPER * RT_TILE_BYTES)), "l"(scratch + (size_t)j * (RT_SUPER * RT_TILE_BYTES)), "r"(bytes), "r"(bar_af + ss * 8u) : "memory");
            };
            if ((G0 & 3u) == (uint32_t)grp) load(0);
            if (nst > 1 && ((G0 + 1u) & 3u) == (uint32_t)grp) load(1);
            for (int i = 0; i < ntiles; ++i) {
                const int j = i / RT_SUPER, sub = i % RT_SUPER;
                const uint32_t G = G0 + (uint32_t)j, ss = G & 1u, gi = g0 + (uint32_t)i, buf = gi & 1u;
#ifdef RS_TRACE
                const bool trace_on = blockIdx.x == 0 && g0 >= 80u && g0 < 120u; const int warp = threadIdx.x >> 5;
#endif
                RTR(0, i);
                if (sub == 0) mbar_wait(&sm.a_full[ss], (G >> 1) & 1u);               // the super-tile has landed
                RTR(1, i);
                if (gi >= 2u) mbar_wait(&sm.acc_empty[grp][buf], ((gi >> 1) & 1u) ^ 1u);   // the group's four warps have pulled the buffer's previous tile out of TMEM
                RTR(2, i);
                const bool last = sub == RT_SUPER - 1 || i == ntiles - 1;
                rt_issue_tile(tmem_grp + buf * 64u, a_lo0 + (ss * RT_SUPER + (uint32_t)sub) * (RT_TILE_BYTES >> 4), b1_lo, b2_lo, bar_cf + buf * 8u, last ? bar_ae + ss * 8u : 0u);
                RTR(3, i);
                // behind the second issue of super-tile j: the copy of super-tile j + 1 into the other super-stage (free once every group is
                // done with super-tile j - 1, which by now they normally are; the slowest group's warp never waits here)
                if (sub == 1 && j + 1 < nst && j + 1 >= 2 && ((G + 1u) & 3u) == (uint32_t)grp) load(j + 1);
            }
        }
        G0 += (uint32_t)nst;
        g0 += (uint32_t)ntiles;
        __syncwarp();
    }
}

// Out of line (rare): correspondence c against the (up to) 10 hypotheses of one chunk, queue entries qbase .. qbase + 9 (the first
// `nvalid` exist); w[k] = approximate d2 - d2max.  Every pair inside the band is re-evaluated with the oracle's FP32 chain.
// Returns bit k: exact says inlier but the filter said no; bit 16 + k: the filter said inlier but exact says no.
__device__ __noinline__ uint32_t tc_recheck(uint32_t c, int qbase, int nvalid, float d2max, float band, const float (&w)[10])
{
    const RsSmem& sm = rs_smem();
    float s[3], q[3];
    load_record_smem(sm, c, s, q);
    uint32_t fix = 0u;
#pragma unroll 1
    for (int k = 0; k < 10; ++k) {
        const float wk = k == 0 ? w[0] : k == 1 ? w[1] : k == 2 ? w[2] : k == 3 ? w[3] : k == 4 ? w[4] : k == 5 ? w[5] : k == 6 ? w[6] : k == 7 ? w[7] : k == 8 ? w[8] : w[9];
        if (fabsf(wk) <= band && k < nvalid) {
            float R[9], tt[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) R[e] = sm.u.tc.q[e][qbase + k];
#pragma unroll
            for (int e = 0; e < 3; ++e) tt[e] = sm.u.tc.q[9 + e][qbase + k];
            const bool exact_in = resid2(R, tt, s[0], s[1], s[2], q[0], q[1], q[2]) < d2max;
            const bool approx_in = (__float_as_uint(wk) >> 31) != 0u;
            if (exact_in && !approx_in) fix |= 1u << k;
            if (!exact_in && approx_in) fix |= 1u << (16 + k);
        }
    }
    return fix;
}

// Score the queued hypotheses [base, base + n), n <= RT_FLUSH, on the tensor cores.  Warp group g (warps 4g .. 4g + 3 = the 128 TMEM lanes
// of a tile; thread = correspondence) owns flush-local hypotheses 20g .. 20g + 19: their 3 components are the 64 columns of the group's
// accumulator tiles (2 chunks of 32 columns; a chunk holds 10 hypotheses as 5 pairs (x_a x_b y_a y_b z_a z_b) so that one FFMA2 squares a
// component of two hypotheses).  Every group has its own pipeline: two ping-pong buffers, its own barriers and its own tensor-core warp.
// While the group works on the tile of A tile i, the tensor core fills the other buffer with A tile i + 1.  A warp pulls its 64 columns
// into registers, hands the buffer back (the group's tensor-core warp issues A tile i + 2 once all four warps have), and only then does
// the arithmetic.  `tile0` = A tiles consumed before this flush (the mbarrier phase clock).
// Returns, in threads < n, the exact inlier count of hypothesis base + threadIdx.x; bit 30 is set (in every thread) if the tensor-core path
// ran, i.e. ceil(K / 128) A tiles were consumed.  Out of line: ransac_item flushes from three places.
__device__ __noinline__ int tc_flush(int K, int base, int n, float d2max, float s1max, float qmax, uint32_t tile0)
{
    RsSmem& sm = rs_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef RS_TRACE
    const bool trace_on = blockIdx.x == 0 && tile0 >= 80u && tile0 < 120u;
#endif
#ifdef RS_TIMING
    const long long tq0 = clock64();
#endif
    // ---- B operands of the flush + the largest |t_i| ----
    float tm = 0.0f; bool bad = false;
    if ((int)threadIdx.x < 3 * n) {
        const int hl = (int)threadIdx.x / 3, comp = (int)threadIdx.x - 3 * hl, qi = base + hl;
        const float r0 = sm.u.tc.q[3 * comp][qi], r1 = sm.u.tc.q[3 * comp + 1][qi], r2 = sm.u.tc.q[3 * comp + 2][qi], tt = sm.u.tc.q[9 + comp][qi];
        tm = fabsf(tt);
        bad = !(tm <= RT_RANGE) || !(fabsf(r0) <= 1.001f) || !(fabsf(r1) <= 1.001f) || !(fabsf(r2) <= 1.001f);
        __half h0, l0, h1, l1, h2, l2, ht, lt;
        split_f16(r0, h0, l0); split_f16(r1, h1, l1); split_f16(r2, h2, l2); split_f16(tt, ht, lt);
        const __half zero = __float2half_rn(0.0f), m1 = __float2half_rn(-1.0f);
        const __half e0 = comp == 0 ? m1 : zero, e1 = comp == 1 ? m1 : zero, e2 = comp == 2 ? m1 : zero;
        const int grp = hl / RT_HT, hh = hl - grp * RT_HT, ch = hh / 10, k = hh - ch * 10;
        static_assert(RT_HT == 20, "two chunks of ten hypotheses per group");
        const int col = grp * 64 + ch * 32 + (k >> 1) * 6 + comp * 2 + (k & 1);
        unsigned char* b1 = sm.u.tc.b_op[0] + rt_row_offset(col);
        unsigned char* b2 = sm.u.tc.b_op[1] + rt_row_offset(col);
        *reinterpret_cast<uint4*>(b1) = make_uint4(pack_h2(h0, h1), pack_h2(h2, h0), pack_h2(h1, h2), pack_h2(ht, e0));
        *reinterpret_cast<uint4*>(b1 + 128) = make_uint4(pack_h2(e1, e2), pack_h2(e0, e1), pack_h2(e2, zero), 0u);
        *reinterpret_cast<uint4*>(b2) = make_uint4(pack_h2(l0, l1), pack_h2(l2, zero), 0u, pack_h2(lt, zero));
        *reinterpret_cast<uint4*>(b2 + 128) = make_uint4(0u, 0u, 0u, 0u);
    }
    tm = warp_max(tm);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { atomicMax(&sm.stat[2], __float_as_uint(tm)); if (bad) sm.stat[3] = 1u; }
    fence_proxy_async();                        // the operand rows are read by the async proxy (tcgen05.mma)
    rs_sync();
    const float tmax = __uint_as_float(sm.stat[2]);
    const bool fallback = sm.stat[3] != 0u;
    // |x~ - x_oracle| <= delta = (2^-19 + 2^-22) B' (tensor-core path + the four roundings of the FP32 chain), B' <= 1.001 |s|_1 + |t_i| + |q_i|;
    // then |d2~ - d2_oracle| <= delta (2 sqrt(3) sqrt(d2) + 3 delta) + 2^-20 d2max as long as d2 <= 2 d2max, and beyond 2 d2max the sign of
    // d2~ - d2max cannot flip if delta <= sqrt(d2max) / 64.  1.02: rounding of this very computation.
    const float bp = 1.001f * s1max + tmax + qmax;
    const float delta = 1.02f * 2.1457672e-6f * bp;
    const float rt = sqrtf(d2max);
    const float band = 1.02f * (delta * (4.8989795f * rt + 3.0f * delta) + 9.5367432e-7f * d2max);
    const bool ok = !fallback && n >= RT_MIN_TC && delta <= 0.015625f * rt && d2max > 0.0f && d2max <= 1.0e6f;
    if (!ok) {                                  // block-uniform: score these hypotheses exactly
        rs_sync();
        if (threadIdx.x == 0) { sm.stat[2] = 0u; sm.stat[3] = 0u; }
        uint32_t h; int count, hi;
        int* partial = reinterpret_cast<int*>(sm.u.tc.a_ring[0]);    // the ring is idle (queue 1 may still hold entries)
        int* res = partial + RS_THREADS;
        const bool mine = score_queue<true, true>(sm, nullptr, K, base, n, d2max, partial, h, count, hi);
        rs_sync();
        if (mine) res[hi] = count;
        rs_sync();
        const int r = (int)threadIdx.x < n ? res[threadIdx.x] : 0;
        rs_sync();
        return r;
    }
#ifdef RS_TIMING
    const long long tq1 = clock64();
#endif
    // ---- epilogue: thread = correspondence ----
    if (threadIdx.x == 0) mbar_arrive(&sm.flush_go);    // the tensor-core warps start on this flush (every worker is past the rs_sync above)
    const int grp = warp >> 2, qd = warp & 3;
    const int ntiles = (K + RT_TILE - 1) / RT_TILE;
    const uint32_t taddr = sm.tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(grp * 128);
    const f32x2 negT = pack2(-d2max, -d2max);
    const int hl0 = grp * RT_HT;                                      // flush-local index of this thread's first hypothesis
    // one 32-column chunk: 5 hypothesis pairs; sign bit of d2~ - d2max -> count, |.| -> band test
    auto process = [&](const float (&v)[32], int (&c10)[10], int i, int ch) {
        float m = 3.0e38f;
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
            const f32x2 x = pack2(v[6 * pr], v[6 * pr + 1]), y = pack2(v[6 * pr + 2], v[6 * pr + 3]), z = pack2(v[6 * pr + 4], v[6 * pr + 5]);
            const f32x2 d = fma2(x, x, fma2(y, y, fma2(z, z, negT)));                // d2~ - d2max of two hypotheses
            float wa, wb;
            unpack2(d, wa, wb);
            c10[2 * pr] += (int)(__float_as_uint(wa) >> 31);
            c10[2 * pr + 1] += (int)(__float_as_uint(wb) >> 31);
            m = min3abs(m, wa, wb);
        }
        if (m <= band) {                                             // rare: some (h, c) of this chunk is too close to call -> the oracle's chain decides
            const int c = i * RT_TILE + qd * 32 + lane;
            if (c < K) {
                float w[10];
#pragma unroll
                for (int pr = 0; pr < 5; ++pr) {
                    const f32x2 x = pack2(v[6 * pr], v[6 * pr + 1]), y = pack2(v[6 * pr + 2], v[6 * pr + 3]), z = pack2(v[6 * pr + 4], v[6 * pr + 5]);
                    unpack2(fma2(x, x, fma2(y, y, fma2(z, z, negT))), w[2 * pr], w[2 * pr + 1]);
                }
                const uint32_t fix = tc_recheck((uint32_t)c, base + hl0 + ch * 10, n - (hl0 + ch * 10), d2max, band, w);
#pragma unroll
                for (int k = 0; k < 10; ++k) c10[k] += (int)((fix >> k) & 1u) - (int)((fix >> (16 + k)) & 1u);
            }
        }
    };
    int c0[10], c1[10];                                               // inlier counts of this thread's 2 x 10 hypotheses over its correspondences
#pragma unroll
    for (int j = 0; j < 10; ++j) { c0[j] = 0; c1[j] = 0; }
    auto acquire = [&](int i) {                                       // the accumulator of A tile i is complete
        const uint32_t gi = tile0 + (uint32_t)i;
        mbar_wait(&sm.acc_full[grp][gi & 1u], (gi >> 1) & 1u);
        tc_fence_after();
        __syncwarp();
    };
    auto release = [&](int i) {                                       // all TMEM reads of A tile i have landed: the buffer goes back to the group's tensor-core warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.acc_empty[grp][(tile0 + (uint32_t)i) & 1u]);
    };
    auto col0 = [&](int i) { return taddr + ((tile0 + (uint32_t)i) & 1u) * 64u; };
    if (warp < RT_EPI_WARPS) {
        float va[32], vb[32];
        for (int i = 0; i < ntiles; ++i) {
            RTR(0, i);
            acquire(i);
            RTR(1, i);
            tmem_ld32_issue(col0(i), va); tmem_ld32_issue(col0(i) + 32u, vb);
            tmem_ld_wait(va); tmem_ld_pin(vb);
            RTR(2, i);
            release(i);
            RTR(4, i);
            process(va, c0, i, 0);
            process(vb, c1, i, 1);
            RTR(5, i);
        }
        // ---- totals over the 128 lanes x tiles ----
#pragma unroll
        for (int j = 0; j < RT_HT; ++j) {
            const int tot = __reduce_add_sync(0xffffffffu, j < 10 ? c0[j % 10] : c1[j % 10]);
            if (lane == j) sm.u.tc.wcnt[warp][j] = tot;
        }
    }
#ifdef RS_TIMING
    const long long tq2 = clock64();
#endif
    rs_sync();
#ifdef RS_TRACE
    if (trace_on) for (int k = threadIdx.x; k < 20 * RTR_EV * RTR_TILES; k += RS_THREADS) g_trace[k] = (&sm.trace[0][0])[k];
    else if (blockIdx.x == 0 && tile0 < 40u) for (int k = threadIdx.x; k < 20 * RTR_EV * RTR_TILES; k += RS_THREADS) (&sm.trace[0][0])[k] = 0u;
#endif
    int r = 0;                                                        // hypothesis t of the flush: group t / 20, sum over the group's four warps
    if ((int)threadIdx.x < n) {
        const int g4 = 4 * ((int)threadIdx.x / RT_HT), j = (int)threadIdx.x % RT_HT;
        r = sm.u.tc.wcnt[g4][j] + sm.u.tc.wcnt[g4 + 1][j] + sm.u.tc.wcnt[g4 + 2][j] + sm.u.tc.wcnt[g4 + 3][j];
    }
    if (threadIdx.x == 0) { sm.stat[2] = 0u; sm.stat[3] = 0u; }
    rs_sync();
#ifdef RS_TIMING
    if (threadIdx.x == 0) {
        atomicAdd(&g_rt_dbg[0], (unsigned long long)(tq1 - tq0)); atomicAdd(&g_rt_dbg[1], (unsigned long long)(tq2 - tq1)); atomicAdd(&g_rt_dbg[4], (unsigned long long)(clock64() - tq2));
        atomicAdd(&g_rt_dbg[5], 1ull); atomicAdd(&g_rt_dbg[6], (unsigned long long)ntiles);
    }
#endif
    return r | (1 << 30);
}

// stage 2 on n queue-1 entries starting at `base` (thread i takes entry base + i): Kabsch + distance check on dense warps, survivors -> queue 2
template <bool RES, bool TC>
BFR_DEVINL void fit_block(RsSmem& sm, const float4* __restrict__ corr_p, uint32_t K, uint64_t seed, uint32_t pair_id, int base, int n, float d2max)
{
    const int lane = threadIdx.x & 31;
    float R[9], t[3];
    bool ok = false; uint32_t h = 0;
    if ((int)threadIdx.x < n) {
        h = sm.q1[base + threadIdx.x];
        uint32_t id[3];
        sample3(seed, pair_id, h, K, id);                             // regenerated: queue 1 only keeps the hypothesis index
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
            for (int k = 0; k < 9; ++k) q2<TC>(sm, k, pos) = R[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) q2<TC>(sm, 9 + k, pos) = t[k];
            q2h<TC>(sm, pos) = h;
        }
    }
}

struct RsTc { unsigned char* scratch; float s1max, qmax; uint32_t tile0; };

// one work item: hypotheses [hb, he) of pair p.  CONF: Open3D confidence rule (hb = the pair's first hypothesis, one item per pair).
// TC: queue 2 is scored by the tensor-core filter (resident pairs only).
template <bool RES, bool CONF, bool TC>
BFR_DEVINL void ransac_item(RsSmem& sm, const float4* __restrict__ corr_p, int K, uint64_t seed, uint32_t pair_id, uint32_t hb, uint32_t he,
                            float d2max, float sim2, float confidence, RsTc& tc, unsigned long long& best, int& n_scored, long long (&tm)[3])
{
    static_assert(!TC || (RES && !CONF), "the tensor-core filter serves resident pairs without the confidence rule");
    const int lane = threadIdx.x & 31;
    constexpr int FLUSH = TC ? RT_FLUSH : RS_THREADS;                 // hypotheses scored per pass over the correspondences

    if (!CONF) {
        // score blocks of FLUSH hypotheses off the top of queue 2 while it holds that many (`all`: whatever is left)
        auto drain = [&](bool all) {
            for (;;) {
                const int qn = sm.qcount;
                rs_sync();                                            // everyone has read qcount
                const int n = qn >= FLUSH ? FLUSH : (all ? qn : 0);
                if (n == 0) break;
                const int base = qn - n;
                if (TC) {
                    int count;
                    RST(tm[2], count = tc_flush(K, base, n, d2max, tc.s1max, tc.qmax, tc.tile0));
                    if (count & (1 << 30)) { tc.tile0 += (uint32_t)((K + RT_TILE - 1) / RT_TILE); count &= ~(1 << 30); }
                    if ((int)threadIdx.x < n) { const unsigned long long pk = pack_count(count, sm.u.tc.qh[base + threadIdx.x]); best = pk > best ? pk : best; }
                } else {
                    uint32_t h; int count, hi;
                    bool mine;
                    RST(tm[2], mine = (score_queue<RES, false>(sm, corr_p, K, base, n, d2max, reinterpret_cast<int*>(sm.q1), h, count, hi)));
                    if (mine) { const unsigned long long pk = pack_count(count, h); best = pk > best ? pk : best; }
                }
                n_scored += n;
                if (threadIdx.x == 0) sm.qcount = base;
                rs_sync();
            }
        };
        auto fit_and_score = [&](int base, int n) {
            RST(tm[1], (fit_block<RES, TC>(sm, corr_p, (uint32_t)K, seed, pair_id, base, n, d2max)));
            rs_sync();
            drain(false);
        };
        for (uint32_t base = hb; base < he; base += RS_S1 * RS_THREADS) {
            // stage 1: RS_S1 independent hypotheses per thread (their sample gathers overlap), cheap checks only (~10 % survive at 70 % outliers)
            uint32_t hh[RS_S1];
            bool ok[RS_S1];
#ifdef RS_TIMING
            const long long ts_ = clock64();
#endif
#pragma unroll
            for (int u = 0; u < RS_S1; ++u) {
                hh[u] = base + (uint32_t)u * RS_THREADS + threadIdx.x;
                ok[u] = RES ? (precheck<RES>(sm, corr_p, (uint32_t)K, seed, pair_id, hh[u], sim2) & (hh[u] < he))        // (a hypothesis index past the range draws valid samples too)
                            : (hh[u] < he && precheck<RES>(sm, corr_p, (uint32_t)K, seed, pair_id, hh[u], sim2));
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
                    if (ok[u]) sm.q1[pos] = hh[u];
                }
            }
            rs_sync();
            const int n1 = sm.q1count;
            rs_sync();                             // everyone has read q1count before it changes
            if (n1 >= RS_THREADS) {
                int done = 0;
                for (; n1 - done >= RS_THREADS; done += RS_THREADS) fit_and_score(done, RS_THREADS);
                const int rem = n1 - done;                                 // < RS_THREADS leftovers move to the front
                uint32_t mv = 0u;
                if ((int)threadIdx.x < rem) mv = sm.q1[done + threadIdx.x];
                rs_sync();
                if ((int)threadIdx.x < rem) sm.q1[threadIdx.x] = mv;
                if (threadIdx.x == 0) sm.q1count = rem;
                rs_sync();
            }
        }
        {
            const int n1 = sm.q1count;             // flush queue 1, then queue 2
            rs_sync();
            if (n1 > 0) fit_and_score(0, n1);
            if (threadIdx.x == 0) sm.q1count = 0;
            rs_sync();
        }
        drain(true);
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
                const bool ok = hh < he && precheck<RES>(sm, corr_p, (uint32_t)K, seed, pair_id, hh, sim2);
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                uint32_t* q1tail = reinterpret_cast<uint32_t*>(sm.q1) + 4 * RS_THREADS;    // srt_c's slice is free at this point
                if (bal) {
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(&sm.q1count, __popc(bal));
                    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
                    if (ok) q1tail[pos] = hh;
                }
            }
            rs_sync();
            const int n1 = sm.q1count;
            rs_sync();
            if (n1 > 0) fit_block<RES, false>(sm, corr_p, (uint32_t)K, seed, pair_id, 4 * RS_THREADS, n1, d2max);
            rs_sync();
            const int qn = sm.qcount;
            rs_sync();
            if (threadIdx.x == 0) { sm.q1count = 0; sm.qcount = 0; }
            if (qn > 0) {
                uint32_t h; int count, hi;
                const bool mine = score_queue<RES, false>(sm, corr_p, K, 0, qn, d2max, reinterpret_cast<int*>(sm.q1), h, count, hi);
                n_scored += qn;
                if (mine) { res_h[hi] = h - hb; res_c[hi] = count; }          // iteration number relative to the pair's first hypothesis
                rs_sync();
                if ((int)threadIdx.x < qn) {                                   // rank by iteration number (all distinct)
                    const uint32_t my = res_h[threadIdx.x];
                    int rank = 0;
                    for (int k = 0; k < qn; ++k) rank += (res_h[k] < my) ? 1 : 0;
                    srt_h[rank] = my; srt_c[rank] = res_c[threadIdx.x];
                }
                rs_sync();
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
            rs_sync();
            if (sm.seq_stop || (base - hb) + RS_THREADS >= sm.seq_bound) break;   // block-uniform
        }
        rs_sync();
        if (threadIdx.x == 0) best = sm.seq_best;
    }
}

__global__ void __launch_bounds__(RS_LAUNCH, 1)
ransac_kernel(const float4* __restrict__ corr, const int32_t* __restrict__ corr_off, const int32_t* __restrict__ corr_cnt, int P, int splits,
              uint64_t seed, uint32_t pair_id_base, uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, float confidence,
              unsigned long long* __restrict__ best_packed, int32_t* __restrict__ valid_count, unsigned char* __restrict__ tc_scratch)
{
    RsSmem& sm = rs_smem();                                           // no static shared memory in this kernel: the dynamic base is 1024-aligned (no integer
    if ((smem_u32(&sm) & 127u) != 0u) __trap();                       // round-up of the pointer: it would turn every access into a generic load)
    const float d2max = __fmul_rn(dist_th, dist_th), sim2 = __fmul_rn(similar_th, similar_th);
    const bool conf = confidence > 0.0f && confidence < 1.0f;
    const bool tc_on = tc_scratch != nullptr && !conf;                // kernel-uniform: the tensor-core warps serve this launch
    const uint32_t nh = h_end - h_begin;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (tc_on) {
        if (threadIdx.x == 0) {
            mbar_init(&sm.flush_go, 1);
            for (int t = 0; t < 2 * RT_GROUPS; ++t) { mbar_init(&sm.acc_empty[t >> 1][t & 1], 4); mbar_init(&sm.acc_full[t >> 1][t & 1], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], RT_GROUPS); }   // a super-stage is free once every group's MMAs have read it
            mbar_fence_init();
            sm.tc_ntiles = 0;
        }
        if (warp == RS_WARPS) tmem_alloc512(&sm.tmem_base);
        static_assert(RT_GROUPS == 4 && RS_LAUNCH == RS_THREADS + 32 * RT_GROUPS, "one tensor-core warp per epilogue warp group");
        for (int i = threadIdx.x; i < (int)sizeof(sm.u.tc.b_op) / 16; i += RS_LAUNCH) reinterpret_cast<uint4*>(sm.u.tc.b_op)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < 4) sm.stat[threadIdx.x] = 0u;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp >= RS_WARPS) {                                           // the tensor-core warps
        if (tc_on) {
            tc_warp_loop(sm, tc_scratch + (size_t)blockIdx.x * RT_MAX_TILES * RT_TILE_BYTES, warp - RS_WARPS);
            tc_fence_before();
            asm volatile("bar.sync 3, 128;" ::: "memory");            // every tensor-core warp is done with tensor memory
            if (warp == RS_WARPS) tmem_dealloc512(sm.tmem_base);
        }
        return;
    }

    RsTc tc;
    tc.scratch = tc_on ? tc_scratch + (size_t)blockIdx.x * RT_MAX_TILES * RT_TILE_BYTES : nullptr;
    tc.s1max = 0.0f; tc.qmax = 0.0f; tc.tile0 = 0u;
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
        rs_sync();                                                    // the previous item is done with shared memory
        if (threadIdx.x == 0) { sm.qcount = 0; sm.q1count = 0; sm.seq_best = 0ull; sm.seq_bound = he - hb; sm.seq_stop = 0; }
        if (K <= RS_CHUNK) {
            load_chunk<3, RS_CHUNK, RS_THREADS>(sm.chunk, corr_p, K, 0);
            rs_sync();
            if (conf) ransac_item<true, true, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, tc, best, n_scored, tm);
            else if (tc_on && tc_prepare_pair(sm, K, tc.scratch, tc.s1max, tc.qmax))
                ransac_item<true, false, true>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, tc, best, n_scored, tm);
            else ransac_item<true, false, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, tc, best, n_scored, tm);
        } else {
            rs_sync();
            if (conf) ransac_item<false, true, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, tc, best, n_scored, tm);
            else      ransac_item<false, false, false>(sm, corr_p, K, seed, pair_id, hb, he, d2max, sim2, confidence, tc, best, n_scored, tm);
        }
        if (valid_count && threadIdx.x == 0 && n_scored) atomicAdd(valid_count + p, n_scored);
#ifdef RS_TIMING
        scored_total += n_scored;
#endif
        // block max -> one atomic
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
        rs_sync();
        if (lane == 0) sm.red[warp] = best;
        rs_sync();
        if (threadIdx.x == 0) {
            unsigned long long b = 0ull;
            for (int w = 0; w < RS_WARPS; ++w) b = sm.red[w] > b ? sm.red[w] : b;
            if (b) atomicMax(best_packed + p, b);
        }
    }
    if (tc_on) {                                                      // release the tensor-core warps
        rs_sync();
        if (threadIdx.x == 0) { sm.tc_ntiles = -1; mbar_arrive(&sm.flush_go); }
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

#ifdef RS_TRACE
}
extern "C" __attribute__((visibility("default"))) unsigned bfr_dbg_ransac_trace(unsigned* out) { cudaMemcpyFromSymbol(out, bfr::g_trace, sizeof(unsigned) * 20 * 8 * 12); return 20 * 8 * 12; }
namespace bfr {
#endif
#ifdef RS_TIMING
}
extern "C" __attribute__((visibility("default"))) void bfr_dbg_ransac_counters(unsigned long long* out) { cudaMemcpyFromSymbol(out, bfr::g_rs_dbg, 64); unsigned long long z[8] = {0}; cudaMemcpyToSymbol(bfr::g_rs_dbg, z, 64); }
extern "C" __attribute__((visibility("default"))) void bfr_dbg_ransac_tc_counters(unsigned long long* out) { cudaMemcpyFromSymbol(out, bfr::g_rt_dbg, 64); unsigned long long z[8] = {0}; cudaMemcpyToSymbol(bfr::g_rt_dbg, z, 64); }
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

size_t ransac_scratch_bytes() { return (size_t)RS_MAX_CTAS * RT_MAX_TILES * RT_TILE_BYTES; }

// -1 = not set (default: tensor-core scoring whenever scratch is available); 0 = exact FP32 scoring only; per thread like the K1 algorithm switch
static thread_local int g_ransac_tc = -1;
void ransac_set_tc(int on) { g_ransac_tc = on; }
int ransac_get_tc() { return g_ransac_tc != 0 ? 1 : 0; }

// tc_scratch: ransac_scratch_bytes() of device memory owned by this call (16-byte aligned), or nullptr / too small: every hypothesis is
// scored by the exact FP32 loop.  The results are identical either way.
cudaError_t ransac_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                          uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, float confidence, int splits,
                          unsigned long long* best_packed, int32_t* valid_count, void* tc_scratch, size_t tc_scratch_bytes, cudaStream_t stream)
{
    static std::atomic<unsigned long long> attr_done{0};
    const size_t smem = sizeof(RsSmem);
    cudaError_t e = ensure_dyn_smem((const void*)ransac_kernel, (int)smem, attr_done);
    if (e != cudaSuccess) return e;
    if (P <= 0 || h_end <= h_begin) return cudaSuccess;
    const bool conf = confidence > 0.0f && confidence < 1.0f;
    if (splits < 1 || conf) splits = 1;                               // the convergence rule is sequential in the hypothesis index: one CTA per pair
    const long long items = (long long)P * splits;
    int sms = sm_count();
    if (sms > RS_MAX_CTAS) sms = RS_MAX_CTAS;
    const unsigned grid = (unsigned)(items < sms ? items : sms);      // persistent: one CTA per SM, items dealt round-robin
    unsigned char* scratch = reinterpret_cast<unsigned char*>(((uintptr_t)tc_scratch + 15) & ~(uintptr_t)15);
    if (!tc_scratch || g_ransac_tc == 0 || (size_t)(scratch - reinterpret_cast<unsigned char*>(tc_scratch)) + (size_t)grid * RT_MAX_TILES * RT_TILE_BYTES > tc_scratch_bytes)
        scratch = nullptr;
    ransac_kernel<<<grid, RS_LAUNCH, smem, stream>>>(reinterpret_cast<const float4*>(corr), corr_off, corr_cnt, P, splits, seed, pair_id_base,
                                                     h_begin, h_end, dist_th, similar_th, confidence, best_packed, valid_count, scratch);
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
