// mutual_nn_tc.cu — K1-TC: tensor-core *filter* + exact FP32 re-check for the mutual-NN argmax (sm_100a, tcgen05/TMEM/TMA).
//
// Same contract and bit-exact results as k1_mutual_nn_kernel (mutual_nn.cu / oracle orc_mutual_nn), ~10x less FP32 work:
//   1. S~ = A B^T with tcgen05.mma kind::f16 on f16 copies of the descriptors (written by k1_prep_kernel; M = 128, N = 256,
//      K = 16 per instruction, FP32 accumulators in TMEM).  f16 keeps 11 significand bits (or an absolute error <= 2^-25 per element
//      below its normal range), so |S~_ij - a_i.b_j| <= (2^-10 + 2^-22) |a_i||b_j| + 2^-25 sqrt(32) (|a_i| + |b_j|); with
//      eps = 1.0625 * 2^-10 |a_i| max_j|b_j| + 2^-22 (|a_i| + max_j|b_j|) every column whose exact score could be the row maximum
//      satisfies  S~_ij + hn(b_j) >= max_j(S~_ij + hn(b_j)) - 2 eps.  (bf16 operands, tried first, need a band 8x wider: 14 % slower on
//      the benchmark workload, 29 % on descriptors without clear matches.)  A pair holding an element f16 cannot represent (|x| > 65504,
//      non-finite) is flagged by the prep kernel; its rows skip the filter result and are scanned exactly.
//   2. The epilogue (one thread per row, TMEM -> registers with tcgen05.ld) keeps the running approximate maximum and a short list of
//      "events": 32-column chunks whose maximum was inside the 2-eps band when they streamed past, with their eight 4-column group
//      maxima (compacted when the list runs low).
//   3. Each thread re-evaluates the groups that are still in band at the end with the EXACT FP32 chain of the oracle (acc = hn(b);
//      acc = fma(a_k, b_k, acc), k ascending) and publishes (key << 32 | ~index) with the same 64-bit RED.MAX as the FP32 kernel, so
//      ties still go to the lowest index and the result is bit-identical by construction.  If an in-band event ever had to be
//      dropped (many near-duplicate descriptors) the row falls back to an exact scan.
// The column direction (tgt -> src) is the same kernel with the roles of the two descriptor sets swapped.
//
// Warp roles (320 threads, 1 CTA/SM, 256 own rows = 2 row halves, 512 TMEM columns = one 128 x 256 accumulator tile per half):
//   warp 0   TMA producer: the CTA's 256 "own" rows once, then 256-row tiles of the streamed side through a 3-stage ring
//            (cp.async.bulk.tensor.2d, SWIZZLE_64B, mbarrier expect-tx)
//   warp 1   TMEM allocator + MMA issuer: per tile and half 2 x tcgen05.mma (K = 2 x 16), tcgen05.commit -> accumulator full / stage free
//   warps 2-9 epilogue (one thread per own row): two batches of 4 x tcgen05.ld.32x32b.x32, (add hn(b_j),) FMNMX3 tree per 4-column
//            group, predicated append of the chunks inside the band; the tile goes back to the MMA warp as soon as the second
//            batch is in registers.  Why the tile is 128 x 256 and single-buffered per half: see the note in the epilogue.
// -DTC_TIMING adds clock64 phase timers and event counters (tools/k1_bench.py prints them); it is never defined in the product build.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include "bfr_tcgen05.cuh"
#include <cuda.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <type_traits>

namespace bfr {

constexpr int TC_BM = 256;                      // own rows per CTA = two M=128 accumulator halves sharing every streamed tile
constexpr int TC_BN = 256;                      // streamed rows per tile = MMA N (the largest cta_group::1 shape: see the note on accumulator switches below)
constexpr int TC_D = 32;
constexpr int TC_MMAK = TC_D / 16;                // K = 16 per f16 MMA
constexpr int TC_STAGES = 3;                     // streamed tiles in flight (3 x 16 KB)
constexpr int TC_EPI_WARPS = 8;                 // one epilogue thread per own row
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SUB = 4;                       // candidate granularity: 4-column groups
constexpr int TC_CAP = 13;                       // band events (32-column chunks with their 8 group maxima) kept per row
constexpr int TC_GCAP = 32;                     // surviving 4-column groups per row handed to the exact re-check
constexpr int TC_MAX_TILES = 24;                // streamed tiles per CTA (hn cache = 6144 floats)

#ifdef TC_TIMING
__device__ unsigned long long g_dbg[8];
#endif

struct TcSmem {
    uint16_t a[TC_BM * TC_D];                   // 16 KB f16, SWIZZLE_64B K-major (one 64-byte row per descriptor); rows 128.. = second half
    uint16_t b[TC_STAGES][TC_BN * TC_D];        // 3 x 16 KB (after the last MMA: uint32_t glist[TC_GCAP][TC_BM], the groups to re-check)
    float hn[TC_MAX_TILES * TC_BN];             // -|b_j|^2/2 of the CTA's streamed columns (-inf beyond the pair)
    struct Ev {                                 // band events of one epilogue warp (once consumed: the warp's 16 KB staging area of the re-check)
        float4 cmg[TC_CAP][2][32];              //   [slot][half][lane]: the eight 4-column group maxima of the chunk
        uint2 cid[TC_CAP][32];                  //   {chunk maximum (float bits), CTA-local first streamed column of the chunk}
    } ev[TC_EPI_WARPS];
    float red[TC_THREADS / 32], red2[TC_THREADS / 32];
    uint64_t a_full, full[TC_STAGES], empty[TC_STAGES], acc_full[2], acc_empty[2];     // acc_*[h]: the accumulator tile of row half h
    uint32_t tmem_base;
};
static_assert(sizeof(uint16_t) * TC_STAGES * TC_BN * TC_D >= sizeof(uint32_t) * TC_GCAP * TC_BM, "group list must fit in the TMA ring");
static_assert(sizeof(TcSmem::Ev) >= 4 * 256 * sizeof(float4) + 32 * sizeof(unsigned long long), "a warp's event slice doubles as its 16 KB staging area (4 candidate columns x 32 rows x 128 B) + 32 packed row maxima");
static_assert(TC_STAGES >= 3 && TC_GCAP * TC_BM * sizeof(uint32_t) <= 2 * TC_BN * TC_D * sizeof(uint16_t), "the group list must leave the third TMA stage free for the staged own rows");
constexpr uint32_t TC_CMG_HALF = sizeof(float4) * 32;           // 512: second half of an event's group maxima
constexpr uint32_t TC_CMG_SLOT = 2 * TC_CMG_HALF;               // 1024
constexpr uint32_t TC_CID_SLOT = sizeof(uint2) * 32;            // 256

// Hot-loop append, fully predicated (no branch): if (c >= thr) { event slot <- (mg[0..7], c, id); advance both slot pointers }.
// p16 / p8 are the shared-memory addresses of this row's next free slot in cmg / cid.
BFR_DEVINL void append_if_in_band(float c, float thr, uint32_t& p16, uint32_t& p8, const float (&mg)[8], uint32_t id)
{
    asm volatile("{\n\t.reg .pred q;\n\t"
                 "setp.ge.f32 q, %2, %3;\n\t"
                 "@q st.shared.v4.f32 [%0], {%4, %5, %6, %7};\n\t"
                 "@q st.shared.v4.f32 [%0+512], {%8, %9, %10, %11};\n\t"
                 "@q st.shared.v2.b32 [%1], {%12, %13};\n\t"
                 "@q add.u32 %0, %0, 1024;\n\t"
                 "@q add.u32 %1, %1, 256;\n\t}"
                 : "+r"(p16), "+r"(p8)
                 : "f"(c), "f"(thr), "f"(mg[0]), "f"(mg[1]), "f"(mg[2]), "f"(mg[3]), "f"(mg[4]), "f"(mg[5]), "f"(mg[6]), "f"(mg[7]),
                   "r"(__float_as_uint(c)), "r"(id)
                 : "memory");
}
static_assert(TC_CMG_HALF == 512 && TC_CMG_SLOT == 1024 && TC_CID_SLOT == 256, "append_if_in_band hard-codes these strides");

BFR_DEVINL void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// asynchronous L2 prefetch of a contiguous global range (bytes % 16 == 0, 16-byte aligned)
BFR_DEVINL void l2_prefetch(const void* gptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
BFR_DEVINL uint64_t umma_desc_sw64(const void* smem)
{   // K-major, SWIZZLE_64B: 8-row groups 512 B apart (SBO = 32 x 16 B), LBO ignored (1), descriptor version 1 (Blackwell), layout 4 = SW64
    return (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// exact score of (own row, candidate row) — the oracle's chain.  COLDIR = the "own" side is the target set.
BFR_DEVINL float exact_score(bool COLDIR, const float4 (&own)[8], float own_hn, const float* __restrict__ cand_row, float cand_hn)
{
    float acc = COLDIR ? own_hn : cand_hn;
    const float4* c4 = reinterpret_cast<const float4*>(cand_row);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const float4 c = __ldg(c4 + k4);
        acc = __fmaf_rn(own[k4].x, c.x, acc); acc = __fmaf_rn(own[k4].y, c.y, acc);
        acc = __fmaf_rn(own[k4].z, c.z, acc); acc = __fmaf_rn(own[k4].w, c.w, acc);
    }
    return COLDIR ? __fadd_rn(acc, cand_hn) : acc;
}

// exact scan of one row by a whole warp (lane l takes streamed columns j_begin + l, + 32, ...): the fallback for rows whose candidate list
// overflowed.  `o` = the own row (every lane holds the same values), returns the packed best (key << 32 | ~index) in every lane.  Kept out of
// line: it runs for a handful of rows per launch and must not cost the main path registers or instruction-cache lines.
__device__ __noinline__ unsigned long long warp_exact_scan(bool COLDIR, const float4* __restrict__ own_row_smem, int swz, float ohn, const float* __restrict__ xs,
                                                           const float* __restrict__ hn_str_p, int j_begin, int j_end)
{
    const int lane = threadIdx.x & 31;
    float4 o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = own_row_smem[c ^ swz];
    float sb = -INFINITY; int sj = 0x7fffffff;
#pragma unroll 2
    for (int j = j_begin + lane; j < j_end; j += 32) {
        const float e = exact_score(COLDIR, o, ohn, xs + (size_t)j * TC_D, hn_str_p[j]);
        if (e > sb) { sb = e; sj = j; }                               // ascending j per lane: strict > keeps the lowest index
    }
    unsigned long long pk = (sj != 0x7fffffff) ? pack_best(float_key(sb), (uint32_t)sj) : 0ull;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, pk, off); pk = other > pk ? other : pk; }
    return pk;
}

// Balanced part of the exact re-check, run by one whole warp, out of line (it only runs for warps in which some row has more than one
// in-band group, and must not cost the main path registers).  The rows' first groups are done by the caller (lane = row); the remaining
// (row, group) pairs are flattened level by level (level k = the rows that have a k-th group, k >= 1) and dealt out evenly, one
// (group, column) per lane and staging buffer, so a warp does not need as many rounds as its busiest lane.  Per-row maxima meet in
// bestS (shared-memory atom.max on the packed key).  n = this lane's group count, glist_col = &glist[0][tile row of lane 0] (row stride
// TC_BM), ownS = the warp's 32 staged own rows (XOR-swizzled float4 chunks), stage = 4 staging buffers of 32 rows.
__device__ __noinline__ void warp_recheck_balanced(bool COLDIR, int n, int nmax, const uint32_t* glist_col, const float4* ownS, float4* stage, unsigned long long* bestS,
                                                   float own_hn, const float* __restrict__ xs, const float* hn_smem, int col0, int j_end)
{
    const int lane = threadIdx.x & 31, sub = lane >> 3, chunk = lane & 7;
    int total = 0;                                                    // groups beyond the first, over the whole warp
    for (int kk = 1; kk < nmax; ++kk) total += __popc(__ballot_sync(0xffffffffu, n > kk));
    for (int base = 0; base < TC_SUB * total; base += TC_SUB * 32) {
        int jc[TC_SUB], rl[TC_SUB];
        float4 reg[TC_SUB][8];
#pragma unroll
        for (int u = 0; u < TC_SUB; ++u) {
            const int item = base + u * 32 + lane;
            int rem = item >> 2, src = -1, kf = 0;                   // item -> (level kf, rem-th row of that level, column item & 3)
            for (int kk = 1; kk < nmax; ++kk) {
                unsigned m = __ballot_sync(0xffffffffu, n > kk);
                const int cnt = __popc(m);
                if (src < 0) {
                    if (rem < cnt) { for (int q2 = 0; q2 < rem; ++q2) m &= m - 1u; src = __ffs((int)m) - 1; kf = kk; }
                    else rem -= cnt;
                }
            }
            const int j0 = (src >= 0) ? (int)glist_col[kf * TC_BM + src] : -0x40000000;
            jc[u] = (j0 >= 0 && j0 + (item & 3) < j_end) ? j0 + (item & 3) : -1;
            rl[u] = src >= 0 ? src : 0;
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) {                          // cooperative gather: 8 lanes per 128-byte row, 4 rows per instruction
                const int jr = __shfl_sync(0xffffffffu, jc[u], 4 * s8 + sub);
                reg[u][s8] = (jr >= 0) ? __ldg(reinterpret_cast<const float4*>(xs + (size_t)jr * TC_D) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < TC_SUB; ++u)
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) { const int rid = 4 * s8 + sub; stage[u * 256 + rid * 8 + (chunk ^ (rid & 7))] = reg[u][s8]; }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < TC_SUB; ++u) {
            const float ohn = __shfl_sync(0xffffffffu, own_hn, rl[u]);
            if (jc[u] >= 0) {
                const float cand_hn = hn_smem[jc[u] - col0];
                float acc = COLDIR ? ohn : cand_hn;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 o = ownS[rl[u] * 8 + (k4 ^ (rl[u] & 7))], c = stage[u * 256 + lane * 8 + (k4 ^ (lane & 7))];
                    acc = __fmaf_rn(o.x, c.x, acc); acc = __fmaf_rn(o.y, c.y, acc); acc = __fmaf_rn(o.z, c.z, acc); acc = __fmaf_rn(o.w, c.w, acc);
                }
                const float e = COLDIR ? __fadd_rn(acc, cand_hn) : acc;
                atomicMax(&bestS[rl[u]], pack_best(float_key(e), (uint32_t)jc[u]));
            }
        }
        __syncwarp();
    }
}

// One launch covers both directions of every pair: row blocks [0, nblk_src) of a pair own source rows and stream the target set
// (-> row_packed), row blocks [nblk_src, gridDim.x) own target rows and stream the source set (-> col_packed).  The CTAs of a pair are
// adjacent in launch order, so the second direction finds the pair's descriptors (FP32 and f16) in L2.
__global__ void __launch_bounds__(TC_THREADS, 1)
k1_tc_kernel(const __grid_constant__ CUtensorMap map_src_own, const __grid_constant__ CUtensorMap map_tgt_str,
             const __grid_constant__ CUtensorMap map_tgt_own, const __grid_constant__ CUtensorMap map_src_str,
             const float* __restrict__ src, const float* __restrict__ tgt,
             const int32_t* __restrict__ src_off, const int32_t* __restrict__ tgt_off,
             const float* __restrict__ hn_src, const float* __restrict__ hn_tgt, int pad_src, int pad_tgt,
             unsigned long long* __restrict__ row_packed, unsigned long long* __restrict__ col_packed,
             const int32_t* __restrict__ out_of_range, int num_pairs, int nblk_src, int splits_src, int splits_tgt, int blk0_src, int blk0_tgt)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];      // no static shared memory in this kernel: base is 1024-aligned
    TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();                 // swizzle atoms need (at least) 512-byte alignment

    const bool COLDIR = (int)blockIdx.x >= nblk_src;                  // the "own" side is the target set
    const int bx = (int)blockIdx.x - (COLDIR ? nblk_src : 0), nbx = COLDIR ? (int)gridDim.x - nblk_src : nblk_src;
    const int splits = COLDIR ? splits_tgt : splits_src;
    if ((int)blockIdx.y >= splits) return;
    const CUtensorMap* map_own = COLDIR ? &map_tgt_own : &map_src_own;
    const CUtensorMap* map_str = COLDIR ? &map_src_str : &map_tgt_str;
    const float* __restrict__ x_own = COLDIR ? tgt : src;
    const float* __restrict__ x_str = COLDIR ? src : tgt;
    const int32_t* __restrict__ off_own = COLDIR ? tgt_off : src_off;
    const int32_t* __restrict__ off_str = COLDIR ? src_off : tgt_off;
    const float* __restrict__ hn_own = COLDIR ? hn_tgt : hn_src;
    const float* __restrict__ hn_str = COLDIR ? hn_src : hn_tgt;
    const int pad_own = COLDIR ? pad_tgt : pad_src, pad_str = COLDIR ? pad_src : pad_tgt;
    unsigned long long* __restrict__ out_packed = COLDIR ? col_packed : row_packed;

    const int p = blockIdx.z;
    const int oo = off_own[p], M = min(off_own[p + 1] - oo, pad_own);   // a pair larger than the caller's max_M / max_N bound is truncated to it
    const int os = off_str[p], N = min(off_str[p + 1] - os, pad_str);
    const int row0 = ((COLDIR ? blk0_tgt : blk0_src) + bx) * TC_BM;     // blk0_*: first row block of this launch's partition (multi-GPU row split)
    if (row0 >= M || N <= 0) return;
    // a descriptor of this pair does not fit f16 (|x| > 65504 or non-finite): the filter's scores are meaningless (inf / NaN), so every row
    // of the pair is scanned exactly by its warp.  The flags are only consumed after the main loop: their load latency costs nothing.
    const int32_t oor_a = __ldg(out_of_range + p), oor_b = __ldg(out_of_range + num_pairs + p);
    const int ntiles_all = (N + TC_BN - 1) / TC_BN;
    const int t_begin = (int)(((long long)blockIdx.y * ntiles_all) / splits);
    const int t_end = (int)(((long long)(blockIdx.y + 1) * ntiles_all) / splits);
    const int ntiles = t_end - t_begin;
    if (ntiles <= 0) return;
    const int halves = (M - row0 > 128) ? 2 : 1;                      // second accumulator half only if it has rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* hn_str_p = hn_str + (size_t)p * pad_str;
#ifdef TC_TIMING
    long long tv_glist = 0, tv_own = 0, tv_rounds = 0, tv_lat = 0; long long tq0 = clock64(), tq1 = 0, tq2 = 0, tq3 = 0, tq4 = 0, tq_wait = 0, tq_ld = 0, tq_proc = 0, tq_cmp = 0;
#define TCT(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#else
#define TCT(acc, stmt) { stmt; }
#endif

    // ---- setup: barriers, TMEM, streamed half-norms (+ their minimum = largest streamed norm) ----------------------
    if (threadIdx.x == 0) {
        mbar_init(&sm.a_full, 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&sm.acc_full[a], 1); mbar_init(&sm.acc_empty[a], 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float hmin = 0.0f, hmax = -INFINITY;
    {   // 128-bit loads: the tile range starts on a 1 KB boundary of the padded array (-inf beyond N)
        const float4* src4 = reinterpret_cast<const float4*>(hn_str_p + (size_t)t_begin * TC_BN);
        float4* dst4 = reinterpret_cast<float4*>(sm.hn);
        for (int i = threadIdx.x; i < ntiles * (TC_BN / 4); i += TC_THREADS) {
            const float4 h = __ldg(src4 + i);
            dst4[i] = h;
            const float hv[4] = { h.x, h.y, h.z, h.w };
#pragma unroll
            for (int c = 0; c < 4; ++c) if (hv[c] > -INFINITY) { hmin = fminf(hmin, hv[c]); hmax = fmaxf(hmax, hv[c]); }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { hmin = fminf(hmin, __shfl_xor_sync(0xffffffffu, hmin, o)); hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, o)); }
    if (lane == 0) { sm.red[warp] = hmin; sm.red2[warp] = hmax; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
#pragma unroll
    for (int w = 0; w < TC_THREADS / 32; ++w) { hmin = fminf(hmin, sm.red[w]); hmax = fmaxf(hmax, sm.red2[w]); }
    const float str_max_sq = -2.0f * hmin;                            // max_j |b_j|^2 over this CTA's columns
    const float hn_spread = fmaxf(hmax - hmin, 0.0f);                 // 0 for exactly normalised descriptors
#ifdef TC_TIMING
    tq1 = clock64();
#endif

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(&sm.a_full, TC_BM * TC_D * 2);
            tma_load_2d(sm.a, map_own, 0, oo + row0, &sm.a_full);
#ifndef TC_NO_L2_PREFETCH
            // The exact re-check reads FP32 rows the main loop never touches (it streams the f16 copies): the CTA's own rows and a
            // scattered subset of the streamed set.  Pull them towards L2 now - the own rows, and this CTA's share of the streamed rows of
            // the pair (together the row blocks of a pair cover all of them) - so that the re-check waits for L2, not for DRAM.
            {
                const int own_rows = min(TC_BM, M - row0);
                l2_prefetch(x_own + (size_t)(oo + row0) * TC_D, (uint32_t)own_rows * TC_D * 4u);
                const int nblk = nbx, j0 = t_begin * TC_BN, j1 = min(N, t_end * TC_BN);
                const int per = ((j1 - j0 + nblk - 1) / nblk + 7) & ~7;                      // rows per row block, rounded to 1 KB
                const int a0 = j0 + bx * per, a1 = min(j1, a0 + per);
                if (a1 > a0) l2_prefetch(x_str + (size_t)(os + a0) * TC_D, (uint32_t)(a1 - a0) * TC_D * 4u);
            }
#endif
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_STAGES; const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(&sm.empty[s], ph ^ 1u);
                mbar_expect_tx(&sm.full[s], TC_BN * TC_D * 2);
                tma_load_2d(sm.b[s], map_str, 0, os + (t_begin + it) * TC_BN, &sm.full[s]);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor (kind::f16): D = F32, A = B = F16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t adesc0 = umma_desc_sw64(sm.a), adesc1 = umma_desc_sw64(sm.a + 128 * TC_D);
            mbar_wait(&sm.a_full, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_STAGES; const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(&sm.full[s], ph);
                const uint64_t bdesc = umma_desc_sw64(sm.b[s]);
                for (int h = 0; h < halves; ++h) {                    // one accumulator tile (128 rows x 256 streamed columns) per row half
                    mbar_wait(&sm.acc_empty[h], ((uint32_t)it & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(h * TC_BN);
                    const uint64_t ad = h ? adesc1 : adesc0;
#pragma unroll
                    for (int k = 0; k < TC_MMAK; ++k)              // K = 16 per instruction: +32 bytes (2 x 16 B) per step inside the 64-byte row
                        umma_f16(d, ad + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
                    umma_commit(&sm.acc_full[h]);
                }
                umma_commit(&sm.empty[s]);                            // the stage is free once every MMA that reads it has completed
            }
        }
    } else if (warp - 2 < 4 * halves) {
        // ================= epilogue: one thread per own row =================
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                             // which M = 128 accumulator
        const int r = half * 128 + q * 32 + lane;                     // row inside the CTA tile
        const int row = row0 + r;
        const bool valid = row < M;
        const float own_hn = valid ? hn_own[(size_t)p * pad_own + row] : 0.0f;
        const float own_norm = sqrtf(fmaxf(-2.0f * own_hn, 0.0f)), str_norm = sqrtf(fmaxf(str_max_sq, 0.0f));
        // 2 eps, eps = 1.0625 * 2^-10 |a| max|b| + 2^-22 (|a| + max|b|): f16 operands carry 11 significand bits (relative error 2^-11 each),
        // or an absolute error of at most 2^-25 per element below the normal range
        // ... plus the rounding of the exact FP32 chain itself, which defines the winner: 33 round-offs of partial sums bounded by
        // |b|^2/2 + |a||b| <= 1.5 max(|a|, max|b|)^2 (2^-24 each) -- negligible for normalised rows, but it keeps the band airtight when a row's
        // norm is far below the largest streamed norm
        const float big_norm = fmaxf(own_norm, str_norm);
        const float two_eps = 0.0020752f * own_norm * str_norm + 4.77e-7f * (own_norm + str_norm) + 6.0e-6f * big_norm * big_norm + 1e-30f;
        // Streamed norms (nearly) uniform -- L2-normalised descriptors, BUFFER's case: rank full tiles on the raw dot products
        // (no hn add) and widen the band by the spread of hn; scores of hn-adjusted (partial) tiles are shifted by -hmax to match.
        const bool uniform = hn_spread <= 0.0009765625f * str_max_sq;   // CTA-uniform (the branch below contains warp-collective TMEM loads)
        const float band = !valid ? -INFINITY : uniform ? two_eps + hn_spread : two_eps;     // rows beyond M never record an event
        float m_run = -INFINITY, dropped_max = -INFINITY;
        const int ew = warp - 2;                                      // epilogue warp index: owns sm.ev[ew].cmg / sm.ev[ew].cid
        const uint32_t p16_base = smem_u32(&sm.ev[ew].cmg[0][0][lane]), p8_base = smem_u32(&sm.ev[ew].cid[0][lane]);
        uint32_t p16 = p16_base, p8 = p8_base;                        // next free event slot of this row
        const uint32_t p8_high = p8_base + (uint32_t)(TC_CAP - 4) * TC_CID_SLOT;   // a batch of 128 columns appends at most 4 events

        // one 32-column chunk: (add hn(b_j),) 4-column group maxima -> chunk maximum
        auto reduce_chunk = [&](float (&v)[32], int colbase, float (&mg)[8], auto raw_tag) -> float {
            constexpr bool RAW = decltype(raw_tag)::value;            // RAW: v stays the bare dot product, compared in "dot + hmax" units
            if (!RAW) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 h = *reinterpret_cast<const float4*>(&sm.hn[colbase + 4 * c4]);
                    f32x2 lo = add2(pack2(v[4 * c4], v[4 * c4 + 1]), pack2(h.x, h.y)), hi = add2(pack2(v[4 * c4 + 2], v[4 * c4 + 3]), pack2(h.z, h.w));
                    unpack2(lo, v[4 * c4], v[4 * c4 + 1]); unpack2(hi, v[4 * c4 + 2], v[4 * c4 + 3]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {                             // maxima of the eight 4-column groups, in raw dot-product units
                mg[k] = fmaxf(max3(v[4 * k], v[4 * k + 1], v[4 * k + 2]), v[4 * k + 3]);
                if (!RAW) mg[k] -= hmax;                              // hn-adjusted tiles: shift so that both kinds of tile compare
            }
            return fmaxf(max3(max3(mg[0], mg[1], mg[2]), mg[3], mg[4]), max3(mg[5], mg[6], mg[7]));
        };
        // one batch of four chunks (128 streamed columns), in two steps so that the next TMEM load can be issued in between: the four
        // reductions are independent; the running maximum is raised by the whole batch first, so all four chunks are tested against one
        // (tighter) threshold; a chunk whose maximum is inside the band keeps its eight group maxima as one event (predicated stores)
        float g0[8], g1[8], g2[8], g3[8], c0, c1, c2, c3;
        auto reduce_batch = [&](float (&v0)[32], float (&v1)[32], float (&v2)[32], float (&v3)[32], int colbase, auto raw_tag) {
            c0 = reduce_chunk(v0, colbase, g0, raw_tag); c1 = reduce_chunk(v1, colbase + 32, g1, raw_tag);
            c2 = reduce_chunk(v2, colbase + 64, g2, raw_tag); c3 = reduce_chunk(v3, colbase + 96, g3, raw_tag);
        };
        auto append_batch = [&](int colbase) {
            m_run = fmaxf(max3(m_run, c0, c1), fmaxf(c2, c3));
            const float thr = m_run - band;
            append_if_in_band(c0, thr, p16, p8, g0, (uint32_t)colbase);
            append_if_in_band(c1, thr, p16, p8, g1, (uint32_t)(colbase + 32));
            append_if_in_band(c2, thr, p16, p8, g2, (uint32_t)(colbase + 64));
            append_if_in_band(c3, thr, p16, p8, g3, (uint32_t)(colbase + 96));
        };
        // rare (per lane, once the slots run low): drop the events that fell out of the band.  If more than TC_CAP - 4 survive (the running
        // maximum is still creeping up through typical chunk maxima) the lowest ones are dropped too and their maximum remembered: the row
        // only needs the exact scan if that maximum is still inside the band at the very end.
        auto compact = [&]() {
            const float thr = m_run - band;
            const int cnt = (int)((p8 - p8_base) / TC_CID_SLOT);
            int n = 0;
            for (int k = 0; k < cnt; ++k) {
                const uint2 e = sm.ev[ew].cid[k][lane];
                if (__uint_as_float(e.x) >= thr) {
                    if (n != k) { sm.ev[ew].cid[n][lane] = e; sm.ev[ew].cmg[n][0][lane] = sm.ev[ew].cmg[k][0][lane]; sm.ev[ew].cmg[n][1][lane] = sm.ev[ew].cmg[k][1][lane]; }
                    ++n;
                }
            }
#ifdef TC_TIMING
            atomicAdd(&g_dbg[0], 1ull); atomicAdd(&g_dbg[1], (unsigned long long)cnt); atomicAdd(&g_dbg[2], (unsigned long long)n);
            if (n > TC_CAP - 4) atomicAdd(&g_dbg[valid ? 3 : 4], 1ull);
#endif
            while (n > TC_CAP - 4) {
                int jmin = 0; float cmin = __uint_as_float(sm.ev[ew].cid[0][lane].x);
                for (int k = 1; k < n; ++k) { const float ck = __uint_as_float(sm.ev[ew].cid[k][lane].x); if (ck < cmin) { cmin = ck; jmin = k; } }
                dropped_max = fmaxf(dropped_max, cmin);
                --n;
                if (jmin != n) { sm.ev[ew].cid[jmin][lane] = sm.ev[ew].cid[n][lane]; sm.ev[ew].cmg[jmin][0][lane] = sm.ev[ew].cmg[n][0][lane]; sm.ev[ew].cmg[jmin][1][lane] = sm.ev[ew].cmg[n][1][lane]; }
            }
            p16 = p16_base + (uint32_t)n * TC_CMG_SLOT; p8 = p8_base + (uint32_t)n * TC_CID_SLOT;
        };
        // The tensor pipe only streams back-to-back MMAs that accumulate into the SAME tile; every switch to another accumulator tile
        // drains it (~350 cycles, independent of N: tools/microbench/mma_bubble.cu).  With K = 32 there are just two MMAs per tile, so
        // the tile is made as large as the instruction allows (128 x 256) and each row half owns ONE such tile (2 x 256 = all 512 TMEM
        // columns).  The halves alternate: while the four warps of one half pull their tile into registers and reduce it, the tensor
        // core fills the tile of the other half.  A warp hands its tile back as soon as the last 128 columns are in registers.
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * TC_BN);
        float va[32], vb[32], vc[32], vd[32];
        for (int it = 0; it < ntiles; ++it) {
            const int cb = it * TC_BN;                                // CTA-local first streamed column of this tile
#ifdef TC_TIMING
            const long long tw0 = clock64();
#endif
            mbar_wait(&sm.acc_full[half], (uint32_t)it & 1u);
#ifdef TC_TIMING
            tq_wait += clock64() - tw0;
#endif
            tc_fence_after();
            __syncwarp();
            // two batches of 128 columns; the second load is in flight while the first batch's events are appended
            const bool raw0 = uniform && t_begin * TC_BN + cb + 128 <= N;        // no padding columns to mask and uniform norms
            const bool raw1 = uniform && t_begin * TC_BN + cb + 256 <= N;
            tmem_ld32_issue(t0, va); tmem_ld32_issue(t0 + 32, vb); tmem_ld32_issue(t0 + 64, vc); tmem_ld32_issue(t0 + 96, vd);
            TCT(tq_ld, tmem_ld_wait(va); tmem_ld_pin(vb); tmem_ld_pin(vc); tmem_ld_pin(vd));
            TCT(tq_proc,
            if (raw0) reduce_batch(va, vb, vc, vd, cb, std::true_type{});
            else reduce_batch(va, vb, vc, vd, cb, std::false_type{});)
            tmem_ld32_issue(t0 + 128, va); tmem_ld32_issue(t0 + 160, vb); tmem_ld32_issue(t0 + 192, vc); tmem_ld32_issue(t0 + 224, vd);
            TCT(tq_proc, append_batch(cb));
            TCT(tq_cmp, if (p8 > p8_high) compact());
            TCT(tq_ld, tmem_ld_wait(va); tmem_ld_pin(vb); tmem_ld_pin(vc); tmem_ld_pin(vd));
            tc_fence_before();                                        // all TMEM reads of this tile are complete
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.acc_empty[half]);
            TCT(tq_proc,
            if (raw1) reduce_batch(va, vb, vc, vd, cb + 128, std::true_type{});
            else reduce_batch(va, vb, vc, vd, cb + 128, std::false_type{});
            append_batch(cb + 128);)
            TCT(tq_cmp, if (p8 > p8_high) compact());
        }

        // The group list below reuses the TMA ring: every MMA of BOTH row halves must have read its stage first.  A warp knows that only
        // for its own half (its last acc_full wait), so the epilogue warps meet once here (named barrier 1; warps 0 and 1 are not part).
        asm volatile("bar.sync 1, %0;" ::"r"(128 * halves) : "memory");
#ifdef TC_TIMING
        tq2 = clock64();
#endif
        // ---- exact FP32 re-check of the surviving groups (or of the whole row after an overflow) ---------------------------
        // One lane = one row, but the candidate rows are fetched cooperatively: 8 lanes read the 8 float4 of one 128-byte row
        // (coalesced, 4 rows per warp-wide load) and transpose through an XOR-swizzled per-warp staging area (the warp's own event
        // slice, dead once the group list is built), so that the L1 sees 4 wavefronts per load instruction instead of 32.  The four
        // columns of a group are fetched together (one memory round trip per group, usually one per row).
        {
            const float thr = m_run - band;
            bool overflow = (oor_a | oor_b) != 0 || !(dropped_max < thr);    // a dropped event could still hold the maximum (or thr is NaN): exact scan of the row
            const int j_end = min(N, t_end * TC_BN);
            const int sub = lane >> 3, chunk = lane & 7;
            // gather 32 rows (one per lane, row index `want`, -1 = none) of `base` into registers, 4 rows per instruction
            auto fetch = [&](const float* __restrict__ base, int want, float4 (&reg)[8]) {
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const int rid = 4 * s8 + sub;
                    const int jr = __shfl_sync(0xffffffffu, want, rid);
                    reg[s8] = (jr >= 0) ? __ldg(reinterpret_cast<const float4*>(base + (size_t)jr * TC_D) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 reg[TC_SUB][8];
            fetch(x_own + (size_t)oo * TC_D, valid ? row : -1, reg[0]);      // own row: in flight while the events are decoded
            // surviving events -> list of in-band 4-column groups (global streamed column of each), in the idle TMA ring
            uint32_t (*glist)[TC_BM] = reinterpret_cast<uint32_t (*)[TC_BM]>(&sm.b[0][0]);
            int n = 0;
            if (valid && !overflow) {
                const int cnt = (int)((p8 - p8_base) / TC_CID_SLOT);
                uint32_t q_ev = 0;                                    // events whose chunk maximum is still inside the band
#pragma unroll
                for (int k = 0; k < TC_CAP; ++k)
                    if (k < cnt && __uint_as_float(sm.ev[ew].cid[k][lane].x) >= thr) q_ev |= 1u << k;
                while (q_ev) {                                        // lanes walk their own (few) events in lock step
                    const int k = __ffs((int)q_ev) - 1; q_ev &= q_ev - 1;
                    const float4 g0 = sm.ev[ew].cmg[k][0][lane], g1 = sm.ev[ew].cmg[k][1][lane];
                    const uint32_t colbase = (uint32_t)(t_begin * TC_BN) + sm.ev[ew].cid[k][lane].y;
                    uint32_t q_g = (g0.x >= thr ? 1u : 0u) | (g0.y >= thr ? 2u : 0u) | (g0.z >= thr ? 4u : 0u) | (g0.w >= thr ? 8u : 0u) |
                                   (g1.x >= thr ? 16u : 0u) | (g1.y >= thr ? 32u : 0u) | (g1.z >= thr ? 64u : 0u) | (g1.w >= thr ? 128u : 0u);
                    while (q_g) {
                        const int g = __ffs((int)q_g) - 1; q_g &= q_g - 1;
                        if (n < TC_GCAP) glist[n][r] = colbase + 4u * (uint32_t)g;
                        ++n;
                    }
                }
                if (n > TC_GCAP) { overflow = true; n = 0; }
            }
#ifdef TC_TIMING
            tv_glist = clock64();
#endif
            __syncwarp();                                             // every lane is done with its events: the slice becomes the staging area
            float4* stage = reinterpret_cast<float4*>(&sm.ev[ew]);    // TC_SUB buffers x (32 rows x 8 float4)
            auto put = [&](int buf, const float4 (&rg)[8]) {
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) { const int rid = 4 * s8 + sub; stage[buf * 256 + rid * 8 + (chunk ^ (rid & 7))] = rg[s8]; }
            };
            auto get = [&](int buf, float4 (&rowv)[8]) {
#pragma unroll
                for (int c = 0; c < 8; ++c) rowv[c] = stage[buf * 256 + lane * 8 + (c ^ (lane & 7))];
            };
            // own rows: staged once per warp in operand memory that is dead by now (the f16 own tile / the third TMA stage), so that any
            // lane can score any row of the warp in the balanced part and in the overflow scan
            float4* ownS = (ew < 4) ? reinterpret_cast<float4*>(sm.a) + ew * 256 : reinterpret_cast<float4*>(sm.b[2]) + (ew - 4) * 256;
            unsigned long long* bestS = reinterpret_cast<unsigned long long*>(stage + TC_SUB * 256);      // per-row maxima (packed)
            float4 own[8];
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) { const int rid = 4 * s8 + sub; ownS[rid * 8 + (chunk ^ (rid & 7))] = reg[0][s8]; }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) own[c] = ownS[lane * 8 + (c ^ (lane & 7))];
            const int nmax = __reduce_max_sync(0xffffffffu, n);
#ifdef TC_TIMING
            tv_own = clock64(); tv_rounds = nmax;
#endif
            const float* xs = x_str + (size_t)os * TC_D;
            float best = -INFINITY; int best_j = 0x7fffffff;
            // first group of every row: lane = row, the four columns of the group in one memory round trip (the common case ends here)
            if (nmax > 0) {
                const int j0 = (n > 0) ? (int)glist[0][r] : -0x40000000;
                int jc[TC_SUB];
#pragma unroll
                for (int u = 0; u < TC_SUB; ++u) { jc[u] = (j0 >= 0 && j0 + u < j_end) ? j0 + u : -1; fetch(xs, jc[u], reg[u]); }
#ifdef TC_TIMING
                const long long tr0 = clock64();
#endif
#pragma unroll
                for (int u = 0; u < TC_SUB; ++u) put(u, reg[u]);
                __syncwarp();
#ifdef TC_TIMING
                tv_lat += clock64() - tr0;
#endif
#pragma unroll
                for (int u = 0; u < TC_SUB; ++u) {
                    if (jc[u] >= 0) {
                        float4 c[8];
                        get(u, c);
                        const float cand_hn = sm.hn[jc[u] - t_begin * TC_BN];
                        float acc = COLDIR ? own_hn : cand_hn;
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4) {
                            acc = __fmaf_rn(own[k4].x, c[k4].x, acc); acc = __fmaf_rn(own[k4].y, c[k4].y, acc);
                            acc = __fmaf_rn(own[k4].z, c[k4].z, acc); acc = __fmaf_rn(own[k4].w, c[k4].w, acc);
                        }
                        const float e = COLDIR ? __fadd_rn(acc, cand_hn) : acc;
                        if (e > best || (e == best && jc[u] < best_j)) { best = e; best_j = jc[u]; }
                    }
                }
                __syncwarp();
            }
            bestS[lane] = (best_j != 0x7fffffff) ? pack_best(float_key(best), (uint32_t)best_j) : 0ull;
            __syncwarp();
            if (nmax > 1)                                             // rows without a clear winner: their further groups, dealt out evenly
                warp_recheck_balanced(COLDIR, n, nmax, &glist[0][r - lane], ownS, stage, bestS, own_hn, xs, sm.hn, t_begin * TC_BN, j_end);
            __syncwarp();
            // pathological rows (an in-band event had to be dropped: many near-duplicates, or ten chunks within the band of the maximum): exact
            // scan of the row, done by the whole warp (lane l takes columns l, l + 32, ...) - a single lane would hold its SM for milliseconds
            unsigned long long scan_best = 0ull;
            for (unsigned ovm = __ballot_sync(0xffffffffu, valid && overflow); ovm; ovm &= ovm - 1u) {
                const int src = __ffs((int)ovm) - 1;
                const float ohn = __shfl_sync(0xffffffffu, own_hn, src);
                const unsigned long long pk = warp_exact_scan(COLDIR, ownS + src * 8, src & 7, ohn, xs, hn_str_p, t_begin * TC_BN, j_end);
                if (lane == src) scan_best = pk;
            }
            unsigned long long fin = bestS[lane];
            fin = scan_best > fin ? scan_best : fin;
            if (valid && fin != 0ull) red_max_u64(out_packed + (size_t)p * pad_own + row, fin);
        }
    }

#ifdef TC_TIMING
    tq3 = clock64();
#endif
    tc_fence_before();
    __syncthreads();
#ifdef TC_TIMING
    tq4 = clock64();
    if (blockIdx.z == 3 && bx == 1 && lane == 0 && (warp == 2 || warp == 9)) printf("dir %d cta %d warp %d: setup %lld main %lld (acc wait %lld ldwait %lld proc %lld compact %lld) verify %lld (glist %lld own %lld nmax %lld loadwait %lld) tailwait %lld\n", (int)COLDIR, bx, warp, tq1 - tq0, tq2 - tq1, tq_wait, tq_ld, tq_proc, tq_cmp, tq3 - tq2, tv_glist - tq2, tv_own - tv_glist, tv_rounds, tv_lat, tq4 - tq3);
#endif
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

#ifdef TC_TIMING
}
extern "C" __attribute__((visibility("default"))) void bfr_dbg_counters(unsigned long long* out) { cudaMemcpyFromSymbol(out, bfr::g_dbg, 64); unsigned long long z[8] = {0}; cudaMemcpyToSymbol(bfr::g_dbg, z, 64); }
namespace bfr {
#endif
// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_map(CUtensorMap* map, const void* base, long long rows, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = { (cuuint64_t)TC_D, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)TC_D * 2 };
    cuuint32_t box[2] = { (cuuint32_t)TC_D, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool k1_tc_supported(int D, long long total_M, long long total_N) { return D == TC_D && total_M > 0 && total_N > 0 && encode_fn() != nullptr; }

// both directions; hna/hnb/row_packed/col_packed are the (prepared, zeroed) workspace arrays of k1_launch
cudaError_t k1_tc_launch(const float* src, const float* tgt, const void* src_f16, const void* tgt_f16, const int32_t* out_of_range, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                         long long total_M, long long total_N, const float* hna, const float* hnb, int padM, int padN,
                         unsigned long long* row_packed, unsigned long long* col_packed, int part, int nparts, cudaStream_t stream)
{
    CUtensorMap ms_own, ms_str, mt_own, mt_str;
    if (!make_map(&ms_own, src_f16, total_M, TC_BM) || !make_map(&ms_str, src_f16, total_M, TC_BN) ||
        !make_map(&mt_own, tgt_f16, total_N, TC_BM) || !make_map(&mt_str, tgt_f16, total_N, TC_BN)) return cudaErrorNotSupported;
    const size_t smem = sizeof(TcSmem) + 1024;
    {
        static std::atomic<unsigned long long> attr_done{0};
        cudaError_t e = ensure_dyn_smem((const void*)k1_tc_kernel, (int)smem, attr_done);
        if (e != cudaSuccess) return e;
    }
    // one launch, both directions: x = row blocks of the source side, then of the target side; y = column splits (the larger of the two
    // directions' counts; the other direction's surplus CTAs exit at once); z = pair
    const int nblk_src = (max_M + TC_BM - 1) / TC_BM, nblk_tgt = (max_N + TC_BM - 1) / TC_BM;
    const int tiles_tgt = (max_N + TC_BN - 1) / TC_BN, tiles_src = (max_M + TC_BN - 1) / TC_BN;
    // column splits: as many as the hn cache demands (TC_MAX_TILES tiles per CTA); for small batches (the reference registers one pair per
    // forward) more, so that the row blocks of one or two pairs still cover the 148 SMs in one wave (>= 4 tiles per CTA, at most 8 splits);
    // rows merge across splits through the packed RED.MAX like across CTAs
    // multi-GPU row split: this launch owns row blocks [b0, b1) of each direction (nparts = 1: all of them)
    const int b0s = (int)(((long long)part * nblk_src) / nparts), b1s = (int)(((long long)(part + 1) * nblk_src) / nparts);
    const int b0t = (int)(((long long)part * nblk_tgt) / nparts), b1t = (int)(((long long)(part + 1) * nblk_tgt) / nparts);
    const long long ctas = (long long)P * ((b1s - b0s) + (b1t - b0t));
    const int fill = ctas > 0 && ctas < 148 ? (int)(148 / ctas) : 1;
    auto pick = [&](int tiles) {
        int need = (tiles + TC_MAX_TILES - 1) / TC_MAX_TILES; if (need < 1) need = 1;
        int want = fill; if (want > tiles / 4) want = tiles / 4; if (want > 8) want = 8;
        return want > need ? want : need;
    };
    const int splits_src = pick(tiles_tgt);     // src rows own, tgt streamed
    const int splits_tgt = pick(tiles_src);     // tgt rows own, src streamed
    if ((b1s - b0s) + (b1t - b0t) <= 0) return cudaSuccess;
    dim3 grid((unsigned)((b1s - b0s) + (b1t - b0t)), (unsigned)(splits_src > splits_tgt ? splits_src : splits_tgt), (unsigned)P);
    k1_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(ms_own, mt_str, mt_own, ms_str, src, tgt, src_off, tgt_off, hna, hnb, padM, padN,
                                                     row_packed, col_packed, out_of_range, P, b1s - b0s, splits_src, splits_tgt, b0s, b0t);
    return cudaGetLastError();
}

}  // namespace bfr
