// mutual_nn.cu — K1: fused descriptor L2 distance + mutual-nearest-neighbour argmin for sm_100a.
//
// Replaces buffer.mutual_matching (reference models/BUFFER.py:335-359): two knn_cuda.KNN(k=1) brute-force passes that
// materialise the N x M distance matrix in HBM, two device->host syncs and a numpy mutual check.  Here one pass
// computes every a_i.b_j once, reduces row maxima (src -> tgt NN) and column maxima (tgt -> src NN) on the fly and
// never stores the matrix.  Arithmetic (bit-exact with oracle/bfr_oracle.c orc_mutual_nn):
//     acc_ij = -|b_j|^2/2, then acc_ij = fma(a_ik, b_jk, acc_ij) for k = 0..31;   nn_s[i] = argmax_j acc_ij,
//     nn_t[j] = argmax_i (acc_ij - |a_i|^2/2);   ties -> lowest index (packed 64-bit max of (key << 32 | ~index)).
//
// Design ("row-stationary"): a thread keeps 4 source rows x 32 dims in registers as two packed row pairs, so the
// inner product is FFMA2 (fma.rn.f32x2: two rows per instruction) against a target value that every lane of the warp
// reads from the same shared-memory address (LDS.128 broadcast: 4 dims per load, no bank conflicts, no swizzle).
// Target rows stream through a 4-stage shared-memory ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier).
// Row maxima are thread-local (value-only FMNMX3 per 8-column chunk, index recovered only when the chunk improves the
// running maximum); column maxima are one CREDUX (redux.sync.max.f32) per column per warp, the winning lane publishes
// with a 64-bit atomicMax.  Issue-slot accounting and the measured roofline are in DESIGN.md §K1.
#include "bfr_common.cuh"
#include "bfr_kernels.h"
#include <atomic>
#include <cfloat>
#include <cmath>

namespace bfr {

constexpr int K1_D = 32;                      // descriptor length (BUFFER: Cylindrical_Net dim=32, models/patchnet.py:69-85)
constexpr int K1_WARPS = 4;
constexpr int K1_THREADS = K1_WARPS * 32;
constexpr int K1_RPT = 4;                     // source rows per thread (two FFMA2 row pairs)
constexpr int K1_ROWS_PER_WARP = 32 * K1_RPT; // 128
constexpr int K1_ROWS = K1_THREADS * K1_RPT;  // 512 source rows per CTA
constexpr int K1_TILE = 64;                   // target rows per pipeline stage
constexpr int K1_STAGES = 4;
constexpr int K1_JC = 8;                      // columns per register chunk

struct __align__(128) K1Smem {
    float b[K1_STAGES][K1_TILE][K1_D];        // 4 x 8 KB target tiles
    float hb[K1_STAGES][K1_TILE];             // -|b_j|^2/2 of the tile's columns
    uint64_t full[K1_STAGES];                 // TMA completion (expect_tx) barriers
    int done[K1_STAGES];                      // warps that have finished reading the stage; the last one refills it
};

// ---- prep: half squared norms into the padded workspace (-inf in the padding), zeroed packed bests and -- for the tensor-core
// path -- an f16 (round-to-nearest-even) copy of the descriptors, row o + i of the concatenated array -> row o + i of xb.  A pair with an
// element that f16 cannot hold (|x| > 65504 or non-finite) is flagged in oor[p]: the tensor-core kernel then scans that pair exactly ------
__global__ void __launch_bounds__(256) k1_prep_kernel(const float* __restrict__ x, const int32_t* __restrict__ off, int P, int D, int pad,
                                                      float* __restrict__ hn, unsigned long long* __restrict__ packed, uint4* __restrict__ xb,
                                                      int32_t* __restrict__ oor)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P * pad) return;
    const int p = (int)(gid / pad), i = (int)(gid % pad);
    const int o = off[p], n = off[p + 1] - o;
    float h = -INFINITY;
    if (i < n) {
        const float* row = x + (size_t)(o + i) * D;
        float s = 0.0f;
        if (D == K1_D) {
            float4 v[K1_D / 4];
#pragma unroll
            for (int k = 0; k < K1_D / 4; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(row) + k);
#pragma unroll
            for (int k = 0; k < K1_D / 4; ++k) {
                s = __fmaf_rn(v[k].x, v[k].x, s); s = __fmaf_rn(v[k].y, v[k].y, s); s = __fmaf_rn(v[k].z, v[k].z, s); s = __fmaf_rn(v[k].w, v[k].w, s);
            }
            if (xb) {
                float amax = 0.0f;
#pragma unroll
                for (int k = 0; k < K1_D / 4; ++k) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[k].x), fabsf(v[k].y)), fmaxf(fabsf(v[k].z), fabsf(v[k].w))));
                if (!(amax <= 65504.0f) || !(s == s)) atomicOr(&oor[p], 1);                  // (NaN elements: fmaxf drops them, the norm does not)
                uint4* dst = xb + (size_t)(o + i) * (K1_D / 8);
#pragma unroll
                for (int k = 0; k < K1_D / 8; ++k) {
                    uint4 w;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(v[2 * k].y), "f"(v[2 * k].x));       // low half = first element
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(v[2 * k].w), "f"(v[2 * k].z));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.z) : "f"(v[2 * k + 1].y), "f"(v[2 * k + 1].x));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w.w) : "f"(v[2 * k + 1].w), "f"(v[2 * k + 1].z));
                    dst[k] = w;
                }
            }
        } else if ((D & 3) == 0) {
            for (int k = 0; k < D; k += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
                s = __fmaf_rn(v.x, v.x, s); s = __fmaf_rn(v.y, v.y, s); s = __fmaf_rn(v.z, v.z, s); s = __fmaf_rn(v.w, v.w, s);
            }
        } else {
            for (int k = 0; k < D; ++k) { const float v = __ldg(row + k); s = __fmaf_rn(v, v, s); }
        }
        h = __fmul_rn(-0.5f, s);
    }
    hn[gid] = h;
    packed[gid] = 0ull;
}

// ---- main kernel ----------------------------------------------------------------------------------------------
// One 8-column chunk of the current tile against this thread's 4 stationary rows.  FULL = every column of the chunk is
// a real target row (no bounds checks); the partial last tile of a pair uses FULL = false.
template <bool FULL>
BFR_DEVINL void k1_chunk(const K1Smem& sm, int stage, int j0, int ncols, int col0, const f32x2 (&ap)[2][K1_D], f32x2 hap0, f32x2 hap1,
                         float (&rbest)[K1_RPT], int (&ridx)[K1_RPT], unsigned long long* __restrict__ colp, const uint32_t (&nrow)[K1_RPT])
{
    f32x2 acc[2][K1_JC];
    {
        const float4 h0 = *reinterpret_cast<const float4*>(&sm.hb[stage][j0]);
        const float4 h1 = *reinterpret_cast<const float4*>(&sm.hb[stage][j0 + 4]);
        const float hb[K1_JC] = { h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w };
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) { acc[0][jj] = pack2(hb[jj], hb[jj]); acc[1][jj] = acc[0][jj]; }
    }
    // 16 independent accumulator pairs per k step: consecutive FFMA2 never depend on each other
#pragma unroll
    for (int k4 = 0; k4 < K1_D / 4; ++k4) {
        float4 b[K1_JC];
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) b[jj] = *reinterpret_cast<const float4*>(&sm.b[stage][j0 + jj][4 * k4]);   // warp-uniform address
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) { acc[0][jj] = fma2(ap[0][4 * k4 + 0], pack2(b[jj].x, b[jj].x), acc[0][jj]); acc[1][jj] = fma2(ap[1][4 * k4 + 0], pack2(b[jj].x, b[jj].x), acc[1][jj]); }
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) { acc[0][jj] = fma2(ap[0][4 * k4 + 1], pack2(b[jj].y, b[jj].y), acc[0][jj]); acc[1][jj] = fma2(ap[1][4 * k4 + 1], pack2(b[jj].y, b[jj].y), acc[1][jj]); }
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) { acc[0][jj] = fma2(ap[0][4 * k4 + 2], pack2(b[jj].z, b[jj].z), acc[0][jj]); acc[1][jj] = fma2(ap[1][4 * k4 + 2], pack2(b[jj].z, b[jj].z), acc[1][jj]); }
#pragma unroll
        for (int jj = 0; jj < K1_JC; ++jj) { acc[0][jj] = fma2(ap[0][4 * k4 + 3], pack2(b[jj].w, b[jj].w), acc[0][jj]); acc[1][jj] = fma2(ap[1][4 * k4 + 3], pack2(b[jj].w, b[jj].w), acc[1][jj]); }
    }
    // ---- row direction: value-only chunk maxima; one rarely-taken branch recovers indices for the rows that improved --
    float v[K1_RPT][K1_JC], cm[K1_RPT];
#pragma unroll
    for (int jj = 0; jj < K1_JC; ++jj) { unpack2(acc[0][jj], v[0][jj], v[1][jj]); unpack2(acc[1][jj], v[2][jj], v[3][jj]); }
#pragma unroll
    for (int r = 0; r < K1_RPT; ++r)
        cm[r] = fmaxf(max3(v[r][0], v[r][1], v[r][2]), max3(v[r][3], v[r][4], max3(v[r][5], v[r][6], v[r][7])));
    if ((cm[0] > rbest[0]) | (cm[1] > rbest[1]) | (cm[2] > rbest[2]) | (cm[3] > rbest[3])) {
#pragma unroll
        for (int r = 0; r < K1_RPT; ++r) {
            if (cm[r] > rbest[r]) {
                int sel = K1_JC - 1;
#pragma unroll
                for (int jj = K1_JC - 2; jj >= 0; --jj) sel = (v[r][jj] == cm[r]) ? jj : sel;
                rbest[r] = cm[r]; ridx[r] = col0 + j0 + sel;
            }
        }
    }
    // ---- column direction: add -|a_i|^2/2, one warp-wide max per column (8 independent CREDUX chains), the lanes that
    //      hold the maximum publish (value key << 32 | ~row) with a predicated 64-bit RED.MAX -------------------------------
    float m[K1_JC], wm[K1_JC];
#pragma unroll
    for (int jj = 0; jj < K1_JC; ++jj) {
        acc[0][jj] = add2(acc[0][jj], hap0);
        acc[1][jj] = add2(acc[1][jj], hap1);
        unpack2(acc[0][jj], v[0][jj], v[1][jj]); unpack2(acc[1][jj], v[2][jj], v[3][jj]);
        m[jj] = fmaxf(max3(v[0][jj], v[1][jj], v[2][jj]), v[3][jj]);
    }
#pragma unroll
    for (int jj = 0; jj < K1_JC; ++jj) wm[jj] = warp_max(m[jj]);
#pragma unroll
    for (int jj = 0; jj < K1_JC; ++jj) {
        if (FULL || j0 + jj < ncols) {
            uint32_t lo = nrow[3];                                  // lowest row of this thread that attains the maximum
            lo = (v[2][jj] == wm[jj]) ? nrow[2] : lo;
            lo = (v[1][jj] == wm[jj]) ? nrow[1] : lo;
            lo = (v[0][jj] == wm[jj]) ? nrow[0] : lo;
            red_max_u64_if_eq(colp + col0 + j0 + jj, ((unsigned long long)float_key(wm[jj]) << 32) | lo, m[jj], wm[jj]);
        }
    }
}

__global__ void __launch_bounds__(K1_THREADS, 2)
k1_mutual_nn_kernel(const float* __restrict__ src, const float* __restrict__ tgt,
                    const int32_t* __restrict__ src_off, const int32_t* __restrict__ tgt_off,
                    const float* __restrict__ hna, const float* __restrict__ hnb, int padM, int padN,
                    unsigned long long* __restrict__ row_packed, unsigned long long* __restrict__ col_packed, int splits, int blk0)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    K1Smem& sm = *reinterpret_cast<K1Smem*>(smem_raw);

    const int p = blockIdx.z;
    const int so = src_off[p], M = min(src_off[p + 1] - so, padM);     // a pair larger than the caller's max_M / max_N bound is truncated to it,
    const int to = tgt_off[p], N = min(tgt_off[p + 1] - to, padN);     // never read past its workspace slice
    const int row0 = (blk0 + (int)blockIdx.x) * K1_ROWS;
    if (row0 >= M || N <= 0) return;
    const int ntiles = (N + K1_TILE - 1) / K1_TILE;
    const int t_begin = (int)(((long long)blockIdx.y * ntiles) / splits);
    const int t_end = (int)(((long long)(blockIdx.y + 1) * ntiles) / splits);
    if (t_begin >= t_end) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int active_warps = min(K1_WARPS, (M - row0 + K1_ROWS_PER_WARP - 1) / K1_ROWS_PER_WARP);

    // zero the ring once so that rows a partial tile does not overwrite stay finite
    for (int i = threadIdx.x; i < K1_STAGES * K1_TILE * K1_D / 4; i += K1_THREADS)
        reinterpret_cast<float4*>(&sm.b[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < K1_STAGES; ++s) { mbar_init(&sm.full[s], 1); sm.done[s] = 0; }
        mbar_fence_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (warp >= active_warps) return;

    const float* tgt_p = tgt + (size_t)to * K1_D;
    const float* hnb_p = hnb + (size_t)p * padN;
    auto issue_tile = [&](int t, int stage) {
        const int nrows = min(K1_TILE, N - t * K1_TILE);
        mbar_expect_tx(&sm.full[stage], (uint32_t)(nrows * K1_D * 4 + K1_TILE * 4));
        tma_load_1d(&sm.b[stage][0][0], tgt_p + (size_t)t * K1_TILE * K1_D, (uint32_t)(nrows * K1_D * 4), &sm.full[stage]);
        tma_load_1d(&sm.hb[stage][0], hnb_p + (size_t)t * K1_TILE, (uint32_t)(K1_TILE * 4), &sm.full[stage]);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < K1_STAGES && t_begin + s < t_end; ++s) issue_tile(t_begin + s, s);

    // ---- this thread's 4 source rows, packed as two row pairs (stationary for the whole kernel) ------------------
    const int i0 = row0 + warp * K1_ROWS_PER_WARP + lane * K1_RPT;
    f32x2 ap[2][K1_D];
#pragma unroll
    for (int rp = 0; rp < 2; ++rp) {
        const int ia = i0 + 2 * rp, ib = ia + 1;
        const float4* ra = reinterpret_cast<const float4*>(src + (size_t)(so + ia) * K1_D);
        const float4* rb = reinterpret_cast<const float4*>(src + (size_t)(so + ib) * K1_D);
#pragma unroll
        for (int k4 = 0; k4 < K1_D / 4; ++k4) {
            const float4 x = (ia < M) ? __ldg(ra + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 y = (ib < M) ? __ldg(rb + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
            ap[rp][4 * k4 + 0] = pack2(x.x, y.x); ap[rp][4 * k4 + 1] = pack2(x.y, y.y);
            ap[rp][4 * k4 + 2] = pack2(x.z, y.z); ap[rp][4 * k4 + 3] = pack2(x.w, y.w);
        }
    }
    const float* hna_p = hna + (size_t)p * padM;
    const f32x2 hap0 = pack2(hna_p[i0], hna_p[i0 + 1]), hap1 = pack2(hna_p[i0 + 2], hna_p[i0 + 3]);   // -inf beyond M

    float rbest[K1_RPT]; int ridx[K1_RPT]; uint32_t nrow[K1_RPT];           // nrow = low word of the packed best: ~row index
#pragma unroll
    for (int r = 0; r < K1_RPT; ++r) { rbest[r] = -INFINITY; ridx[r] = 0; nrow[r] = 0xFFFFFFFFu - (uint32_t)(i0 + r); }
    unsigned long long* colp = col_packed + (size_t)p * padN;

#pragma unroll 1
    for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, stage = it % K1_STAGES;
        mbar_wait(&sm.full[stage], (uint32_t)(it / K1_STAGES) & 1u);
        const int ncols = min(K1_TILE, N - t * K1_TILE), col0 = t * K1_TILE;
        if (ncols == K1_TILE) {
#pragma unroll 1
            for (int j0 = 0; j0 < K1_TILE; j0 += K1_JC) k1_chunk<true>(sm, stage, j0, ncols, col0, ap, hap0, hap1, rbest, ridx, colp, nrow);
        } else {
#pragma unroll 1
            for (int j0 = 0; j0 < ncols; j0 += K1_JC) k1_chunk<false>(sm, stage, j0, ncols, col0, ap, hap0, hap1, rbest, ridx, colp, nrow);
        }
        // release the stage; the last warp out re-arms it with tile t + STAGES (no warp ever waits for a free slot)
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&sm.done[stage], 1) == active_warps - 1) {
                sm.done[stage] = 0;
                if (t + K1_STAGES < t_end) { fence_proxy_async(); issue_tile(t + K1_STAGES, stage); }
            }
        }
    }
    unsigned long long* rowp = row_packed + (size_t)p * padM;
#pragma unroll
    for (int r = 0; r < K1_RPT; ++r)
        if (i0 + r < M) red_max_u64(rowp + i0 + r, pack_best(float_key(rbest[r]), (uint32_t)ridx[r]));
}

// ---- select: decode packed bests, mutual check, ascending compaction, optional gather of matched keypoints --------
// Mirrors models/BUFFER.py:356-357 (s_mids ascending, t_mids = nn_s[s_mids]) and :284,:287 (ss_kpts = kpts1[s_mids],
// tt_kpts = kpts2[t_mids]) written as 8-float records {sx sy sz 0 qx qy qz 0}.  Two phases so that one huge pair (BASELINE config 5:
// 100k x 100k) is not decoded by a single CTA: k1_decode_kernel (one CTA per 1024-row block) writes nn / distances and the number of
// mutual matches of its block; k1_compact_kernel adds up the counts of the blocks before its own (<= a few hundred ints) and scatters.
constexpr int SEL_THREADS = 256;
constexpr int SEL_ROWS = 1024;                 // rows per CTA (4 per thread)

BFR_DEVINL bool mutual_flag(const unsigned long long* __restrict__ rowp, const unsigned long long* __restrict__ colp, int i, int M, int N, uint32_t& j)
{
    if (i >= M) { j = 0; return false; }
    j = packed_index(rowp[i]);
    return (N > 0) && (j < (uint32_t)N) && (packed_index(colp[j]) == (uint32_t)i);
}

__global__ void __launch_bounds__(SEL_THREADS) k1_decode_kernel(const int32_t* __restrict__ src_off, const int32_t* __restrict__ tgt_off,
                                                                const unsigned long long* __restrict__ row_packed, const unsigned long long* __restrict__ col_packed,
                                                                const float* __restrict__ hna, int padM, int padN, int nblk,
                                                                int64_t* __restrict__ nn_s, int64_t* __restrict__ nn_t, float* __restrict__ d_s, float* __restrict__ d_t,
                                                                int32_t* __restrict__ blk_cnt)
{
    __shared__ int warp_cnt[SEL_THREADS / 32];
    const int p = blockIdx.y, b = blockIdx.x;
    const int so = src_off[p], M = min(src_off[p + 1] - so, padM), to = tgt_off[p], N = min(tgt_off[p + 1] - to, padN);
    const unsigned long long* rowp = row_packed + (size_t)p * padM;
    const unsigned long long* colp = col_packed + (size_t)p * padN;
    if (nn_t || d_t)
        for (int j = b * SEL_ROWS + threadIdx.x; j < min(N, (b + 1) * SEL_ROWS); j += SEL_THREADS) {
            const unsigned long long c = colp[j];
            if (nn_t) nn_t[to + j] = (int64_t)packed_index(c);
            if (d_t) { const float d2 = __fmul_rn(-2.0f, key_float((uint32_t)(c >> 32))); d_t[to + j] = __fsqrt_rn(d2 > 0.0f ? d2 : 0.0f); }
        }
    int cnt = 0;
    for (int i = b * SEL_ROWS + threadIdx.x; i < min(M, (b + 1) * SEL_ROWS); i += SEL_THREADS) {
        uint32_t j;
        cnt += mutual_flag(rowp, colp, i, M, N, j) ? 1 : 0;
        if (nn_s) nn_s[so + i] = (int64_t)j;
        if (d_s) { const float d2 = __fmul_rn(-2.0f, __fadd_rn(key_float((uint32_t)(rowp[i] >> 32)), hna[(size_t)p * padM + i])); d_s[so + i] = __fsqrt_rn(d2 > 0.0f ? d2 : 0.0f); }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) tot += warp_cnt[w];
        if (b < nblk) blk_cnt[(size_t)p * nblk + b] = tot;            // (blocks beyond the source side only decode target rows)
    }
}

__global__ void __launch_bounds__(SEL_THREADS) k1_compact_kernel(const int32_t* __restrict__ src_off, const int32_t* __restrict__ tgt_off,
                                                                 const unsigned long long* __restrict__ row_packed, const unsigned long long* __restrict__ col_packed,
                                                                 int padM, int padN, int nblk, const int32_t* __restrict__ blk_cnt,
                                                                 const float* __restrict__ src_xyz, const float* __restrict__ tgt_xyz,
                                                                 int64_t* __restrict__ s_mids, int64_t* __restrict__ t_mids, int32_t* __restrict__ n_mutual,
                                                                 float4* __restrict__ corr)
{
    __shared__ int warp_cnt[SEL_THREADS / 32];
    __shared__ int base_s;
    const int p = blockIdx.y, b = blockIdx.x;
    const int so = src_off[p], M = min(src_off[p + 1] - so, padM), to = tgt_off[p], N = min(tgt_off[p + 1] - to, padN);
    const unsigned long long* rowp = row_packed + (size_t)p * padM;
    const unsigned long long* colp = col_packed + (size_t)p * padN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // matches in the blocks before this one (and, for the last block, the pair's total)
        int before = 0, total = 0;
        for (int k = threadIdx.x; k < nblk; k += SEL_THREADS) { const int c = blk_cnt[(size_t)p * nblk + k]; total += c; if (k < b) before += c; }
        before = __reduce_add_sync(0xffffffffu, before); total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0) { warp_cnt[warp] = before; }
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < SEL_THREADS / 32; ++w) s += warp_cnt[w]; base_s = s; }
        __syncthreads();
        if (n_mutual && b == nblk - 1) {
            if (lane == 0) warp_cnt[warp] = total;
            __syncthreads();
            if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < SEL_THREADS / 32; ++w) s += warp_cnt[w]; n_mutual[p] = s; }
        }
        __syncthreads();
    }
    if (!s_mids && !corr) return;
    for (int i0 = b * SEL_ROWS; i0 < min(M, (b + 1) * SEL_ROWS); i0 += SEL_THREADS) {
        const int i = i0 + threadIdx.x;
        uint32_t j;
        const bool flag = mutual_flag(rowp, colp, i, M, N, j);
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int pre = base_s, tot = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { const int c = warp_cnt[w]; if (w < warp) pre += c; tot += c; }
        if (flag) {
            const int pos = pre + __popc(bal & ((1u << lane) - 1u));
            if (s_mids) { s_mids[so + pos] = (int64_t)i; t_mids[so + pos] = (int64_t)j; }
            if (corr) {
                const float* s = src_xyz + (size_t)(so + i) * 3; const float* q = tgt_xyz + (size_t)(to + j) * 3;
                corr[2 * (size_t)(so + pos)] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.0f);
                corr[2 * (size_t)(so + pos) + 1] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.0f);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base_s += tot;
        __syncthreads();
    }
}

// gather correspondences given explicit index pairs (the Open3D-style API: pcd0, pcd1, corr)
__global__ void gather_corr_kernel(const float* __restrict__ src_xyz, const float* __restrict__ tgt_xyz, const int64_t* __restrict__ s_ids,
                                   const int64_t* __restrict__ t_ids, int K, float4* __restrict__ corr)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    const float* s = src_xyz + (size_t)s_ids[c] * 3; const float* q = tgt_xyz + (size_t)t_ids[c] * 3;
    corr[2 * (size_t)c] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.0f);
    corr[2 * (size_t)c + 1] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.0f);
}

// ---- measurement aids ------------------------------------------------------------------------------------------
// Pure FFMA2 stream (16 independent packed accumulators per thread): the FP32 issue peak the K1 roofline is
// normalised against, measured live by bench.py on the GPU it runs on (MEASURED_PEAKS.json has no FP32 figure).
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* __restrict__ out, const float* __restrict__ in, int iters)
{
    f32x2 acc[16], a[4]; float b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = pack2(in[(threadIdx.x + i) & 63], in[(threadIdx.x + i + 7) & 63]);
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = pack2(in[64 + i + (threadIdx.x & 1)], in[68 + i + (threadIdx.x & 1)]); b[i] = in[72 + i + (threadIdx.x & 1)]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma2(a[(i + u) & 3], pack2(b[u], b[u]), acc[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; unpack2(acc[i], x, y); s += x + y; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// flops executed by one probe launch = grid * 256 threads * iters * 64 FFMA2 * 4
cudaError_t fp32_probe_launch(int grid, int iters, float* scratch, cudaStream_t stream)
{
    // scratch: >= grid*256 + 128 floats; the first 128 are the (arbitrary, finite) inputs
    fp32_probe_kernel<<<grid, 256, 0, stream>>>(scratch + 128, scratch, iters);
    return cudaGetLastError();
}

static thread_local cudaEvent_t g_k1_ev0 = nullptr, g_k1_ev1 = nullptr;
void k1_set_events(cudaEvent_t e0, cudaEvent_t e1) { g_k1_ev0 = e0; g_k1_ev1 = e1; }

// ---- host launchers -------------------------------------------------------------------------------------------
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

int k1_pad_rows(int max_rows) { return round_up(max_rows > 0 ? max_rows : 1, K1_ROWS); }   // multiple of 512 (and of 64)

// workspace carve-up: packed u64 arrays first (one contiguous region: what a multi-GPU row split max-reduces), then the float arrays
// (256-byte aligned for TMA), the f16 copies of the tensor-core path, the range flags and the per-block match counts of the select phase
struct K1Ws {
    unsigned long long* row_packed; unsigned long long* col_packed; float* hna; float* hnb; uint4* src_h; uint4* tgt_h; int32_t* oor; int32_t* blk_cnt;
    int padM, padN, sel_blocks;
};
static inline int k1_sel_blocks(int padM) { return (padM + SEL_ROWS - 1) / SEL_ROWS; }

static K1Ws k1_carve(void* ws, int P, int max_M, int max_N)
{
    K1Ws k;
    k.padM = k1_pad_rows(max_M); k.padN = k1_pad_rows(max_N); k.sel_blocks = k1_sel_blocks(k.padM);
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    k.row_packed = reinterpret_cast<unsigned long long*>(w); w += (size_t)P * k.padM * 8;
    k.col_packed = reinterpret_cast<unsigned long long*>(w); w += (size_t)P * k.padN * 8;
    k.hna = reinterpret_cast<float*>(w); w += (size_t)P * k.padM * 4;
    k.hnb = reinterpret_cast<float*>(w); w += (size_t)P * k.padN * 4;
    w = reinterpret_cast<unsigned char*>(((uintptr_t)w + 255) & ~(uintptr_t)255);
    k.src_h = reinterpret_cast<uint4*>(w); w += (size_t)P * k.padM * K1_D * 2;          // total_M <= P * padM rows
    k.tgt_h = reinterpret_cast<uint4*>(w); w += (size_t)P * k.padN * K1_D * 2;
    w = reinterpret_cast<unsigned char*>(((uintptr_t)w + 255) & ~(uintptr_t)255);
    k.oor = reinterpret_cast<int32_t*>(w); w += (size_t)P * 8;                           // [2][P]: source side, target side
    w = reinterpret_cast<unsigned char*>(((uintptr_t)w + 255) & ~(uintptr_t)255);
    k.blk_cnt = reinterpret_cast<int32_t*>(w);
    return k;
}

size_t k1_workspace_bytes(int P, int max_M, int max_N)
{
    const size_t padM = (size_t)k1_pad_rows(max_M), padN = (size_t)k1_pad_rows(max_N);
    // packed bests + half norms + (tensor-core path) f16 copies of both descriptor sets (<= P * pad rows of 64 bytes each) + range flags + block counts
    return (size_t)P * (padM + padN) * (sizeof(float) + sizeof(unsigned long long) + (size_t)K1_D * 2) + (size_t)P * 8 +
           (size_t)P * k1_sel_blocks((int)padM) * 4 + 2048;
}

void k1_packed_view(void* ws, int P, int max_M, int max_N, unsigned long long** packed, size_t* count)
{
    const K1Ws k = k1_carve(ws, P, max_M, max_N);
    *packed = k.row_packed;                                           // row_packed [P * padM] immediately followed by col_packed [P * padN]
    *count = (size_t)P * ((size_t)k.padM + (size_t)k.padN);
}

static thread_local int g_k1_algo = 1;          // 0 = FP32 FFMA2 kernel, 1 = tensor-core filter + exact re-check (mutual_nn_tc.cu); per calling thread
void k1_set_algo(int algo) { g_k1_algo = algo; }
int k1_get_algo() { return g_k1_algo; }

cudaError_t ensure_dyn_smem(const void* fn, int bytes, std::atomic<unsigned long long>& done)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);   // a per-device attribute; setting it twice is harmless
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// phase 1 + 2: half norms / f16 copies / zeroed packed bests for ALL rows, then the main kernel on row-block partition `part` of `nparts`
// (both directions are partitioned by row blocks of their own side).  nparts = 1: the whole job.
cudaError_t k1_partial_launch(const float* src, const float* tgt, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                              long long total_M, long long total_N, int D, int col_splits, int part, int nparts, void* ws, cudaStream_t stream)
{
    if (D != K1_D || nparts < 1 || part < 0 || part >= nparts) return cudaErrorInvalidValue;
    const K1Ws k = k1_carve(ws, P, max_M, max_N);
    const bool tc = g_k1_algo == 1 && max_M > 0 && max_N > 0 && k1_tc_supported(D, total_M, total_N);
    {
        const long long na = (long long)P * k.padM, nb = (long long)P * k.padN;
        if (tc) { cudaError_t e = cudaMemsetAsync(k.oor, 0, (size_t)P * 8, stream); if (e != cudaSuccess) return e; }
        k1_prep_kernel<<<(unsigned)((na + 255) / 256), 256, 0, stream>>>(src, src_off, P, D, k.padM, k.hna, k.row_packed, tc ? k.src_h : nullptr, k.oor);
        k1_prep_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(tgt, tgt_off, P, D, k.padN, k.hnb, k.col_packed, tc ? k.tgt_h : nullptr, k.oor + P);
    }
    if (col_splits < 1) col_splits = 1;
    if (g_k1_ev0) cudaEventRecord(g_k1_ev0, stream);
    if (max_M > 0 && max_N > 0) {
        if (tc) {
            cudaError_t e = k1_tc_launch(src, tgt, k.src_h, k.tgt_h, k.oor, src_off, tgt_off, P, max_M, max_N, total_M, total_N, k.hna, k.hnb, k.padM, k.padN,
                                         k.row_packed, k.col_packed, part, nparts, stream);
            if (e != cudaSuccess) return e;
        } else {
            static std::atomic<unsigned long long> attr_done{0};
            cudaError_t e = ensure_dyn_smem((const void*)k1_mutual_nn_kernel, (int)sizeof(K1Smem), attr_done);
            if (e != cudaSuccess) return e;
            const int nblk = (max_M + K1_ROWS - 1) / K1_ROWS;
            const int b0 = (int)(((long long)part * nblk) / nparts), b1 = (int)(((long long)(part + 1) * nblk) / nparts);
            if (b1 > b0) {
                dim3 grid((unsigned)(b1 - b0), (unsigned)col_splits, (unsigned)P);
                k1_mutual_nn_kernel<<<grid, K1_THREADS, sizeof(K1Smem), stream>>>(src, tgt, src_off, tgt_off, k.hna, k.hnb, k.padM, k.padN, k.row_packed, k.col_packed,
                                                                                  col_splits, b0);
            }
        }
    }
    if (g_k1_ev1) cudaEventRecord(g_k1_ev1, stream);
    return cudaGetLastError();
}

// phase 3: decode the (complete) packed bests
cudaError_t k1_select_launch(const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N, void* ws,
                             int64_t* nn_s, int64_t* nn_t, float* d_s, float* d_t, const float* src_xyz, const float* tgt_xyz,
                             int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr, cudaStream_t stream)
{
    const K1Ws k = k1_carve(ws, P, max_M, max_N);
    const int nblk_m = k.sel_blocks, nblk_n = k1_sel_blocks(k.padN);
    dim3 grid((unsigned)(nblk_m > nblk_n ? nblk_m : nblk_n), (unsigned)P);   // target rows beyond the source block count are decoded by extra blocks
    k1_decode_kernel<<<grid, SEL_THREADS, 0, stream>>>(src_off, tgt_off, k.row_packed, k.col_packed, k.hna, k.padM, k.padN, nblk_m, nn_s, nn_t, d_s, d_t, k.blk_cnt);
    if (s_mids || corr || n_mutual) {
        dim3 grid2((unsigned)nblk_m, (unsigned)P);
        k1_compact_kernel<<<grid2, SEL_THREADS, 0, stream>>>(src_off, tgt_off, k.row_packed, k.col_packed, k.padM, k.padN, nblk_m, k.blk_cnt, src_xyz, tgt_xyz,
                                                             s_mids, t_mids, n_mutual, reinterpret_cast<float4*>(corr));
    }
    return cudaGetLastError();
}

cudaError_t k1_launch(const float* src, const float* tgt, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                      long long total_M, long long total_N, int D, int col_splits, void* ws, int64_t* nn_s, int64_t* nn_t, float* d_s, float* d_t,
                      const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr,
                      cudaStream_t stream)
{
    cudaError_t e = k1_partial_launch(src, tgt, src_off, tgt_off, P, max_M, max_N, total_M, total_N, D, col_splits, 0, 1, ws, stream);
    if (e != cudaSuccess) return e;
    return k1_select_launch(src_off, tgt_off, P, max_M, max_N, ws, nn_s, nn_t, d_s, d_t, src_xyz, tgt_xyz, s_mids, t_mids, n_mutual, corr, stream);
}

cudaError_t gather_corr_launch(const float* src_xyz, const float* tgt_xyz, const int64_t* s_ids, const int64_t* t_ids, int K, float* corr, cudaStream_t stream)
{
    if (K > 0) gather_corr_kernel<<<(K + 255) / 256, 256, 0, stream>>>(src_xyz, tgt_xyz, s_ids, t_ids, K, reinterpret_cast<float4*>(corr));
    return cudaGetLastError();
}

}  // namespace bfr
