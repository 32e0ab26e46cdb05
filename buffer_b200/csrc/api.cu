// api.cu — the extern "C" boundary of libbuffer_b200.so (declared in include/buffer_b200.h).
#include "../../include/buffer_b200.h"
#include "bfr_kernels.h"

using namespace bfr;

namespace {

inline int cu(cudaError_t e) { return e == cudaSuccess ? BFR_OK : (int)e; }
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
inline cudaStream_t st(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__global__ void uniform_offsets_kernel(int32_t* off_m, int32_t* off_n, int P, int M, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= P) { off_m[i] = i * M; off_n[i] = i * N; }
}

// carve-up of the bfr_register_batched workspace
struct RegWs {
    void* k1; float* corr; unsigned long long* best; float* T0; void* rs;
};
inline size_t reg_ws_bytes(int P, int max_M, int max_N, int total_M)
{
    return up256(k1_workspace_bytes(P, max_M, max_N)) + up256((size_t)(total_M > 0 ? total_M : 1) * 32) + up256((size_t)P * 8) + up256((size_t)P * 64) +
           up256(ransac_scratch_bytes()) + 256;
}
inline RegWs reg_ws_carve(void* ws, int P, int max_M, int max_N, int total_M)
{
    unsigned char* w = reinterpret_cast<unsigned char*>(up256((size_t)(uintptr_t)ws));
    RegWs r;
    r.k1 = w; w += up256(k1_workspace_bytes(P, max_M, max_N));
    r.corr = reinterpret_cast<float*>(w); w += up256((size_t)(total_M > 0 ? total_M : 1) * 32);
    r.best = reinterpret_cast<unsigned long long*>(w); w += up256((size_t)P * 8);
    r.T0 = reinterpret_cast<float*>(w); w += up256((size_t)P * 64);
    r.rs = w;
    return r;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int bfr_version(void) { return 200; }

const char* bfr_error_string(int code)
{
    switch (code) {
        case BFR_OK: return "ok";
        case BFR_E_NULL: return "required pointer is NULL";
        case BFR_E_SIZE: return "negative or inconsistent size";
        case BFR_E_DIM: return "unsupported descriptor length (this build supports D == 32)";
        case BFR_E_WORKSPACE: return "workspace too small";
        case BFR_E_ALIGN: return "pointer not 16-byte aligned";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

size_t bfr_mutual_nn_workspace_bytes(int P, int max_M, int max_N) { return k1_workspace_bytes(P < 0 ? 0 : P, max_M, max_N); }

int bfr_mutual_matching_batched(const float* src_des, const float* tgt_des, const int32_t* src_off, const int32_t* tgt_off,
                                int P, int max_M, int max_N, int total_M, int total_N, int D, int col_splits,
                                int64_t* nn_s, int64_t* nn_t, float* dist_s, float* dist_t,
                                const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr_xyz,
                                void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!src_des || !tgt_des || !src_off || !tgt_off || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_M < 0 || max_N < 0 || total_M < 0 || total_N < 0) return BFR_E_SIZE;
    if (D != 32) return BFR_E_DIM;
    if (!aligned16(src_des) || !aligned16(tgt_des)) return BFR_E_ALIGN;
    if (ws_bytes < k1_workspace_bytes(P, max_M, max_N)) return BFR_E_WORKSPACE;
    if (corr_xyz && (!src_xyz || !tgt_xyz)) return BFR_E_NULL;
    if ((s_mids == nullptr) != (t_mids == nullptr)) return BFR_E_NULL;
    return cu(k1_launch(src_des, tgt_des, src_off, tgt_off, P, max_M, max_N, total_M, total_N, D, col_splits, ws, nn_s, nn_t, dist_s, dist_t,
                        src_xyz, tgt_xyz, s_mids, t_mids, n_mutual, corr_xyz, st(stream)));
}

int bfr_mutual_nn_partial(const float* src_des, const float* tgt_des, const int32_t* src_off, const int32_t* tgt_off,
                          int P, int max_M, int max_N, int total_M, int total_N, int D, int col_splits, int part, int nparts,
                          void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!src_des || !tgt_des || !src_off || !tgt_off || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_M < 0 || max_N < 0 || total_M < 0 || total_N < 0 || nparts < 1 || part < 0 || part >= nparts) return BFR_E_SIZE;
    if (D != 32) return BFR_E_DIM;
    if (!aligned16(src_des) || !aligned16(tgt_des)) return BFR_E_ALIGN;
    if (ws_bytes < k1_workspace_bytes(P, max_M, max_N)) return BFR_E_WORKSPACE;
    return cu(k1_partial_launch(src_des, tgt_des, src_off, tgt_off, P, max_M, max_N, total_M, total_N, D, col_splits, part, nparts, ws, st(stream)));
}

int bfr_mutual_nn_packed(void* ws, size_t ws_bytes, int P, int max_M, int max_N, uint64_t** packed, size_t* count)
{
    if (!ws || !packed || !count) return BFR_E_NULL;
    if (P < 0 || max_M < 0 || max_N < 0) return BFR_E_SIZE;
    if (ws_bytes < k1_workspace_bytes(P, max_M, max_N)) return BFR_E_WORKSPACE;
    unsigned long long* ptr = nullptr;
    k1_packed_view(ws, P, max_M, max_N, &ptr, count);
    *packed = reinterpret_cast<uint64_t*>(ptr);
    return BFR_OK;
}

int bfr_mutual_select(const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                      int64_t* nn_s, int64_t* nn_t, float* dist_s, float* dist_t,
                      const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr_xyz,
                      void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!src_off || !tgt_off || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_M < 0 || max_N < 0) return BFR_E_SIZE;
    if (ws_bytes < k1_workspace_bytes(P, max_M, max_N)) return BFR_E_WORKSPACE;
    if (corr_xyz && (!src_xyz || !tgt_xyz)) return BFR_E_NULL;
    if ((s_mids == nullptr) != (t_mids == nullptr)) return BFR_E_NULL;
    return cu(k1_select_launch(src_off, tgt_off, P, max_M, max_N, ws, nn_s, nn_t, dist_s, dist_t, src_xyz, tgt_xyz, s_mids, t_mids, n_mutual, corr_xyz, st(stream)));
}

int bfr_gather_corr(const float* src_xyz, const float* tgt_xyz, const int64_t* s_ids, const int64_t* t_ids, int K, float* corr_xyz, void* stream)
{
    if (K == 0) return BFR_OK;
    if (!src_xyz || !tgt_xyz || !s_ids || !t_ids || !corr_xyz) return BFR_E_NULL;
    if (K < 0) return BFR_E_SIZE;
    return cu(gather_corr_launch(src_xyz, tgt_xyz, s_ids, t_ids, K, corr_xyz, st(stream)));
}

size_t bfr_ransac_workspace_bytes(void) { return ransac_scratch_bytes(); }

int bfr_ransac_batched(const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                       uint64_t seed, uint32_t pair_id_base, uint32_t h_begin, uint32_t h_end,
                       float dist_th, float similar_th, float confidence, int splits, uint64_t* best_packed, int32_t* valid_count,
                       void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!corr_xyz || !corr_off || !corr_cnt || !best_packed) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || h_end < h_begin) return BFR_E_SIZE;
    if (!aligned16(corr_xyz)) return BFR_E_ALIGN;
    return cu(ransac_launch(corr_xyz, corr_off, corr_cnt, P, seed, pair_id_base, h_begin, h_end, dist_th, similar_th, confidence, splits,
                            reinterpret_cast<unsigned long long*>(best_packed), valid_count, ws, ws ? ws_bytes : 0, st(stream)));
}

int bfr_ransac_finalize_batched(const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                                uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, const uint64_t* best_packed,
                                float* T, int32_t* inliers, int64_t* best_h, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!corr_xyz || !corr_off || !corr_cnt || !best_packed || !T) return BFR_E_NULL;
    if (P < 0) return BFR_E_SIZE;
    return cu(ransac_finalize_launch(corr_xyz, corr_off, corr_cnt, P, seed, pair_id_base, dist_th, similar_th,
                                     reinterpret_cast<const unsigned long long*>(best_packed), T, inliers, best_h, st(stream)));
}

int bfr_lrf_hypotheses(const float* cs, const float* ss_R, const float* tt_R, const float* ss_kpts, const float* tt_kpts, int A,
                       float* R_out, float* t_out, void* stream)
{
    if (A == 0) return BFR_OK;
    if (!cs || !ss_R || !tt_R || !ss_kpts || !tt_kpts || !R_out || !t_out) return BFR_E_NULL;
    if (A < 0) return BFR_E_SIZE;
    return cu(lrf_hypotheses_launch(cs, ss_R, tt_R, ss_kpts, tt_kpts, A, R_out, t_out, st(stream)));
}

size_t bfr_score_workspace_bytes(int C) { return score_workspace_bytes(C); }

int bfr_score_hypotheses(const float* R, const float* t, int H, const float* src, const float* tgt, int C, const float* thr, float thr_scalar,
                         int32_t* counts, uint64_t* best_packed, int64_t* best_idx, uint8_t* mask, void* ws, size_t ws_bytes, void* stream)
{
    if (!best_packed || !ws) return BFR_E_NULL;
    if (H < 0 || C < 0) return BFR_E_SIZE;
    if (H > 0 && (!R || !t)) return BFR_E_NULL;
    if (C > 0 && (!src || !tgt)) return BFR_E_NULL;
    if (ws_bytes < score_workspace_bytes(C)) return BFR_E_WORKSPACE;
    return cu(score_hypotheses_launch(R, t, H, src, tgt, C, thr, thr_scalar, counts, reinterpret_cast<unsigned long long*>(best_packed),
                                      best_idx, mask, ws, st(stream)));
}

namespace {
struct VoteWs { unsigned long long* vote_best; unsigned long long* best; int32_t* sub_cnt; float* T0; float* sub_corr; void* rs; };
inline size_t vote_ws_bytes(int P, int total_rows)
{
    return 2 * up256((size_t)P * 8) + up256((size_t)P * 4) + up256((size_t)P * 64) + up256((size_t)(total_rows > 0 ? total_rows : 1) * 32) +
           (total_rows > 0 ? up256(ransac_scratch_bytes()) : 0) + 256;
}
inline VoteWs vote_ws_carve(void* ws, int P, int total_rows = 0)
{
    unsigned char* w = reinterpret_cast<unsigned char*>(up256((size_t)(uintptr_t)ws));
    VoteWs v;
    v.vote_best = reinterpret_cast<unsigned long long*>(w); w += up256((size_t)P * 8);
    v.best = reinterpret_cast<unsigned long long*>(w); w += up256((size_t)P * 8);
    v.sub_cnt = reinterpret_cast<int32_t*>(w); w += up256((size_t)P * 4);
    v.T0 = reinterpret_cast<float*>(w); w += up256((size_t)P * 64);
    v.sub_corr = reinterpret_cast<float*>(w); w += up256((size_t)(total_rows > 0 ? total_rows : 1) * 32);
    v.rs = total_rows > 0 ? w : nullptr;
    return v;
}
}  // namespace

size_t bfr_vote_workspace_bytes(int P, int total_rows) { return vote_ws_bytes(P < 0 ? 0 : P, total_rows); }

int bfr_lrf_vote_batched(float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, int total_rows,
                         const float* ind, const float* ss_R, const float* tt_R, float azi_n, float inlier_th,
                         int32_t* inlier_num, int64_t* best_ind, float* sub_corr, int32_t* sub_cnt, int64_t* inlier_ind,
                         void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!corr_xyz || !corr_off || !corr_cnt || !ind || !ss_R || !tt_R || !sub_corr || !sub_cnt || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_count < 0 || total_rows < 0 || !(azi_n > 0.0f)) return BFR_E_SIZE;
    if (!aligned16(corr_xyz) || !aligned16(sub_corr)) return BFR_E_ALIGN;
    if (ws_bytes < vote_ws_bytes(P, 0)) return BFR_E_WORKSPACE;
    VoteWs v = vote_ws_carve(ws, P);
    return cu(lrf_vote_launch(corr_xyz, corr_off, corr_cnt, P, max_count, ind, ss_R, tt_R, azi_n, inlier_th, inlier_num, v.vote_best, sub_corr, sub_cnt,
                              best_ind, inlier_ind, st(stream)));
}

int bfr_pose_from_votes_batched(float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, int total_rows,
                                const float* ind, const float* ss_R, const float* tt_R, float azi_n, float inlier_th,
                                int hypotheses, uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, float confidence,
                                float refine_thr, int refine_iters, int ransac_splits,
                                float* T_out, int32_t* n_vote_inliers, int32_t* n_inliers, void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!corr_xyz || !corr_off || !corr_cnt || !ind || !ss_R || !tt_R || !T_out || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_count < 0 || total_rows < 0 || hypotheses < 0 || refine_iters < 0 || !(azi_n > 0.0f)) return BFR_E_SIZE;
    if (!aligned16(corr_xyz)) return BFR_E_ALIGN;
    if (ws_bytes < vote_ws_bytes(P, total_rows)) return BFR_E_WORKSPACE;
    cudaStream_t s = st(stream);
    VoteWs v = vote_ws_carve(ws, P, total_rows);
    int32_t* sub_cnt = n_vote_inliers ? n_vote_inliers : v.sub_cnt;
    cudaError_t e = cudaMemsetAsync(v.best, 0, (size_t)P * 8, s);
    if (e != cudaSuccess) return (int)e;
    // models/BUFFER.py:294-311: LRF vote -> inlier_ind (compacted on device) ...
    e = lrf_vote_launch(corr_xyz, corr_off, corr_cnt, P, max_count, ind, ss_R, tt_R, azi_n, inlier_th, nullptr, v.vote_best, v.sub_corr, sub_cnt, nullptr, nullptr, s);
    if (e != cudaSuccess) return (int)e;
    // ... :313-326: RANSAC on that subset ...
    e = ransac_launch(v.sub_corr, corr_off, sub_cnt, P, seed, pair_id_base, 0u, (uint32_t)hypotheses, dist_th, similar_th, confidence, ransac_splits, v.best, nullptr,
                      v.rs, v.rs ? ransac_scratch_bytes() : 0, s);
    if (e != cudaSuccess) return (int)e;
    float* T_ransac = refine_iters > 0 ? v.T0 : T_out;
    e = ransac_finalize_launch(v.sub_corr, corr_off, sub_cnt, P, seed, pair_id_base, dist_th, similar_th, v.best, T_ransac, n_inliers, nullptr, s);
    if (e != cudaSuccess) return (int)e;
    // ... :327-329: post_refinement on ALL mutual matches
    if (refine_iters > 0) e = post_refinement_launch(v.T0, corr_xyz, corr_off, corr_cnt, P, refine_thr, refine_iters, T_out, nullptr, nullptr, max_count, s);
    return cu(e);
}

int bfr_rigid_transform_3d(const float* A, const float* B, const float* w, int bs, int n, float weight_threshold, float* T, void* stream)
{
    if (bs == 0) return BFR_OK;
    if (!A || !B || !T) return BFR_E_NULL;
    if (bs < 0 || n < 0) return BFR_E_SIZE;
    return cu(rigid_transform_launch(A, B, w, bs, n, weight_threshold, T, st(stream)));
}

int bfr_post_refinement_batched(const float* T0, const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                                float thr, int max_iter, int max_count, float* T_out, int32_t* iters, int32_t* inliers, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!T0 || !corr_xyz || !corr_off || !corr_cnt || !T_out) return BFR_E_NULL;
    if (P < 0 || max_iter < 0) return BFR_E_SIZE;
    return cu(post_refinement_launch(T0, corr_xyz, corr_off, corr_cnt, P, thr, max_iter, T_out, iters, inliers, max_count, st(stream)));
}

size_t bfr_register_workspace_bytes(int P, int max_M, int max_N, int total_M, int total_N)
{
    (void)total_N;
    return reg_ws_bytes(P < 0 ? 0 : P, max_M, max_N, total_M);
}

int bfr_register_batched(const float* src_des, const float* src_xyz, const int32_t* src_off,
                         const float* tgt_des, const float* tgt_xyz, const int32_t* tgt_off,
                         int P, int max_M, int max_N, int total_M, int total_N, int D,
                         int hypotheses, uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, float confidence,
                         float refine_thr, int refine_iters, int ransac_splits,
                         float* T_out, int32_t* n_mutual, int32_t* n_inliers, void* ws, size_t ws_bytes, void* stream)
{
    (void)total_N;
    if (P == 0) return BFR_OK;
    if (!src_des || !src_xyz || !src_off || !tgt_des || !tgt_xyz || !tgt_off || !T_out || !n_mutual || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || max_M < 0 || max_N < 0 || total_M < 0 || total_N < 0 || hypotheses < 0 || refine_iters < 0) return BFR_E_SIZE;
    if (D != 32) return BFR_E_DIM;
    if (!aligned16(src_des) || !aligned16(tgt_des)) return BFR_E_ALIGN;
    if (ws_bytes < reg_ws_bytes(P, max_M, max_N, total_M)) return BFR_E_WORKSPACE;
    cudaStream_t s = st(stream);
    RegWs w = reg_ws_carve(ws, P, max_M, max_N, total_M);
    cudaError_t e = cudaMemsetAsync(w.best, 0, (size_t)P * 8, s);
    if (e != cudaSuccess) return (int)e;
    e = k1_launch(src_des, tgt_des, src_off, tgt_off, P, max_M, max_N, total_M, total_N, D, 1, w.k1, nullptr, nullptr, nullptr, nullptr,
                  src_xyz, tgt_xyz, nullptr, nullptr, n_mutual, w.corr, s);
    if (e != cudaSuccess) return (int)e;
    e = ransac_launch(w.corr, src_off, n_mutual, P, seed, pair_id_base, 0u, (uint32_t)hypotheses, dist_th, similar_th, confidence, ransac_splits, w.best, nullptr,
                      w.rs, ransac_scratch_bytes(), s);
    if (e != cudaSuccess) return (int)e;
    float* T_ransac = refine_iters > 0 ? w.T0 : T_out;
    e = ransac_finalize_launch(w.corr, src_off, n_mutual, P, seed, pair_id_base, dist_th, similar_th, w.best, T_ransac, n_inliers, nullptr, s);
    if (e != cudaSuccess) return (int)e;
    if (refine_iters > 0) e = post_refinement_launch(w.T0, w.corr, src_off, n_mutual, P, refine_thr, refine_iters, T_out, nullptr, nullptr, max_M < max_N ? max_M : max_N, s);
    return cu(e);
}

size_t bfr_register_host_workspace_bytes(int P, int M, int N, int D)
{
    if (P < 0 || M < 0 || N < 0 || D < 0) return 0;
    const size_t tm = (size_t)P * M, tn = (size_t)P * N;
    if (tm > (size_t)INT32_MAX || tn > (size_t)INT32_MAX) return 0;      // row offsets are int32 (bfr_register_uniform_host returns BFR_E_SIZE)
    return reg_ws_bytes(P, M, N, (int)tm) + up256(tm * D * 4) + up256(tn * D * 4) + up256(tm * 12) + up256(tn * 12) +
           2 * up256((size_t)(P + 1) * 4) + up256((size_t)P * 64) + 2 * up256((size_t)P * 4) + 512;
}

int bfr_register_uniform_host(const float* src_des_host, const float* src_xyz_host, const float* tgt_des_host, const float* tgt_xyz_host,
                              int P, int M, int N, int D, int hypotheses, uint64_t seed, uint32_t pair_id_base,
                              float dist_th, float similar_th, float confidence, float refine_thr, int refine_iters, int ransac_splits,
                              float* T_out_host, int32_t* n_mutual_host, int32_t* n_inliers_host, void* ws, size_t ws_bytes, void* stream)
{
    if (P == 0) return BFR_OK;
    if (!src_des_host || !src_xyz_host || !tgt_des_host || !tgt_xyz_host || !T_out_host || !ws) return BFR_E_NULL;
    if (P < 0 || P > BFR_MAX_PAIRS || M < 0 || N < 0) return BFR_E_SIZE;
    if ((size_t)P * M > (size_t)INT32_MAX || (size_t)P * N > (size_t)INT32_MAX) return BFR_E_SIZE;     // int32 row offsets: checked before any copy is queued
    if (D != 32) return BFR_E_DIM;
    if (ws_bytes < bfr_register_host_workspace_bytes(P, M, N, D)) return BFR_E_WORKSPACE;
    cudaStream_t s = st(stream);
    const size_t tm = (size_t)P * M, tn = (size_t)P * N;
    unsigned char* w = reinterpret_cast<unsigned char*>(up256((size_t)(uintptr_t)ws));
    float* d_sdes = reinterpret_cast<float*>(w); w += up256(tm * D * 4);
    float* d_tdes = reinterpret_cast<float*>(w); w += up256(tn * D * 4);
    float* d_sxyz = reinterpret_cast<float*>(w); w += up256(tm * 12);
    float* d_txyz = reinterpret_cast<float*>(w); w += up256(tn * 12);
    int32_t* d_offm = reinterpret_cast<int32_t*>(w); w += up256((size_t)(P + 1) * 4);
    int32_t* d_offn = reinterpret_cast<int32_t*>(w); w += up256((size_t)(P + 1) * 4);
    float* d_T = reinterpret_cast<float*>(w); w += up256((size_t)P * 64);
    int32_t* d_nm = reinterpret_cast<int32_t*>(w); w += up256((size_t)P * 4);
    int32_t* d_ni = reinterpret_cast<int32_t*>(w); w += up256((size_t)P * 4);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_sdes, src_des_host, tm * D * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyAsync(d_tdes, tgt_des_host, tn * D * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyAsync(d_sxyz, src_xyz_host, tm * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyAsync(d_txyz, tgt_xyz_host, tn * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
    uniform_offsets_kernel<<<(P + 256) / 256, 256, 0, s>>>(d_offm, d_offn, P, M, N);
    const size_t inner = reg_ws_bytes(P, M, N, (int)tm);
    int rc = bfr_register_batched(d_sdes, d_sxyz, d_offm, d_tdes, d_txyz, d_offn, P, M, N, (int)tm, (int)tn, D, hypotheses, seed, pair_id_base,
                                  dist_th, similar_th, confidence, refine_thr, refine_iters, ransac_splits, d_T, d_nm, d_ni, w, inner, stream);
    if (rc != BFR_OK) return rc;
    if ((e = cudaMemcpyAsync(T_out_host, d_T, (size_t)P * 64, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
    if (n_mutual_host && (e = cudaMemcpyAsync(n_mutual_host, d_nm, (size_t)P * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
    if (n_inliers_host && (e = cudaMemcpyAsync(n_inliers_host, d_ni, (size_t)P * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
    return BFR_OK;
}

int bfr_register_uniform_host_chunked(const float* src_des_host, const float* src_xyz_host, const float* tgt_des_host, const float* tgt_xyz_host,
                                      int P, int M, int N, int D, int chunk_pairs, int hypotheses, uint64_t seed, uint32_t pair_id_base,
                                      float dist_th, float similar_th, float confidence, float refine_thr, int refine_iters, int ransac_splits,
                                      float* T_out_host, int32_t* n_mutual_host, int32_t* n_inliers_host,
                                      void* const* ws, size_t ws_bytes_each, void* const* streams, int n_streams)
{
    if (P == 0) return BFR_OK;
    if (!ws || !streams) return BFR_E_NULL;
    if (P < 0 || chunk_pairs < 1 || n_streams < 1 || M < 0 || N < 0) return BFR_E_SIZE;
    // chunk c goes to stream c % n_streams with that stream's workspace: the copies of one chunk overlap the kernels of the other streams'
    // chunks, and the reuse of a workspace is ordered by its stream
    int c = 0;
    for (int p0 = 0; p0 < P; p0 += chunk_pairs, ++c) {
        const int n = P - p0 < chunk_pairs ? P - p0 : chunk_pairs;
        const int k = c % n_streams;
        const int rc = bfr_register_uniform_host(src_des_host + (size_t)p0 * M * D, src_xyz_host + (size_t)p0 * M * 3, tgt_des_host + (size_t)p0 * N * D,
                                                 tgt_xyz_host + (size_t)p0 * N * 3, n, M, N, D, hypotheses, seed, pair_id_base + (uint32_t)p0, dist_th, similar_th,
                                                 confidence, refine_thr, refine_iters, ransac_splits, T_out_host ? T_out_host + (size_t)p0 * 16 : nullptr,
                                                 n_mutual_host ? n_mutual_host + p0 : nullptr, n_inliers_host ? n_inliers_host + p0 : nullptr,
                                                 ws[k], ws_bytes_each, streams[k]);
        if (rc != BFR_OK) return rc;
    }
    return BFR_OK;
}

size_t bfr_get_matching_indices_workspace_bytes(int N) { return knn3_workspace_bytes(N); }

int bfr_get_matching_indices(const float* source, int N, const float* target, int M, const float* relt_pose, float search_voxel_size,
                             int64_t* match_inds, int32_t* count, int64_t* nn, float* dist, void* ws, size_t ws_bytes, void* stream)
{
    if (!match_inds || !count || !ws || !relt_pose) return BFR_E_NULL;
    if (N < 0 || M < 0) return BFR_E_SIZE;
    if (N > 0 && !source) return BFR_E_NULL;
    if (M > 0 && !target) return BFR_E_NULL;
    if (ws_bytes < knn3_workspace_bytes(N)) return BFR_E_WORKSPACE;
    return cu(get_matching_indices_launch(source, N, target, M, relt_pose, search_voxel_size, match_inds, count, nn, dist, ws, st(stream)));
}

int bfr_furthest_point_sample(const float* xyz, int B, int N, int npoint, int32_t* idx, float* temp, void* stream)
{
    if (B == 0 || npoint == 0) return BFR_OK;
    if (!xyz || !idx || !temp) return BFR_E_NULL;
    if (B < 0 || N <= 0 || npoint < 0) return BFR_E_SIZE;
    return cu(fps_launch(xyz, B, N, npoint, idx, temp, st(stream)));
}

int bfr_svd3_batched(const float* x, int B, float* u, float* s, float* v, void* stream)
{
    if (B == 0) return BFR_OK;
    if (!x || !u || !s || !v) return BFR_E_NULL;
    if (B < 0) return BFR_E_SIZE;
    return cu(svd3_launch(x, B, u, s, v, st(stream)));
}

int bfr_config_set(int key, int value)
{
    if (key == BFR_CFG_K1_ALGO) { if (value != 0 && value != 1) return BFR_E_SIZE; k1_set_algo(value); return BFR_OK; }
    if (key == BFR_CFG_RANSAC_TC) { if (value != 0 && value != 1) return BFR_E_SIZE; ransac_set_tc(value); return BFR_OK; }
    return BFR_E_SIZE;
}

int bfr_config_get(int key)
{
    if (key == BFR_CFG_K1_ALGO) return k1_get_algo();
    if (key == BFR_CFG_RANSAC_TC) return ransac_get_tc();
    return BFR_E_SIZE;
}

int bfr_fp32_probe(int grid, int iters, float* scratch, void* stream)
{
    if (!scratch) return BFR_E_NULL;
    if (grid <= 0 || iters < 0) return BFR_E_SIZE;
    return cu(fp32_probe_launch(grid, iters, scratch, st(stream)));
}

int bfr_debug_set_k1_events(void* ev_start, void* ev_stop)
{
    k1_set_events(reinterpret_cast<cudaEvent_t>(ev_start), reinterpret_cast<cudaEvent_t>(ev_stop));
    return BFR_OK;
}

#pragma GCC visibility pop
}  // extern "C"
