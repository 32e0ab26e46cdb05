/*
 * buffer_b200.h — C ABI of libbuffer_b200.so: the B200-native (sm_100a) correspondence-and-pose back end of BUFFER.
 *
 * This is the drop-in boundary.  The reference (The-Learning-And-Vision-Atelier-LAVA/BUFFER) has no FFI layer of its
 * own for this path — callers reach it through Python attribute lookup — so every entry point below names the
 * reference symbol (file:line under /root/reference) whose work it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *  - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns all memory including the
 *    workspace (`ws`, size from the matching *_workspace_bytes); the library never allocates device memory and every
 *    call is asynchronous on `stream` (no host sync inside) unless stated.  State kept by the library: the per-THREAD
 *    tuning knobs of bfr_config_set / the per-thread debug events (each host thread sees only its own), and one
 *    "shared-memory opt-in done" bit per kernel and device ordinal (idempotent) - calls from several host threads
 *    and on several devices of one process are safe.  Launches go to the CURRENT device of the calling thread: make
 *    the device that owns the buffers current before calling.
 *  - Batches are "varlen": pair p owns rows [off[p], off[p+1]) of the concatenated arrays; offsets are int32 DEVICE
 *    arrays of P+1 entries; max_M / max_N are host-side upper bounds of the per-pair row counts (grid sizing).
 *  - Row-major float32 everywhere.  Descriptor / keypoint base pointers must be 16-byte aligned (TMA bulk copies).
 *  - Return value: 0 = ok; negative = argument error (BFR_E_*); positive = cudaError_t of the failing launch.
 *    No exception crosses the ABI.  Failure convention of the path itself follows the reference: fewer than three
 *    correspondences or no valid hypothesis yields the identity transform (ThreeDMatch/test.py:242-245).
 *  - Correspondence records ("corr") are 8 floats per correspondence: sx sy sz w | qx qy qz 0 (w unused by RANSAC).
 */
#ifndef BUFFER_B200_H
#define BUFFER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFR_OK 0
#define BFR_E_NULL (-1)      /* required pointer is NULL */
#define BFR_E_SIZE (-2)      /* negative / inconsistent size, or more than BFR_MAX_PAIRS pairs in one call */
#define BFR_MAX_PAIRS 65535   /* pairs per batched call (one grid dimension); larger batches are split by the caller */
#define BFR_E_DIM (-3)       /* unsupported descriptor length (this build: D == 32) */
#define BFR_E_WORKSPACE (-4) /* workspace too small */
#define BFR_E_ALIGN (-5)     /* pointer not 16-byte aligned */

int bfr_version(void);
const char* bfr_error_string(int code);

/* Tuning knobs of the CALLING THREAD (thread-local; default 1 in every thread).  BFR_CFG_K1_ALGO selects the mutual-NN
 * implementation: 0 = FP32 FFMA2 kernel (every product in FP32), 1 = tensor-core (tcgen05, f16 operands) filter followed by an
 * exact FP32 re-check of the near-best candidates.  Both produce bit-identical outputs. */
#define BFR_CFG_K1_ALGO 1
/* BFR_CFG_RANSAC_TC selects how RANSAC scores its hypotheses on pairs of up to 5120 correspondences: 1 = tensor-core (tcgen05, 2-level f16
 * operand splits) residual filter with an exact FP32 re-check of the (h, c) pairs too close to the threshold to call (needs the scratch
 * workspace of bfr_ransac_workspace_bytes), 0 = every residual in FP32.  Both produce bit-identical outputs. */
#define BFR_CFG_RANSAC_TC 2
int bfr_config_set(int key, int value);
int bfr_config_get(int key);

/* ---- K1: fused L2 distance + mutual nearest neighbour -------------------------------------------------------------
 * Replaces buffer.mutual_matching (models/BUFFER.py:335-359) and its two knn_cuda.KNN(k=1) calls (:347, :352).
 * Outputs (each optional, may be NULL): nn_s[i] = nearest tgt row of src row i, nn_t[j] = nearest src row of tgt row
 * j (pair-local indices, int64 like knn_cuda); dist_* = Euclidean distances; s_mids/t_mids = the mutual matches of pair
 * p, ascending in s, written at s_mids[src_off[p] ...] with n_mutual[p] entries (models/BUFFER.py:356-357); corr_xyz =
 * the matched keypoints gathered into correspondence records at the same offsets (models/BUFFER.py:284,287; needs
 * src_xyz/tgt_xyz [rows][3]).  col_splits >= 1 splits the target rows of each pair over that many CTAs (use > 1 when P
 * is too small to fill 148 SMs).  total_M / total_N = rows of the concatenated descriptor arrays (TMA tensor-map bounds).
 * max_M / max_N must be >= every pair's row count: a larger pair is truncated to the bound (never read out of its workspace slice). */
size_t bfr_mutual_nn_workspace_bytes(int P, int max_M, int max_N);
int bfr_mutual_matching_batched(const float* src_des, const float* tgt_des, const int32_t* src_off, const int32_t* tgt_off,
                                int P, int max_M, int max_N, int total_M, int total_N, int D, int col_splits,
                                int64_t* nn_s, int64_t* nn_t, float* dist_s, float* dist_t,
                                const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr_xyz,
                                void* ws, size_t ws_bytes, void* stream);

/* The same in phases, for ONE HUGE PAIR whose rows are split over several GPUs (BASELINE config 5: 100k x 100k keypoints):
 *   bfr_mutual_nn_partial  half norms / operand copies for all rows, then the fused distance + arg-max kernel on row-block
 *                          partition `part` of `nparts` (source rows of the src->tgt direction and target rows of the tgt->src
 *                          direction are both partitioned); the packed bests (key << 32 | ~index, 0 = untouched) stay in `ws`;
 *   bfr_mutual_nn_packed   the contiguous uint64 region of `ws` holding them (row bests then column bests, `count` entries):
 *                          the caller max-reduces it across ranks as UNSIGNED 64-bit (NCCL all-reduce over NVLink; with a signed
 *                          reduction flip bit 63 before and after);
 *   bfr_mutual_select      decode + mutual check + ordered compaction (+ gather) from the reduced bests - same outputs as
 *                          bfr_mutual_matching_batched, which is exactly partial(0, 1) followed by select. */
int bfr_mutual_nn_partial(const float* src_des, const float* tgt_des, const int32_t* src_off, const int32_t* tgt_off,
                          int P, int max_M, int max_N, int total_M, int total_N, int D, int col_splits, int part, int nparts,
                          void* ws, size_t ws_bytes, void* stream);
int bfr_mutual_nn_packed(void* ws, size_t ws_bytes, int P, int max_M, int max_N, uint64_t** packed, size_t* count);
int bfr_mutual_select(const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                      int64_t* nn_s, int64_t* nn_t, float* dist_s, float* dist_t,
                      const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr_xyz,
                      void* ws, size_t ws_bytes, void* stream);

/* Gather explicit index pairs into correspondence records: the (pcd0, pcd1, corr) arguments of the Open3D call at
 * models/BUFFER.py:314-316.  s_ids/t_ids: int64 [K]. */
int bfr_gather_corr(const float* src_xyz, const float* tgt_xyz, const int64_t* s_ids, const int64_t* t_ids, int K, float* corr_xyz, void* stream);

/* ---- K2 + K3: RANSAC (Philox sampling, 3-point Kabsch, checkers, inlier scoring, packed max) ---------------------------
 * Replaces o3d.pipelines.registration.registration_ransac_based_on_correspondence as called at models/BUFFER.py:318-324
 * (ransac_n = 3, point-to-point without scaling, edge-length checker `similar_th`, distance checker `dist_th`,
 * max_correspondence_distance = dist_th).  Evaluates hypotheses [h_begin, h_end) of every pair — hypothesis h of pair p
 * draws Philox4x32-10(key = seed, counter = (h, pair_id_base + p, 0, 0)) — and max-accumulates into best_packed[p] =
 * (inlier count << 32) | (0xFFFFFFFF - h): the caller zeroes best_packed before the first call and may split the
 * hypothesis range over several calls, streams or GPUs (all-reduce MAX) before finalising.  corr_cnt[p] < 3 leaves 0.
 * valid_count (optional, [P] int32, zeroed by the caller) accumulates how many hypotheses passed every checker and were
 * scored — the H_valid of the 28*H_valid*C scoring-work model.
 * confidence: RANSACConvergenceCriteria(iter_n, confidence) of models/BUFFER.py:323-324.  >= 1 (KITTI/config.py:65) or <= 0:
 * every hypothesis of the range is evaluated.  In (0, 1) (ThreeDMatch/config.py:65 = 0.999): Open3D's rule, evaluated as by ONE
 * sequential thread - iteration i (= h - h_begin) only runs while i < min(h_end - h_begin, ceil(log(1 - confidence) /
 * log(1 - (best_count / K)^3))) of the best hypothesis before it; the result is the best of exactly those iterations.  The rule
 * is sequential in h, so one call must cover the pair's whole range (no split over calls / GPUs) and `splits` is ignored.
 * ws / ws_bytes: bfr_ransac_workspace_bytes() of device scratch owned by this call until it has completed (the f16 operand tiles
 * of the tensor-core scoring filter, which stay in L2).  ws = NULL (or too small): every hypothesis is scored by the exact FP32 loop -
 * slower, identical results. */
size_t bfr_ransac_workspace_bytes(void);
int bfr_ransac_batched(const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                       uint64_t seed, uint32_t pair_id_base, uint32_t h_begin, uint32_t h_end,
                       float dist_th, float similar_th, float confidence, int splits, uint64_t* best_packed, int32_t* valid_count,
                       void* ws, size_t ws_bytes, void* stream);
/* Decode best_packed and regenerate the winning minimal-sample fit: T [P][16] row-major 4x4 (result.transformation,
 * models/BUFFER.py:326), inlier count and hypothesis index (-1 if none; T = identity). */
int bfr_ransac_finalize_batched(const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                                uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, const uint64_t* best_packed,
                                float* T, int32_t* inliers, int64_t* best_h, void* stream);

/* ---- a3 / a4: per-correspondence LRF hypotheses and their scoring --------------------------------------------------
 * bfr_lrf_hypotheses replaces models/BUFFER.py:294-301: R = tt_R Rz(angle) ss_R^T, t = tt_kpts - R ss_kpts; cs[i] =
 * {cos(angle_i), sin(angle_i)} is supplied by the caller.  bfr_score_hypotheses replaces models/BUFFER.py:303-311:
 * counts[h] = #{c : sqrt(|R_h s_c + t_h - q_c|^2) < thr_c} (the reference's test on the rooted distance, :305-308); best_idx =
 * first maximum (torch.argmax); mask = inliers of the best.  thr: [C] per-correspondence thresholds or NULL (then thr_scalar). */
int bfr_lrf_hypotheses(const float* cs, const float* ss_R, const float* tt_R, const float* ss_kpts, const float* tt_kpts, int A,
                       float* R_out, float* t_out, void* stream);
size_t bfr_score_workspace_bytes(int C);
int bfr_score_hypotheses(const float* R, const float* t, int H, const float* src, const float* tgt, int C, const float* thr, float thr_scalar,
                         int32_t* counts, uint64_t* best_packed, int64_t* best_idx, uint8_t* mask, void* ws, size_t ws_bytes, void* stream);

/* The two blocks fused and batched over pairs - the LRF vote, models/BUFFER.py:294-311 - with everything on the device:
 * corr_xyz = the records of ALL mutual matches of each pair (K1's corr_xyz output; the 4th float of every record is overwritten
 * with the vote threshold), ind [rows] = the inlier head's azimuth index per match (:291-292), ss_R / tt_R [rows][9] = the matched
 * local reference frames (s_R[s_mids], t_R[t_mids], :286,:289), all row-aligned with corr_xyz.  Per pair: proposal c is
 * R = tt_R[c] Rz(ind_c 2 pi / azi_n + 1e-6) ss_R[c]^T, t = q_c - R s_c (cos / sin by a pinned polynomial, mirrored in the oracle; R, t
 * only exist in registers); inlier_num (optional, [rows]) = inliers of every proposal under thr_c = |s_c| pi / azi_n inlier_th;
 * best_ind [P] (optional) = first maximum; sub_corr [rows][8] / sub_cnt [P] = the inliers of the winner compacted in ascending
 * order at the pair's offset - the `corr` of the Open3D call (:311-316); inlier_ind (optional, [rows]) their indices.
 * bfr_pose_from_votes_batched = that vote -> RANSAC on the voted subset (:313-326) -> post_refinement on ALL matches (:327-329),
 * one call, no host sync: the reference's stage flow after the inlier head. */
size_t bfr_vote_workspace_bytes(int P, int total_rows);
int bfr_lrf_vote_batched(float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, int total_rows,
                         const float* ind, const float* ss_R, const float* tt_R, float azi_n, float inlier_th,
                         int32_t* inlier_num, int64_t* best_ind, float* sub_corr, int32_t* sub_cnt, int64_t* inlier_ind,
                         void* ws, size_t ws_bytes, void* stream);
int bfr_pose_from_votes_batched(float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, int total_rows,
                                const float* ind, const float* ss_R, const float* tt_R, float azi_n, float inlier_th,
                                int hypotheses, uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, float confidence,
                                float refine_thr, int refine_iters, int ransac_splits,
                                float* T_out, int32_t* n_vote_inliers, int32_t* n_inliers, void* ws, size_t ws_bytes, void* stream);

/* ---- K4: weighted Kabsch and post-refinement ------------------------------------------------------------------------
 * bfr_rigid_transform_3d replaces rigid_transform_3d (models/BUFFER.py:424-464): A, B [bs][n][3], w [bs][n] or NULL,
 * T [bs][16].  bfr_post_refinement_batched replaces buffer.post_refinement (models/BUFFER.py:382-418) for P pairs:
 * T0/T_out [P][16]; thr = 0.10 (3DMatch/3DLoMatch/ETH) or 1.2 (KITTI), max_iter = 20 in the reference.  max_count = host-side
 * upper bound of corr_cnt (0 = unknown): above 16384 every pair gets an 8-CTA thread-block cluster instead of one CTA (the result
 * does not depend on it: both kernels walk the same fixed reduction tree). */
int bfr_rigid_transform_3d(const float* A, const float* B, const float* w, int bs, int n, float weight_threshold, float* T, void* stream);
int bfr_post_refinement_batched(const float* T0, const float* corr_xyz, const int32_t* corr_off, const int32_t* corr_cnt, int P,
                                float thr, int max_iter, int max_count, float* T_out, int32_t* iters, int32_t* inliers, void* stream);

/* ---- whole back end: descriptors + keypoints -> pose, one call, no host sync -----------------------------------------
 * The test branch of buffer.forward after the descriptors exist, without the learned inlier head:
 * mutual_matching (:283) -> gather (:284-287) -> RANSAC on all mutual matches (:313-326) -> post_refinement (:327-329).
 * refine_iters = 0 skips refinement (KITTI config, pose_refine=False).  total_M/total_N = rows of the concatenated arrays.
 * confidence: see bfr_ransac_batched (1.0 = evaluate all `hypotheses`).  With the learned inlier head in the loop use
 * bfr_mutual_matching_batched -> (head) -> bfr_pose_from_votes_batched instead. */
size_t bfr_register_workspace_bytes(int P, int max_M, int max_N, int total_M, int total_N);
int bfr_register_batched(const float* src_des, const float* src_xyz, const int32_t* src_off,
                         const float* tgt_des, const float* tgt_xyz, const int32_t* tgt_off,
                         int P, int max_M, int max_N, int total_M, int total_N, int D,
                         int hypotheses, uint64_t seed, uint32_t pair_id_base, float dist_th, float similar_th, float confidence,
                         float refine_thr, int refine_iters, int ransac_splits,
                         float* T_out, int32_t* n_mutual, int32_t* n_inliers, void* ws, size_t ws_bytes, void* stream);

/* Same, for a uniform batch (every pair has M source and N target rows) whose inputs live in HOST memory (pinned for
 * truly asynchronous copies): copies inputs host->device on `stream`, runs the back end, copies T / n_mutual /
 * n_inliers back to host buffers.  Still asynchronous: the caller synchronises `stream` before reading the outputs.
 * The device staging area is part of `ws` (bfr_register_host_workspace_bytes). */
size_t bfr_register_host_workspace_bytes(int P, int M, int N, int D);
int bfr_register_uniform_host(const float* src_des_host, const float* src_xyz_host, const float* tgt_des_host, const float* tgt_xyz_host,
                              int P, int M, int N, int D, int hypotheses, uint64_t seed, uint32_t pair_id_base,
                              float dist_th, float similar_th, float confidence, float refine_thr, int refine_iters, int ransac_splits,
                              float* T_out_host, int32_t* n_mutual_host, int32_t* n_inliers_host, void* ws, size_t ws_bytes, void* stream);

/* The whole job in ONE call: the P pairs are cut into chunks of chunk_pairs pairs, chunk c is enqueued (copies + kernels + copies back,
 * exactly bfr_register_uniform_host with pair ids pair_id_base + first pair of the chunk) on streams[c % n_streams] with workspace
 * ws[c % n_streams] (each of ws_bytes_each >= bfr_register_host_workspace_bytes(chunk_pairs, M, N, D)); with two or more streams the
 * host<->device copies of one chunk overlap the kernels of another.  Asynchronous: synchronise every stream before reading the outputs. */
int bfr_register_uniform_host_chunked(const float* src_des_host, const float* src_xyz_host, const float* tgt_des_host, const float* tgt_xyz_host,
                                      int P, int M, int N, int D, int chunk_pairs, int hypotheses, uint64_t seed, uint32_t pair_id_base,
                                      float dist_th, float similar_th, float confidence, float refine_thr, int refine_iters, int ransac_splits,
                                      float* T_out_host, int32_t* n_mutual_host, int32_t* n_inliers_host,
                                      void* const* ws, size_t ws_bytes_each, void* const* streams, int n_streams);

/* ---- "next" rows ------------------------------------------------------------------------------------
 * bfr_get_matching_indices replaces buffer.get_matching_indices (models/BUFFER.py:361-380; twins ThreeDMatch/dataset.py:14-22,
 * ThreeDMatch/trainer.py:38-54): source [N][3] is transformed by relt_pose (DEVICE pointer to a row-major 4x4), every
 * transformed point gets its nearest target [M][3] point, and the pairs with distance < search_voxel_size are written in
 * ascending source order to match_inds [N][2] (int64, like the reference) with their number in count[0]; nn [N] / dist [N]
 * (optional) receive every source point's nearest index and distance.
 * bfr_svd3_batched is the torch_batch_svd.svd contract of cal_Z_axis (utils/common.py:715): x [B][9] -> u [B][9], s [B][3]
 * (descending), v [B][9], x = u diag(s) v^T. */
size_t bfr_get_matching_indices_workspace_bytes(int N);
int bfr_get_matching_indices(const float* source, int N, const float* target, int M, const float* relt_pose, float search_voxel_size,
                             int64_t* match_inds, int32_t* count, int64_t* nn, float* dist, void* ws, size_t ws_bytes, void* stream);
int bfr_svd3_batched(const float* x, int B, float* u, float* s, float* v, void* stream);
/* bfr_furthest_point_sample: pointnet2_ops.furthest_point_sample as called at models/BUFFER.py:266-267 (xyz [B][N][3] ->
 * idx [B][npoint] int32, first index 0, points with |p|^2 <= 1e-3 skipped, ties -> lowest index); temp = [B][N] float scratch. */
int bfr_furthest_point_sample(const float* xyz, int B, int N, int npoint, int32_t* idx, float* temp, void* stream);

/* ---- measurement aids (used by bench.py only) ---------------------------------------------------------------------
 * bfr_fp32_probe: a pure FFMA2 stream on `grid` CTAs of 256 threads, executing grid*256*iters*256 FMAs (2 flop each);
 * scratch: device floats, >= grid*256 + 128, first 128 initialised by the caller to finite values.  The measured rate
 * is the FP32 issue peak the mutual-NN roofline is normalised against.
 * bfr_debug_set_k1_events: cudaEvent_t pair (or NULLs to disable) recorded immediately before / after the main
 * mutual-NN kernel by subsequent calls made from THIS host thread; lets bench.py time that kernel inside the full step. */
int bfr_fp32_probe(int grid, int iters, float* scratch, void* stream);
int bfr_debug_set_k1_events(void* ev_start, void* ev_stop);

#ifdef __cplusplus
}
#endif
#endif /* BUFFER_B200_H */
