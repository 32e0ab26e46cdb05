// bfr_kernels.h — internal launcher declarations (host side) shared by the .cu files and the C ABI (api.cu).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace bfr {

// K1 (mutual_nn.cu)
int k1_pad_rows(int max_rows);
size_t k1_workspace_bytes(int P, int max_M, int max_N);
void k1_set_algo(int algo);
int k1_get_algo();
bool k1_tc_supported(int D, long long total_M, long long total_N);
cudaError_t k1_tc_launch(const float* src, const float* tgt, const void* src_f16, const void* tgt_f16, const int32_t* out_of_range, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                         long long total_M, long long total_N, const float* hna, const float* hnb, int padM, int padN,
                         unsigned long long* row_packed, unsigned long long* col_packed, int part, int nparts, cudaStream_t stream);
// per-device, one-time opt-in to more than 48 KB of dynamic shared memory (`done`: one bit per device ordinal)
cudaError_t ensure_dyn_smem(const void* fn, int bytes, std::atomic<unsigned long long>& done);
void k1_packed_view(void* ws, int P, int max_M, int max_N, unsigned long long** packed, size_t* count);
cudaError_t k1_partial_launch(const float* src, const float* tgt, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                              long long total_M, long long total_N, int D, int col_splits, int part, int nparts, void* ws, cudaStream_t stream);
cudaError_t k1_select_launch(const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N, void* ws,
                             int64_t* nn_s, int64_t* nn_t, float* d_s, float* d_t, const float* src_xyz, const float* tgt_xyz,
                             int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr, cudaStream_t stream);
cudaError_t k1_launch(const float* src, const float* tgt, const int32_t* src_off, const int32_t* tgt_off, int P, int max_M, int max_N,
                      long long total_M, long long total_N, int D, int col_splits, void* ws, int64_t* nn_s, int64_t* nn_t, float* d_s, float* d_t,
                      const float* src_xyz, const float* tgt_xyz, int64_t* s_mids, int64_t* t_mids, int32_t* n_mutual, float* corr,
                      cudaStream_t stream);
cudaError_t gather_corr_launch(const float* src_xyz, const float* tgt_xyz, const int64_t* s_ids, const int64_t* t_ids, int K, float* corr, cudaStream_t stream);

cudaError_t fp32_probe_launch(int grid, int iters, float* scratch, cudaStream_t stream);
void k1_set_events(cudaEvent_t e0, cudaEvent_t e1);

// K2 + K3 (ransac.cu)
cudaError_t ransac_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                          uint32_t h_begin, uint32_t h_end, float dist_th, float similar_th, float confidence, int splits,
                          unsigned long long* best_packed, int32_t* valid_count, void* tc_scratch, size_t tc_scratch_bytes, cudaStream_t stream);
size_t ransac_scratch_bytes();
void ransac_set_tc(int on);
int ransac_get_tc();
cudaError_t lrf_vote_launch(float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, int max_count, const float* ind, const float* ss_R, const float* tt_R,
                            float azi_n, float inlier_th, int32_t* counts, unsigned long long* vote_best, float* sub_corr, int32_t* sub_cnt,
                            int64_t* best_idx, int64_t* inlier_ind, cudaStream_t stream);
cudaError_t ransac_finalize_launch(const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, uint64_t seed, uint32_t pair_id_base,
                                   float dist_th, float similar_th, const unsigned long long* best_packed, float* T, int32_t* inliers, int64_t* best_h, cudaStream_t stream);
cudaError_t lrf_hypotheses_launch(const float* cs, const float* ss_R, const float* tt_R, const float* ss_kpts, const float* tt_kpts, int A,
                                  float* R_out, float* t_out, cudaStream_t stream);
size_t score_workspace_bytes(int C);
cudaError_t score_hypotheses_launch(const float* R, const float* t, int H, const float* src, const float* tgt, int C, const float* thr, float thr_scalar,
                                    int32_t* counts, unsigned long long* best_packed, int64_t* best_idx, uint8_t* mask, void* ws, cudaStream_t stream);

// K4 (refine.cu)
cudaError_t rigid_transform_launch(const float* A, const float* B, const float* w, int bs, int n, float weight_threshold, float* T, cudaStream_t stream);
cudaError_t post_refinement_launch(const float* T0, const float* corr, const int32_t* corr_off, const int32_t* corr_cnt, int P, float thr, int max_iter,
                                   float* Tout, int32_t* iters_out, int32_t* inliers_out, int max_count, cudaStream_t stream);

// "next" rows (extras.cu)
size_t knn3_workspace_bytes(int N);
cudaError_t get_matching_indices_launch(const float* source, int N, const float* target, int M, const float* T, float voxel,
                                        int64_t* pairs, int32_t* count, int64_t* nn_out, float* dist_out, void* ws, cudaStream_t stream);
cudaError_t fps_launch(const float* xyz, int B, int N, int npoint, int32_t* idx, float* temp, cudaStream_t stream);
cudaError_t svd3_launch(const float* x, int B, float* u, float* s, float* v, cudaStream_t stream);

}  // namespace bfr
