"""Host-side (Python/PyTorch) mirror of BUFFER's correspondence-and-pose back end on top of libbuffer_b200.so.

PyTorch is plumbing only: it owns device memory and streams; every operation below is one or more hand-written
sm_100a kernels reached through the C ABI (include/buffer_b200.h).  Function names, argument meaning and failure
behaviour follow the reference (file:line under /root/reference):

  mutual_matching(src_des, tgt_des)                         models/BUFFER.py:335-359
  rigid_transform_3d(A, B, weights=None, weight_threshold=0) models/BUFFER.py:424-464
  post_refinement(initial_trans, src_keypts, tgt_keypts)     models/BUFFER.py:382-418
  registration_ransac_based_on_correspondence(...)           the Open3D call at models/BUFFER.py:318-324
  lrf_hypotheses / score_hypotheses                          the inline blocks models/BUFFER.py:294-301 / 303-311
  register_batched(...)                                      the whole stage, batched over pairs, no host sync

``*_device`` / ``*_batched`` variants take and return device tensors and never synchronise; the reference-compatible
wrappers do the final ``.cpu().numpy()`` the reference does.
"""
import math

import numpy as np
import torch

from . import _lib

DESC_DIM = 32
_workspaces = {}
K1_FP32, K1_TENSOR_FILTER = 0, 1
RANSAC_FP32, RANSAC_TENSOR_FILTER = 0, 1


def set_k1_algo(algo):
    """select the mutual-NN kernel: K1_FP32 (all products in FP32) or K1_TENSOR_FILTER (tcgen05 f16 filter + exact FP32
    re-check); outputs are bit-identical"""
    _lib.check(_lib.lib().bfr_config_set(1, int(algo)), "bfr_config_set")


def get_k1_algo():
    return _lib.lib().bfr_config_get(1)


def set_ransac_scoring(algo):
    """select how RANSAC scores hypotheses on pairs of up to 5120 correspondences: RANSAC_FP32 (every residual in FP32) or
    RANSAC_TENSOR_FILTER (tcgen05 residual filter on f16 operand splits + exact FP32 re-check of borderline residuals, the default);
    outputs are bit-identical"""
    _lib.check(_lib.lib().bfr_config_set(2, int(algo)), "bfr_config_set")


def get_ransac_scoring():
    return _lib.lib().bfr_config_get(2)


def _stream(device=None):
    """raw handle of torch's current stream on `device` (default: the current device)"""
    return torch.cuda.current_stream(device).cuda_stream


def _guard(fn):
    """The library launches on the CURRENT device of the calling thread (include/buffer_b200.h): make the device that owns the
    first CUDA tensor argument current for the duration of the call, so tensors on a non-current GPU work and `_stream()` is that
    device's current stream."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        for a in list(args) + list(kw.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kw)
                break
        return fn(*args, **kw)
    return wrapped


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _ws(nbytes, device, tag="default"):
    """grow-only scratch buffer per (device, stream, tag); caller-owned memory for the C ABI"""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError("buffer_b200: %s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _offsets(counts, device):
    off = torch.zeros(len(counts) + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.as_tensor(counts, dtype=torch.int64), 0).to(torch.int32)
    return off.to(device, non_blocking=True)


# ---------------------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------------------
@_guard
def mutual_matching_batched(src_des, tgt_des, src_off, tgt_off, max_M, max_N, src_xyz=None, tgt_xyz=None,
                            want_nn=True, want_dist=False, want_mids=True, col_splits=None):
    """Batched varlen mutual matching, all on device, no sync.

    src_des [totM, 32], tgt_des [totN, 32] float32; src_off/tgt_off int32 device [P+1].  max_M / max_N must be >= every
    pair's row count (a larger pair is truncated to the bound by the kernels, never read out of its workspace slice).
    Returns dict: nn_s [totM] / nn_t [totN] int64 (pair-local), s_mids/t_mids [totM] int64 (pair p's matches at
    src_off[p] .. src_off[p]+n_mutual[p]), n_mutual [P] int32, corr [totM, 8] float32 (if keypoints given),
    dist_s/dist_t (if want_dist).
    """
    src_des = _f32c(src_des, "src_des"); tgt_des = _f32c(tgt_des, "tgt_des")
    dev = src_des.device
    P = src_off.numel() - 1
    totM, totN = src_des.shape[0], tgt_des.shape[0]
    if src_des.shape[1] != DESC_DIM or tgt_des.shape[1] != DESC_DIM:
        raise RuntimeError("buffer_b200: descriptors must be [n, %d]" % DESC_DIM)
    L = _lib.lib()
    out = {}
    i64 = dict(dtype=torch.int64, device=dev)
    out["nn_s"] = torch.empty(totM, **i64) if want_nn else None
    out["nn_t"] = torch.empty(totN, **i64) if want_nn else None
    out["dist_s"] = torch.empty(totM, dtype=torch.float32, device=dev) if want_dist else None
    out["dist_t"] = torch.empty(totN, dtype=torch.float32, device=dev) if want_dist else None
    out["s_mids"] = torch.empty(totM, **i64) if want_mids else None
    out["t_mids"] = torch.empty(totM, **i64) if want_mids else None
    out["n_mutual"] = torch.empty(P, dtype=torch.int32, device=dev)
    corr = None
    if src_xyz is not None:
        src_xyz = _f32c(src_xyz, "src_xyz"); tgt_xyz = _f32c(tgt_xyz, "tgt_xyz")
        corr = torch.empty(max(totM, 1), 8, dtype=torch.float32, device=dev)
    out["corr"] = corr
    if col_splits is None:                      # fill 148 SMs x 2 CTAs when the batch is small
        row_blocks = max(1, P * ((max_M + 511) // 512))
        col_splits = max(1, min((max_N + 63) // 64, (296 + row_blocks - 1) // row_blocks))
    nbytes = L.bfr_mutual_nn_workspace_bytes(P, max_M, max_N)
    ws = _ws(nbytes, dev, "k1")
    _lib.check(L.bfr_mutual_matching_batched(src_des.data_ptr(), tgt_des.data_ptr(), src_off.data_ptr(), tgt_off.data_ptr(),
                                             P, max_M, max_N, totM, totN, DESC_DIM, col_splits,
                                             _ptr(out["nn_s"]), _ptr(out["nn_t"]), _ptr(out["dist_s"]), _ptr(out["dist_t"]),
                                             _ptr(src_xyz), _ptr(tgt_xyz), _ptr(out["s_mids"]), _ptr(out["t_mids"]),
                                             out["n_mutual"].data_ptr(), _ptr(corr), ws.data_ptr(), ws.numel(), _stream()),
               "bfr_mutual_matching_batched")
    return out


@_guard
def mutual_matching_device(src_des, tgt_des, src_xyz=None, tgt_xyz=None, want_dist=False):
    """single pair, device tensors in / out, no sync (s_mids/t_mids padded to M; valid prefix = n_mutual[0])"""
    dev = src_des.device
    M, N = src_des.shape[0], tgt_des.shape[0]
    so = torch.tensor([0, M], dtype=torch.int32).to(dev, non_blocking=True)
    to = torch.tensor([0, N], dtype=torch.int32).to(dev, non_blocking=True)
    return mutual_matching_batched(src_des, tgt_des, so, to, M, N, src_xyz, tgt_xyz, want_dist=want_dist)


def mutual_matching(src_des, tgt_des):
    """Drop-in for buffer.mutual_matching (models/BUFFER.py:335-359): returns (s_mids, t_mids) numpy int64,
    s_mids ascending.  (The reference's method takes `self` first; see install.py for the bound version.)"""
    r = mutual_matching_device(src_des, tgt_des)
    n = int(r["n_mutual"].item())               # the reference syncs here too (:348, :353)
    return r["s_mids"][:n].cpu().numpy(), r["t_mids"][:n].cpu().numpy()


def knn1(ref, query):
    """k=1 Euclidean nearest neighbour of every query row among ref rows (knn_cuda.KNN(k=1, transpose_mode=True)
    contract for [1,R,32]/[1,Q,32] inputs, models/BUFFER.py:347): -> (dist [1,Q,1] float32, idx [1,Q,1] int64)."""
    r = mutual_matching_device(query[0], ref[0], want_dist=True)
    return r["dist_s"][None, :, None], r["nn_s"][None, :, None]


# ---------------------------------------------------------------------------------------------------------------
# K2 + K3
# ---------------------------------------------------------------------------------------------------------------
@_guard
def gather_corr(src_xyz, tgt_xyz, s_ids, t_ids):
    """(pcd0, pcd1, corr) of the Open3D call (models/BUFFER.py:314-316) -> correspondence records [K, 8]"""
    src_xyz = _f32c(src_xyz, "src_xyz"); tgt_xyz = _f32c(tgt_xyz, "tgt_xyz")
    s_ids = s_ids.to(torch.int64).contiguous(); t_ids = t_ids.to(torch.int64).contiguous()
    K = s_ids.numel()
    corr = torch.empty(max(K, 1), 8, dtype=torch.float32, device=src_xyz.device)
    _lib.check(_lib.lib().bfr_gather_corr(src_xyz.data_ptr(), tgt_xyz.data_ptr(), s_ids.data_ptr(), t_ids.data_ptr(), K, corr.data_ptr(), _stream()),
               "bfr_gather_corr")
    return corr[:K] if K else corr[:0]


@_guard
def ransac_batched(corr, corr_off, corr_cnt, hypotheses, dist_th, similar_th, seed=0, pair_id_base=0, h_begin=0, h_end=None,
                   splits=None, best_packed=None, valid_count=None, confidence=1.0):
    """Evaluate hypotheses [h_begin, h_end) of every pair; max-accumulate into best_packed [P] (int64 view of the
    packed uint64 (count << 32) | (0xFFFFFFFF - h)).  Device only, no sync.  confidence in (0, 1): Open3D's
    RANSACConvergenceCriteria early exit, replayed exactly as one sequential thread would run it (models/BUFFER.py:323-324);
    the call must then cover the pair's whole hypothesis range."""
    corr = _f32c(corr, "corr")
    P = corr_cnt.numel()
    h_end = hypotheses if h_end is None else h_end
    if best_packed is None:
        best_packed = torch.zeros(P, dtype=torch.int64, device=corr.device)
    if splits is None:
        splits = _default_splits(P)
    need = _lib.lib().bfr_ransac_workspace_bytes()
    ws = _ws(need, corr.device, "ransac")
    _lib.check(_lib.lib().bfr_ransac_batched(corr.data_ptr(), corr_off.data_ptr(), corr_cnt.data_ptr(), P, int(seed), int(pair_id_base),
                                             int(h_begin), int(h_end), float(dist_th), float(similar_th), float(confidence), int(splits),
                                             best_packed.data_ptr(), _ptr(valid_count), ws.data_ptr(), ws.numel(), _stream()), "bfr_ransac_batched")
    return best_packed


@_guard
def ransac_finalize_batched(corr, corr_off, corr_cnt, best_packed, dist_th, similar_th, seed=0, pair_id_base=0):
    """-> T [P,4,4] float32, inliers [P] int32, best_h [P] int64 (device)"""
    P = corr_cnt.numel(); dev = corr.device
    T = torch.empty(P, 4, 4, dtype=torch.float32, device=dev)
    inl = torch.empty(P, dtype=torch.int32, device=dev)
    bh = torch.empty(P, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().bfr_ransac_finalize_batched(corr.data_ptr(), corr_off.data_ptr(), corr_cnt.data_ptr(), P, int(seed), int(pair_id_base),
                                                      float(dist_th), float(similar_th), best_packed.data_ptr(), T.data_ptr(), inl.data_ptr(),
                                                      bh.data_ptr(), _stream()), "bfr_ransac_finalize_batched")
    return T, inl, bh


def _default_splits(P, sms=None):
    """work items per pair of the persistent RANSAC kernel (one CTA per SM): the split count s in 1..64 that minimises
    ceil(P s / SMs) (1 / s + 0.06) - rounds of items times the work of an item, 0.06 of a pair's work being what an item costs up front
    (loading the correspondences, building the operand tiles).  1 623 pairs -> 1 (10.97 rounds), 203 pairs (one of 8 ranks) -> 2, one pair -> 64."""
    if sms is None:
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148
    P = max(int(P), 1)
    return min(range(1, 65), key=lambda s: (math.ceil(P * s / sms) * (1.0 / s + 0.06), s))


class RansacResult:
    """mimics the fields of open3d.pipelines.registration.RegistrationResult the reference reads (models/BUFFER.py:326)"""

    def __init__(self, transformation, fitness, inlier_count, best_hypothesis):
        self.transformation = transformation
        self.fitness = fitness
        self.inlier_count = inlier_count
        self.best_hypothesis = best_hypothesis


@_guard
def registration_ransac_based_on_correspondence(src_kpts, tgt_kpts, corr, max_correspondence_distance, similar_th,
                                                iter_n=50000, confidence=1.0, seed=0, pair_id=0):
    """Drop-in for the Open3D call at models/BUFFER.py:318-324 with the checkers the reference passes.

    src_kpts [A,3], tgt_kpts [A,3] CUDA tensors, corr [K,2] integer tensor/array.  Returns an object whose
    ``.transformation`` is a 4x4 float64 numpy array (minimal-sample fit of the best hypothesis; identity if K < 3 or
    nothing valid).  `confidence` = RANSACConvergenceCriteria's second argument: 1.0 (KITTI/config.py:65) evaluates all
    `iter_n` hypotheses, a value in (0, 1) (ThreeDMatch/config.py:65 = 0.999) stops like Open3D's sequential loop does.
    best = max inlier count, ties -> lowest hypothesis index."""
    dev = src_kpts.device
    corr = torch.as_tensor(np.asarray(corr) if not torch.is_tensor(corr) else corr).to(dev).to(torch.int64).reshape(-1, 2)
    K = corr.shape[0]
    rec = gather_corr(src_kpts, tgt_kpts, corr[:, 0], corr[:, 1]) if K else torch.zeros(1, 8, device=dev)
    off = torch.tensor([0, K], dtype=torch.int32).to(dev); cnt = torch.tensor([K], dtype=torch.int32).to(dev)
    best = ransac_batched(rec, off, cnt, iter_n, max_correspondence_distance, similar_th, seed, pair_id, confidence=confidence)
    T, inl, bh = ransac_finalize_batched(rec, off, cnt, best, max_correspondence_distance, similar_th, seed, pair_id)
    n = int(inl.item())
    return RansacResult(T[0].double().cpu().numpy(), n / max(K, 1), n, int(bh.item()))


# ---------------------------------------------------------------------------------------------------------------
# a3 / a4
# ---------------------------------------------------------------------------------------------------------------
def lrf_hypotheses(ind, ss_R, tt_R, ss_kpts, tt_kpts, azi_n=20):
    """models/BUFFER.py:294-301 -> R [A,3,3], t [A,3].  cos/sin of angle = ind*2*pi/azi_n + 1e-6 are evaluated in
    float64 by torch and rounded to float32 (the kernel takes them as input)."""
    ang = ind.double() * 2 * math.pi / azi_n + 1e-6
    cs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous()
    return lrf_hypotheses_cs(cs, ss_R, tt_R, ss_kpts, tt_kpts)


@_guard
def lrf_hypotheses_cs(cs, ss_R, tt_R, ss_kpts, tt_kpts):
    cs = _f32c(cs, "cs"); ss_R = _f32c(ss_R, "ss_R"); tt_R = _f32c(tt_R, "tt_R")
    ss_kpts = _f32c(ss_kpts, "ss_kpts"); tt_kpts = _f32c(tt_kpts, "tt_kpts")
    A = ss_kpts.shape[0]
    R = torch.empty(A, 3, 3, dtype=torch.float32, device=cs.device); t = torch.empty(A, 3, dtype=torch.float32, device=cs.device)
    _lib.check(_lib.lib().bfr_lrf_hypotheses(cs.data_ptr(), ss_R.data_ptr(), tt_R.data_ptr(), ss_kpts.data_ptr(), tt_kpts.data_ptr(), A,
                                             R.data_ptr(), t.data_ptr(), _stream()), "bfr_lrf_hypotheses")
    return R, t


@_guard
def score_hypotheses(R, t, src, tgt, thr):
    """models/BUFFER.py:303-311 -> (inlier_num [H] int32, best_ind [1] int64, inlier_mask [C] bool), device, no sync.
    thr: python float or [C] tensor."""
    R = _f32c(R, "R"); t = _f32c(t, "t"); src = _f32c(src, "src"); tgt = _f32c(tgt, "tgt")
    dev = R.device
    H, Cn = R.shape[0], src.shape[0]
    thr_t = None if not torch.is_tensor(thr) else _f32c(thr, "thr")
    counts = torch.empty(H, dtype=torch.int32, device=dev); best = torch.zeros(1, dtype=torch.int64, device=dev)
    best_idx = torch.empty(1, dtype=torch.int64, device=dev); mask = torch.zeros(Cn, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    ws = _ws(L.bfr_score_workspace_bytes(Cn), dev, "score")
    _lib.check(L.bfr_score_hypotheses(R.data_ptr(), t.data_ptr(), H, src.data_ptr(), tgt.data_ptr(), Cn, _ptr(thr_t),
                                      0.0 if thr_t is not None else float(thr), counts.data_ptr(), best.data_ptr(), best_idx.data_ptr(),
                                      mask.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "bfr_score_hypotheses")
    return counts, best_idx, mask.bool()


def inlier_threshold(ss_kpts, azi_n=20, inlier_th=1 / 3):
    """thr_c = |ss_c| * pi / azi_n * inlier_th  (models/BUFFER.py:306-307); plain torch, it is an input of the path"""
    return torch.sqrt(torch.sum(ss_kpts ** 2, dim=-1)) * np.pi / azi_n * inlier_th


# ---------------------------------------------------------------------------------------------------------------
# K4
# ---------------------------------------------------------------------------------------------------------------
@_guard
def rigid_transform_3d(A, B, weights=None, weight_threshold=0):
    """Drop-in for rigid_transform_3d (models/BUFFER.py:424-464): A, B [bs,n,3], weights [bs,n] -> [bs,4,4] on
    A's device.  Like the reference it zeroes weights below the threshold IN PLACE (:437)."""
    A_ = _f32c(A, "A"); B_ = _f32c(B, "B")
    bs, n = A_.shape[0], A_.shape[1]
    w = None
    if weights is not None:
        weights[weights < weight_threshold] = 0
        w = _f32c(weights, "weights")
    T = torch.empty(bs, 4, 4, dtype=torch.float32, device=A_.device)
    _lib.check(_lib.lib().bfr_rigid_transform_3d(A_.data_ptr(), B_.data_ptr(), _ptr(w), bs, n, float(weight_threshold), T.data_ptr(), _stream()),
               "bfr_rigid_transform_3d")
    return T


def refine_threshold(dataset):
    """inlier threshold list of post_refinement (models/BUFFER.py:395-398)"""
    return 0.10 if dataset in ("3DMatch", "3DLoMatch", "ETH") else 1.2


@_guard
def post_refinement_batched(T0, corr, corr_off, corr_cnt, thr, max_iter=20, max_count=None):
    """-> T [P,4,4], iters [P], inliers [P] (device, no sync).  max_count: host-side upper bound of corr_cnt (default: the rows
    of `corr`); above 16384 every pair is refined by an 8-CTA cluster (same reduction tree, same bits)."""
    if max_count is None:
        max_count = corr.shape[0]
    T0 = _f32c(T0, "T0").reshape(-1, 16); corr = _f32c(corr, "corr")
    P = corr_cnt.numel(); dev = corr.device
    T = torch.empty(P, 4, 4, dtype=torch.float32, device=dev)
    it = torch.empty(P, dtype=torch.int32, device=dev); inl = torch.empty(P, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().bfr_post_refinement_batched(T0.data_ptr(), corr.data_ptr(), corr_off.data_ptr(), corr_cnt.data_ptr(), P, float(thr),
                                                      int(max_iter), int(max_count), T.data_ptr(), it.data_ptr(), inl.data_ptr(), _stream()),
               "bfr_post_refinement_batched")
    return T, it, inl


@_guard
def post_refinement(initial_trans, src_keypts, tgt_keypts, weights=None, dataset="3DMatch"):
    """Drop-in for buffer.post_refinement (models/BUFFER.py:382-418): [1,4,4], [1,n,3], [1,n,3] -> [1,4,4].
    `weights` is ignored exactly as in the reference; the whole <=20-round loop runs in one kernel, no host sync."""
    assert initial_trans.shape[0] == 1
    dev = src_keypts.device
    n = src_keypts.shape[1]
    rec = torch.zeros(max(n, 1), 8, dtype=torch.float32, device=dev)
    rec[:n, 0:3] = src_keypts[0]; rec[:n, 4:7] = tgt_keypts[0]
    off = torch.tensor([0, n], dtype=torch.int32).to(dev); cnt = torch.tensor([n], dtype=torch.int32).to(dev)
    T, _, _ = post_refinement_batched(initial_trans.to(dev), rec, off, cnt, refine_threshold(dataset), 20)
    return T


# ---------------------------------------------------------------------------------------------------------------
# "next" rows
# ---------------------------------------------------------------------------------------------------------------
@_guard
def get_matching_indices_device(source, target, relt_pose, search_voxel_size):
    """no-sync variant -> (match_inds [N,2] int64 padded, count [1] int32, nn [N] int64, dist [N] float32)"""
    source = _f32c(source, "source"); target = _f32c(target, "target"); T = _f32c(relt_pose, "relt_pose").reshape(16)
    dev = source.device
    N, M = source.shape[0], target.shape[0]
    pairs = torch.empty(max(N, 1), 2, dtype=torch.int64, device=dev); count = torch.zeros(1, dtype=torch.int32, device=dev)
    nn = torch.empty(max(N, 1), dtype=torch.int64, device=dev); dist = torch.empty(max(N, 1), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _ws(L.bfr_get_matching_indices_workspace_bytes(N), dev, "knn3")
    _lib.check(L.bfr_get_matching_indices(source.data_ptr(), N, target.data_ptr(), M, T.data_ptr(), float(search_voxel_size), pairs.data_ptr(),
                                          count.data_ptr(), nn.data_ptr(), dist.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "bfr_get_matching_indices")
    return pairs, count, nn[:N], dist[:N]


def get_matching_indices(source, target, relt_pose, search_voxel_size):
    """Drop-in for buffer.get_matching_indices (models/BUFFER.py:361-380): -> match_inds [C,2] int64 CUDA tensor"""
    pairs, count, _, _ = get_matching_indices_device(source, target, relt_pose, search_voxel_size)
    return pairs[: int(count.item())]


@_guard
def svd(x):
    """torch_batch_svd.svd drop-in (utils/common.py:10,715): x [B,3,3] CUDA -> (u [B,3,3], s [B,3] descending, v [B,3,3])"""
    x = _f32c(x, "x")
    if x.dim() != 3 or x.shape[1:] != (3, 3):
        raise RuntimeError("buffer_b200.svd: expected [B,3,3]")
    B = x.shape[0]
    u = torch.empty(B, 3, 3, dtype=torch.float32, device=x.device); s = torch.empty(B, 3, dtype=torch.float32, device=x.device); v = torch.empty_like(u)
    _lib.check(_lib.lib().bfr_svd3_batched(x.data_ptr(), B, u.data_ptr(), s.data_ptr(), v.data_ptr(), _stream()), "bfr_svd3_batched")
    return u, s, v


@_guard
def furthest_point_sample(xyz, npoint):
    """pointnet2_ops.furthest_point_sample drop-in (models/BUFFER.py:266-267): xyz [B,N,3] CUDA -> idx [B,npoint] int32"""
    xyz = _f32c(xyz, "xyz")
    B, N = xyz.shape[0], xyz.shape[1]
    idx = torch.zeros(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.empty(B, max(N, 1), dtype=torch.float32, device=xyz.device)
    _lib.check(_lib.lib().bfr_furthest_point_sample(xyz.data_ptr(), B, N, int(npoint), idx.data_ptr(), temp.data_ptr(), _stream()), "bfr_furthest_point_sample")
    return idx


def gather_operation(features, idx):
    """pointnet2_ops.gather_operation (models/BUFFER.py:268-271): features [B,C,N], idx [B,npoint] -> [B,C,npoint]; plain torch gather"""
    return torch.gather(features, 2, idx.long()[:, None, :].expand(-1, features.shape[1], -1))


def _records(ss_kpts, tt_kpts):
    """[A,3] + [A,3] -> correspondence records [A,8] (sx sy sz 0 | qx qy qz 0)"""
    A = ss_kpts.shape[0]
    rec = torch.zeros(max(A, 1), 8, dtype=torch.float32, device=ss_kpts.device)
    rec[:A, 0:3] = ss_kpts; rec[:A, 4:7] = tt_kpts
    return rec


@_guard
def lrf_vote_batched(corr, corr_off, corr_cnt, ind, ss_R, tt_R, max_count, azi_n=20, inlier_th=1 / 3, want_counts=True, want_ind=True):
    """The LRF vote (models/BUFFER.py:294-311) for P pairs in one call, everything on the device, no sync: R and t only exist in
    registers.  corr [rows,8] = records of ALL mutual matches (K1's `corr`; the 4th float of each record is overwritten with the
    vote threshold), ind [rows] float32, ss_R / tt_R [rows,3,3], row-aligned with corr.
    -> dict: inlier_num [rows] int32, best_ind [P] int64, sub_corr [rows,8] (each pair's voted subset at its offset: RANSAC's
    input), sub_cnt [P] int32, inlier_ind [rows] int64 (pair-local indices of the subset, ascending)."""
    corr = _f32c(corr, "corr"); ind = _f32c(ind, "ind"); ss_R = _f32c(ss_R, "ss_R"); tt_R = _f32c(tt_R, "tt_R")
    dev = corr.device
    rows, P = corr.shape[0], corr_cnt.numel()
    out = {"inlier_num": torch.zeros(rows, dtype=torch.int32, device=dev) if want_counts else None,
           "best_ind": torch.empty(P, dtype=torch.int64, device=dev),
           "sub_corr": torch.empty(max(rows, 1), 8, dtype=torch.float32, device=dev),
           "sub_cnt": torch.empty(P, dtype=torch.int32, device=dev),
           "inlier_ind": torch.empty(max(rows, 1), dtype=torch.int64, device=dev) if want_ind else None}
    L = _lib.lib()
    ws = _ws(L.bfr_vote_workspace_bytes(P, 0), dev, "vote")
    _lib.check(L.bfr_lrf_vote_batched(corr.data_ptr(), corr_off.data_ptr(), corr_cnt.data_ptr(), P, int(max_count), rows, ind.data_ptr(), ss_R.data_ptr(),
                                      tt_R.data_ptr(), float(azi_n), float(inlier_th), _ptr(out["inlier_num"]), out["best_ind"].data_ptr(),
                                      out["sub_corr"].data_ptr(), out["sub_cnt"].data_ptr(), _ptr(out["inlier_ind"]), ws.data_ptr(), ws.numel(), _stream()),
               "bfr_lrf_vote_batched")
    return out


def lrf_vote(ind, ss_R, tt_R, ss_kpts, tt_kpts, azi_n=20, inlier_th=1 / 3):
    """lines 294-311 of models/BUFFER.py for one pair in one fused call (no R / t in memory, no .cpu() at :311):
    -> inlier_num [A] int32, best_ind [1] int64, inlier_ind [A] int64 (first n_inliers entries valid), n_inliers [1] int32"""
    A = ss_kpts.shape[0]
    dev = ss_kpts.device
    rec = _records(_f32c(ss_kpts, "ss_kpts"), _f32c(tt_kpts, "tt_kpts"))
    off = torch.tensor([0, A], dtype=torch.int32).to(dev); cnt = torch.tensor([A], dtype=torch.int32).to(dev)
    r = lrf_vote_batched(rec, off, cnt, ind, ss_R, tt_R, A, azi_n, inlier_th)
    return r["inlier_num"][:A], r["best_ind"], r["inlier_ind"][:A], r["sub_cnt"]


@_guard
def pose_from_votes_batched(corr, corr_off, corr_cnt, ind, ss_R, tt_R, max_count, hypotheses=50000, dist_th=0.10, similar_th=0.8, confidence=1.0,
                            refine_thr=0.10, refine_iters=20, seed=0, pair_id_base=0, azi_n=20, inlier_th=1 / 3, ransac_splits=None):
    """The reference's stage flow after the inlier head, models/BUFFER.py:291-329, for P pairs in ONE call with zero host syncs:
    LRF vote on all mutual matches -> inlier_ind compacted on the device -> RANSAC on that subset -> post_refinement on ALL matches.
    Arguments as lrf_vote_batched.  -> T [P,4,4] float32, n_vote_inliers [P] int32 (= len(inlier_ind)), n_inliers [P] int32."""
    corr = _f32c(corr, "corr"); ind = _f32c(ind, "ind"); ss_R = _f32c(ss_R, "ss_R"); tt_R = _f32c(tt_R, "tt_R")
    dev = corr.device
    rows, P = corr.shape[0], corr_cnt.numel()
    T = torch.empty(P, 4, 4, dtype=torch.float32, device=dev)
    nv = torch.empty(P, dtype=torch.int32, device=dev); ni = torch.empty(P, dtype=torch.int32, device=dev)
    if ransac_splits is None:
        ransac_splits = _default_splits(P)
    L = _lib.lib()
    ws = _ws(L.bfr_vote_workspace_bytes(P, rows), dev, "vote")
    _lib.check(L.bfr_pose_from_votes_batched(corr.data_ptr(), corr_off.data_ptr(), corr_cnt.data_ptr(), P, int(max_count), rows, ind.data_ptr(),
                                             ss_R.data_ptr(), tt_R.data_ptr(), float(azi_n), float(inlier_th), int(hypotheses), int(seed), int(pair_id_base),
                                             float(dist_th), float(similar_th), float(confidence), float(refine_thr), int(refine_iters), int(ransac_splits),
                                             T.data_ptr(), nv.data_ptr(), ni.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "bfr_pose_from_votes_batched")
    return T, nv, ni


# ---------------------------------------------------------------------------------------------------------------
# K1 in phases: one huge pair whose rows are split over several GPUs (BASELINE config 5)
# ---------------------------------------------------------------------------------------------------------------
class MutualNNSplit:
    """bfr_mutual_nn_partial -> (all-reduce MAX of the packed bests across ranks, done by the caller) -> bfr_mutual_select.
    `packed` is an int64 view of the uint64 workspace region (untouched entries are 0); flipping bit 63 maps the unsigned order
    onto the signed one, which is what NCCL's int64 MAX reduces."""

    def __init__(self, src_des, tgt_des, src_off, tgt_off, max_M, max_N):
        self.src_des = _f32c(src_des, "src_des"); self.tgt_des = _f32c(tgt_des, "tgt_des")
        self.src_off, self.tgt_off, self.max_M, self.max_N = src_off, tgt_off, int(max_M), int(max_N)
        self.P = src_off.numel() - 1
        self.dev = self.src_des.device
        L = _lib.lib()
        self.nbytes = L.bfr_mutual_nn_workspace_bytes(self.P, self.max_M, self.max_N)
        self.ws = torch.empty(self.nbytes + 1024, dtype=torch.uint8, device=self.dev)
        import ctypes as C
        ptr, cnt = C.c_void_p(0), C.c_size_t(0)
        _lib.check(L.bfr_mutual_nn_packed(self.ws.data_ptr(), self.ws.numel(), self.P, self.max_M, self.max_N, C.byref(ptr), C.byref(cnt)), "bfr_mutual_nn_packed")
        start = ptr.value - self.ws.data_ptr()
        self.packed = self.ws[start:start + 8 * cnt.value].view(torch.int64)

    def partial(self, part, nparts, col_splits=1):
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().bfr_mutual_nn_partial(self.src_des.data_ptr(), self.tgt_des.data_ptr(), self.src_off.data_ptr(), self.tgt_off.data_ptr(), self.P,
                                                        self.max_M, self.max_N, self.src_des.shape[0], self.tgt_des.shape[0], DESC_DIM, int(col_splits),
                                                        int(part), int(nparts), self.ws.data_ptr(), self.ws.numel(), _stream()), "bfr_mutual_nn_partial")
        return self.packed

    def all_reduce_max(self, group=None):
        """unsigned 64-bit MAX across ranks with NCCL's signed int64 MAX (flip bit 63 before and after)"""
        import torch.distributed as dist
        self.packed ^= -0x8000000000000000
        dist.all_reduce(self.packed, op=dist.ReduceOp.MAX, group=group)
        self.packed ^= -0x8000000000000000

    def select(self, src_xyz=None, tgt_xyz=None, want_nn=True, want_mids=True):
        dev, P = self.dev, self.P
        totM, totN = self.src_des.shape[0], self.tgt_des.shape[0]
        i64 = dict(dtype=torch.int64, device=dev)
        out = {"nn_s": torch.empty(totM, **i64) if want_nn else None, "nn_t": torch.empty(totN, **i64) if want_nn else None,
               "s_mids": torch.empty(totM, **i64) if want_mids else None, "t_mids": torch.empty(totM, **i64) if want_mids else None,
               "n_mutual": torch.empty(P, dtype=torch.int32, device=dev), "corr": None}
        if src_xyz is not None:
            src_xyz = _f32c(src_xyz, "src_xyz"); tgt_xyz = _f32c(tgt_xyz, "tgt_xyz")
            out["corr"] = torch.empty(max(totM, 1), 8, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfr_mutual_select(self.src_off.data_ptr(), self.tgt_off.data_ptr(), P, self.max_M, self.max_N, _ptr(out["nn_s"]), _ptr(out["nn_t"]), 0, 0,
                                                    _ptr(src_xyz), _ptr(tgt_xyz), _ptr(out["s_mids"]), _ptr(out["t_mids"]), out["n_mutual"].data_ptr(), _ptr(out["corr"]),
                                                    self.ws.data_ptr(), self.ws.numel(), _stream()), "bfr_mutual_select")
        return out


# ---------------------------------------------------------------------------------------------------------------
# whole back end
# ---------------------------------------------------------------------------------------------------------------
@_guard
def register_batched(src_des, src_xyz, src_off, tgt_des, tgt_xyz, tgt_off, max_M, max_N, hypotheses=50000, dist_th=0.10, similar_th=0.8,
                     refine_thr=0.10, refine_iters=20, seed=0, pair_id_base=0, ransac_splits=None, confidence=1.0):
    """mutual matching -> RANSAC on ALL mutual matches -> post-refinement for P pairs, one C call, no host sync.
    -> T [P,4,4] float32, n_mutual [P] int32, n_inliers [P] int32 (device).
    This is the test branch of buffer.forward WITHOUT the learned inlier head: the reference hands RANSAC only the LRF-vote
    subset (models/BUFFER.py:303-316).  With the head in the loop use mutual_matching_batched -> (head) ->
    pose_from_votes_batched, which reproduces the reference's flow exactly."""
    src_des = _f32c(src_des, "src_des"); tgt_des = _f32c(tgt_des, "tgt_des"); src_xyz = _f32c(src_xyz, "src_xyz"); tgt_xyz = _f32c(tgt_xyz, "tgt_xyz")
    dev = src_des.device
    P = src_off.numel() - 1
    totM, totN = src_des.shape[0], tgt_des.shape[0]
    T = torch.empty(P, 4, 4, dtype=torch.float32, device=dev)
    nm = torch.empty(P, dtype=torch.int32, device=dev); ni = torch.empty(P, dtype=torch.int32, device=dev)
    if ransac_splits is None:
        ransac_splits = _default_splits(P)
    L = _lib.lib()
    ws = _ws(L.bfr_register_workspace_bytes(P, max_M, max_N, totM, totN), dev, "register")
    _lib.check(L.bfr_register_batched(src_des.data_ptr(), src_xyz.data_ptr(), src_off.data_ptr(), tgt_des.data_ptr(), tgt_xyz.data_ptr(),
                                      tgt_off.data_ptr(), P, max_M, max_N, totM, totN, DESC_DIM, int(hypotheses), int(seed), int(pair_id_base),
                                      float(dist_th), float(similar_th), float(confidence), float(refine_thr), int(refine_iters), int(ransac_splits),
                                      T.data_ptr(), nm.data_ptr(), ni.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "bfr_register_batched")
    return T, nm, ni


def register_uniform(src_des, src_xyz, tgt_des, tgt_xyz, **kw):
    """[P,N,32] / [P,N,3] device tensors (every pair the same size) -> register_batched"""
    P, M = src_des.shape[0], src_des.shape[1]
    N = tgt_des.shape[1]
    dev = src_des.device
    so = (torch.arange(P + 1, dtype=torch.int32) * M).to(dev, non_blocking=True)
    to = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev, non_blocking=True)
    return register_batched(src_des.reshape(P * M, -1), src_xyz.reshape(P * M, 3), so, tgt_des.reshape(P * N, -1), tgt_xyz.reshape(P * N, 3), to,
                            M, N, **kw)


class HostRegistrar:
    """End-to-end path for inputs that live in (pinned) HOST memory: chunks of pairs are copied host->device, processed
    and their poses copied back on two alternating CUDA streams, so copies overlap compute.  The chunk loop runs inside the library
    (ONE bfr_register_uniform_host_chunked call per batch).  This is the call bench.py times for the `e2e` number."""

    def __init__(self, chunk_pairs, M, N, device, hypotheses=50000, dist_th=0.10, similar_th=0.8, refine_thr=0.10, refine_iters=20,
                 seed=0, ransac_splits=None, n_streams=2, confidence=1.0, pair_id_base=0):
        self.chunk, self.M, self.N, self.dev = chunk_pairs, M, N, torch.device(device)
        self.pair_id_base = pair_id_base
        self.kw = dict(hypotheses=hypotheses, dist_th=dist_th, similar_th=similar_th, refine_thr=refine_thr, refine_iters=refine_iters, seed=seed,
                       confidence=confidence)
        self.splits = ransac_splits if ransac_splits is not None else _default_splits(chunk_pairs)
        L = _lib.lib()
        nbytes = L.bfr_register_host_workspace_bytes(chunk_pairs, M, N, DESC_DIM)
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
        self.ws = [torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.dev) for _ in range(n_streams)]

    def run(self, src_des, src_xyz, tgt_des, tgt_xyz, T_out, n_mutual_out=None, n_inliers_out=None):
        """all arguments are HOST tensors ([P,M,32], [P,M,3], [P,N,32], [P,N,3] float32; outputs [P,4,4] float32 and
        optional [P] int32), ideally pinned.  Returns after everything has landed in the output tensors."""
        with torch.cuda.device(self.dev):
            return self._run(src_des, src_xyz, tgt_des, tgt_xyz, T_out, n_mutual_out, n_inliers_out)

    def _run(self, src_des, src_xyz, tgt_des, tgt_xyz, T_out, n_mutual_out, n_inliers_out):
        import ctypes as C
        L = _lib.lib()
        P = src_des.shape[0]
        k = self.kw
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            s.wait_stream(cur)
        ns = len(self.streams)
        ws = (C.c_void_p * ns)(*[w.data_ptr() for w in self.ws])
        st = (C.c_void_p * ns)(*[s.cuda_stream for s in self.streams])
        _lib.check(L.bfr_register_uniform_host_chunked(src_des.data_ptr(), src_xyz.data_ptr(), tgt_des.data_ptr(), tgt_xyz.data_ptr(), P, self.M, self.N, DESC_DIM,
                                                       int(self.chunk), int(k["hypotheses"]), int(k["seed"]), int(self.pair_id_base), float(k["dist_th"]),
                                                       float(k["similar_th"]), float(k["confidence"]), float(k["refine_thr"]), int(k["refine_iters"]), int(self.splits),
                                                       T_out.data_ptr(), 0 if n_mutual_out is None else n_mutual_out.data_ptr(),
                                                       0 if n_inliers_out is None else n_inliers_out.data_ptr(), ws, min(w.numel() for w in self.ws), st, ns),
                   "bfr_register_uniform_host_chunked")
        for s in self.streams:
            s.synchronize()
        return T_out
