"""ctypes binding of the CPU oracle (oracle/bfr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under buffer_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbfr_oracle.so")


def build(force=False):
    """Compile oracle/bfr_oracle.c -> oracle/libbfr_oracle.so (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "bfr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbfr_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_ransac.restype = C.c_uint64
        _lib.orc_mutual_select.restype = C.c_int
        _lib.orc_hypothesis.restype = C.c_int
        _lib.orc_count_inliers.restype = C.c_int32
        _lib.orc_kabsch_rotation.restype = C.c_int
        _lib.orc_ransac_confidence.restype = C.c_uint64
        _lib.orc_sqrt_threshold.restype = C.c_float
        _lib.orc_vote_threshold.restype = C.c_float
        _lib.orc_det_log.restype = C.c_double
        _lib.orc_ransac_exit_bound.restype = C.c_uint32
        _lib.orc_lrf_vote.restype = C.c_int32
    return _lib


def set_num_threads(n):
    """threads used by the batched (OpenMP over pairs) entry points; returns the effective count"""
    return int(lib().orc_set_num_threads(C.c_int(int(n))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def sample3(seed, pair_id, h, K):
    o = np.zeros(3, np.uint32)
    lib().orc_sample3(C.c_uint64(seed), C.c_uint32(pair_id), C.c_uint32(h), C.c_uint32(K), _p(o))
    return o


def half_sqnorms(x):
    x = _f32(x); out = np.zeros(x.shape[0], np.float32)
    lib().orc_half_sqnorms(_p(x), C.c_int(x.shape[0]), C.c_int(x.shape[1]), _p(out))
    return out


def mutual_nn(src, tgt, want_dist=False):
    """-> nn_s [M] int64, nn_t [N] int64 (, dist_s, dist_t)"""
    src = _f32(src); tgt = _f32(tgt)
    M, D = src.shape; N = tgt.shape[0]
    nn_s = np.zeros(M, np.int64); nn_t = np.zeros(N, np.int64)
    ds = np.zeros(M, np.float32) if want_dist else None
    dt = np.zeros(N, np.float32) if want_dist else None
    lib().orc_mutual_nn(_p(src), C.c_int(M), _p(tgt), C.c_int(N), C.c_int(D), _p(nn_s), _p(nn_t), _p(ds), _p(dt))
    return (nn_s, nn_t, ds, dt) if want_dist else (nn_s, nn_t)


def mutual_select(nn_s, nn_t):
    nn_s = np.ascontiguousarray(nn_s, np.int64); nn_t = np.ascontiguousarray(nn_t, np.int64)
    s = np.zeros(len(nn_s), np.int64); t = np.zeros(len(nn_s), np.int64)
    A = lib().orc_mutual_select(_p(nn_s), C.c_int(len(nn_s)), _p(nn_t), C.c_int(len(nn_t)), _p(s), _p(t))
    return s[:A].copy(), t[:A].copy()


def mutual_matching(src_des, tgt_des):
    """oracle of buffer.mutual_matching (models/BUFFER.py:335-359) -> (s_mids, t_mids) int64"""
    nn_s, nn_t = mutual_nn(src_des, tgt_des)
    return mutual_select(nn_s, nn_t)


def gather_corr(src_xyz, tgt_xyz, s_ids, t_ids):
    src_xyz = _f32(src_xyz); tgt_xyz = _f32(tgt_xyz)
    s_ids = np.ascontiguousarray(s_ids, np.int64); t_ids = np.ascontiguousarray(t_ids, np.int64)
    K = len(s_ids); corr = np.zeros((K, 8), np.float32)
    lib().orc_gather_corr(_p(src_xyz), _p(tgt_xyz), _p(s_ids), _p(t_ids), C.c_int(K), _p(corr))
    return corr


def kabsch_rotation(H):
    H = _f32(H).reshape(9); R = np.zeros(9, np.float32)
    ok = lib().orc_kabsch_rotation(_p(H), _p(R))
    return R.reshape(3, 3), bool(ok)


def hypothesis(corr, seed, pair_id, h, dist_th, similar_th):
    corr = _f32(corr); R = np.zeros(9, np.float32); t = np.zeros(3, np.float32)
    ok = lib().orc_hypothesis(_p(corr), C.c_uint32(corr.shape[0]), C.c_uint64(seed), C.c_uint32(pair_id), C.c_uint32(h),
                              C.c_float(dist_th), C.c_float(similar_th), _p(R), _p(t))
    return bool(ok), R.reshape(3, 3), t


def count_inliers(corr, R, t, dist_th):
    corr = _f32(corr); R = _f32(R).reshape(9); t = _f32(t)
    return int(lib().orc_count_inliers(_p(corr), C.c_uint32(corr.shape[0]), _p(R), _p(t), C.c_float(dist_th)))


def ransac(corr, seed, pair_id, H, dist_th, similar_th, h_begin=0, h_end=None, want_counts=False):
    """-> packed best (python int) [, counts per hypothesis (-1 = rejected)]"""
    corr = _f32(corr)
    h_end = H if h_end is None else h_end
    counts = np.zeros(h_end - h_begin, np.int32) if want_counts else None
    best = lib().orc_ransac(_p(corr), C.c_uint32(corr.shape[0]), C.c_uint64(seed), C.c_uint32(pair_id), C.c_uint32(h_begin),
                            C.c_uint32(h_end), C.c_float(dist_th), C.c_float(similar_th), _p(counts))
    return (int(best), counts) if want_counts else int(best)


def ransac_confidence(corr, seed, pair_id, H, dist_th, similar_th, confidence, h_begin=0):
    """Open3D's confidence early exit replayed sequentially (models/BUFFER.py:323-324) -> (packed best, iterations run)"""
    corr = _f32(corr); it = C.c_uint32(0)
    best = lib().orc_ransac_confidence(_p(corr), C.c_uint32(corr.shape[0]), C.c_uint64(seed), C.c_uint32(pair_id), C.c_uint32(h_begin),
                                       C.c_uint32(h_begin + H), C.c_float(dist_th), C.c_float(similar_th), C.c_float(confidence), C.byref(it))
    return int(best), int(it.value)


def sqrt_threshold(thr):
    return float(lib().orc_sqrt_threshold(C.c_float(thr)))


def det_sincos(a):
    sn = C.c_float(0); cs = C.c_float(0)
    lib().orc_det_sincos(C.c_float(a), C.byref(sn), C.byref(cs))
    return sn.value, cs.value


def det_log(v):
    return float(lib().orc_det_log(C.c_double(v)))


def ransac_exit_bound(count, K, confidence, max_iter):
    return int(lib().orc_ransac_exit_bound(C.c_uint32(count), C.c_uint32(K), C.c_double(det_log(1.0 - float(np.float32(confidence)))), C.c_uint32(max_iter)))


def lrf_vote(corr, ind, ss_R, tt_R, azi_n=20.0, inlier_th=1.0 / 3.0):
    """models/BUFFER.py:294-311 for one pair -> inlier_num [A] int32, best_ind, inlier_ind [K] int64"""
    corr = _f32(corr); ind = _f32(ind); ss_R = _f32(ss_R).reshape(-1, 9); tt_R = _f32(tt_R).reshape(-1, 9)
    A = corr.shape[0]
    counts = np.zeros(A, np.int32); best = C.c_int64(-1); sel = np.zeros(max(A, 1), np.int64)
    k = lib().orc_lrf_vote(_p(corr), C.c_int(A), _p(ind), _p(ss_R), _p(tt_R), C.c_float(azi_n), C.c_float(inlier_th), _p(counts), C.byref(best), _p(sel))
    return counts, best.value, sel[:k].copy()


def pose_from_votes_batched(corr, corr_off, corr_cnt, ind, ss_R, tt_R, H, seed, pair_id_base, dist_th, similar_th, confidence=1.0,
                            refine_thr=0.1, refine_iters=20, azi_n=20.0, inlier_th=1.0 / 3.0):
    """vote -> RANSAC on the voted subset -> refinement on all matches (models/BUFFER.py:291-329), OpenMP over pairs
    -> T [P,4,4], n_vote_inliers [P], n_inliers [P]"""
    corr = _f32(corr); ind = _f32(ind); ss_R = _f32(ss_R).reshape(-1, 9); tt_R = _f32(tt_R).reshape(-1, 9)
    corr_off = np.ascontiguousarray(corr_off, np.int32); corr_cnt = np.ascontiguousarray(corr_cnt, np.int32)
    P = len(corr_cnt)
    T = np.zeros((P, 16), np.float32); nv = np.zeros(P, np.int32); ni = np.zeros(P, np.int32)
    lib().orc_pose_from_votes_batched(_p(corr), _p(corr_off), _p(corr_cnt), C.c_int(P), _p(ind), _p(ss_R), _p(tt_R), C.c_float(azi_n), C.c_float(inlier_th),
                                      C.c_int(H), C.c_uint64(seed), C.c_uint32(pair_id_base), C.c_float(dist_th), C.c_float(similar_th), C.c_float(confidence),
                                      C.c_float(refine_thr), C.c_int(refine_iters), _p(T), _p(nv), _p(ni))
    return T.reshape(P, 4, 4), nv, ni


def ransac_finalize(corr, seed, pair_id, best, dist_th, similar_th):
    """-> T [4,4] float32, inlier count, best hypothesis index (-1 if none)"""
    corr = _f32(corr); T = np.zeros(16, np.float32); cnt = C.c_int32(0); bh = C.c_int64(0)
    lib().orc_ransac_finalize(_p(corr), C.c_uint32(corr.shape[0]), C.c_uint64(seed), C.c_uint32(pair_id), C.c_uint64(best),
                              C.c_float(dist_th), C.c_float(similar_th), _p(T), C.byref(cnt), C.byref(bh))
    return T.reshape(4, 4), cnt.value, bh.value


def lrf_hypotheses(cs, ss_R, tt_R, ss_kpts, tt_kpts):
    cs = _f32(cs); ss_R = _f32(ss_R); tt_R = _f32(tt_R); ss_kpts = _f32(ss_kpts); tt_kpts = _f32(tt_kpts)
    A = ss_kpts.shape[0]; R = np.zeros((A, 3, 3), np.float32); t = np.zeros((A, 3), np.float32)
    lib().orc_lrf_hypotheses(_p(cs), _p(ss_R), _p(tt_R), _p(ss_kpts), _p(tt_kpts), C.c_int(A), _p(R), _p(t))
    return R, t


def score_hypotheses(R, t, src, tgt, thr):
    """-> counts [H] int32, best index, inlier mask [C] bool.  thr: scalar or [C]."""
    R = _f32(R); t = _f32(t); src = _f32(src); tgt = _f32(tgt)
    H = R.shape[0]; Cn = src.shape[0]
    thr_arr = None if np.isscalar(thr) else _f32(thr)
    counts = np.zeros(H, np.int32); best = C.c_int64(-1); mask = np.zeros(Cn, np.uint8)
    lib().orc_score_hypotheses(_p(R), _p(t), C.c_int(H), _p(src), _p(tgt), C.c_int(Cn), _p(thr_arr),
                               C.c_float(float(thr) if thr_arr is None else 0.0), _p(counts), C.byref(best), _p(mask))
    return counts, best.value, mask.astype(bool)


def rigid_transform_3d(A, B, weights=None, weight_threshold=0.0):
    """A, B: [bs, n, 3]; weights [bs, n] or None -> [bs, 4, 4] float32 (models/BUFFER.py:424-464)"""
    A = _f32(A); B = _f32(B)
    bs, n, _ = A.shape
    w = None if weights is None else _f32(weights)
    T = np.zeros((bs, 16), np.float32)
    for b in range(bs):
        lib().orc_rigid_transform_3d(_p(A[b]), _p(B[b]), _p(w[b]) if w is not None else None, C.c_int(n),
                                     C.c_float(weight_threshold), _p(T[b]))
    return T.reshape(bs, 4, 4)


def post_refinement(T0, corr, thr, max_iter=20):
    """-> T [4,4], iterations run, last inlier count (models/BUFFER.py:382-418)"""
    T0 = _f32(T0).reshape(16); corr = _f32(corr); T = np.zeros(16, np.float32); it = C.c_int32(0); li = C.c_int32(0)
    lib().orc_post_refinement(_p(T0), _p(corr), C.c_int(corr.shape[0]), C.c_float(thr), C.c_int(max_iter), _p(T), C.byref(it), C.byref(li))
    return T.reshape(4, 4), it.value, li.value


def get_matching_indices(source, target, relt_pose, search_voxel_size, want_nn=False):
    """oracle of buffer.get_matching_indices (models/BUFFER.py:361-380) -> match_inds [C,2] int64 (, nn [N], dist [N])"""
    source = _f32(source); target = _f32(target); T = _f32(relt_pose).reshape(16)
    N, M = source.shape[0], target.shape[0]
    pairs = np.zeros((max(N, 1), 2), np.int64); nn = np.zeros(N, np.int64); dist = np.zeros(N, np.float32)
    lib().orc_get_matching_indices.restype = C.c_int
    c = lib().orc_get_matching_indices(_p(source), C.c_int(N), _p(target), C.c_int(M), _p(T), C.c_float(search_voxel_size), _p(pairs), _p(nn), _p(dist))
    return (pairs[:c].copy(), nn, dist) if want_nn else pairs[:c].copy()


def furthest_point_sample(xyz, npoint):
    """pointnet2_ops.furthest_point_sample restated (models/BUFFER.py:266-267): xyz [B,N,3] -> idx [B,npoint] int32"""
    xyz = _f32(xyz); B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    for b in range(B):
        lib().orc_furthest_point_sample(_p(xyz[b]), C.c_int(N), C.c_int(npoint), _p(idx[b]))
    return idx


def svd3(x):
    """batched 3x3 SVD (torch_batch_svd contract, utils/common.py:715): x [B,3,3] -> u [B,3,3], s [B,3] descending, v [B,3,3]"""
    x = _f32(x).reshape(-1, 9); B = x.shape[0]
    u = np.zeros((B, 9), np.float32); s = np.zeros((B, 3), np.float32); v = np.zeros((B, 9), np.float32)
    for b in range(B):
        lib().orc_svd3(_p(x[b]), _p(u[b]), _p(s[b]), _p(v[b]))
    return u.reshape(B, 3, 3), s, v.reshape(B, 3, 3)


def register_batched(src_des, src_xyz, src_off, tgt_des, tgt_xyz, tgt_off, H, seed, pair_id_base, dist_th, similar_th,
                     refine_thr, refine_iters=20, confidence=1.0):
    """whole back end on the CPU for a batch of pairs (OpenMP over pairs) -> T [P,4,4], n_mutual [P], n_inliers [P]"""
    src_des = _f32(src_des); tgt_des = _f32(tgt_des); src_xyz = _f32(src_xyz); tgt_xyz = _f32(tgt_xyz)
    src_off = np.ascontiguousarray(src_off, np.int32); tgt_off = np.ascontiguousarray(tgt_off, np.int32)
    P = len(src_off) - 1
    T = np.zeros((P, 16), np.float32); nm = np.zeros(P, np.int32); ni = np.zeros(P, np.int32)
    lib().orc_register_batched(_p(src_des), _p(src_xyz), _p(src_off), _p(tgt_des), _p(tgt_xyz), _p(tgt_off), C.c_int(P),
                               C.c_int(src_des.shape[1]), C.c_int(H), C.c_uint64(seed), C.c_uint32(pair_id_base),
                               C.c_float(dist_th), C.c_float(similar_th), C.c_float(confidence), C.c_float(refine_thr), C.c_int(refine_iters),
                               _p(T), _p(nm), _p(ni))
    return T.reshape(P, 4, 4), nm, ni
