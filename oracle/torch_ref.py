"""The reference's CPU *torch* path of the correspondence-and-pose stage, restated so that it can be TIMED on a box without
/root/reference (BASELINE config 1: one 3DMatch-sized pair on the host cores).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `cpu_baseline_torch` leg and tests/): nothing under buffer_b200/ imports it.
It keeps the COMPUTATIONAL STRUCTURE of the reference - that is what is being timed:

  mutual_matching      models/BUFFER.py:335-359   two brute-force k=1 searches over the full M x N distance matrix (KNN_CUDA 0.2 is a
                                                  CUDA-only third-party wheel: on the CPU the same search is torch.cdist + min)
  kabsch               models/BUFFER.py:424-464   weighted centroids with the +1e-6 denominator, diag_embed(weights) [n x n], torch.svd,
                                                  det fix on the last column
  post_refinement      models/BUFFER.py:382-418   <= 20 rounds, one host sync per round, Kabsch on the boolean-gathered inliers
  ransac               models/BUFFER.py:313-326   Open3D 0.13's loop as a per-hypothesis Python loop on the torch Kabsch
                                                  above - Open3D itself (C++/OpenMP) is a third-party wheel that is absent here, so
                                                  the loop is timed on a bounded number of hypotheses and scaled to H (stated in the output)

tests/test_oracle_golden.py checks these restatements against the fixtures produced by the unmodified reference.
"""
import time

import numpy as np
import torch


def warp(points, T):
    """utils/SE3.py:43-57 for one [n,3] cloud and one [4,4] transform"""
    return (T[:3, :3] @ points.T + T[:3, 3:4]).T


def mutual_matching(src_des, tgt_des):
    """-> (s_mids, t_mids) numpy int64, s_mids ascending"""
    d = torch.cdist(src_des[None], tgt_des[None])[0]
    nn_s = d.min(dim=1)[1].numpy()                     # first minimum on ties, like the stub of SURVEY App. A
    nn_t = d.min(dim=0)[1].numpy()
    s_mids = np.where(nn_t[nn_s] == np.arange(len(nn_s)))[0]
    return s_mids, nn_s[s_mids]


def kabsch(A, B, weights=None, weight_threshold=0.0):
    """[bs,n,3] x2 (+ [bs,n]) -> [bs,4,4]; same operations as rigid_transform_3d incl. the O(n^2) diag_embed"""
    bs = A.shape[0]
    w = torch.ones_like(A[:, :, 0]) if weights is None else weights
    w[w < weight_threshold] = 0
    den = w.sum(dim=1, keepdim=True)[:, :, None] + 1e-6
    cA = (A * w[:, :, None]).sum(dim=1, keepdim=True) / den
    cB = (B * w[:, :, None]).sum(dim=1, keepdim=True) / den
    H = (A - cA).transpose(1, 2) @ torch.diag_embed(w) @ (B - cB)
    U, _, V = torch.svd(H)
    D = torch.eye(3).repeat(bs, 1, 1)
    D[:, 2, 2] = torch.det(V @ U.transpose(1, 2))
    R = V @ D @ U.transpose(1, 2)
    T = torch.eye(4).repeat(bs, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3:4] = cB.transpose(1, 2) - R @ cA.transpose(1, 2)
    return T


def post_refinement(T0, src, tgt, thr=0.10, rounds=20):
    """[4,4], [n,3], [n,3] -> [4,4]"""
    T = T0.clone()
    prev = 0
    for _ in range(rounds):
        dist = torch.norm(warp(src, T) - tgt, dim=-1)
        keep = dist < thr
        n = int(keep.sum())                            # the reference's per-round host sync (:406)
        if n == prev:
            break
        prev = n
        T = kabsch(src[None, keep], tgt[None, keep], (1 / (1 + (dist / thr) ** 2))[None, keep])[0]
    return T


def ransac(ss, tt, samples, dist_th, similar_th):
    """Open3D-semantics loop over the given minimal sets [H,3] (with replacement; repeated index = invalid) on all K correspondences
    ss[i] <-> tt[i].  -> (T_best [4,4], best_count, best_h, valid hypotheses)"""
    best_n, best_h, best_T, valid = -1, -1, torch.eye(4), 0
    for h in range(len(samples)):
        i, j, k = (int(x) for x in samples[h])
        if i == j or i == k or j == k:
            continue
        ps, pt = ss[[i, j, k]], tt[[i, j, k]]
        ok = True
        for a, b in ((0, 1), (0, 2), (1, 2)):          # CorrespondenceCheckerBasedOnEdgeLength
            ds, dt = float(torch.norm(ps[a] - ps[b])), float(torch.norm(pt[a] - pt[b]))
            if ds < dt * similar_th or dt < ds * similar_th:
                ok = False
                break
        if not ok:
            continue
        T = kabsch(ps[None], pt[None])[0]
        if float(torch.norm(warp(ps, T) - pt, dim=-1).max()) > dist_th:     # CorrespondenceCheckerBasedOnDistance
            continue
        valid += 1
        n = int((torch.norm(warp(ss, T) - tt, dim=-1) < dist_th).sum())
        if n > best_n:
            best_n, best_h, best_T = n, h, T
    return best_T, max(best_n, 0), best_h, valid


def time_pair(src_des, tgt_des, src_xyz, tgt_xyz, samples_fn, hypotheses, timed_hypotheses, dist_th, similar_th, refine_thr, threads):
    """One pair through the whole stage on the host cores.  RANSAC runs `timed_hypotheses` of the `hypotheses` iterations and its
    time is scaled linearly (every iteration costs the same on average: no early exit, BASELINE.md par. 4).
    samples_fn(K, n) -> [n,3] minimal sets.  -> dict of seconds and the pose."""
    torch.set_num_threads(threads)
    best = None
    for _ in range(3):                                 # matching: best of 3 (first call warms the allocator / thread pool)
        t0 = time.perf_counter()
        s_mids, t_mids = mutual_matching(src_des, tgt_des)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ss, tt = src_xyz[s_mids], tgt_xyz[t_mids]
    samples = samples_fn(len(s_mids), timed_hypotheses)
    t0 = time.perf_counter()
    T, n, h, valid = ransac(ss, tt, samples, dist_th, similar_th)
    t_ransac = time.perf_counter() - t0
    t0 = time.perf_counter()
    Tr = post_refinement(T, ss, tt, refine_thr)
    t_refine = time.perf_counter() - t0
    scaled = t_ransac * hypotheses / max(timed_hypotheses, 1)
    return {"match_s": best, "ransac_s_measured": t_ransac, "ransac_hypotheses_measured": int(timed_hypotheses), "ransac_s_scaled": scaled,
            "refine_s": t_refine, "total_s": best + scaled + t_refine, "mutual": int(len(s_mids)), "valid_hypotheses": int(valid),
            "best_count": int(n), "T": Tr.numpy()}
