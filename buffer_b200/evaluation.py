"""Evaluation side of the back end (second half): Redwood / 3DMatch trajectory files and registration recall.

Host-side mirror of the evaluation helpers the reference keeps in its test scripts, same names and argument meaning so a caller can
switch imports:

* ``read_trajectory`` / ``read_trajectory_info``     - ThreeDMatch/test.py:18-56, :59-89   (Redwood ``.log`` / ``.info`` readers)
* ``computeTransformationErr``                       - ThreeDMatch/test.py:92-110          (RMSE proxy ``e^T info e / info[0,0]``)
* ``evaluate_registration``                          - ThreeDMatch/test.py:113-173         (precision / recall over non-consecutive pairs)
* ``extract_corresponding_trajectors``               - ThreeDMatch/test.py:176-197
* ``write_trajectory_entry``                         - the inline writer ThreeDMatch/test.py:250-261 (appends ``inv(T_est)`` to a ``.log``)
* ``pair_errors`` / ``dgr_recall``                   - the inline "recall of DGR" block ThreeDMatch/test.py:263-283 (KITTI: KITTI/test.py:65-87,
                                                      ETH: generalization/ThreeD2ETH/test.py:66-67) - thresholds per dataset below

``mat2quat`` restates ``nibabel.quaternions.mat2quat`` (third-party, unpinned in the reference's README; Bar-Itzhack's eigenvector method:
the unit quaternion (w, x, y, z) is the eigenvector of the largest eigenvalue of the symmetric 4x4 ``K`` matrix, sign chosen so that w >= 0).

Everything here is numpy on the host: the per-pair work is a 6-vector quadratic form, and the poses it consumes are the [P,4,4] array the CUDA
back end already returns - ``evaluate_poses`` evaluates a whole batch of device poses with one copy.
"""
import math
import os

import numpy as np

# success thresholds of the inline recall blocks (rte in metres, rre in degrees)
RECALL_THRESHOLDS = {
    "3DMatch": (0.3, 15.0),      # ThreeDMatch/test.py:264-265
    "3DLoMatch": (0.3, 15.0),
    "KITTI": (0.3, 1.0),         # as written in KITTI/test.py:66-67
    "ETH": (0.3, 2.0),           # generalization/ThreeD2ETH/test.py:66-67
}


def mat2quat(M):
    """3x3 rotation (or near-rotation) matrix -> unit quaternion (w, x, y, z), w >= 0 (nibabel.quaternions.mat2quat)."""
    M = np.asarray(M, dtype=np.float64)
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = M.flat
    K = np.array([[Qxx - Qyy - Qzz, 0, 0, 0],
                  [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                  [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                  [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)               # uses the lower triangle
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return q


def read_trajectory(filename, dim=4):
    """Redwood ``.log``: blocks of one key line ``i \\t j \\t n`` and ``dim`` matrix rows.  Returns (keys [n,3] array of strings,
    traj [n,dim,dim] float32) - ThreeDMatch/test.py:18-56."""
    with open(filename) as f:
        lines = f.readlines()
    keys = [[c.strip() for c in line.split("\t")[0:3]] for line in lines[0::dim + 1]]
    rows = [line.split("\t")[0:dim] for i, line in enumerate(lines) if i % (dim + 1) != 0]
    traj = np.asarray(rows, dtype=np.float32).reshape(-1, dim, dim)
    return np.asarray(keys), traj


def read_trajectory_info(filename, dim=6):
    """Redwood ``.info``: blocks of one key line ``i j n_frame`` and six rows of the 6x6 information matrix.  Returns
    (n_frame of the last block, info [n,dim,dim] float32) - ThreeDMatch/test.py:59-89."""
    with open(filename) as f:
        contents = f.readlines()
    n_pairs = len(contents) // 7
    assert len(contents) == 7 * n_pairs
    n_frame, info = 0, []
    for i in range(n_pairs):
        _, _, n_frame = [int(x) for x in contents[i * 7].strip().split()]
        info.append(np.concatenate([np.array(row.split(), dtype=np.float64).reshape(1, -1) for row in contents[i * 7 + 1:i * 7 + 7]], axis=0))
    return n_frame, np.asarray(info, dtype=np.float32).reshape(-1, dim, dim)


def _error_vectors(trans):
    """[n,4,4] -> [n,6] rows (tx, ty, tz, qx, qy, qz): translation and vector part of the rotation quaternion"""
    trans = np.asarray(trans, dtype=np.float64).reshape(-1, 4, 4)
    quat_xyz = np.stack([mat2quat(T[:3, :3])[1:] for T in trans]) if len(trans) else np.zeros((0, 3))
    return np.concatenate([trans[:, :3, 3], quat_xyz], axis=1)


def _information_errors(trans, info):
    """batched quadratic form ``e^T I e / I[0,0]`` with ``e = _error_vectors(trans)``; trans [n,4,4], info [n,6,6] -> [n]"""
    e = _error_vectors(trans)
    info = np.asarray(info, dtype=np.float64).reshape(-1, 6, 6)
    return np.einsum("ni,nij,nj->n", e, info, e) / info[:, 0, 0]


def computeTransformationErr(trans, info):
    """``e^T info e / info[0,0]`` for the 6-vector ``e`` = (translation, quaternion xyz) of the 4x4 ``trans`` (ThreeDMatch/test.py:92-110)."""
    return float(_information_errors(np.asarray(trans)[None], np.asarray(info)[None])[0])


def evaluate_registration(num_fragment, result, result_pairs, gt_pairs, gt, gt_info, err2=0.2):
    """3DMatch / Redwood protocol (ThreeDMatch/test.py:113-173): only non-consecutive ground-truth pairs count; an estimate is good when
    the information-weighted error of ``inv(gt) @ pose`` is at most ``err2**2``.  Returns (precision, recall, flags, transformation_errors);
    flags: 0 good, 1 wrong, 2 pair not in the ground truth.  Like the reference, ground-truth entry 0 can never be hit (its index doubles as
    the "no pair" marker of the lookup table).  Vectorised: one table lookup for all estimates, one batched quadratic form."""
    gt_pairs = np.asarray(gt_pairs); result_pairs = np.asarray(result_pairs)
    gi, gj = gt_pairs[:, 0].astype(np.int64), gt_pairs[:, 1].astype(np.int64)
    far = (gj - gi) > 1
    lookup = np.zeros((num_fragment, num_fragment), dtype=np.int64)
    lookup[gi[far], gj[far]] = np.nonzero(far)[0]                    # later duplicates overwrite earlier ones, as a sequential fill would
    ri, rj = result_pairs[:, 0].astype(np.int64), result_pairs[:, 1].astype(np.int64)
    hit = lookup[ri, rj]                                              # ground-truth row of every estimate, 0 = not evaluated
    known = hit > 0
    errors = np.full(len(result_pairs), np.nan)
    if known.any():
        rel = np.linalg.inv(np.asarray(gt)[hit[known]]) @ np.asarray(result)[known]      # in the trajectories' own dtype (float32 from the readers), like the reference
        errors[known] = _information_errors(rel, np.asarray(gt_info)[hit[known]])
    ok = known & (errors <= err2 ** 2)
    flags = np.where(ok, 0, np.where(known, 1, 2)).tolist()
    evaluated = int(known.sum()) if known.any() else 1e6
    return ok.sum() * 1.0 / evaluated, ok.sum() * 1.0 / np.count_nonzero(lookup), flags, errors


def extract_corresponding_trajectors(est_pairs, gt_pairs, gt_traj):
    """ground-truth transforms of exactly the estimated pairs (ThreeDMatch/test.py:176-197); like the reference it overwrites the third
    column of ``est_pairs`` with the scene's fragment count.  One broadcast comparison instead of a search per pair; an estimated pair
    without a ground-truth row keeps a zero matrix."""
    gt_pairs = np.asarray(gt_pairs)
    est_pairs[:, 2] = gt_pairs[0][2]
    same = (np.asarray(est_pairs)[:, None, :] == gt_pairs[None, :, :]).all(axis=2)          # [n_est, n_gt]
    out = np.zeros((len(est_pairs), 4, 4))
    found = same.any(axis=1)
    out[found] = np.asarray(gt_traj)[same.argmax(axis=1)[found]]
    return out


def write_trajectory_entry(path, src_id, tgt_id, trans_est):
    """append one estimate to a Redwood ``.log`` exactly as ThreeDMatch/test.py:250-261 does: key line ``src \\t tgt \\t 1`` and the rows of
    ``inv(trans_est)`` (``None`` -> identity, :242-245), values printed with Python's shortest-repr float formatting."""
    trans_est = np.eye(4, 4) if trans_est is None else np.asarray(trans_est)
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        os.makedirs(d)
    trans = np.linalg.inv(trans_est)
    with open(path, "a+") as f:
        f.write(f"{src_id}\t {tgt_id}\t  1\n")
        for r in range(4):
            f.write(f"{trans[r, 0]}\t {trans[r, 1]}\t {trans[r, 2]}\t {trans[r, 3]}\t \n")


def pair_errors(trans_est, trans_gt):
    """(rte [m], rre [deg]) of one pair as the inline block ThreeDMatch/test.py:266-269 computes them (float64)."""
    trans_est, trans_gt = np.asarray(trans_est, dtype=np.float64), np.asarray(trans_gt, dtype=np.float64)
    rte = np.linalg.norm(trans_est[:3, 3] - trans_gt[:3, 3])
    rre = np.arccos(np.clip((np.trace(trans_est[:3, :3].T @ trans_gt[:3, :3]) - 1) / 2, -1 + 1e-16, 1 - 1e-16)) * 180 / math.pi
    return rte, rre


def dgr_recall(trans_est, trans_gt, dataset="3DMatch"):
    """Recall / TE / RE of a batch of poses (ThreeDMatch/test.py:263-283): a pair succeeds iff ``rte < rte_thresh and rre < rre_thresh``;
    TE and RE are means over the successful pairs.  ``trans_est``, ``trans_gt``: [P,4,4].  Returns (recall, te, re, states [P,3])."""
    rte_thresh, rre_thresh = RECALL_THRESHOLDS[dataset]
    states = []
    for Te, Tg in zip(np.asarray(trans_est), np.asarray(trans_gt)):
        rte, rre = pair_errors(Te, Tg)
        states.append(np.array([rte < rte_thresh and rre < rre_thresh, rte, rre]))
    states = np.array(states).reshape(-1, 3)
    ok = states[:, 0] == 1
    recall = states[:, 0].sum() / max(states.shape[0], 1)
    te = states[ok, 1].mean() if ok.any() else float("nan")
    re = states[ok, 2].mean() if ok.any() else float("nan")
    return recall, te, re, states


def evaluate_poses(T_device, src_ids, tgt_ids, log_path=None, trans_gt=None, dataset="3DMatch"):
    """Glue for the CUDA back end: ``T_device`` is the [P,4,4] pose tensor returned by ``backend.register_*`` (device or host).  One copy
    to the host, then optionally the Redwood ``.log`` entries (``log_path``) and, when ground truth is given, the DGR recall."""
    T = T_device.detach().cpu().numpy() if hasattr(T_device, "detach") else np.asarray(T_device)
    if log_path is not None:
        for p in range(T.shape[0]):
            write_trajectory_entry(log_path, src_ids[p], tgt_ids[p], T[p])
    if trans_gt is None:
        return None
    gt = trans_gt.detach().cpu().numpy() if hasattr(trans_gt, "detach") else np.asarray(trans_gt)
    return dgr_recall(T, gt, dataset)
