"""T1 semantic oracle: import the UNMODIFIED reference (models/BUFFER.py, utils/SE3.py) on CPU.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container), so it is used by
oracle/gen_golden.py to produce tests/golden/ fixtures and by tests that skip when the reference is absent.

The reference's third-party natives (knn_cuda, open3d, pointnet2_ops, kornia, torch_batch_svd, matplotlib) are not
installed; they are registered as stub modules.  The only stub with behaviour is
knn_cuda.KNN: brute-force Euclidean k-NN restated as torch.cdist + min (first-index ties), which is what
KNN_CUDA 0.2 computes at the call sites models/BUFFER.py:347,352.
"""
import math
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("BUFFER_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "BUFFER.py"))


class _KNN:
    """knn_cuda.KNN(k, transpose_mode=True)(ref[1,R,D], query[1,Q,D]) -> (dist[1,Q,k], idx[1,Q,k])"""

    def __init__(self, k, transpose_mode=False):
        assert k == 1 and transpose_mode
        self.k = k

    def __call__(self, ref, query):
        d = torch.cdist(query.double(), ref.double())          # exact-ish distances; ties -> first index
        dist, idx = d.min(dim=-1, keepdim=True)
        return dist.float(), idx


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_B = None


def load():
    """returns the reference's models.BUFFER module (cached)"""
    global _B
    if _B is not None:
        return _B
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _mod("open3d")
    p2 = _mod("pointnet2_ops"); p2.pointnet2_utils = _mod("pointnet2_ops.pointnet2_utils")
    mpl = _mod("matplotlib"); mpl.colors = _mod("matplotlib.colors"); mpl.cm = _mod("matplotlib.cm"); mpl.pyplot = _mod("matplotlib.pyplot")
    k = _mod("kornia"); k.geometry = _mod("kornia.geometry"); k.geometry.conversions = _mod("kornia.geometry.conversions")
    _mod("torch_batch_svd", svd=lambda x: torch.svd(x))
    _mod("knn_cuda", KNN=_KNN)
    for name in ("easydict", "nibabel", "tensorboardX"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _mod(name, EasyDict=dict, SummaryWriter=object)
    import models.BUFFER as B      # noqa: E402  (the reference, unmodified)
    _B = B
    return B


class _Cfg:
    def __init__(self, dataset):
        self.data = types.SimpleNamespace(dataset=dataset)


def self_stub(dataset="3DMatch"):
    return types.SimpleNamespace(config=_Cfg(dataset))


# --- thin callers of the reference's own functions ----------------------------------------------------------

def mutual_matching(src_des, tgt_des):
    """reference buffer.mutual_matching (models/BUFFER.py:335-359), descriptors as numpy float32"""
    B = load()
    s, t = B.buffer.mutual_matching(self_stub(), torch.from_numpy(np.asarray(src_des, np.float32)),
                                    torch.from_numpy(np.asarray(tgt_des, np.float32)))
    return np.asarray(s, np.int64), np.asarray(t, np.int64)


def rigid_transform_3d(A, Bp, weights=None, weight_threshold=0):
    B = load()
    w = None if weights is None else torch.from_numpy(np.array(weights, np.float32))
    return B.rigid_transform_3d(torch.from_numpy(np.asarray(A, np.float32)), torch.from_numpy(np.asarray(Bp, np.float32)),
                                w, weight_threshold).numpy()


def post_refinement(T0, src, tgt, dataset="3DMatch"):
    B = load()
    out = B.buffer.post_refinement(self_stub(dataset), torch.from_numpy(np.asarray(T0, np.float32))[None],
                                   torch.from_numpy(np.asarray(src, np.float32))[None], torch.from_numpy(np.asarray(tgt, np.float32))[None])
    return out[0].numpy()


def se3():
    load()
    import utils.SE3 as S
    return S


# --- restatement of the two inline blocks of buffer.forward (not callable in the reference) -------------------

def lrf_hypotheses(ind, ss_R, tt_R, ss_kpts, tt_kpts, azi_n=20):
    """models/BUFFER.py:294-301 verbatim in torch; kornia's angle_axis_to_rotation_matrix restated for a z axis
    (Rodrigues with kornia's eps=1e-6 normalisation)."""
    ind = torch.as_tensor(ind, dtype=torch.float32); ss_R = torch.as_tensor(ss_R, dtype=torch.float32)
    tt_R = torch.as_tensor(tt_R, dtype=torch.float32); ss_kpts = torch.as_tensor(ss_kpts, dtype=torch.float32)
    tt_kpts = torch.as_tensor(tt_kpts, dtype=torch.float32)
    angle = ind * 2 * np.pi / azi_n + 1e-6
    angle_axis = torch.zeros_like(ss_kpts)
    angle_axis[:, -1] = 1
    angle_axis = angle_axis * angle[:, None]
    theta = angle_axis.norm(dim=-1)
    wz = angle_axis[:, 2] / (theta + 1e-6)
    c, s = torch.cos(theta), torch.sin(theta)
    azi_R = torch.zeros(len(ind), 3, 3)
    azi_R[:, 0, 0] = c; azi_R[:, 0, 1] = -wz * s; azi_R[:, 1, 0] = wz * s; azi_R[:, 1, 1] = c
    azi_R[:, 2, 2] = c + wz * wz * (1 - c)
    R = tt_R @ azi_R @ ss_R.transpose(-1, -2)
    t = tt_kpts - (R @ ss_kpts.unsqueeze(-1)).squeeze()
    return R.numpy(), t.numpy()


def score_hypotheses(R, t, ss_kpts, tt_kpts, azi_n=20, inlier_th=1 / 3):
    """models/BUFFER.py:303-311 verbatim -> inlier_num [A], best_ind, inlier_ind, thr [A]"""
    R = torch.as_tensor(R, dtype=torch.float32); t = torch.as_tensor(t, dtype=torch.float32)
    ss_kpts = torch.as_tensor(ss_kpts, dtype=torch.float32); tt_kpts = torch.as_tensor(tt_kpts, dtype=torch.float32)
    tss_kpts = ss_kpts[None] @ R.transpose(-1, -2) + t[:, None]
    diffs = torch.sqrt(torch.sum((tss_kpts - tt_kpts[None]) ** 2, dim=-1))
    thr = torch.sqrt(torch.sum(ss_kpts ** 2, dim=-1)) * np.pi / azi_n * inlier_th
    sign = diffs < thr[None]
    inlier_num = torch.sum(sign, dim=-1)
    best_ind = torch.argmax(inlier_num)
    inlier_ind = torch.where(sign[best_ind] == True)[0].numpy()     # noqa: E712
    return inlier_num.numpy(), int(best_ind), inlier_ind, thr.numpy()


def ransac_open3d_semantics(src_kpts, tgt_kpts, corr, dist_th, similar_th, samples):
    """Open3D 0.13 RegistrationRANSACBasedOnCorrespondence semantics as a per-hypothesis
    Python loop that calls the REFERENCE's rigid_transform_3d / transform; `samples` [H,3] are the minimal sets
    (drawn by the shared Philox stream).  All hypotheses are evaluated; best = max count, ties -> lowest index.
    -> (T_best [4,4] float32, best_count, best_h, counts [H] with -1 for rejected)"""
    B = load()
    src = torch.as_tensor(src_kpts, dtype=torch.float32); tgt = torch.as_tensor(tgt_kpts, dtype=torch.float32)
    corr = np.asarray(corr)
    s_all = src[corr[:, 0]]; t_all = tgt[corr[:, 1]]
    best = (-1, -1, None); counts = np.full(len(samples), -1, np.int64)
    for h, smp in enumerate(samples):
        if len(set(int(x) for x in smp)) < 3:
            continue
        ps = s_all[smp]; pt = t_all[smp]
        ok = True
        for a in range(3):
            for b in range(a + 1, 3):
                ds = float((ps[a] - ps[b]).norm()); dt = float((pt[a] - pt[b]).norm())
                if ds < dt * similar_th or dt < ds * similar_th:
                    ok = False
        if not ok:
            continue
        T = B.rigid_transform_3d(ps[None], pt[None])[0]
        if float((B.transform(ps, T) - pt).norm(dim=-1).max()) > dist_th:
            continue
        d = (B.transform(s_all, T) - t_all).norm(dim=-1)
        n = int((d < dist_th).sum())
        counts[h] = n
        if n > best[0]:
            best = (n, h, T.numpy())
    if best[2] is None:
        return np.eye(4, dtype=np.float32), 0, -1, counts
    return best[2], best[0], best[1], counts


def recall_3dmatch(T_est, T_gt, rte_thresh=0.3, rre_thresh=15.0):
    """ThreeDMatch/test.py:263-270 success criterion -> (success, rte, rre_deg)"""
    rte = np.linalg.norm(T_est[:3, 3] - T_gt[:3, 3])
    rre = np.arccos(np.clip((np.trace(T_est[:3, :3].T @ T_gt[:3, :3]) - 1) / 2, -1 + 1e-16, 1 - 1e-16)) * 180 / math.pi
    return bool(rte < rte_thresh and rre < rre_thresh), float(rte), float(rre)
