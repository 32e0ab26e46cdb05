"""Synthetic fragment-pair generator: keypoints + 32-d unit descriptors with a planted SE(3).

Everything is float32 and drawn from ``torch.Generator().manual_seed(1000003 * cfg_id + first_pair)`` on the CPU (or on
the given device), vectorised over the pairs of one call.  For pair p and source keypoint i:

* ``src_xyz ~ U(box)``; ground truth ``R`` from a normalised N(0, I4) quaternion, ``t ~ U(-1, 1)^3 * t_scale``
* ``src_des = normalize(N(0, I_D))``; ``tgt_des[pi(i)] = normalize(src_des[i] + 0.05 N(0, I_D))`` for a random
  permutation ``pi`` (so practically every row is a mutual match)
* inliers (a random ``1 - outlier_ratio`` fraction): ``tgt_xyz[pi(i)] = R src_xyz[i] + t + sigma N(0, I3)``;
  outliers: ``tgt_xyz[pi(i)] = R u + t`` with a fresh ``u ~ U(box)``
"""
from dataclasses import dataclass

import torch

BOX_3DMATCH = ((-1.5, 1.5), (-1.5, 1.5), (0.0, 3.0))
BOX_KITTI = ((-50.0, 50.0), (-50.0, 50.0), (-3.0, 3.0))


@dataclass
class PairBatch:
    """P pairs with the same keypoint count N (row-major, contiguous)."""
    src_des: torch.Tensor    # [P, N, D]
    tgt_des: torch.Tensor    # [P, N, D]
    src_xyz: torch.Tensor    # [P, N, 3]
    tgt_xyz: torch.Tensor    # [P, N, 3]
    T_gt: torch.Tensor       # [P, 4, 4]
    perm: torch.Tensor       # [P, N] int64, target row of source row i
    inlier: torch.Tensor     # [P, N] bool, per source row

    @property
    def num_pairs(self):
        return self.src_des.shape[0]

    def to(self, device, non_blocking=False):
        return PairBatch(*[getattr(self, f).to(device, non_blocking=non_blocking) for f in
                           ("src_des", "tgt_des", "src_xyz", "tgt_xyz", "T_gt", "perm", "inlier")])


def quat_to_rot(q):
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(q.shape[:-1] + (3, 3))


def make_pairs(num_pairs, num_kpts, cfg_id=2, first_pair=0, desc_dim=32, outlier_ratio=0.70, outlier_ratio_hi=None,
               box=BOX_3DMATCH, sigma=0.01, t_scale=1.0, desc_noise=0.05, device="cpu"):
    """Generate ``num_pairs`` pairs.  ``outlier_ratio_hi`` (if given) draws the ratio per pair from
    U(outlier_ratio, outlier_ratio_hi) (config 3, low overlap)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(1000003 * cfg_id + first_pair)
    P, N, D = num_pairs, num_kpts, desc_dim
    f32 = dict(dtype=torch.float32, device=dev)
    lo = torch.tensor([b[0] for b in box], **f32)
    hi = torch.tensor([b[1] for b in box], **f32)

    def ubox(*shape):
        return lo + (hi - lo) * torch.rand(*shape, 3, generator=g, **f32)

    src_xyz = ubox(P, N)
    R = quat_to_rot(torch.randn(P, 4, generator=g, **f32))
    t = (torch.rand(P, 3, generator=g, **f32) * 2 - 1) * t_scale
    src_des = torch.nn.functional.normalize(torch.randn(P, N, D, generator=g, **f32), dim=-1)
    perm = torch.argsort(torch.rand(P, N, generator=g, **f32), dim=-1)
    noisy = torch.nn.functional.normalize(src_des + desc_noise * torch.randn(P, N, D, generator=g, **f32), dim=-1)
    tgt_des = torch.empty_like(noisy)
    tgt_des.scatter_(1, perm[:, :, None].expand(P, N, D), noisy)
    if outlier_ratio_hi is None:
        rho = torch.full((P, 1), float(outlier_ratio), **f32)
    else:
        rho = outlier_ratio + (outlier_ratio_hi - outlier_ratio) * torch.rand(P, 1, generator=g, **f32)
    rank = torch.argsort(torch.argsort(torch.rand(P, N, generator=g, **f32), dim=-1), dim=-1)
    inlier = rank < torch.round((1 - rho) * N).long()
    moved = src_xyz @ R.transpose(-1, -2) + t[:, None] + sigma * torch.randn(P, N, 3, generator=g, **f32)
    clutter = ubox(P, N) @ R.transpose(-1, -2) + t[:, None]
    tgt_by_src = torch.where(inlier[:, :, None], moved, clutter)
    tgt_xyz = torch.empty_like(tgt_by_src)
    tgt_xyz.scatter_(1, perm[:, :, None].expand(P, N, 3), tgt_by_src)
    T = torch.eye(4, **f32).repeat(P, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return PairBatch(src_des.contiguous(), tgt_des.contiguous(), src_xyz.contiguous(), tgt_xyz.contiguous(), T, perm, inlier)


def make_lrf_votes(R_gt, inlier, azi_n=20, seed=0, device="cpu"):
    """Synthetic inputs of the LRF vote (models/BUFFER.py:286-292) for `rows` matched keypoints: the inlier head's azimuth index
    `ind` [rows] (integers in [0, azi_n) as float32), source frames ss_R [rows,3,3] (random rotations) and target frames tt_R with
    tt_R Rz(ind 2 pi / azi_n) ss_R^T = R_gt for the inlier rows (random for the others).  R_gt: [rows,3,3] or [3,3]; inlier [rows] bool."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(7919 * seed + 13)
    rows = inlier.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    ss_R = quat_to_rot(torch.randn(rows, 4, generator=g, **f32))
    ind = torch.randint(0, azi_n, (rows,), generator=g, device=dev).float()
    ang = ind.double() * (2 * 3.141592653589793 / azi_n)
    Rz = torch.zeros(rows, 3, 3, dtype=torch.float64, device=dev)
    Rz[:, 0, 0] = torch.cos(ang); Rz[:, 0, 1] = -torch.sin(ang); Rz[:, 1, 0] = torch.sin(ang); Rz[:, 1, 1] = torch.cos(ang); Rz[:, 2, 2] = 1
    tt_R = (R_gt.to(dev).double().expand(rows, 3, 3) @ ss_R.double() @ Rz.transpose(-1, -2)).float()
    rnd = quat_to_rot(torch.randn(rows, 4, generator=g, **f32))
    tt_R = torch.where(inlier.to(dev)[:, None, None], tt_R, rnd)
    return ind.contiguous(), ss_R.contiguous(), tt_R.contiguous()


# BASELINE.json configs (sizes are the build's): keyword sets for make_pairs + RANSAC parameters
CONFIGS = {
    1: dict(gen=dict(cfg_id=1, num_kpts=5000, outlier_ratio=0.70), num_pairs=1, hypotheses=50000, dist_th=0.10, similar_th=0.8, refine_thr=0.10),
    2: dict(gen=dict(cfg_id=2, num_kpts=5000, outlier_ratio=0.70), num_pairs=1623, hypotheses=50000, dist_th=0.10, similar_th=0.8, refine_thr=0.10),
    3: dict(gen=dict(cfg_id=3, num_kpts=5000, outlier_ratio=0.90, outlier_ratio_hi=0.97), num_pairs=1623, hypotheses=500000, dist_th=0.10, similar_th=0.8, refine_thr=0.10),
    4: dict(gen=dict(cfg_id=4, num_kpts=20000, outlier_ratio=0.70, box=BOX_KITTI, sigma=0.1, t_scale=10.0), num_pairs=555, hypotheses=50000, dist_th=0.6, similar_th=0.9, refine_thr=1.2),
    5: dict(gen=dict(cfg_id=5, num_kpts=5000, outlier_ratio=0.70), num_pairs=16384, hypotheses=50000, dist_th=0.10, similar_th=0.8, refine_thr=0.10),
}


def rotation_error_rad(Ra, Rb):
    """geodesic angle between rotations, robust near 0 (uses the skew part, not acos of the trace)"""
    d = Ra.transpose(-1, -2).double() @ Rb.double()
    s = torch.stack([d[..., 2, 1] - d[..., 1, 2], d[..., 0, 2] - d[..., 2, 0], d[..., 1, 0] - d[..., 0, 1]], -1).norm(dim=-1) / 2
    c = (d.diagonal(dim1=-2, dim2=-1).sum(-1) - 1) / 2
    return torch.atan2(s, c)


def registration_recall(T_est, T_gt, rte_thresh=0.3, rre_thresh_deg=15.0):
    """Recall criterion of ThreeDMatch/test.py:263-283 (rte < 0.3 m and rre < 15 deg), vectorised over pairs."""
    rte = (T_est[:, :3, 3].double() - T_gt[:, :3, 3].double()).norm(dim=-1)
    rre = rotation_error_rad(T_est[:, :3, :3], T_gt[:, :3, :3]) * 180.0 / 3.141592653589793
    ok = (rte < rte_thresh) & (rre < rre_thresh_deg)
    return ok.double().mean().item(), rte, rre
