"""SE(3) helpers with the call surface of the reference's utils/SE3.py (file:line cited per function).

Kept on the host in plain torch / numpy: they are O(1)-sized glue around the kernels (4x4 matrices), not a hot loop;
point transforms on the hot path happen inside the CUDA kernels (resid2 in csrc/bfr_common.cuh).
Like the reference module, this one is meant to be star-imported and deliberately re-exports ``torch``, ``np`` and
``random`` (reference consumers such as models/BUFFER.py:8,19,28 rely on that leak).

Behaviour notes (SURVEY.md §7 "quirks"): ``integrate_trans`` always builds on ``torch.eye(4)`` / ``np.eye(4)`` so torch
results are float32; the reference's batched *numpy* branches call tensor-only methods (.view/.permute) and raise —
here they work (a superset, nothing that worked before changes).
"""
import random  # noqa: F401  (re-exported on purpose)

import numpy as np  # noqa: F401
import torch  # noqa: F401

__all__ = ["torch", "np", "random", "rotation_matrix", "translation_matrix", "transform", "decompose_trans",
           "integrate_trans", "concatenate"]


def _is_torch(x):
    return isinstance(x, torch.Tensor)


def _swap_last(x):
    return x.transpose(-1, -2) if _is_torch(x) else np.swapaxes(x, -1, -2)


def rotation_matrix(num_axis, augment_rotation):
    """utils/SE3.py:5-30 — random rotation; num_axis 0 -> identity, 1 -> about z only, 3 -> Rx Ry Rz.
    Angles are ``np.random.rand(3) * 2 pi * augment_rotation`` (same RNG stream as the reference)."""
    if num_axis not in (0, 1, 3):
        raise AssertionError("num_axis must be 0, 1 or 3")
    if num_axis == 0:
        return np.eye(3)
    ax, ay, az = np.random.rand(3) * 2 * np.pi * augment_rotation
    cz, sz = np.cos(az), np.sin(az)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    if num_axis == 1:
        return Rz
    cx, sx, cy, sy = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return Rx @ Ry @ Rz


def translation_matrix(augment_translation):
    """utils/SE3.py:32-41 — random translation in [0, augment_translation)^3 as a [3,1] column."""
    return (np.random.rand(3) * augment_translation).reshape(3, 1)


def transform(pts, trans):
    """utils/SE3.py:43-57 — R @ p + t for pts [n,3] with trans [4,4], or pts [bs,n,3] with trans [bs,4,4]."""
    R, t = decompose_trans(trans)
    return _swap_last(R @ _swap_last(pts) + t)


def decompose_trans(trans):
    """utils/SE3.py:59-71 — (R [..,3,3], t [..,3,1]) views of a [4,4] or [bs,4,4] transform."""
    return trans[..., :3, :3], trans[..., :3, 3:4]


def integrate_trans(R, t):
    """utils/SE3.py:73-96 — assemble [4,4] / [bs,4,4] from R and t ([3,1], or anything reshapeable to [bs,3,1])."""
    batched = R.ndim == 3
    if _is_torch(R):
        out = torch.eye(4).to(R.device)
        if batched:
            out = out[None].repeat(R.shape[0], 1, 1)
    else:
        out = np.eye(4)
        if batched:
            out = np.tile(out[None], (R.shape[0], 1, 1))
    out[..., :3, :3] = R
    out[..., :3, 3:4] = t.reshape(-1, 3, 1) if batched else t
    return out


def concatenate(trans1, trans2):
    """utils/SE3.py:98-112 — trans1 @ trans2 composed through (R, t)."""
    R1, t1 = decompose_trans(trans1)
    R2, t2 = decompose_trans(trans2)
    return integrate_trans(R1 @ R2, R1 @ t2 + t1)
