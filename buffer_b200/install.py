"""Drop the B200 back end into a loaded reference module (models/BUFFER.py) without editing it.

``install(B)`` rebinds, on the reference module object ``B`` (``import models.BUFFER as B``):

  B.buffer.mutual_matching   -> K1 (models/BUFFER.py:335-359)
  B.buffer.post_refinement   -> K4 loop (models/BUFFER.py:382-418), threshold chosen from self.config.data.dataset
  B.buffer.get_matching_indices -> 3-D nearest neighbour + voxel filter (models/BUFFER.py:361-380)
  B.rigid_transform_3d       -> K4 weighted Kabsch (models/BUFFER.py:424-464)
  B.KNN                      -> K1-backed k=1 nearest neighbour for 32-d descriptors (models/BUFFER.py:347,352)
  B.o3d.pipelines.registration.registration_ransac_based_on_correspondence (+ the estimator / checker / criteria
                               constructors the call site builds, models/BUFFER.py:318-324) -> K2+K3

so ``buffer.forward`` runs its test branch unchanged on top of the CUDA kernels.  INTEGRATION.md has the details.
"""
import types

import numpy as np
import torch

from . import backend


class _Tagged:
    def __init__(self, kind, *args):
        self.kind, self.args = kind, args


class _KNN1:
    def __init__(self, k=1, transpose_mode=True):
        if k != 1 or not transpose_mode:
            raise NotImplementedError("buffer_b200 KNN: only k=1, transpose_mode=True (the reference's hot-path use)")

    def __call__(self, ref, query):
        return backend.knn1(ref, query)


def _points(pcd):
    pts = getattr(pcd, "points", pcd)
    return torch.as_tensor(np.asarray(pts), dtype=torch.float32).cuda()


def _ransac_shim(pcd0, pcd1, corr, max_correspondence_distance, estimation=None, ransac_n=3, checkers=(), criteria=None, seed=0):
    if ransac_n != 3:
        raise NotImplementedError("buffer_b200 RANSAC: ransac_n must be 3 (models/BUFFER.py:321)")
    similar_th, iter_n, confidence = 0.0, 100000, 0.999
    for c in checkers:
        if getattr(c, "kind", "") == "edge":
            similar_th = c.args[0]
    if criteria is not None:
        iter_n, confidence = criteria.args[0], criteria.args[1]
    return backend.registration_ransac_based_on_correspondence(_points(pcd0), _points(pcd1), np.asarray(corr), max_correspondence_distance,
                                                               similar_th, iter_n, confidence, seed)


def open3d_shim():
    """a minimal `o3d` namespace exposing exactly what models/BUFFER.py:313-324 touches"""
    reg = types.SimpleNamespace(
        registration_ransac_based_on_correspondence=_ransac_shim,
        TransformationEstimationPointToPoint=lambda with_scaling=False: _Tagged("p2p", with_scaling),
        CorrespondenceCheckerBasedOnEdgeLength=lambda th=0.9: _Tagged("edge", th),
        CorrespondenceCheckerBasedOnDistance=lambda th: _Tagged("dist", th),
        RANSACConvergenceCriteria=lambda max_iteration=100000, confidence=0.999: _Tagged("crit", max_iteration, confidence))
    util = types.SimpleNamespace(Vector2iVector=lambda a: np.asarray(a), Vector3dVector=lambda a: np.asarray(a))
    return types.SimpleNamespace(pipelines=types.SimpleNamespace(registration=reg), utility=util)


def install(B, replace_open3d=True):
    """patch the reference module in place; returns B"""
    def mutual_matching(self, src_des, tgt_des):
        return backend.mutual_matching(src_des, tgt_des)

    def post_refinement(self, initial_trans, src_keypts, tgt_keypts, weights=None):
        return backend.post_refinement(initial_trans, src_keypts, tgt_keypts, weights, dataset=self.config.data.dataset)

    def get_matching_indices(self, source, target, relt_pose, search_voxel_size):
        return backend.get_matching_indices(source, target, relt_pose, search_voxel_size)

    B.buffer.mutual_matching = mutual_matching
    B.buffer.post_refinement = post_refinement
    B.buffer.get_matching_indices = get_matching_indices
    B.rigid_transform_3d = backend.rigid_transform_3d
    B.KNN = _KNN1
    if replace_open3d:
        B.o3d = open3d_shim()
        B.make_open3d_point_cloud = lambda xyz, color=None: types.SimpleNamespace(points=np.asarray(xyz))
    return B
