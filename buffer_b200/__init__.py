"""buffer_b200 — B200-native (sm_100a) correspondence-and-pose back end for BUFFER.

Scope: mutual-NN descriptor matching, Philox RANSAC with closed-form 3-point Kabsch, SE(3) inlier
scoring, weighted-Kabsch post-refinement and the SE3 helpers — the stage of the reference's ``buffer.forward`` after
keypoints and descriptors exist.  Python/PyTorch host code over hand-written CUDA behind a C ABI; no CPU fallback.
"""
from . import SE3, synthetic  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):          # backend (needs the built .so) is imported lazily
    if name in ("backend", "install", "dist", "evaluation"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
