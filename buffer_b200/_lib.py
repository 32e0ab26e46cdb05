"""ctypes loader of libbuffer_b200.so (the C ABI of include/buffer_b200.h).

The product path has NO fallback: if the CUDA library is missing or fails to load, importing an op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libbuffer_b200.so")

_vp, _i, _u32, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/buffer_b200.h one to one
SIGNATURES = {
    "bfr_version": (_i, []),
    "bfr_error_string": (C.c_char_p, [_i]),
    "bfr_config_set": (_i, [_i, _i]),
    "bfr_config_get": (_i, [_i]),
    "bfr_mutual_nn_workspace_bytes": (_sz, [_i, _i, _i]),
    "bfr_mutual_matching_batched": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_mutual_nn_partial": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "bfr_mutual_nn_packed": (_i, [_vp, _sz, _i, _i, _i, _vp, _vp]),
    "bfr_mutual_select": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_gather_corr": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "bfr_ransac_workspace_bytes": (_sz, []),
    "bfr_ransac_batched": (_i, [_vp, _vp, _vp, _i, _u64, _u32, _u32, _u32, _f, _f, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "bfr_ransac_finalize_batched": (_i, [_vp, _vp, _vp, _i, _u64, _u32, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "bfr_lrf_hypotheses": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "bfr_score_workspace_bytes": (_sz, [_i]),
    "bfr_score_hypotheses": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_vote_workspace_bytes": (_sz, [_i, _i]),
    "bfr_lrf_vote_batched": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_pose_from_votes_batched": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _i, _u64, _u32, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_rigid_transform_3d": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "bfr_post_refinement_batched": (_i, [_vp, _vp, _vp, _vp, _i, _f, _i, _i, _vp, _vp, _vp, _vp]),
    "bfr_register_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "bfr_register_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _u64, _u32, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_register_host_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "bfr_register_uniform_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u64, _u32, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_register_uniform_host_chunked": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _u64, _u32, _f, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp, _i]),
    "bfr_get_matching_indices_workspace_bytes": (_sz, [_i]),
    "bfr_get_matching_indices": (_i, [_vp, _i, _vp, _i, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bfr_svd3_batched": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "bfr_furthest_point_sample": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "bfr_fp32_probe": (_i, [_i, _i, _vp, _vp]),
    "bfr_debug_set_k1_events": (_i, [_vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "buffer_b200: %s is missing - build it with `python -m buffer_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % SO_PATH)
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError here = ABI mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().bfr_error_string(rc).decode()
        raise RuntimeError("buffer_b200 %s failed: %s (code %d)" % (what, msg, rc))
