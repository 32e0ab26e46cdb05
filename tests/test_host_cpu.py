"""CPU tests of the host-side logic: SE3 mirror vs the reference fixtures, C-ABI exports, synthetic generator,
install shim wiring, and the multi-process (gloo, world_size 2) sharding / split-hypothesis logic."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


# ---- utils/SE3.py mirror --------------------------------------------------------------------------------------
def test_se3_matches_reference_fixture():
    from buffer_b200 import SE3
    g = np.load(os.path.join(G, "se3.npz"))
    T1 = SE3.integrate_trans(torch.from_numpy(g["R1"]), torch.from_numpy(g["t1"]))
    assert T1.dtype == torch.float32 and np.array_equal(T1.numpy(), g["T1"])
    T1n = SE3.integrate_trans(g["R1"].astype(np.float64), g["t1"].astype(np.float64))
    assert isinstance(T1n, np.ndarray) and np.array_equal(T1n, g["T1_np"])
    assert np.allclose(SE3.transform(torch.from_numpy(g["pts"]), T1).numpy(), g["transform_torch"], atol=1e-6)
    assert np.allclose(SE3.transform(g["pts"].astype(np.float64), g["T1"].astype(np.float64)), g["transform_numpy"], atol=1e-12)
    assert np.allclose(SE3.concatenate(T1, torch.from_numpy(g["T2"])).numpy(), g["concat_torch"], atol=1e-6)
    Tb = SE3.integrate_trans(torch.from_numpy(g["Rb"]), torch.from_numpy(g["tb"]))
    assert np.array_equal(Tb.numpy(), g["Tb"])
    assert np.allclose(SE3.transform(torch.from_numpy(g["ptsb"]), Tb).numpy(), g["transform_batched"], atol=1e-6)
    assert np.allclose(SE3.concatenate(Tb, Tb.clone()).numpy(), g["concat_batched"], atol=1e-6)
    R, t = SE3.decompose_trans(Tb)
    assert R.shape == (4, 3, 3) and t.shape == (4, 3, 1)
    R, t = SE3.decompose_trans(g["T1"])
    assert R.shape == (3, 3) and t.shape == (3, 1)
    np.random.seed(7)                                                  # same numpy stream as the fixture
    assert np.allclose(SE3.rotation_matrix(3, 1.0), g["rm3"]) and np.allclose(SE3.rotation_matrix(1, 0.5), g["rm1"])
    assert np.array_equal(SE3.rotation_matrix(0, 1.0), g["rm0"]) and np.allclose(SE3.translation_matrix(0.5), g["tm"])
    with pytest.raises(AssertionError):
        SE3.rotation_matrix(2, 1.0)


def test_se3_star_import_reexports():
    ns = {}
    exec("from buffer_b200.SE3 import *", ns)
    for name in ("torch", "np", "random", "transform", "integrate_trans", "decompose_trans", "concatenate", "rotation_matrix", "translation_matrix"):
        assert name in ns
    # batched numpy works here (raises in the reference)
    Tb = ns["integrate_trans"](np.tile(np.eye(3), (2, 1, 1)), np.ones((2, 3, 1)))
    assert Tb.shape == (2, 4, 4) and np.allclose(ns["transform"](np.zeros((2, 5, 3)), Tb), 1.0)


# ---- C ABI ---------------------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "buffer_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bfr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    from buffer_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.SIGNATURES) == declared                         # the ctypes table mirrors the header one to one
    L = _lib.lib()
    assert L.bfr_version() >= 100
    assert L.bfr_error_string(0) == b"ok" and b"NULL" in L.bfr_error_string(-1)
    # size queries and argument validation run without a GPU
    assert L.bfr_mutual_nn_workspace_bytes(2, 5000, 5000) >= 2 * 2 * 5120 * 12
    assert L.bfr_register_workspace_bytes(2, 100, 100, 200, 200) > 0 and L.bfr_score_workspace_bytes(10) >= 320
    assert L.bfr_mutual_matching_batched(None, None, None, None, 1, 1, 1, 1, 1, 32, 1, None, None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert L.bfr_ransac_batched(None, None, None, 0, 0, 0, 0, 0, 0.1, 0.8, 1.0, 1, None, None, None, 0, None) == 0     # P == 0 is a no-op
    assert L.bfr_rigid_transform_3d(None, None, None, 3, 3, 0.0, None, None) == -1
    # RANSAC scratch for the tensor-core scoring filter: one 4 KB A tile per 128 correspondences, 40 tiles per resident pair, per CTA
    assert L.bfr_ransac_workspace_bytes() >= 148 * 40 * 4096
    assert L.bfr_register_workspace_bytes(2, 100, 100, 200, 200) >= L.bfr_ransac_workspace_bytes()
    assert L.bfr_config_get(2) == 1 and L.bfr_config_set(2, 0) == 0 and L.bfr_config_get(2) == 0 and L.bfr_config_set(2, 1) == 0
    assert L.bfr_config_set(2, 7) == -2 and L.bfr_config_set(99, 0) == -2
    # more than BFR_MAX_PAIRS pairs in one call is an argument error (BFR_E_SIZE), reported before anything touches the device
    buf = ctypes.create_string_buffer(64)
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    assert L.bfr_mutual_matching_batched(ptr, ptr, ptr, ptr, 70000, 1, 1, 1, 1, 32, 1, None, None, None, None, None, None, None, None, None, None, ptr, 1 << 40, None) == -2
    assert L.bfr_ransac_batched(ptr, ptr, ptr, 70000, 0, 0, 0, 10, 0.1, 0.8, 1.0, 1, ptr, None, None, 0, None) == -2
    # int32 row offsets: P * M beyond INT32_MAX is refused up front (no copy is queued), the size query reports 0
    assert L.bfr_register_host_workspace_bytes(60000, 40000, 40000, 32) == 0
    assert L.bfr_register_uniform_host(ptr, ptr, ptr, ptr, 60000, 40000, 40000, 32, 10, 0, 0, 0.1, 0.8, 1.0, 0.1, 20, 1, ptr, None, None, ptr, 1 << 40, None) == -2
    assert L.bfr_vote_workspace_bytes(3, 100) >= 100 * 32 and L.bfr_lrf_vote_batched(None, None, None, 1, 1, 1, None, None, None, 20.0, 0.3, None, None, None, None, None, None, 0, None) == -1


def test_product_never_imports_oracle():
    """the product package must not reference oracle/ (only tests, smoke and bench's cpu_baseline may)"""
    pkg = os.path.join(ROOT, "buffer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "libbfr_oracle" not in txt and "bfr_oracle.c\"" not in txt, f


# ---- synthetic generator ---------------------------------------------------------------------------------------
def test_synthetic_pairs_are_deterministic_and_well_formed():
    from buffer_b200 import synthetic as S
    a = S.make_pairs(3, 200, cfg_id=2, first_pair=5); b = S.make_pairs(3, 200, cfg_id=2, first_pair=5)
    assert torch.equal(a.src_des, b.src_des) and torch.equal(a.tgt_xyz, b.tgt_xyz)
    assert torch.allclose(a.src_des.norm(dim=-1), torch.ones(3, 200), atol=1e-5)
    assert int(a.inlier.sum(-1)[0]) == 60                               # (1 - 0.7) * 200
    moved = a.src_xyz @ a.T_gt[:, :3, :3].transpose(-1, -2) + a.T_gt[:, None, :3, 3]
    err = (torch.gather(a.tgt_xyz, 1, a.perm[:, :, None].expand(-1, -1, 3)) - moved).norm(dim=-1)
    assert float(err[a.inlier].max()) < 0.06 and float(err[~a.inlier].median()) > 0.5
    lo = S.make_pairs(4, 200, cfg_id=3, outlier_ratio=0.9, outlier_ratio_hi=0.97)
    assert 5 <= int(lo.inlier.sum(-1).min()) and int(lo.inlier.sum(-1).max()) <= 21
    rec, rte, rre = S.registration_recall(a.T_gt, a.T_gt)
    assert rec == 1.0 and float(rre.max()) < 1e-3


# ---- install shim wiring (no kernels are launched) ----------------------------------------------------------------
def test_install_rebinds_reference_symbols():
    import types
    from buffer_b200 import install, backend
    B = types.SimpleNamespace(buffer=type("buffer", (), {}), rigid_transform_3d=None, KNN=None, o3d=None)
    install.install(B)
    assert B.rigid_transform_3d is backend.rigid_transform_3d
    reg = B.o3d.pipelines.registration
    chk = [reg.CorrespondenceCheckerBasedOnEdgeLength(0.8), reg.CorrespondenceCheckerBasedOnDistance(0.1)]
    crit = reg.RANSACConvergenceCriteria(50000, 0.999)
    assert chk[0].kind == "edge" and chk[0].args == (0.8,) and crit.args == (50000, 0.999)
    assert callable(B.buffer.mutual_matching) and callable(B.buffer.post_refinement) and callable(reg.registration_ransac_based_on_correspondence)
    with pytest.raises(NotImplementedError):
        B.KNN(k=2)
    with pytest.raises(RuntimeError):                                   # CPU tensors: there is no CPU path
        backend.mutual_matching(torch.zeros(4, 32), torch.zeros(4, 32))


# ---- multi-process host logic on gloo ----------------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from buffer_b200 import dist as D, synthetic as S
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    P, N, H = 5, 300, 1500
    b = S.make_pairs(P, N, cfg_id=31)

    def register_fn(sd, sx, td, tx, pair_id_base=0, **kw):           # CPU stand-in for backend.register_uniform
        n = sd.shape[0]
        off = np.arange(n + 1, dtype=np.int32) * N
        T, nm, ni = O.register_batched(sd.reshape(-1, 32).numpy(), sx.reshape(-1, 3).numpy(), off, td.reshape(-1, 32).numpy(), tx.reshape(-1, 3).numpy(),
                                       off, H, 5, pair_id_base, 0.1, 0.8, 0.1, 20)
        return torch.from_numpy(T), torch.from_numpy(nm), torch.from_numpy(ni)

    T, nm, ni = D.register_sharded(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, register_fn=register_fn)
    # split-hypothesis mode on pair 0, replicated on both ranks
    s, t = O.mutual_matching(b.src_des[0].numpy(), b.tgt_des[0].numpy())
    corr = O.gather_corr(b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy(), s, t)

    def ransac_fn(c, off, cnt, Hh, d, sim, seed=0, pair_id_base=0, h_begin=0, h_end=None):
        return torch.tensor([O.ransac(c, seed, pair_id_base, Hh, d, sim, h_begin, h_end)], dtype=torch.int64)

    def finalize_fn(c, off, cnt, best, d, sim, seed=0, pair_id_base=0):
        T_, cnt_, bh = O.ransac_finalize(c, seed, pair_id_base, int(best.item()), d, sim)
        return torch.from_numpy(T_)[None], torch.tensor([cnt_]), torch.tensor([bh])

    T2, c2, bh2 = D.ransac_split_hypotheses(corr, None, None, 4000, 0.1, 0.8, seed=9, pair_id_base=3, ransac_fn=ransac_fn, finalize_fn=finalize_fn)
    torch.save({"T": T, "nm": nm, "ni": ni, "T2": T2, "c2": c2, "bh2": bh2}, os.path.join(tmp, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_split_hypotheses_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    from buffer_b200 import dist as D, synthetic as S
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [D.shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt")); r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    for k in ("T", "nm", "ni", "T2", "c2", "bh2"):
        assert torch.equal(r0[k], r1[k]), k                            # every rank ends with the same full result
    P, N, H = 5, 300, 1500
    b = S.make_pairs(P, N, cfg_id=31)
    off = np.arange(P + 1, dtype=np.int32) * N
    T, nm, ni = oracle.register_batched(b.src_des.reshape(-1, 32).numpy(), b.src_xyz.reshape(-1, 3).numpy(), off, b.tgt_des.reshape(-1, 32).numpy(),
                                        b.tgt_xyz.reshape(-1, 3).numpy(), off, H, 5, 0, 0.1, 0.8, 0.1, 20)
    assert np.array_equal(r0["T"].numpy(), T) and np.array_equal(r0["nm"].numpy(), nm) and np.array_equal(r0["ni"].numpy(), ni)
    s, t = oracle.mutual_matching(b.src_des[0].numpy(), b.tgt_des[0].numpy())
    corr = oracle.gather_corr(b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy(), s, t)
    best = oracle.ransac(corr, 9, 3, 4000, 0.1, 0.8)
    T1, c1, bh1 = oracle.ransac_finalize(corr, 9, 3, best, 0.1, 0.8)
    assert np.array_equal(r0["T2"][0].numpy(), T1) and int(r0["c2"]) == c1 and int(r0["bh2"]) == bh1
