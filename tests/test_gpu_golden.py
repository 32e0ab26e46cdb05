"""GPU vs the REFERENCE directly: the CUDA path (through the C ABI) on the inputs of every tests/golden/*.npz fixture against the outputs the
unmodified reference produced for them (oracle/gen_golden.py), at the tolerances of tests/test_oracle_golden.py - no oracle in between."""
import os

import numpy as np
import pytest
import torch

from buffer_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(DEV)


def rot_err(Ra, Rb):
    return float(S.rotation_error_rad(torch.from_numpy(np.asarray(Ra, np.float64)), torch.from_numpy(np.asarray(Rb, np.float64))))


def close_pose(T, Tr, tol=1e-5):
    return rot_err(T[:3, :3], Tr[:3, :3]) < tol and float(np.abs(np.asarray(T)[:3, 3] - Tr[:3, 3]).max()) < tol


@pytest.mark.parametrize("algo", [0, 1])
def test_mutual_matching_fixture(backend, algo):
    g = load("mutual_matching")
    backend.set_k1_algo(algo)
    try:
        s, t = backend.mutual_matching(dev(g["src_des"]), dev(g["tgt_des"]))
    finally:
        backend.set_k1_algo(1)
    assert s.dtype == np.int64 and np.array_equal(s, g["s_mids"]) and np.array_equal(t, g["t_mids"])


@pytest.mark.parametrize("case", ["n3", "n3_noise", "n200_w", "n200_wthr", "n2000", "mirror"])
def test_rigid_transform_3d_fixture(backend, case):
    g = load("rigid_transform_3d")
    w = dev(g[case + "_w"].copy()) if case + "_w" in g.files else None
    T = backend.rigid_transform_3d(dev(g[case + "_A"]), dev(g[case + "_B"]), w, float(g[case + "_thr"])).cpu().numpy()
    for b in range(T.shape[0]):
        assert close_pose(T[b], g[case + "_T"][b])
        assert abs(np.linalg.det(T[b, :3, :3].astype(np.float64)) - 1) < 1e-5


def test_post_refinement_fixture(backend):
    g = load("post_refinement")
    for key, ds in (("T_3dmatch", "3DMatch"), ("T_kitti", "KITTI")):
        T = backend.post_refinement(dev(g["T0"])[None], dev(g["src"])[None], dev(g["tgt"])[None], dataset=ds)[0].cpu().numpy()
        assert close_pose(T, g[key])
    T = backend.post_refinement(dev(g["T0_far"])[None], dev(g["src"])[None], dev(g["tgt"])[None])[0].cpu().numpy()
    assert np.array_equal(T, g["T_far"])                               # no inliers: returned unchanged (models/BUFFER.py:406-407)


def test_lrf_scoring_fixture(backend):
    g = load("lrf_scoring")
    R, t = backend.lrf_hypotheses(dev(g["ind"]), dev(g["ss_R"]), dev(g["tt_R"]), dev(g["ss"]), dev(g["tt"]))
    assert float(np.abs(R.cpu().numpy() - g["R"]).max()) < 5e-6 and float(np.abs(t.cpu().numpy() - g["t"]).max()) < 2e-5
    # a4 on the REFERENCE's hypotheses: counts, argmax and inlier set equal the reference's exactly (sqrt(d2) < thr)
    counts, best, mask = backend.score_hypotheses(dev(g["R"]), dev(g["t"]), dev(g["ss"]), dev(g["tt"]), dev(g["thr"]))
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), g["inlier_num"])
    assert int(best.item()) == int(g["best_ind"]) and np.array_equal(np.nonzero(mask.cpu().numpy())[0], g["inlier_ind"])
    # the fused vote (angle, cos / sin and thresholds computed on the device): same winner and inlier set
    num, bi, sel, n = backend.lrf_vote(dev(g["ind"]), dev(g["ss_R"]), dev(g["tt_R"]), dev(g["ss"]), dev(g["tt"]))
    assert int(bi.item()) == int(g["best_ind"]) and int(n.item()) == len(g["inlier_ind"])
    assert np.array_equal(sel[: int(n.item())].cpu().numpy(), g["inlier_ind"])
    d = num.cpu().numpy().astype(np.int64) - g["inlier_num"]
    assert np.abs(d).max() <= 1 and np.mean(d != 0) < 0.01


def test_ransac_fixture(backend):
    g = load("ransac")
    K = len(g["ss"]); H = int(g["H"]); seed = int(g["seed"]); pid = int(g["pair_id"])
    corr = np.stack([np.arange(K), np.arange(K)], 1)
    res = backend.registration_ransac_based_on_correspondence(dev(g["ss"]), dev(g["tt"]), corr, float(g["dist_th"]), float(g["similar_th"]),
                                                              iter_n=H, confidence=1.0, seed=seed, pair_id=pid)
    assert res.best_hypothesis == int(g["best_h"]) and abs(res.inlier_count - int(g["best_count"])) <= 1
    assert close_pose(res.transformation, g["T_best"])
    rec = backend._records(dev(g["ss"]), dev(g["tt"]))
    off = torch.tensor([0, K], dtype=torch.int32, device=DEV); cnt = torch.tensor([K], dtype=torch.int32, device=DEV)
    nv = torch.zeros(1, dtype=torch.int32, device=DEV)
    backend.ransac_batched(rec, off, cnt, H, float(g["dist_th"]), float(g["similar_th"]), seed=seed, pair_id_base=pid, valid_count=nv)
    assert abs(int(nv.item()) - int((g["counts"] >= 0).sum())) <= max(1, H // 100)      # checker decisions agree (borderline fits may flip)


def test_next_rows_fixture(backend):
    g = load("next_rows")
    m = backend.get_matching_indices(dev(g["gm_src"]), dev(g["gm_tgt"]), dev(g["gm_T"]), float(g["gm_voxel"]))
    assert np.array_equal(m.cpu().numpy(), g["gm_pairs"])
    u, s, v = backend.svd(dev(g["cov"]))
    assert float(np.abs(s.cpu().numpy() - g["svd_s"]).max()) < 2e-4 * float(g["svd_s"].max())
    z = u[:, :, -1].cpu().numpy()
    z = np.where((np.sum(-z * g["ref_point"], axis=1) < 0)[:, None], -z, z)
    assert float(np.abs(z - g["z_axis"]).max()) < 1e-4


def test_config1_full_size_fixture(backend):
    """BASELINE config 1 (one 5 000 x 5 000 pair) against the reference's own outputs at full size: matching exact, the RANSAC winner of
    the shared Philox stream, the refined pose within 1e-5, the vote's winner and inlier set exact, and the reference's whole stage flow
    (vote -> RANSAC on the subset -> refinement on all matches) in one call."""
    from test_oracle_golden import config1_inputs
    g, b, ind, ss_R, tt_R = config1_inputs()
    bd = b.to(DEV)
    r = backend.mutual_matching_device(bd.src_des[0], bd.tgt_des[0], bd.src_xyz[0], bd.tgt_xyz[0])
    n = int(r["n_mutual"].item())
    assert np.array_equal(r["s_mids"][:n].cpu().numpy(), g["s_mids"]) and np.array_equal(r["t_mids"][:n].cpu().numpy(), g["t_mids"])
    H, seed, pid = int(g["H"]), int(g["seed"]), int(g["pair_id"])
    off = torch.tensor([0, n], dtype=torch.int32, device=DEV); cnt = r["n_mutual"]
    nv = torch.zeros(1, dtype=torch.int32, device=DEV)
    best = backend.ransac_batched(r["corr"], off, cnt, H, 0.1, 0.8, seed=seed, pair_id_base=pid, valid_count=nv)
    T, inl, bh = backend.ransac_finalize_batched(r["corr"], off, cnt, best, 0.1, 0.8, seed=seed, pair_id_base=pid)
    assert int(bh.item()) == int(g["best_h"]) and abs(int(inl.item()) - int(g["best_count"])) <= 1
    assert abs(int(nv.item()) - int((g["counts"] >= 0).sum())) <= 10
    assert close_pose(T[0].cpu().numpy(), g["T_best"])
    Tr, _, _ = backend.post_refinement_batched(T, r["corr"], off, cnt, 0.10)
    assert close_pose(Tr[0].cpu().numpy(), g["T_refined"])
    ok, rte, rre = S.registration_recall(Tr.cpu(), b.T_gt)
    assert ok == 1.0
    v = backend.lrf_vote_batched(r["corr"].clone(), off, cnt, ind.to(DEV), ss_R.to(DEV), tt_R.to(DEV), n)
    k = int(v["sub_cnt"].item())
    assert int(v["best_ind"].item()) == int(g["vote_best_ind"]) and np.array_equal(v["inlier_ind"][:k].cpu().numpy(), g["vote_inlier_ind"])
    d = v["inlier_num"][:n].cpu().numpy().astype(np.int64) - g["vote_inlier_num"]
    assert np.abs(d).max() <= 1 and np.mean(d != 0) < 0.01
    T2, nvote, ni = backend.pose_from_votes_batched(r["corr"].clone(), off, cnt, ind.to(DEV), ss_R.to(DEV), tt_R.to(DEV), n, hypotheses=H, seed=seed, pair_id_base=pid)
    assert int(nvote.item()) == len(g["vote_inlier_ind"]) and abs(int(ni.item()) - int(g["sub_best_count"])) <= 1
    assert close_pose(T2[0].cpu().numpy(), g["sub_T_refined"])
