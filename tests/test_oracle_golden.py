"""CPU tests (no GPU): the oracle (oracle/bfr_oracle.c) against the golden fixtures generated from the UNMODIFIED reference
(tests/golden/*.npz, oracle/gen_golden.py) and against the Philox known-answer vectors.  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from buffer_b200 import synthetic as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def rot_err(Ra, Rb):
    return float(S.rotation_error_rad(torch.from_numpy(np.asarray(Ra, np.float64)), torch.from_numpy(np.asarray(Rb, np.float64))))


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors: philox4x32 10 rounds
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kats:
        assert [int(x) for x in oracle.philox4x32_10(ctr, key)] == out


def test_sample3_is_mulhi_of_philox(oracle):
    for h, K in ((0, 5000), (123456, 1501), (49999, 3)):
        r = oracle.philox4x32_10([h, 77, 0, 0], [0xDEADBEEF, 0x12345678])
        want = [(int(x) * K) >> 32 for x in r[:3]]
        assert [int(x) for x in oracle.sample3(0x12345678DEADBEEF, 77, h, K)] == want


def test_mutual_matching_equals_reference(oracle):
    g = load("mutual_matching")
    s, t = oracle.mutual_matching(g["src_des"], g["tgt_des"])
    assert s.dtype == np.int64 and np.array_equal(s, g["s_mids"]) and np.array_equal(t, g["t_mids"])
    assert np.all(np.diff(s) > 0)


@pytest.mark.parametrize("case", ["n3", "n3_noise", "n200_w", "n200_wthr", "n2000", "mirror"])
def test_rigid_transform_3d_equals_reference(oracle, case):
    g = load("rigid_transform_3d")
    w = g[case + "_w"].copy() if case + "_w" in g.files else None
    T = oracle.rigid_transform_3d(g[case + "_A"], g[case + "_B"], w, float(g[case + "_thr"]))
    Tr = g[case + "_T"]
    assert T.shape == Tr.shape
    for b in range(T.shape[0]):
        assert rot_err(T[b, :3, :3], Tr[b, :3, :3]) < 1e-5            # north_star tolerance: 1e-5 rad
        assert np.abs(T[b, :3, 3] - Tr[b, :3, 3]).max() < 1e-5        # 1e-5 m
        assert abs(np.linalg.det(T[b, :3, :3].astype(np.float64)) - 1) < 1e-5
    assert np.array_equal(T[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (T.shape[0], 1)))


def test_post_refinement_equals_reference(oracle):
    g = load("post_refinement")
    corr = np.zeros((len(g["src"]), 8), np.float32); corr[:, :3] = g["src"]; corr[:, 4:7] = g["tgt"]
    for key, thr in (("T_3dmatch", 0.10), ("T_kitti", 1.2)):
        T, it, inl = oracle.post_refinement(g["T0"], corr, thr, 20)
        assert rot_err(T[:3, :3], g[key][:3, :3]) < 1e-5 and np.abs(T[:3, 3] - g[key][:3, 3]).max() < 1e-5
        assert 1 <= it <= 20 and inl > 0
    T, it, inl = oracle.post_refinement(g["T0_far"], corr, 0.10, 20)
    assert np.array_equal(T, g["T_far"]) and it == 0 and inl == 0     # no inliers: returned unchanged (models/BUFFER.py:406-407)


def test_lrf_hypotheses_and_scoring_equal_reference(oracle):
    g = load("lrf_scoring")
    ang = g["ind"].astype(np.float64) * 2 * np.pi / 20 + 1e-6
    cs = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)
    R, t = oracle.lrf_hypotheses(cs, g["ss_R"], g["tt_R"], g["ss"], g["tt"])
    assert np.abs(R - g["R"]).max() < 5e-6 and np.abs(t - g["t"]).max() < 2e-5
    # scoring of the REFERENCE's hypotheses so that only the scoring arithmetic is compared
    counts, best, mask = oracle.score_hypotheses(g["R"], g["t"], g["ss"], g["tt"], g["thr"])
    assert best == int(g["best_ind"])
    assert np.array_equal(np.nonzero(mask)[0], g["inlier_ind"])
    diff = counts.astype(np.int64) - g["inlier_num"]
    print("a4 counts differing from the reference: %d of %d" % (int((diff != 0).sum()), len(diff)))
    assert np.abs(diff).max() == 0                                     # same test as the reference: sqrt(d2) < thr (:305-308)
    # the fused vote (angle, cos / sin, thresholds all computed inside): same winner, same inlier set, counts equal
    corr = np.zeros((len(g["ss"]), 8), np.float32); corr[:, :3] = g["ss"]; corr[:, 4:7] = g["tt"]
    vcounts, vbest, vsel = oracle.lrf_vote(corr, g["ind"], g["ss_R"], g["tt_R"], 20.0, float(g["inlier_th"]) if "inlier_th" in g.files else 1.0 / 3.0)
    assert vbest == int(g["best_ind"]) and np.array_equal(vsel, g["inlier_ind"])
    vdiff = vcounts.astype(np.int64) - g["inlier_num"]
    assert np.abs(vdiff).max() <= 1 and np.mean(vdiff != 0) < 0.01    # polynomial cos/sin vs torch's: borderline residuals only


def test_pinned_helpers(oracle):
    """sqrt_threshold <=> the rooted compare; polynomial sin/cos and log against libm; Open3D's iteration bound"""
    rng = np.random.RandomState(5)
    for thr in list(rng.uniform(1e-3, 3.0, 200).astype(np.float32)) + [np.float32(0.1), np.float32(0.0524), np.float32(1e-20), np.float32(3e19)]:
        T = np.float32(oracle.sqrt_threshold(float(thr)))
        x = T
        for _ in range(3):
            x = np.nextafter(x, np.float32(0))
        for _ in range(7):                                             # floats around T: d2 < T  <=>  sqrt(d2) < thr
            assert (x < T) == (np.sqrt(x, dtype=np.float32) < thr), (thr, x)
            x = np.nextafter(x, np.float32(np.inf))
    assert oracle.sqrt_threshold(0.0) == 0.0 and oracle.sqrt_threshold(-1.0) == 0.0
    for a in np.concatenate([np.arange(0, 21) * np.float32(2 * np.pi / 20) + np.float32(1e-6), rng.uniform(0, 50, 300)]).astype(np.float32):
        sn, cs = oracle.det_sincos(float(a))
        assert abs(sn - np.sin(np.float64(a))) < 4e-7 and abs(cs - np.cos(np.float64(a))) < 4e-7
    for v in list(rng.uniform(1e-9, 1.0, 300)) + [1.0 - 1e-12, 0.001, 0.5, 2.0, 1e300]:
        assert abs(oracle.det_log(v) - np.log(v)) <= 4e-16 * max(1.0, abs(np.log(v)))
    for count, K in ((30, 100), (300, 5000), (1500, 5000), (5000, 5000), (1, 5000), (0, 10)):
        b = oracle.ransac_exit_bound(count, K, 0.999, 50000)
        if count == 0:
            assert b == 50000
        elif count == K:
            assert b == 0
        else:
            ref = np.ceil(np.log(1 - np.float64(np.float32(0.999))) / np.log(1 - (count / K) ** 3))
            assert b == min(50000, int(ref))


def test_ransac_confidence_is_a_prefix_of_the_full_run(oracle):
    """Open3D's early exit evaluates a prefix of the hypothesis sequence: its winner equals the full run restricted to that prefix"""
    b = S.make_pairs(2, 800, cfg_id=61)
    for p in range(2):
        s, t = oracle.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        corr = oracle.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)
        best_c, iters = oracle.ransac_confidence(corr, 3, p, 20000, 0.1, 0.8, 0.999)
        assert 0 < iters < 20000                                       # 30 % inliers: stops after a few hundred iterations
        assert best_c == oracle.ransac(corr, 3, p, 20000, 0.1, 0.8, 0, iters)
        cnt = best_c >> 32
        assert iters >= oracle.ransac_exit_bound(cnt, len(s), 0.999, 20000) or iters == 20000
        best_1, it1 = oracle.ransac_confidence(corr, 3, p, 2000, 0.1, 0.8, 1.0)
        assert it1 == 2000 and best_1 == oracle.ransac(corr, 3, p, 2000, 0.1, 0.8)


def test_ransac_equals_reference_semantics(oracle):
    g = load("ransac")
    K = len(g["ss"]); H = int(g["H"]); seed = int(g["seed"]); pid = int(g["pair_id"])
    corr = np.zeros((K, 8), np.float32); corr[:, :3] = g["ss"]; corr[:, 4:7] = g["tt"]
    for h in range(0, H, 37):                                          # shared Philox stream -> identical minimal sets
        assert [int(x) for x in oracle.sample3(seed, pid, h, K)] == [int(x) for x in g["samples"][h]]
    best, counts = oracle.ransac(corr, seed, pid, H, float(g["dist_th"]), float(g["similar_th"]), want_counts=True)
    ref = g["counts"]
    valid_o, valid_r = counts >= 0, ref >= 0
    assert np.mean(valid_o != valid_r) < 0.01                          # checker decisions agree (borderline fits may flip)
    both = valid_o & valid_r
    assert both.sum() > 5 and np.abs(counts[both] - ref[both]).max() <= 1
    T, cnt, bh = oracle.ransac_finalize(corr, seed, pid, best, float(g["dist_th"]), float(g["similar_th"]))
    assert bh == int(g["best_h"]) and abs(cnt - int(g["best_count"])) <= 1
    assert rot_err(T[:3, :3], g["T_best"][:3, :3]) < 1e-5 and np.abs(T[:3, 3] - g["T_best"][:3, 3]).max() < 1e-5
    ok, rte, rre = S.registration_recall(torch.from_numpy(T)[None], torch.from_numpy(g["T_gt"])[None])
    assert ok == 1.0


def test_next_rows_equal_reference(oracle):
    g = load("next_rows")
    pairs = oracle.get_matching_indices(g["gm_src"], g["gm_tgt"], g["gm_T"], float(g["gm_voxel"]))
    assert pairs.dtype == np.int64 and np.array_equal(pairs, g["gm_pairs"])          # ThreeDMatch/dataset.py:14-22
    u, s, v = oracle.svd3(g["cov"])                                                   # torch_batch_svd contract, utils/common.py:715
    assert np.abs(s - g["svd_s"]).max() < 2e-4 * g["svd_s"].max() and np.all(np.diff(s, axis=1) <= 0)
    rec = np.einsum("bij,bj,bkj->bik", u, s, v)
    assert np.abs(rec - g["cov"]).max() < 1e-4 * np.abs(g["cov"]).max()
    for a, b_ in ((u, g["svd_u"]), (v, g["svd_v"])):                                  # singular vectors agree up to sign
        assert np.abs(np.abs(np.einsum("bik,bik->bk", a, b_)) - 1).max() < 1e-4
    # cal_Z_axis = u[:, :, -1] with the reference's sign disambiguation
    z = u[:, :, -1]
    flip = (np.sum(-z * g["ref_point"], axis=1) < 0)[:, None]
    z = np.where(flip, -z, z)
    assert np.abs(z - g["z_axis"]).max() < 1e-4


def test_oracle_properties(oracle):
    """SE(3) recovery from noiseless inliers, invariance to outlier positions, SO(3) membership, K < 3 -> identity"""
    b = S.make_pairs(2, 600, cfg_id=55, sigma=0.0)
    for p in range(2):
        s, t = oracle.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        assert np.array_equal(t, b.perm[p].numpy()[s]) and len(s) == 600
        corr = oracle.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)
        best = oracle.ransac(corr, 11, p, 3000, 0.1, 0.8)
        T, cnt, bh = oracle.ransac_finalize(corr, 11, p, best, 0.1, 0.8)
        assert cnt == int(b.inlier[p].sum()) and bh >= 0
        assert rot_err(T[:3, :3], b.T_gt[p, :3, :3].numpy()) < 1e-5 and np.abs(T[:3, 3] - b.T_gt[p, :3, 3].numpy()).max() < 1e-5
        R = T[:3, :3].astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6
        # move the outliers somewhere else: same winner, same pose
        corr2 = corr.copy(); out = ~b.inlier[p].numpy()[s]; corr2[out, 4:7] += 7.0
        best2 = oracle.ransac(corr2, 11, p, 3000, 0.1, 0.8)
        assert best2 == best
    for K in (0, 1, 2):
        c = np.random.RandomState(K).randn(max(K, 1), 8).astype(np.float32)[:K]
        assert oracle.ransac(c.reshape(K, 8), 1, 0, 100, 0.1, 0.8) == 0
        T, cnt, bh = oracle.ransac_finalize(c.reshape(K, 8), 1, 0, 0, 0.1, 0.8)
        assert np.array_equal(T, np.eye(4, dtype=np.float32)) and cnt == 0 and bh == -1


def test_oracle_kabsch_reflection_and_degenerate(oracle):
    rng = np.random.RandomState(3)
    for _ in range(50):
        A = rng.randn(30, 3); q = rng.randn(4); q /= np.linalg.norm(q)
        Rg = S.quat_to_rot(torch.from_numpy(q)[None].float())[0].numpy().astype(np.float64)
        Bm = A @ Rg.T; Bm[:, 0] *= -1                                   # improper: needs the reflection fix
        H = ((A - A.mean(0)).T @ (Bm - Bm.mean(0))).astype(np.float32)
        R, ok = oracle.kabsch_rotation(H)
        U, Sv, Vt = np.linalg.svd(H.astype(np.float64)); V = Vt.T
        Rr = V @ np.diag([1, 1, np.linalg.det(V @ U.T)]) @ U.T
        assert ok and abs(np.linalg.det(R.astype(np.float64)) - 1) < 1e-5 and rot_err(R, Rr) < 1e-4
    R, ok = oracle.kabsch_rotation(np.outer([1, 2, 3], [3, 2, 1]).astype(np.float32))   # rank 1: rejected
    assert not ok and np.array_equal(R, np.eye(3, dtype=np.float32))
    R, ok = oracle.kabsch_rotation(np.zeros((3, 3), np.float32))
    assert not ok


def config1_inputs():
    """regenerate the inputs of tests/golden/config1.npz from the seed and verify them against the stored checksums"""
    g = load("config1")
    c = S.CONFIGS[1]
    b = S.make_pairs(1, first_pair=0, **c["gen"])
    s_mids, t_mids = g["s_mids"].astype(np.int64), g["t_mids"].astype(np.int64)
    inl = b.inlier[0].numpy()[s_mids]
    ind, ss_R, tt_R = S.make_lrf_votes(b.T_gt[0, :3, :3], torch.from_numpy(inl), azi_n=20, seed=1)
    chk = np.array([np.float64(x.numpy().astype(np.float64).sum()) for x in (b.src_des[0], b.tgt_des[0], b.src_xyz[0], b.tgt_xyz[0], ss_R, tt_R, ind)])
    assert np.allclose(chk, g["checksums"], rtol=0, atol=1e-9), "synthetic generator drifted: regenerate tests/golden/config1.npz"
    return g, b, ind, ss_R, tt_R


def test_config1_full_size_equals_reference(oracle):
    """BASELINE config 1 (5 000 x 5 000 keypoints): oracle vs the reference's own functions at full size"""
    g, b, ind, ss_R, tt_R = config1_inputs()
    s, t = oracle.mutual_matching(b.src_des[0].numpy(), b.tgt_des[0].numpy())
    assert np.array_equal(s, g["s_mids"]) and np.array_equal(t, g["t_mids"])
    corr = oracle.gather_corr(b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy(), s, t)
    H, seed, pid = int(g["H"]), int(g["seed"]), int(g["pair_id"])
    best, counts = oracle.ransac(corr, seed, pid, H, float(g["dist_th"]), float(g["similar_th"]), want_counts=True)
    ref = g["counts"].astype(np.int64)
    assert np.mean((counts >= 0) != (ref >= 0)) < 0.002
    both = (counts >= 0) & (ref >= 0)
    assert both.sum() > 100 and np.abs(counts[both] - ref[both]).max() <= 2
    T, cnt, bh = oracle.ransac_finalize(corr, seed, pid, best, float(g["dist_th"]), float(g["similar_th"]))
    assert bh == int(g["best_h"]) and abs(cnt - int(g["best_count"])) <= 1
    assert rot_err(T[:3, :3], g["T_best"][:3, :3]) < 1e-5 and np.abs(T[:3, 3] - g["T_best"][:3, 3]).max() < 1e-5
    Tr, it, inl = oracle.post_refinement(T, corr, 0.10, 20)
    assert rot_err(Tr[:3, :3], g["T_refined"][:3, :3]) < 1e-5 and np.abs(Tr[:3, 3] - g["T_refined"][:3, 3]).max() < 1e-5
    # the vote at full size and the reference's flow on the voted subset
    vcounts, vbest, vsel = oracle.lrf_vote(corr, ind.numpy(), ss_R.numpy(), tt_R.numpy())
    assert vbest == int(g["vote_best_ind"]) and np.array_equal(vsel, g["vote_inlier_ind"])
    vd = vcounts.astype(np.int64) - g["vote_inlier_num"]
    assert np.abs(vd).max() <= 1 and np.mean(vd != 0) < 0.01
    off = np.array([0], np.int32); cnt_ = np.array([len(s)], np.int32)
    T2, nv, ni = oracle.pose_from_votes_batched(corr, off, cnt_, ind.numpy(), ss_R.numpy(), tt_R.numpy(), H, seed, pid, 0.1, 0.8, 1.0, 0.1, 20)
    assert nv[0] == len(g["vote_inlier_ind"]) and abs(int(ni[0]) - int(g["sub_best_count"])) <= 1
    assert rot_err(T2[0, :3, :3], g["sub_T_refined"][:3, :3]) < 1e-5 and np.abs(T2[0, :3, 3] - g["sub_T_refined"][:3, 3]).max() < 1e-5


def test_torch_restatement_equals_reference():
    """oracle/torch_ref.py (the reference's CPU torch path restated for timing in bench.py) against the reference's own outputs"""
    from oracle import torch_ref as TR
    g = load("mutual_matching")
    s, t = TR.mutual_matching(torch.from_numpy(g["src_des"]), torch.from_numpy(g["tgt_des"]))
    assert np.array_equal(s, g["s_mids"]) and np.array_equal(t, g["t_mids"])
    g = load("rigid_transform_3d")
    for case in ("n3", "n200_w", "n200_wthr", "mirror"):
        w = torch.from_numpy(g[case + "_w"].copy()) if case + "_w" in g.files else None
        T = TR.kabsch(torch.from_numpy(g[case + "_A"]), torch.from_numpy(g[case + "_B"]), w, float(g[case + "_thr"])).numpy()
        assert np.abs(T - g[case + "_T"]).max() < 1e-5
    g = load("post_refinement")
    T = TR.post_refinement(torch.from_numpy(g["T0"]), torch.from_numpy(g["src"]), torch.from_numpy(g["tgt"]), 0.10).numpy()
    assert np.abs(T - g["T_3dmatch"]).max() < 1e-5
    g = load("ransac")
    T, n, h, valid = TR.ransac(torch.from_numpy(g["ss"]), torch.from_numpy(g["tt"]), g["samples"], float(g["dist_th"]), float(g["similar_th"]))
    assert h == int(g["best_h"]) and n == int(g["best_count"]) and valid == int((g["counts"] >= 0).sum())
    assert np.abs(T.numpy() - g["T_best"]).max() < 1e-5
