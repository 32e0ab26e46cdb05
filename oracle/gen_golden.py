#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py

Every fixture stores the seeded inputs together with the outputs of the reference's own functions
(oracle/ref_import.py documents how the reference is imported and which third-party pieces are stubbed).
The committed .npz files are what pins the CPU oracle (tests/test_oracle_golden.py) — and through it the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O            # noqa: E402  (only for the shared Philox sample stream)
from oracle import ref_import as RI       # noqa: E402
from buffer_b200 import synthetic as S    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rot(g, n):
    return S.quat_to_rot(torch.randn(n, 4, generator=g)).numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    B = RI.load()
    SE3 = RI.se3()
    g = torch.Generator().manual_seed(20231017)

    # ---- utils/SE3.py ---------------------------------------------------------------------------------------
    R1, R2 = rot(g, 1)[0], rot(g, 1)[0]
    t1, t2 = torch.randn(3, 1, generator=g).numpy(), torch.randn(3, 1, generator=g).numpy()
    pts = torch.randn(17, 3, generator=g).numpy()
    T1 = SE3.integrate_trans(torch.from_numpy(R1), torch.from_numpy(t1)).numpy()
    T2 = SE3.integrate_trans(torch.from_numpy(R2), torch.from_numpy(t2)).numpy()
    Rb, tb = rot(g, 4), torch.randn(4, 3, 1, generator=g).numpy()
    Tb = SE3.integrate_trans(torch.from_numpy(Rb), torch.from_numpy(tb)).numpy()
    ptsb = torch.randn(4, 9, 3, generator=g).numpy()
    np.random.seed(7)
    rm3 = SE3.rotation_matrix(3, 1.0); rm1 = SE3.rotation_matrix(1, 0.5); rm0 = SE3.rotation_matrix(0, 1.0); tm = SE3.translation_matrix(0.5)
    np.savez(os.path.join(OUT, "se3.npz"), R1=R1, R2=R2, t1=t1, t2=t2, pts=pts, T1=T1, T2=T2,
             T1_np=SE3.integrate_trans(R1.astype(np.float64), t1.astype(np.float64)),
             transform_torch=SE3.transform(torch.from_numpy(pts), torch.from_numpy(T1)).numpy(),
             transform_numpy=SE3.transform(pts.astype(np.float64), T1.astype(np.float64)),
             concat_torch=SE3.concatenate(torch.from_numpy(T1), torch.from_numpy(T2)).numpy(),
             Rb=Rb, tb=tb, Tb=Tb, ptsb=ptsb,
             transform_batched=SE3.transform(torch.from_numpy(ptsb), torch.from_numpy(Tb)).numpy(),
             concat_batched=SE3.concatenate(torch.from_numpy(Tb), torch.from_numpy(Tb.copy())).numpy(),
             rm3=rm3, rm1=rm1, rm0=rm0, tm=tm)

    # ---- rigid_transform_3d (models/BUFFER.py:424-464) --------------------------------------------------------
    cases = {}
    for name, bs, n, noise, use_w, thr in (("n3", 5, 3, 0.0, False, 0), ("n3_noise", 5, 3, 0.02, False, 0), ("n200_w", 2, 200, 0.01, True, 0),
                                           ("n200_wthr", 2, 200, 0.01, True, 0.4), ("n2000", 1, 2000, 0.05, True, 0)):
        A = torch.randn(bs, n, 3, generator=g)
        Rg = torch.from_numpy(rot(g, bs)); tg = torch.randn(bs, 1, 3, generator=g)
        Bp = A @ Rg.transpose(-1, -2) + tg + noise * torch.randn(bs, n, 3, generator=g)
        w = torch.rand(bs, n, generator=g) if use_w else None
        T = RI.rigid_transform_3d(A.numpy(), Bp.numpy(), None if w is None else w.numpy().copy(), thr)
        cases[name + "_A"] = A.numpy(); cases[name + "_B"] = Bp.numpy(); cases[name + "_T"] = T
        cases[name + "_thr"] = np.float32(thr)
        if w is not None:
            cases[name + "_w"] = w.numpy()
    # reflection case: mirrored target -> det fix must engage
    A = torch.randn(1, 50, 3, generator=g); Bm = A.clone(); Bm[..., 2] = -Bm[..., 2]
    cases["mirror_A"] = A.numpy(); cases["mirror_B"] = Bm.numpy(); cases["mirror_T"] = RI.rigid_transform_3d(A.numpy(), Bm.numpy()); cases["mirror_thr"] = np.float32(0)
    np.savez(os.path.join(OUT, "rigid_transform_3d.npz"), **cases)

    # ---- synthetic pair shared by the remaining fixtures -------------------------------------------------------
    b = S.make_pairs(1, 400, cfg_id=101)
    src_des, tgt_des = b.src_des[0, :380].numpy(), b.tgt_des[0].numpy()        # ragged: 380 x 400
    s_mids, t_mids = RI.mutual_matching(src_des, tgt_des)
    np.savez(os.path.join(OUT, "mutual_matching.npz"), src_des=src_des, tgt_des=tgt_des, s_mids=s_mids, t_mids=t_mids)

    src_xyz, tgt_xyz = b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy()
    ss, tt = src_xyz[s_mids], tgt_xyz[t_mids]

    # ---- post_refinement (models/BUFFER.py:382-418) -----------------------------------------------------------
    T0 = b.T_gt[0].clone(); T0[:3, 3] += torch.tensor([0.03, -0.02, 0.025]); T0 = T0.numpy()
    pr = {"src": ss, "tgt": tt, "T0": T0, "T_3dmatch": RI.post_refinement(T0, ss, tt, "3DMatch"), "T_kitti": RI.post_refinement(T0, ss, tt, "KITTI")}
    Tid = np.eye(4, dtype=np.float32); Tid[:3, 3] = 50.0            # no inliers at all -> returned unchanged
    pr["T_far"] = RI.post_refinement(Tid, ss, tt, "3DMatch"); pr["T0_far"] = Tid
    np.savez(os.path.join(OUT, "post_refinement.npz"), **pr)

    # ---- a3 / a4: LRF hypotheses + scoring (models/BUFFER.py:294-311) --------------------------------------------
    A_ = len(s_mids)
    ss_R = rot(g, A_); ind = (torch.rand(A_, generator=g) * 20).numpy()
    Rg = b.T_gt[0, :3, :3].double().numpy()
    ang = ind.astype(np.float64) * 2 * np.pi / 20 + 1e-6
    Rz = np.zeros((A_, 3, 3)); Rz[:, 0, 0] = np.cos(ang); Rz[:, 0, 1] = -np.sin(ang); Rz[:, 1, 0] = np.sin(ang); Rz[:, 1, 1] = np.cos(ang); Rz[:, 2, 2] = 1
    tt_R = (Rg @ ss_R.astype(np.float64) @ np.transpose(Rz, (0, 2, 1))).astype(np.float32)
    bad = ~b.inlier[0].numpy()[s_mids]
    tt_R[bad] = rot(g, int(bad.sum()))
    R_h, t_h = RI.lrf_hypotheses(ind, ss_R, tt_R, ss, tt, azi_n=20)
    inlier_num, best_ind, inlier_ind, thr = RI.score_hypotheses(R_h, t_h, ss, tt, azi_n=20, inlier_th=1 / 3)
    np.savez(os.path.join(OUT, "lrf_scoring.npz"), ind=ind.astype(np.float32), ss_R=ss_R, tt_R=tt_R, ss=ss, tt=tt, R=R_h, t=t_h,
             inlier_num=inlier_num, best_ind=np.int64(best_ind), inlier_ind=inlier_ind, thr=thr)

    # ---- RANSAC with Open3D semantics on the reference's own Kabsch / transform ------------------------------------
    K = len(s_mids); H = 600; seed = 0xC0FFEE; pair_id = 5
    samples = np.stack([O.sample3(seed, pair_id, h, K) for h in range(H)]).astype(np.int64)
    corr = np.stack([np.arange(K), np.arange(K)], 1)
    T_best, best_cnt, best_h, counts = RI.ransac_open3d_semantics(ss, tt, corr, 0.10, 0.8, samples)
    np.savez(os.path.join(OUT, "ransac.npz"), ss=ss, tt=tt, samples=samples, seed=np.uint64(seed), pair_id=np.uint32(pair_id), H=np.int64(H),
             dist_th=np.float32(0.10), similar_th=np.float32(0.8), T_best=T_best, best_count=np.int64(best_cnt), best_h=np.int64(best_h), counts=counts,
             T_gt=b.T_gt[0].numpy())
    # ---- "next" rows: get_matching_indices (numpy twin ThreeDMatch/dataset.py:14-22) and cal_Z_axis' SVD ---------------------
    import ThreeDMatch.dataset as TD
    import utils.common as UC
    gm_src = (torch.rand(700, 3, generator=g) * 2).numpy()
    Tg = b.T_gt[0].numpy()
    moved = gm_src @ Tg[:3, :3].T + Tg[:3, 3]
    gm_tgt = (moved + 0.02 * torch.randn(700, 3, generator=g).numpy())[torch.randperm(700, generator=g).numpy()][:610].astype(np.float32)
    gm = TD.get_matching_indices(gm_src, gm_tgt, Tg, 0.05)
    local = torch.randn(40, 64, 3, generator=g) * torch.tensor([1.0, 0.6, 0.15])
    local = local @ torch.from_numpy(rot(g, 40)).transpose(-1, -2)
    refp = torch.randn(40, 3, generator=g)
    z_axis = UC.cal_Z_axis(local, ref_point=refp).numpy()                      # uses the torch_batch_svd stub = torch.svd
    cov = torch.matmul(local.transpose(-1, -2), local)
    u_, s_, v_ = torch.svd(cov)
    np.savez(os.path.join(OUT, "next_rows.npz"), gm_src=gm_src, gm_tgt=gm_tgt, gm_T=Tg, gm_voxel=np.float32(0.05), gm_pairs=gm,
             local=local.numpy(), ref_point=refp.numpy(), z_axis=z_axis, cov=cov.numpy(), svd_u=u_.numpy(), svd_s=s_.numpy(), svd_v=v_.numpy())

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


# ---- evaluation helpers (second half): ThreeDMatch/test.py:18-197 ---------------------------------------------
def reference_eval_functions():
    """The reference keeps its evaluator inside a script whose module level loads the model, the data loaders and nibabel.  Only the five
    function definitions are needed, so they are compiled UNMODIFIED from the file's AST and run in a namespace that provides numpy and an
    `nq` stub: nibabel (third-party, absent here) is replaced by the restatement of its `mat2quat` in buffer_b200/evaluation.py, which the
    test additionally cross-checks against scipy's Rotation."""
    import ast
    import types
    from buffer_b200 import evaluation as E
    path = os.path.join(RI.REF_ROOT, "ThreeDMatch", "test.py")
    tree = ast.parse(open(path).read())
    mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef)], type_ignores=[])
    ns = {"np": np, "nq": types.SimpleNamespace(mat2quat=E.mat2quat)}
    exec(compile(mod, path, "exec"), ns)
    return ns


def gen_evaluation():
    import tempfile
    ns = reference_eval_functions()
    rng = np.random.RandomState(31)
    g = torch.Generator().manual_seed(77)
    n_frag = 9
    pairs = [(i, j) for i in range(n_frag) for j in range(i + 1, n_frag) if rng.rand() < 0.55]
    gt = np.zeros((len(pairs), 4, 4)); gt[:, 3, 3] = 1
    gt[:, :3, :3] = rot(g, len(pairs)); gt[:, :3, 3] = rng.randn(len(pairs), 3)
    info = np.zeros((len(pairs), 6, 6))
    for k in range(len(pairs)):
        a = rng.randn(6, 6); info[k] = a @ a.T * 50 + np.eye(6) * 200
    gt_log = "".join("%d\t %d\t %d\n" % (i, j, n_frag) + "".join("\t ".join(repr(float(v)) for v in row) + "\t \n" for row in gt[k])
                     for k, (i, j) in enumerate(pairs))
    gt_info = "".join("%d\t%d\t%d\n" % (i, j, n_frag) + "".join("\t".join(repr(float(v)) for v in row) + "\n" for row in info[k])
                      for k, (i, j) in enumerate(pairs))
    # estimates: most pairs of the ground truth (some perturbed slightly, some grossly) plus two pairs that are not in it
    est_pairs, est = [], []
    for k, (i, j) in enumerate(pairs):
        if rng.rand() < 0.15:
            continue
        T = np.linalg.inv(gt[k])                      # the .log stores inv(T_est): a perfect estimate reproduces gt
        kind = rng.rand()
        if kind < 0.6:
            d = np.eye(4); d[:3, :3] = S.quat_to_rot(torch.tensor([[1.0, 0.004, -0.003, 0.002]])).numpy()[0]; d[:3, 3] = rng.randn(3) * 0.01
            T = T @ d
        elif kind < 0.85:
            d = np.eye(4); d[:3, :3] = rot(g, 1)[0]; d[:3, 3] = rng.randn(3) * 0.5
            T = T @ d
        est_pairs.append((i, j)); est.append(T)
    for (i, j) in ((0, 1), (2, 3)):
        if (i, j) not in pairs:
            est_pairs.append((i, j)); est.append(np.eye(4))
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "gt.log"), "w").write(gt_log)
        open(os.path.join(tmp, "gt.info"), "w").write(gt_info)
        gt_pairs_r, gt_traj_r = ns["read_trajectory"](os.path.join(tmp, "gt.log"))
        n_fr, gt_info_r = ns["read_trajectory_info"](os.path.join(tmp, "gt.info"))
        # the reference's own writer lines (ThreeDMatch/test.py:250-261), verbatim semantics
        est_log = os.path.join(tmp, "est.log")
        with open(est_log, "a+") as f:
            for (i, j), T in zip(est_pairs, est):
                trans = np.linalg.inv(T)
                f.write(f'{i}\t {j}\t  1\n')
                for r in range(4):
                    f.write(f"{trans[r, 0]}\t {trans[r, 1]}\t {trans[r, 2]}\t {trans[r, 3]}\t \n")
        est_log_text = open(est_log).read()
        est_pairs_r, est_traj_r = ns["read_trajectory"](est_log)
        precision, recall, flags, errors = ns["evaluate_registration"](n_fr, est_traj_r, est_pairs_r, gt_pairs_r, gt_traj_r, gt_info_r)
        ext = ns["extract_corresponding_trajectors"](est_pairs_r[:5].copy(), gt_pairs_r, gt_traj_r) if all(tuple(int(x) for x in p[:2]) in pairs for p in est_pairs_r[:5]) else np.zeros((0, 4, 4))
        terr = np.array([ns["computeTransformationErr"](np.linalg.inv(gt_traj_r[k]) @ est_traj_r[min(k, len(est_traj_r) - 1)], gt_info_r[k]) for k in range(min(6, len(pairs)))])
    # the inline "recall of DGR" block (ThreeDMatch/test.py:263-283), restated verbatim on the same estimates
    import math
    states = []
    gt_of_est = {p: gt[k] for k, p in enumerate(pairs)}
    te_in, tg_in = [], []
    for p, T in zip(est_pairs, est):
        if p not in gt_of_est:
            continue
        trans_est, trans = T, np.linalg.inv(gt_of_est[p])
        rte = np.linalg.norm(trans_est[:3, 3] - trans[:3, 3])
        rre = np.arccos(np.clip((np.trace(trans_est[:3, :3].T @ trans[:3, :3]) - 1) / 2, -1 + 1e-16, 1 - 1e-16)) * 180 / math.pi
        states.append(np.array([rte < 0.3 and rre < 15, rte, rre])); te_in.append(trans_est); tg_in.append(trans)
    states = np.array(states)
    np.savez(os.path.join(OUT, "evaluation.npz"), gt_log=gt_log, gt_info=gt_info, est_log=est_log_text, est_pairs=np.array(est_pairs), est=np.array(est),
             gt_pairs_r=gt_pairs_r, gt_traj_r=gt_traj_r, n_fr=n_fr, gt_info_r=gt_info_r, est_pairs_r=est_pairs_r, est_traj_r=est_traj_r,
             precision=precision, recall=recall, flags=np.array(flags), errors=errors, ext=ext, terr=terr,
             dgr_states=states, dgr_est=np.array(te_in), dgr_gt=np.array(tg_in),
             dgr_recall=states[:, 0].sum() / states.shape[0], dgr_te=states[states[:, 0] == 1, 1].mean(), dgr_re=states[states[:, 0] == 1, 2].mean())
    print("evaluation.npz: %d gt pairs, %d estimates, precision %.3f recall %.3f, DGR recall %.3f" % (len(pairs), len(est_pairs), precision, recall, states[:, 0].mean()))


def gen_config1():
    """BASELINE config 1 at full size: ONE 3DMatch-sized pair (5 000 x 5 000 keypoints, 32-d, 70 % outliers) through the reference's own
    functions: buffer.mutual_matching, the inline vote block (:294-311), the Open3D-semantics loop on the reference's Kabsch for the first
    5 000 hypotheses of the shared Philox stream, buffer.post_refinement on all matches.  The inputs are NOT stored (2.6 MB): they are
    regenerated from the seed (buffer_b200.synthetic, CPU generator) and checked against the stored checksums."""
    c = S.CONFIGS[1]
    b = S.make_pairs(1, first_pair=0, **c["gen"])
    N = c["gen"]["num_kpts"]
    src_des, tgt_des = b.src_des[0].numpy(), b.tgt_des[0].numpy()
    src_xyz, tgt_xyz = b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy()
    s_mids, t_mids = RI.mutual_matching(src_des, tgt_des)
    ss, tt = src_xyz[s_mids], tgt_xyz[t_mids]
    K = len(s_mids); H = 5000; seed = 0xB0FFE7; pair_id = 0
    samples = np.stack([O.sample3(seed, pair_id, h, K) for h in range(H)]).astype(np.int64)
    corr = np.stack([np.arange(K), np.arange(K)], 1)
    T_best, best_cnt, best_h, counts = RI.ransac_open3d_semantics(ss, tt, corr, c["dist_th"], c["similar_th"], samples)
    T_ref = RI.post_refinement(T_best, ss, tt, "3DMatch")
    # the vote block on all K matches
    inl = b.inlier[0].numpy()[s_mids]
    ind, ss_R, tt_R = S.make_lrf_votes(b.T_gt[0, :3, :3], torch.from_numpy(inl), azi_n=20, seed=1)
    R_h, t_h = RI.lrf_hypotheses(ind.numpy(), ss_R.numpy(), tt_R.numpy(), ss, tt, azi_n=20)
    inlier_num, best_ind, inlier_ind, thr = RI.score_hypotheses(R_h, t_h, ss, tt, azi_n=20, inlier_th=1 / 3)
    # reference flow :313-329 on the voted subset
    sub = np.stack([inlier_ind, inlier_ind], 1)
    samples2 = np.stack([O.sample3(seed, pair_id, h, len(sub)) for h in range(H)]).astype(np.int64)
    T_best2, best_cnt2, best_h2, counts2 = RI.ransac_open3d_semantics(ss, tt, sub, c["dist_th"], c["similar_th"], samples2)
    T_ref2 = RI.post_refinement(T_best2, ss, tt, "3DMatch")
    chk = np.array([np.float64(x.astype(np.float64).sum()) for x in (src_des, tgt_des, src_xyz, tgt_xyz, ss_R.numpy(), tt_R.numpy(), ind.numpy())])
    np.savez_compressed(os.path.join(OUT, "config1.npz"), checksums=chk, N=np.int64(N), s_mids=s_mids.astype(np.int16), t_mids=t_mids.astype(np.int16),
                        H=np.int64(H), seed=np.uint64(seed), pair_id=np.uint32(pair_id), dist_th=np.float32(c["dist_th"]), similar_th=np.float32(c["similar_th"]),
                        counts=counts.astype(np.int16), best_h=np.int64(best_h), best_count=np.int64(best_cnt), T_best=T_best, T_refined=T_ref, T_gt=b.T_gt[0].numpy(),
                        vote_inlier_num=inlier_num.astype(np.int16), vote_best_ind=np.int64(best_ind), vote_inlier_ind=inlier_ind.astype(np.int16),
                        sub_counts=counts2.astype(np.int16), sub_best_h=np.int64(best_h2), sub_best_count=np.int64(best_cnt2), sub_T_best=T_best2, sub_T_refined=T_ref2)
    print("config1.npz: %d mutual matches, RANSAC winner h=%d with %d inliers (%d valid of %d); vote winner %d with %d inliers; subset RANSAC h=%d with %d inliers; %d bytes"
          % (K, best_h, best_cnt, int((counts >= 0).sum()), H, best_ind, len(inlier_ind), best_h2, best_cnt2, os.path.getsize(os.path.join(OUT, "config1.npz"))))


if __name__ == "__main__":
    if "--config1" in sys.argv:             # `--config1`: only (re)generate config1.npz
        gen_config1()
        sys.exit(0)
    if "--evaluation" not in sys.argv:      # `--evaluation`: only (re)generate evaluation.npz
        main()
        gen_config1()
    gen_evaluation()
