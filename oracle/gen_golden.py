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


if __name__ == "__main__":
    main()
