"""GPU parity tests proper: every result of the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer / index / count outputs must be bit-exact; transforms are compared bit-for-bit where the
arithmetic order is pinned (RANSAC fit, refinement) and within 1e-5 against float64 ground truth otherwise."""
import os

import numpy as np
import pytest
import torch

from buffer_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pairs(P, N, **kw):
    return S.make_pairs(P, N, **kw)


@pytest.fixture(params=[0, 1], ids=["k1_fp32", "k1_tensor_filter"])
def algo(request, backend):
    """run the test with both mutual-NN implementations (FP32 FFMA2 kernel / tcgen05 filter + exact re-check)"""
    backend.set_k1_algo(request.param)
    yield request.param
    backend.set_k1_algo(1)


@pytest.mark.parametrize("M,N", [(1, 1), (3, 5), (64, 64), (65, 63), (511, 513), (512, 512), (700, 1300), (2500, 1500)])
def test_mutual_nn_bit_exact_ragged(oracle, backend, algo, M, N):
    g = torch.Generator().manual_seed(M * 10007 + N)
    src = torch.nn.functional.normalize(torch.randn(M, 32, generator=g), dim=-1)
    tgt = torch.nn.functional.normalize(torch.randn(N, 32, generator=g), dim=-1)
    if M > 10 and N > 10:           # plant duplicates -> exact distance ties, lowest index must win
        tgt[7] = tgt[3]; src[9] = src[2]
        k = min(M, N) // 2
        tgt[:k] = torch.nn.functional.normalize(src[:k] + 0.05 * torch.randn(k, 32, generator=g), dim=-1)
    nn_s, nn_t, ds, dt = oracle.mutual_nn(src.numpy(), tgt.numpy(), want_dist=True)
    r = backend.mutual_matching_device(src.to(DEV), tgt.to(DEV), want_dist=True)
    assert np.array_equal(r["nn_s"].cpu().numpy(), nn_s)
    assert np.array_equal(r["nn_t"].cpu().numpy(), nn_t)
    assert np.array_equal(r["dist_s"].cpu().numpy(), ds)
    assert np.array_equal(r["dist_t"].cpu().numpy(), dt)
    s, t = oracle.mutual_select(nn_s, nn_t)
    n = int(r["n_mutual"].item())
    assert n == len(s)
    assert np.array_equal(r["s_mids"][:n].cpu().numpy(), s) and np.array_equal(r["t_mids"][:n].cpu().numpy(), t)
    s2, t2 = backend.mutual_matching(src.to(DEV), tgt.to(DEV))
    assert s2.dtype == np.int64 and np.array_equal(s2, s) and np.array_equal(t2, t)


def test_mutual_nn_unnormalised_and_splits(oracle, backend, algo):
    g = torch.Generator().manual_seed(5)
    src = torch.randn(900, 32, generator=g) * torch.rand(900, 1, generator=g) * 3
    tgt = torch.randn(1100, 32, generator=g) * torch.rand(1100, 1, generator=g) * 3
    nn_s, nn_t = oracle.mutual_nn(src.numpy(), tgt.numpy())
    so = torch.tensor([0, 900], dtype=torch.int32, device=DEV); to = torch.tensor([0, 1100], dtype=torch.int32, device=DEV)
    for splits in (1, 2, 5, 18):
        r = backend.mutual_matching_batched(src.to(DEV), tgt.to(DEV), so, to, 900, 1100, col_splits=splits)
        assert np.array_equal(r["nn_s"].cpu().numpy(), nn_s), splits
        assert np.array_equal(r["nn_t"].cpu().numpy(), nn_t), splits


def test_mutual_nn_adversarial_inputs(oracle, backend, algo):
    """near-duplicate floods (candidate-list overflow -> exact scan), exact duplicates (ties -> lowest index), zero rows,
    huge / tiny norms, and a target set wide enough to need column splits in the tensor-core kernel (N > 8192)"""
    g = torch.Generator().manual_seed(77)
    nrm = lambda x: torch.nn.functional.normalize(x, dim=-1)
    cases = []
    base = nrm(torch.randn(1, 32, generator=g))
    flood = nrm(base + 1e-4 * torch.randn(600, 32, generator=g))            # 600 targets inside everybody's 2-eps band
    cases.append((nrm(base + 1e-3 * torch.randn(300, 32, generator=g)), torch.cat([flood, nrm(torch.randn(400, 32, generator=g))])))
    dup_t = nrm(torch.randn(50, 32, generator=g)).repeat(20, 1)               # every target row 20 times: exact ties
    cases.append((nrm(torch.randn(700, 32, generator=g)), dup_t))
    z = nrm(torch.randn(500, 32, generator=g)); z[::7] = 0.0                  # zero descriptors
    cases.append((z, nrm(torch.randn(450, 32, generator=g))))
    cases.append((1e3 * torch.randn(400, 32, generator=g), 1e3 * torch.randn(500, 32, generator=g)))
    cases.append((1e-3 * torch.randn(400, 32, generator=g), 1e-3 * torch.randn(500, 32, generator=g) * torch.rand(500, 1, generator=g)))
    # beyond the f16 range of the tensor-core operands (|x| > 65504 -> the pair is scanned exactly), one side only, and far below it
    cases.append((1e5 * torch.randn(300, 32, generator=g), 1e5 * torch.randn(350, 32, generator=g)))
    cases.append((nrm(torch.randn(260, 32, generator=g)), 3e5 * torch.randn(300, 32, generator=g)))
    cases.append((1e-6 * torch.randn(400, 32, generator=g), 1e-7 * torch.randn(500, 32, generator=g)))
    cases.append((nrm(torch.randn(300, 32, generator=g)), 1e-6 * nrm(torch.randn(280, 32, generator=g))))
    big = nrm(torch.randn(400, 32, generator=g)); big[7, 3] = 7e4                                     # a single out-of-range element
    cases.append((big, nrm(torch.randn(380, 32, generator=g))))
    cases.append((nrm(torch.randn(300, 32, generator=g)), nrm(torch.randn(9000, 32, generator=g))))
    cases.append((nrm(torch.randn(9000, 32, generator=g)), nrm(torch.randn(300, 32, generator=g))))
    for src, tgt in cases:
        nn_s, nn_t, ds, dt = oracle.mutual_nn(src.numpy(), tgt.numpy(), want_dist=True)
        r = backend.mutual_matching_device(src.to(DEV), tgt.to(DEV), want_dist=True)
        assert np.array_equal(r["nn_s"].cpu().numpy(), nn_s)
        assert np.array_equal(r["nn_t"].cpu().numpy(), nn_t)
        assert np.array_equal(r["dist_s"].cpu().numpy(), ds) and np.array_equal(r["dist_t"].cpu().numpy(), dt)


def test_mutual_matching_batched_varlen(oracle, backend, algo):
    sizes = [(300, 200), (1, 7), (1025, 513), (64, 640), (5, 5)]
    g = torch.Generator().manual_seed(11)
    srcs = [torch.nn.functional.normalize(torch.randn(m, 32, generator=g), dim=-1) for m, _ in sizes]
    tgts = [torch.nn.functional.normalize(torch.randn(n, 32, generator=g), dim=-1) for _, n in sizes]
    sx = [torch.randn(m, 3, generator=g) for m, _ in sizes]; tx = [torch.randn(n, 3, generator=g) for _, n in sizes]
    so = backend._offsets([m for m, _ in sizes], torch.device(DEV)); to = backend._offsets([n for _, n in sizes], torch.device(DEV))
    r = backend.mutual_matching_batched(torch.cat(srcs).to(DEV), torch.cat(tgts).to(DEV), so, to, 1025, 640,
                                        torch.cat(sx).to(DEV), torch.cat(tx).to(DEV))
    torch.cuda.synchronize()
    so_h = so.cpu().numpy(); to_h = to.cpu().numpy()
    for p, (m, n) in enumerate(sizes):
        nn_s, nn_t = oracle.mutual_nn(srcs[p].numpy(), tgts[p].numpy())
        assert np.array_equal(r["nn_s"][so_h[p]:so_h[p + 1]].cpu().numpy(), nn_s)
        assert np.array_equal(r["nn_t"][to_h[p]:to_h[p + 1]].cpu().numpy(), nn_t)
        s, t = oracle.mutual_select(nn_s, nn_t)
        k = int(r["n_mutual"][p].item())
        assert k == len(s)
        assert np.array_equal(r["s_mids"][so_h[p]:so_h[p] + k].cpu().numpy(), s)
        corr = oracle.gather_corr(sx[p].numpy(), tx[p].numpy(), s, t)
        assert np.array_equal(r["corr"][so_h[p]:so_h[p] + k].cpu().numpy()[:, [0, 1, 2, 4, 5, 6]], corr[:, [0, 1, 2, 4, 5, 6]])


@pytest.mark.parametrize("N,H,rho", [(600, 4000, 0.7), (2500, 6000, 0.7), (1000, 20000, 0.93)])
def test_ransac_counts_and_fit_bit_exact(oracle, backend, N, H, rho):
    b = _pairs(3, N, cfg_id=7, outlier_ratio=rho)
    seed = 0x1234ABCD5678
    for p in range(3):
        s, t = oracle.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        corr = oracle.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)
        best = oracle.ransac(corr, seed, 40 + p, H, 0.1, 0.8)
        T_o, cnt_o, bh_o = oracle.ransac_finalize(corr, seed, 40 + p, best, 0.1, 0.8)
        cd = torch.from_numpy(corr).to(DEV)
        off = torch.tensor([0, len(s)], dtype=torch.int32, device=DEV); cnt = torch.tensor([len(s)], dtype=torch.int32, device=DEV)
        _, counts_o = oracle.ransac(corr, seed, 40 + p, H, 0.1, 0.8, want_counts=True)
        for splits in (1, 7):
            nv = torch.zeros(1, dtype=torch.int32, device=DEV)
            bp = backend.ransac_batched(cd, off, cnt, H, 0.1, 0.8, seed=seed, pair_id_base=40 + p, splits=splits, valid_count=nv)
            assert int(bp.item()) == best, (p, splits)
            assert int(nv.item()) == int((counts_o >= 0).sum())      # same set of hypotheses passes the checkers
        T, inl, bh = backend.ransac_finalize_batched(cd, off, cnt, bp, 0.1, 0.8, seed=seed, pair_id_base=40 + p)
        assert int(inl.item()) == cnt_o and int(bh.item()) == bh_o
        assert np.array_equal(T[0].cpu().numpy(), T_o)            # same closed-form fit, bit for bit
        # split the hypothesis range over two calls (the multi-GPU mode) -> same packed best
        bp2 = backend.ransac_batched(cd, off, cnt, H, 0.1, 0.8, seed=seed, pair_id_base=40 + p, h_begin=0, h_end=H // 3)
        bp2 = backend.ransac_batched(cd, off, cnt, H, 0.1, 0.8, seed=seed, pair_id_base=40 + p, h_begin=H // 3, h_end=H, best_packed=bp2)
        assert int(bp2.item()) == best


def test_ransac_degenerate_inputs(oracle, backend):
    dev = torch.device(DEV)
    for K in (0, 1, 2):
        corr = torch.randn(max(K, 1), 8, device=dev)
        off = torch.tensor([0, K], dtype=torch.int32, device=dev); cnt = torch.tensor([K], dtype=torch.int32, device=dev)
        bp = backend.ransac_batched(corr, off, cnt, 1000, 0.1, 0.8)
        T, inl, bh = backend.ransac_finalize_batched(corr, off, cnt, bp, 0.1, 0.8)
        assert int(bp.item()) == 0 and int(inl.item()) == 0 and int(bh.item()) == -1
        assert torch.equal(T[0].cpu(), torch.eye(4))
    # collinear / identical points: no valid hypothesis -> identity
    corr = torch.zeros(50, 8, device=dev); corr[:, 0] = torch.arange(50, device=dev); corr[:, 4] = torch.arange(50, device=dev)
    off = torch.tensor([0, 50], dtype=torch.int32, device=dev); cnt = torch.tensor([50], dtype=torch.int32, device=dev)
    bp = backend.ransac_batched(corr, off, cnt, 2000, 0.1, 0.8)
    assert int(bp.item()) == oracle.ransac(corr.cpu().numpy(), 0, 0, 2000, 0.1, 0.8)


def test_open3d_style_call(oracle, backend):
    b = _pairs(1, 800, cfg_id=9)
    s, t = oracle.mutual_matching(b.src_des[0].numpy(), b.tgt_des[0].numpy())
    keep = np.arange(0, len(s), 2)
    corr_idx = np.stack([s[keep], t[keep]], 1)
    res = backend.registration_ransac_based_on_correspondence(b.src_xyz[0].to(DEV), b.tgt_xyz[0].to(DEV), corr_idx, 0.1, 0.8, iter_n=5000,
                                                              confidence=0.999, seed=3, pair_id=1)
    corr = oracle.gather_corr(b.src_xyz[0].numpy(), b.tgt_xyz[0].numpy(), s[keep], t[keep])
    best, iters = oracle.ransac_confidence(corr, 3, 1, 5000, 0.1, 0.8, 0.999)      # the reference's 3DMatch criteria (config.py:65)
    assert iters < 5000
    T_o, cnt_o, _ = oracle.ransac_finalize(corr, 3, 1, best, 0.1, 0.8)
    assert res.transformation.dtype == np.float64 and res.transformation.shape == (4, 4)
    assert np.array_equal(res.transformation.astype(np.float32), T_o) and res.inlier_count == cnt_o
    res0 = backend.registration_ransac_based_on_correspondence(b.src_xyz[0].to(DEV), b.tgt_xyz[0].to(DEV), corr_idx[:2], 0.1, 0.8, iter_n=100)
    assert np.array_equal(res0.transformation, np.eye(4))


def test_lrf_hypotheses_and_scoring_bit_exact(oracle, backend):
    A = 700
    g = torch.Generator().manual_seed(21)
    ss_R = S.quat_to_rot(torch.randn(A, 4, generator=g)); ind = torch.rand(A, generator=g) * 20
    Rg = S.quat_to_rot(torch.randn(1, 4, generator=g))[0]; tg = torch.rand(3, generator=g)
    ss = torch.rand(A, 3, generator=g) * 3 - 1.5
    tt = ss @ Rg.T + tg + 0.01 * torch.randn(A, 3, generator=g)
    ang = ind.double() * 2 * np.pi / 20 + 1e-6
    Rz = torch.zeros(A, 3, 3, dtype=torch.float64); Rz[:, 0, 0] = torch.cos(ang); Rz[:, 0, 1] = -torch.sin(ang)
    Rz[:, 1, 0] = torch.sin(ang); Rz[:, 1, 1] = torch.cos(ang); Rz[:, 2, 2] = 1
    tt_R = (Rg.double() @ ss_R.double() @ Rz.transpose(-1, -2)).float()      # so that tt_R Rz ss_R^T = Rg for true matches
    out = torch.rand(A, generator=g) < 0.6
    tt[out] = torch.rand(int(out.sum()), 3, generator=g) * 3
    tt_R[out] = S.quat_to_rot(torch.randn(int(out.sum()), 4, generator=g))
    cs = torch.stack([torch.cos(ang), torch.sin(ang)], -1).float()
    R_o, t_o = oracle.lrf_hypotheses(cs.numpy(), ss_R.numpy(), tt_R.numpy(), ss.numpy(), tt.numpy())
    R_g, t_g = backend.lrf_hypotheses(ind.to(DEV), ss_R.to(DEV), tt_R.to(DEV), ss.to(DEV), tt.to(DEV))
    assert np.array_equal(R_g.cpu().numpy(), R_o) and np.array_equal(t_g.cpu().numpy(), t_o)
    thr = backend.inlier_threshold(ss)
    for th in (thr, 0.08):
        c_o, b_o, m_o = oracle.score_hypotheses(R_o, t_o, ss.numpy(), tt.numpy(), th if np.isscalar(th) else th.numpy())
        c_g, b_g, m_g = backend.score_hypotheses(R_g, t_g, ss.to(DEV), tt.to(DEV), th if np.isscalar(th) else th.to(DEV))
        assert np.array_equal(c_g.cpu().numpy(), c_o) and int(b_g.item()) == b_o and np.array_equal(m_g.cpu().numpy(), m_o)
    assert c_o.max() > 0.25 * A


@pytest.mark.parametrize("bs,n", [(1, 3), (4, 3), (2, 50), (3, 257), (1, 5000)])
def test_rigid_transform_3d_bit_exact(oracle, backend, bs, n):
    g = torch.Generator().manual_seed(bs * 100 + n)
    A = torch.randn(bs, n, 3, generator=g); Rg = S.quat_to_rot(torch.randn(bs, 4, generator=g)); tg = torch.randn(bs, 1, 3, generator=g)
    B = A @ Rg.transpose(-1, -2) + tg + 0.01 * torch.randn(bs, n, 3, generator=g)
    w = torch.rand(bs, n, generator=g)
    for weights, thr in ((None, 0), (w, 0), (w, 0.3)):
        To = oracle.rigid_transform_3d(A.numpy(), B.numpy(), None if weights is None else weights.numpy().copy(), thr)
        wd = None if weights is None else weights.clone().to(DEV)
        Tg = backend.rigid_transform_3d(A.to(DEV), B.to(DEV), wd, thr)
        assert np.array_equal(Tg.cpu().numpy(), To)
        if weights is not None and thr > 0:
            assert float(wd[wd < thr].abs().sum()) == 0.0          # in-place zeroing like the reference
    # reflection case: mirrored target must still give a proper rotation
    Bm = B.clone(); Bm[..., 0] = -Bm[..., 0]
    Tm = backend.rigid_transform_3d(A.to(DEV), Bm.to(DEV)).cpu()
    if n > 3:
        assert torch.allclose(torch.det(Tm[:, :3, :3]), torch.ones(bs), atol=1e-5)


def test_post_refinement_bit_exact(oracle, backend):
    b = _pairs(4, 1500, cfg_id=13)
    for p in range(4):
        s, t = oracle.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        corr = oracle.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)
        T0 = b.T_gt[p].clone(); T0[:3, 3] += 0.03; T0 = T0.numpy()
        To, it_o, inl_o = oracle.post_refinement(T0, corr, 0.10, 20)
        Tg = backend.post_refinement(torch.from_numpy(T0)[None].to(DEV), torch.from_numpy(corr[:, 0:3])[None].to(DEV),
                                     torch.from_numpy(corr[:, 4:7])[None].to(DEV))
        assert Tg.shape == (1, 4, 4) and np.array_equal(Tg[0].cpu().numpy(), To)
        assert it_o >= 1 and inl_o > 300


@pytest.mark.parametrize("N,M", [(0, 5), (1, 1), (300, 257), (4097, 3000)])
def test_get_matching_indices_bit_exact(oracle, backend, N, M):
    g = torch.Generator().manual_seed(N * 31 + M)
    src = torch.rand(N, 3, generator=g) * 3
    T = torch.eye(4); T[:3, :3] = S.quat_to_rot(torch.randn(1, 4, generator=g))[0]; T[:3, 3] = torch.randn(3, generator=g)
    tgt = (torch.rand(M, 3, generator=g) * 3) @ T[:3, :3].T + T[:3, 3]
    k = min(N, M) // 2
    if k:
        tgt[:k] = src[:k] @ T[:3, :3].T + T[:3, 3] + 0.01 * torch.randn(k, 3, generator=g)
        tgt[1] = tgt[0]                                              # exact tie -> lowest index
    pairs_o, nn_o, dist_o = oracle.get_matching_indices(src.numpy(), tgt.numpy(), T.numpy(), 0.05, want_nn=True)
    pairs, count, nn, dist = backend.get_matching_indices_device(src.to(DEV), tgt.to(DEV), T.to(DEV), 0.05)
    assert int(count.item()) == len(pairs_o)
    assert np.array_equal(pairs[: len(pairs_o)].cpu().numpy(), pairs_o)
    assert np.array_equal(nn.cpu().numpy(), nn_o) and np.array_equal(dist.cpu().numpy(), dist_o)
    m = backend.get_matching_indices(src.to(DEV), tgt.to(DEV), T.to(DEV), 0.05)
    assert m.is_cuda and m.dtype == torch.int64 and np.array_equal(m.cpu().numpy(), pairs_o)


def test_batched_svd3_bit_exact(oracle, backend):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1000, 3, 3, generator=g)
    x[:100] = x[:100] @ x[:100].transpose(-1, -2)                   # covariance-like (cal_Z_axis' input)
    x[100] = 0; x[101] = torch.outer(torch.tensor([1., 2., 3.]), torch.tensor([3., 2., 1.])); x[102, :, 2] = x[102, :, 0]
    x[103] = torch.eye(3); x[104] = torch.diag(torch.tensor([3., 3., 1.]))
    uo, so, vo = oracle.svd3(x.numpy())
    u, s, v = backend.svd(x.to(DEV))
    assert np.array_equal(u.cpu().numpy(), uo) and np.array_equal(s.cpu().numpy(), so) and np.array_equal(v.cpu().numpy(), vo)
    rec = u @ torch.diag_embed(s) @ v.transpose(-1, -2)
    assert float((rec.cpu() - x).abs().max()) < 2e-5 and bool((s[:, :-1] >= s[:, 1:]).all())
    eye = torch.eye(3, device=DEV)
    assert float((u.transpose(-1, -2) @ u - eye).abs().max()) < 1e-5 and float((v.transpose(-1, -2) @ v - eye).abs().max()) < 1e-5
    ref = torch.linalg.svdvals(x.double())
    assert float((s.cpu().double() - ref).abs().max()) < 1e-5 * float(ref.max())


def test_furthest_point_sample_bit_exact(oracle, backend):
    g = torch.Generator().manual_seed(12)
    xyz = torch.rand(3, 4000, 3, generator=g) * 3 - 1.5
    xyz[1, :50] = 0.0                                               # |p|^2 <= 1e-3: skipped by pointnet2's kernel
    xyz[2, 100] = xyz[2, 7]                                         # duplicate point
    idx_o = oracle.furthest_point_sample(xyz.numpy(), 300)
    idx = backend.furthest_point_sample(xyz.to(DEV), 300)
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), idx_o)
    assert len(set(idx_o[0].tolist())) == 300 and idx_o[0, 0] == 0
    feats = xyz.transpose(1, 2).contiguous().to(DEV)
    k = backend.gather_operation(feats, idx)
    assert torch.equal(k[0, :, 5], xyz[0, idx_o[0, 5]].to(DEV))
    # spread: sampled points cover the cloud far better than the first 300 points do
    d = torch.cdist(xyz[0], xyz[0][idx_o[0].astype(np.int64)]).min(dim=1)[0].max()
    assert float(d) < 0.45


def _vote_batch(P, N, cfg_id, rho=0.7):
    """P pairs with LRF-vote inputs: records of all mutual matches (from the oracle), ind / ss_R / tt_R row-aligned"""
    from oracle import oracle as O
    b = _pairs(P, N, cfg_id=cfg_id, outlier_ratio=rho)
    corr, ind, ssR, ttR, cnt = [], [], [], [], []
    for p in range(P):
        s, t = O.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        c = O.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)
        i_, sr, tr = S.make_lrf_votes(b.T_gt[p, :3, :3], b.inlier[p][torch.from_numpy(s)], seed=100 * cfg_id + p)
        corr.append(c); ind.append(i_.numpy()); ssR.append(sr.numpy()); ttR.append(tr.numpy()); cnt.append(len(s))
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    return b, np.concatenate(corr), np.concatenate(ind), np.concatenate(ssR), np.concatenate(ttR), off, np.asarray(cnt, np.int32)


def test_lrf_vote_batched_bit_exact(oracle, backend):
    """the fused LRF vote (models/BUFFER.py:294-311): counts of every proposal, the winner and the compacted inlier subset"""
    b, corr, ind, ssR, ttR, off, cnt = _vote_batch(3, 900, 23)
    cnt[2] = 2; ragged = np.concatenate([np.arange(off[p], off[p] + cnt[p]) for p in range(3)])      # a degenerate pair too
    d = lambda a: torch.from_numpy(a).to(DEV)
    r = backend.lrf_vote_batched(d(corr), d(off), d(cnt), d(ind), d(ssR), d(ttR), int(cnt.max()))
    torch.cuda.synchronize()
    for p in range(3):
        sl = slice(off[p], off[p] + cnt[p])
        c_o, b_o, sel_o = oracle.lrf_vote(corr[sl], ind[sl], ssR[sl], ttR[sl])
        k = int(r["sub_cnt"][p].item())
        assert np.array_equal(r["inlier_num"][sl].cpu().numpy(), c_o) and int(r["best_ind"][p].item()) == b_o
        assert k == len(sel_o) and np.array_equal(r["inlier_ind"][off[p]:off[p] + k].cpu().numpy(), sel_o)
        sub = r["sub_corr"][off[p]:off[p] + k].cpu().numpy()
        assert np.array_equal(sub[:, [0, 1, 2, 4, 5, 6]], corr[sl][sel_o][:, [0, 1, 2, 4, 5, 6]])
    assert int(r["sub_cnt"][0].item()) > 200


def test_pose_from_votes_matches_oracle(oracle, backend):
    """the reference's stage flow in one call (vote -> RANSAC on the voted subset -> refinement on ALL matches), with and without
    Open3D's confidence early exit"""
    b, corr, ind, ssR, ttR, off, cnt = _vote_batch(4, 1100, 29)
    d = lambda a: torch.from_numpy(a).to(DEV)
    for conf in (1.0, 0.999):
        To, nv_o, ni_o = oracle.pose_from_votes_batched(corr, off[:-1], cnt, ind, ssR, ttR, 6000, 77, 10, 0.1, 0.8, conf, 0.1, 20)
        T, nv, ni = backend.pose_from_votes_batched(d(corr.copy()), d(off), d(cnt), d(ind), d(ssR), d(ttR), int(cnt.max()), hypotheses=6000, seed=77,
                                                    pair_id_base=10, confidence=conf)
        assert np.array_equal(nv.cpu().numpy(), nv_o) and np.array_equal(ni.cpu().numpy(), ni_o)
        assert np.array_equal(T.cpu().numpy(), To)
        recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt)
        assert recall == 1.0 and float(rte.max()) < 0.01


@pytest.mark.parametrize("N,H,rho,conf", [(700, 5000, 0.7, 0.999), (2500, 20000, 0.7, 0.99), (1500, 30000, 0.9, 0.999), (6000, 20000, 0.8, 0.999)])
def test_ransac_confidence_bit_exact(oracle, backend, N, H, rho, conf):
    """Open3D's RANSACConvergenceCriteria(iter_n, confidence) (models/BUFFER.py:323-324): the GPU replays the sequential rule exactly"""
    b = _pairs(3, N, cfg_id=37, outlier_ratio=rho)
    corrs, cnts = [], []
    for p in range(3):
        s, t = oracle.mutual_matching(b.src_des[p].numpy(), b.tgt_des[p].numpy())
        corrs.append(oracle.gather_corr(b.src_xyz[p].numpy(), b.tgt_xyz[p].numpy(), s, t)); cnts.append(len(s))
    off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int32)
    cd = torch.from_numpy(np.concatenate(corrs)).to(DEV)
    od = torch.from_numpy(off).to(DEV); cn = torch.tensor(cnts, dtype=torch.int32, device=DEV)
    nv = torch.zeros(3, dtype=torch.int32, device=DEV)
    bp = backend.ransac_batched(cd, od, cn, H, 0.1, 0.8, seed=91, pair_id_base=5, h_begin=100, h_end=100 + H, confidence=conf, valid_count=nv)
    full = backend.ransac_batched(cd, od, cn, H, 0.1, 0.8, seed=91, pair_id_base=5, h_begin=100, h_end=100 + H)
    for p in range(3):
        best_o, iters = oracle.ransac_confidence(corrs[p], 91, 5 + p, H, 0.1, 0.8, conf, h_begin=100)
        assert int(bp[p].item()) == best_o, (p, iters)
        assert iters < H and int(nv[p].item()) <= int((oracle.ransac(corrs[p], 91, 5 + p, H, 0.1, 0.8, 100, 100 + H, want_counts=True)[1] >= 0).sum())
        assert int(full[p].item()) >= best_o                          # the full run can only be at least as good


def test_mutual_nn_row_split_equals_whole(oracle, backend, algo):
    """K1 in phases (the multi-GPU row split of one huge pair): partitions run one after the other into separate workspaces, the packed
    bests merged by an unsigned 64-bit max (what the all-reduce does), then select == the one-call path, bit for bit"""
    g = torch.Generator().manual_seed(19)
    nrm = lambda x: torch.nn.functional.normalize(x, dim=-1)
    M, N = 3300, 2900
    src = nrm(torch.randn(M, 32, generator=g)); tgt = nrm(torch.randn(N, 32, generator=g))
    tgt[:1500] = nrm(src[:1500] + 0.05 * torch.randn(1500, 32, generator=g)); tgt[7] = tgt[3]
    sx = torch.randn(M, 3, generator=g); tx = torch.randn(N, 3, generator=g)
    so = torch.tensor([0, M], dtype=torch.int32, device=DEV); to = torch.tensor([0, N], dtype=torch.int32, device=DEV)
    whole = backend.mutual_matching_batched(src.to(DEV), tgt.to(DEV), so, to, M, N, sx.to(DEV), tx.to(DEV))
    for nparts in (2, 3, 8):
        merged = None
        for part in range(nparts):
            sp = backend.MutualNNSplit(src.to(DEV), tgt.to(DEV), so, to, M, N)
            pk = sp.partial(part, nparts).clone() ^ -0x8000000000000000      # unsigned order -> signed order
            merged = pk if merged is None else torch.maximum(merged, pk)
        sp.packed.copy_(merged ^ -0x8000000000000000)
        out = sp.select(sx.to(DEV), tx.to(DEV))
        n = int(out["n_mutual"].item())
        assert n == int(whole["n_mutual"].item())
        for key in ("nn_s", "nn_t"):
            assert torch.equal(out[key], whole[key]), (nparts, key)
        assert torch.equal(out["s_mids"][:n], whole["s_mids"][:n]) and torch.equal(out["corr"][:n], whole["corr"][:n])
    nn_s, nn_t = oracle.mutual_nn(src.numpy(), tgt.numpy())
    assert np.array_equal(whole["nn_s"].cpu().numpy(), nn_s) and np.array_equal(whole["nn_t"].cpu().numpy(), nn_t)


def test_post_refinement_big_pair_cluster_bit_exact(oracle, backend):
    """more than 16384 correspondences: the 8-block reduction tree, run by an 8-CTA cluster (max_count given) or by one CTA - same bits"""
    n = 40000
    g = torch.Generator().manual_seed(6)
    src = torch.rand(n, 3, generator=g) * 3; Rg = S.quat_to_rot(torch.randn(1, 4, generator=g))[0]; tg = torch.rand(3, generator=g)
    tgt = src @ Rg.T + tg + 0.01 * torch.randn(n, 3, generator=g)
    out = torch.rand(n, generator=g) < 0.6; tgt[out] = torch.rand(int(out.sum()), 3, generator=g) * 3
    corr = np.zeros((n, 8), np.float32); corr[:, :3] = src.numpy(); corr[:, 4:7] = tgt.numpy()
    T0 = np.eye(4, dtype=np.float32); T0[:3, :3] = Rg.numpy(); T0[:3, 3] = tg.numpy() + 0.03
    To, it_o, inl_o = oracle.post_refinement(T0, corr, 0.10, 20)
    cd = torch.from_numpy(corr).to(DEV); off = torch.tensor([0, n], dtype=torch.int32, device=DEV); cnt = torch.tensor([n], dtype=torch.int32, device=DEV)
    for mc in (n, 0):                       # cluster kernel / single-CTA kernel
        T, it, inl = backend.post_refinement_batched(torch.from_numpy(T0)[None].to(DEV), cd, off, cnt, 0.10, 20, max_count=mc)
        assert np.array_equal(T[0].cpu().numpy(), To) and int(it.item()) == it_o and int(inl.item()) == inl_o
    # a mixed batch through the cluster kernel: a small pair next to the big one
    corr2 = np.concatenate([corr[:900], corr]); off2 = torch.tensor([0, 900, 900 + n], dtype=torch.int32, device=DEV)
    cnt2 = torch.tensor([900, n], dtype=torch.int32, device=DEV)
    T2, _, _ = backend.post_refinement_batched(torch.from_numpy(np.stack([T0, T0])).to(DEV), torch.from_numpy(corr2).to(DEV), off2, cnt2, 0.10, 20, max_count=n)
    Ts, _, _ = oracle.post_refinement(T0, corr[:900], 0.10, 20)
    assert np.array_equal(T2[0].cpu().numpy(), Ts) and np.array_equal(T2[1].cpu().numpy(), To)


def test_k1_algo_is_per_thread(backend):
    """bfr_config_set is thread-local: a second host thread toggling the algorithm does not disturb this one, and both threads can launch"""
    import threading
    g = torch.Generator().manual_seed(2)
    src = torch.nn.functional.normalize(torch.randn(800, 32, generator=g), dim=-1).to(DEV)
    tgt = torch.nn.functional.normalize(torch.randn(900, 32, generator=g), dim=-1).to(DEV)
    ref = backend.mutual_matching_device(src, tgt)["nn_s"].clone()
    seen, errs = {}, []

    def worker():
        try:
            seen["default"] = backend.get_k1_algo()
            backend.set_k1_algo(backend.K1_FP32)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(20):
                    r = backend.mutual_matching_device(src, tgt)
                st.synchronize()
            seen["same"] = bool(torch.equal(r["nn_s"], ref)); seen["algo"] = backend.get_k1_algo()
        except Exception as e:              # noqa: BLE001
            errs.append(e)

    th = threading.Thread(target=worker); th.start()
    for _ in range(20):
        r = backend.mutual_matching_device(src, tgt)
        assert backend.get_k1_algo() == backend.K1_TENSOR_FILTER
    th.join()
    torch.cuda.synchronize()
    assert not errs and seen == {"default": 1, "same": True, "algo": 0} and torch.equal(r["nn_s"], ref)


def test_register_pipeline_matches_oracle_and_recovers_pose(oracle, backend, algo):
    P, N, H = 6, 1200, 8000
    b = _pairs(P, N, cfg_id=17)
    off = np.arange(P + 1, dtype=np.int32) * N
    To, nm_o, ni_o = oracle.register_batched(b.src_des.reshape(-1, 32).numpy(), b.src_xyz.reshape(-1, 3).numpy(), off,
                                             b.tgt_des.reshape(-1, 32).numpy(), b.tgt_xyz.reshape(-1, 3).numpy(), off, H, 99, 0, 0.1, 0.8, 0.1, 20)
    bd = b.to(DEV)
    T, nm, ni = backend.register_uniform(bd.src_des, bd.src_xyz, bd.tgt_des, bd.tgt_xyz, hypotheses=H, seed=99)
    assert np.array_equal(nm.cpu().numpy(), nm_o) and np.array_equal(ni.cpu().numpy(), ni_o)
    assert np.array_equal(T.cpu().numpy(), To)
    recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt)
    assert recall == 1.0 and float(rte.max()) < 0.01 and float(rre.max()) < 0.5
    # host-buffer entry point (pinned memory, two streams) gives the same poses
    reg = backend.HostRegistrar(4, N, N, DEV, hypotheses=H, seed=99)
    pin = lambda x: x.contiguous().pin_memory()
    Th = torch.empty(P, 4, 4).pin_memory(); nmh = torch.empty(P, dtype=torch.int32).pin_memory()
    reg.run(pin(b.src_des), pin(b.src_xyz), pin(b.tgt_des), pin(b.tgt_xyz), Th, nmh)
    assert np.array_equal(Th.numpy(), To) and np.array_equal(nmh.numpy(), nm_o)


def test_full_size_properties(backend, algo):
    """config-2-sized pairs (5000 keypoints, 50k hypotheses): size-independent properties instead of the slow oracle"""
    P, N = 8, 5000
    b = _pairs(P, N, cfg_id=2).to(DEV)
    r = backend.mutual_matching_batched(b.src_des.reshape(-1, 32), b.tgt_des.reshape(-1, 32),
                                        (torch.arange(P + 1, dtype=torch.int32) * N).to(DEV), (torch.arange(P + 1, dtype=torch.int32) * N).to(DEV), N, N)
    nn_s = r["nn_s"].reshape(P, N); nn_t = r["nn_t"].reshape(P, N)
    assert torch.equal(nn_s, b.perm)                                 # planted permutation recovered
    assert torch.equal(torch.gather(nn_t, 1, nn_s), torch.arange(N, device=DEV).expand(P, N))
    assert int(r["n_mutual"].min()) == N
    sm = r["s_mids"].reshape(P, N)
    assert bool((sm[:, 1:] > sm[:, :-1]).all())                      # ascending
    T, nm, ni = backend.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=50000, seed=1)
    T2, _, ni2 = backend.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=50000, seed=1, ransac_splits=5)
    assert torch.equal(T, T2) and torch.equal(ni, ni2)               # deterministic, independent of the CTA split
    recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt.cpu())
    assert recall == 1.0
    assert float(rte.max()) < 2e-3 and float(S.rotation_error_rad(T[:, :3, :3].cpu(), b.T_gt[:, :3, :3].cpu()).max()) < 2e-3
    assert int(ni.min()) > 0.27 * N                                  # ~30 % planted inliers found by the best hypothesis
    R = T[:, :3, :3].double().cpu()
    assert float((R @ R.transpose(-1, -2) - torch.eye(3, dtype=torch.float64)).abs().max()) < 1e-5
    assert torch.allclose(torch.det(R), torch.ones(P, dtype=torch.float64), atol=1e-5)


def _subset_check(oracle, backend, src, tgt, rows=150):
    """exactness at sizes the oracle cannot scan fully: nearest neighbours of random row subsets against the oracle"""
    g = torch.Generator().manual_seed(3)
    r = backend.mutual_matching_device(src, tgt)
    si = torch.randperm(src.shape[0], generator=g)[:rows]; ti = torch.randperm(tgt.shape[0], generator=g)[:rows]
    nn_s, _ = oracle.mutual_nn(src[si.to(src.device)].cpu().numpy(), tgt.cpu().numpy())
    _, nn_t = oracle.mutual_nn(src.cpu().numpy(), tgt[ti.to(tgt.device)].cpu().numpy())
    assert np.array_equal(r["nn_s"][si.to(src.device)].cpu().numpy(), nn_s)
    assert np.array_equal(r["nn_t"][ti.to(tgt.device)].cpu().numpy(), nn_t)
    return r


def test_config3_low_overlap_many_hypotheses(oracle, backend):
    """BASELINE config 3: ~95 % outliers, 500k hypotheses per pair"""
    c = S.CONFIGS[3]
    b = S.make_pairs(2, 5000, first_pair=0, **{k: v for k, v in c["gen"].items() if k != "num_kpts"}).to(DEV)
    T, nm, ni = backend.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=c["hypotheses"], seed=5)
    recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt.cpu())
    assert recall == 1.0 and float(rte.max()) < 0.02
    assert torch.equal(ni.cpu(), b.inlier.sum(-1).int().cpu()) or int((ni.cpu() - b.inlier.sum(-1).int().cpu()).abs().max()) <= 3
    # exactness of the RANSAC winner on one pair against the oracle (H = 500k takes ~1 s on the CPU)
    s, t = oracle.mutual_matching(b.src_des[0].cpu().numpy(), b.tgt_des[0].cpu().numpy())
    corr = oracle.gather_corr(b.src_xyz[0].cpu().numpy(), b.tgt_xyz[0].cpu().numpy(), s, t)
    best = oracle.ransac(corr, 5, 0, c["hypotheses"], c["dist_th"], c["similar_th"])
    assert (best >> 32) == int(ni[0])


def test_config4_kitti_scale(oracle, backend):
    """BASELINE config 4: 20k keypoints, 0.6 m inlier threshold, KITTI-sized scene"""
    c = S.CONFIGS[4]
    b = S.make_pairs(2, 20000, first_pair=0, **{k: v for k, v in c["gen"].items() if k != "num_kpts"}).to(DEV)
    _subset_check(oracle, backend, b.src_des[0], b.tgt_des[0])
    T, nm, ni = backend.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=c["hypotheses"], dist_th=c["dist_th"],
                                         similar_th=c["similar_th"], refine_thr=c["refine_thr"], seed=2)
    recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt.cpu(), rte_thresh=0.3, rre_thresh_deg=1.0)   # KITTI/test.py:66-67
    assert recall == 1.0 and int(nm.min()) == 20000


def test_config5_huge_pair_split_by_hypothesis(oracle, backend):
    """BASELINE config 5: one 100k x 100k pair; hypotheses evaluated in 4 slices (as 4 ranks would) and max-merged"""
    b = S.make_pairs(1, 100000, cfg_id=5).to(DEV)
    r = _subset_check(oracle, backend, b.src_des[0], b.tgt_des[0], rows=60)
    assert int(r["n_mutual"].item()) == 100000
    rm = backend.mutual_matching_device(b.src_des[0], b.tgt_des[0], b.src_xyz[0], b.tgt_xyz[0])
    K = int(rm["n_mutual"].item())
    off = torch.tensor([0, K], dtype=torch.int32, device=DEV); cnt = rm["n_mutual"]
    H = 50000
    whole = backend.ransac_batched(rm["corr"], off, cnt, H, 0.1, 0.8, seed=7)
    parts = None
    for rnk in range(4):
        h0, h1 = rnk * H // 4, (rnk + 1) * H // 4
        p = backend.ransac_batched(rm["corr"], off, cnt, H, 0.1, 0.8, seed=7, h_begin=h0, h_end=h1)
        parts = p if parts is None else torch.maximum(parts, p)      # what all_reduce(MAX) does across ranks
    assert torch.equal(parts, whole)
    T, inl, bh = backend.ransac_finalize_batched(rm["corr"], off, cnt, whole, 0.1, 0.8, seed=7)
    Tr, it, _ = backend.post_refinement_batched(T, rm["corr"], off, cnt, 0.1)
    recall, rte, rre = S.registration_recall(Tr.cpu(), b.T_gt.cpu())
    assert recall == 1.0 and int(inl) > 29000 and float(rte.max()) < 1e-3


@pytest.mark.gpu
def test_tensor_filter_equals_fp32_kernel_on_stress_inputs(backend):
    """the tensor-core filter path against the all-FP32 kernel (itself bit-exact with the oracle above) on inputs built to stress the exact
    re-check: clusters of near-duplicates (several in-band groups per row -> balanced passes), a dense blob (candidate-list overflow ->
    warp-cooperative scan), unrelated rows, un-normalised rows with duplicates and zeros, ragged sizes, column splits (N > 6144)"""
    g = torch.Generator().manual_seed(4242)
    nrm = lambda x: torch.nn.functional.normalize(x, dim=-1)
    r = lambda *s: torch.randn(*s, generator=g)

    def make(kind, M, N):
        if kind == "unrelated":
            return nrm(r(M, 32)), nrm(r(N, 32))
        if kind == "clusters":
            nc = max(1, N // 9); centers = nrm(r(nc, 32))
            tgt = nrm(centers[torch.randint(0, nc, (N,), generator=g)] + (10.0 ** torch.empty(N, 1).uniform_(-5, -2, generator=g)) * r(N, 32))
            return nrm(centers[torch.randint(0, nc, (M,), generator=g)] + 0.02 * r(M, 32)), tgt
        if kind == "flood":
            c = nrm(r(1, 32))
            return nrm(c + 1e-3 * r(M, 32)), torch.cat([nrm(c + 1e-4 * r(N // 2, 32)), nrm(r(N - N // 2, 32))])
        src = r(M, 32) * 10.0 ** torch.empty(M, 1).uniform_(-2, 2, generator=g)          # "scaled"
        tgt = r(N, 32) * 10.0 ** torch.empty(N, 1).uniform_(-2, 2, generator=g)
        nd = min(tgt[::11].shape[0], tgt[1::11].shape[0]); tgt[::11][:nd] = tgt[1::11][:nd]; src[::13] = 0.0
        return src, tgt

    cases = [("unrelated", 5000, 5000), ("clusters", 3100, 2900), ("flood", 1500, 2600), ("scaled", 2222, 3333), ("clusters", 700, 9000),
             ("unrelated", 257, 12500), ("flood", 300, 700), ("clusters", 4999, 513)]
    try:
        for kind, M, N in cases:
            src, tgt = make(kind, M, N)
            src, tgt = src.to(DEV), tgt.to(DEV)
            out = {}
            for algo in (backend.K1_FP32, backend.K1_TENSOR_FILTER):
                backend.set_k1_algo(algo)
                out[algo] = backend.mutual_matching_device(src, tgt, want_dist=True)
            for key in ("nn_s", "nn_t", "dist_s", "dist_t"):
                assert torch.equal(out[backend.K1_FP32][key], out[backend.K1_TENSOR_FILTER][key]), (kind, M, N, key)
    finally:
        backend.set_k1_algo(backend.K1_TENSOR_FILTER)


def _ransac_both_ways(backend, corr, H, dist_th=0.1, sim_th=0.8, splits=1, seed=3):
    K = corr.shape[0]
    off = torch.tensor([0, K], dtype=torch.int32, device=DEV); cnt = torch.tensor([K], dtype=torch.int32, device=DEV)
    out = []
    try:
        for algo in (backend.RANSAC_TENSOR_FILTER, backend.RANSAC_FP32):
            backend.set_ransac_scoring(algo)
            nv = torch.zeros(1, dtype=torch.int32, device=DEV)
            bp = backend.ransac_batched(corr.to(DEV), off, cnt, H, dist_th, sim_th, seed=seed, pair_id_base=11, splits=splits, valid_count=nv)
            out.append((int(bp.item()), int(nv.item())))
    finally:
        backend.set_ransac_scoring(backend.RANSAC_TENSOR_FILTER)
    return out


@pytest.mark.parametrize("K", [3, 100, 127, 128, 129, 1000, 5119, 5120, 5121])
def test_ransac_tensor_filter_equals_fp32_scoring(oracle, backend, K):
    """RANSAC scores hypotheses either with the exact FP32 loop or with the tcgen05 residual filter (2-level f16 operand splits) + exact
    re-check of the residuals inside the error band: both must return the same packed best and the same number of valid hypotheses, and
    equal the oracle.  K straddles the 128-correspondence A tiles and the 5120-correspondence residency limit (beyond: FP32 only)."""
    g = torch.Generator().manual_seed(K)
    s = torch.rand(K, 3, generator=g) * 3 - 1.5
    q = s + 0.03 * torch.randn(K, 3, generator=g)
    out = torch.rand(K, generator=g) < 0.5
    q[out] = torch.rand(int(out.sum()), 3, generator=g) * 3 - 1.5
    rec = torch.zeros(K, 8); rec[:, :3] = s; rec[:, 4:7] = q
    (b1, n1), (b0, n0) = _ransac_both_ways(backend, rec, 6000)
    assert (b1, n1) == (b0, n0)
    assert b1 == oracle.ransac(rec.numpy(), 3, 11, 6000, 0.1, 0.8)


def test_ransac_tensor_filter_adversarial(oracle, backend):
    """inputs built to sit on the filter's decision boundary or outside its operating range"""
    g = torch.Generator().manual_seed(77)
    K = 2000
    s = torch.rand(K, 3, generator=g) * 2 - 1
    q = s.clone(); q[:, 0] += 0.1                                     # residual == threshold under the identity ...
    q[::3, 0] = torch.nextafter(q[::3, 0], torch.tensor(10.0)); q[1::3, 0] = torch.nextafter(q[1::3, 0], torch.tensor(-10.0))    # ... +- 1 ulp
    q[:600] = s[:600]
    rec = torch.zeros(K, 8); rec[:, :3] = s; rec[:, 4:7] = q
    cases = [("on the threshold", rec, 0.1)]
    far = rec.clone(); far[:, :3] += 20000.0; far[:, 4:7] += 20000.0  # beyond the f16 split range: the pair is scored in FP32
    cases.append(("coordinates 2e4", far, 0.1))
    big = rec.clone(); big[:, :3] *= 100.0; big[:, 4:7] *= 100.0
    cases.append(("coordinates 100, threshold 10", big, 10.0))
    tiny = rec.clone(); tiny[:, :3] *= 1e-3; tiny[:, 4:7] *= 1e-3
    cases.append(("coordinates 1e-3, threshold 1e-4", tiny, 1e-4))
    nan = rec.clone(); nan[17, 5] = float("nan")
    cases.append(("one NaN coordinate", nan, 0.1))
    for name, r, thr in cases:
        for splits in (1, 3):
            (b1, n1), (b0, n0) = _ransac_both_ways(backend, r, 5000, dist_th=thr, splits=splits)
            assert (b1, n1) == (b0, n0), name
        if name != "one NaN coordinate":
            assert b1 == oracle.ransac(r.numpy(), 3, 11, 5000, thr, 0.8), name


def test_ransac_tensor_filter_full_size_pairs(backend):
    """config-2-sized pairs (5000 correspondences, 50000 hypotheses): the two scoring paths agree on every pair of a batch"""
    b = _pairs(6, 5000, cfg_id=2)
    P, N = 6, 5000
    off = (torch.arange(P + 1, dtype=torch.int32) * N).to(DEV)
    rm = backend.mutual_matching_batched(b.src_des.reshape(P * N, 32).to(DEV), b.tgt_des.reshape(P * N, 32).to(DEV), off, off, N, N,
                                         b.src_xyz.reshape(P * N, 3).to(DEV), b.tgt_xyz.reshape(P * N, 3).to(DEV), want_nn=False, want_mids=False)
    res = []
    try:
        for algo in (backend.RANSAC_TENSOR_FILTER, backend.RANSAC_FP32):
            backend.set_ransac_scoring(algo)
            nv = torch.zeros(P, dtype=torch.int32, device=DEV)
            bp = backend.ransac_batched(rm["corr"], off, rm["n_mutual"], 50000, 0.1, 0.8, seed=1, pair_id_base=0, splits=2, valid_count=nv)
            res.append((bp.clone(), nv.clone()))
    finally:
        backend.set_ransac_scoring(backend.RANSAC_TENSOR_FILTER)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert int(res[0][0].min().item()) >> 32 > 1000


def test_full_size_batch_equals_cpu_port(oracle, backend):
    """BASELINE config 2 at full per-pair size (5 000 x 5 000 keypoints, 50 000 hypotheses, 20 refinement rounds): a batch of the bench
    workload's pairs through the one-call back end equals the CPU port pose for pose, bit for bit (the check bench.py makes on 256 pairs)"""
    c = S.CONFIGS[2]; P = 48; N = c["gen"]["num_kpts"]
    b = S.make_pairs(P, first_pair=0, **c["gen"])
    off = np.arange(P + 1, dtype=np.int32) * N
    oracle.set_num_threads(os.cpu_count() or 1)
    To, nm_o, ni_o = oracle.register_batched(b.src_des.reshape(-1, 32).numpy(), b.src_xyz.reshape(-1, 3).numpy(), off, b.tgt_des.reshape(-1, 32).numpy(),
                                             b.tgt_xyz.reshape(-1, 3).numpy(), off, c["hypotheses"], 0, 0, c["dist_th"], c["similar_th"], c["refine_thr"], 20)
    bd = b.to(DEV)
    for scoring in (backend.RANSAC_TENSOR_FILTER, backend.RANSAC_FP32):
        backend.set_ransac_scoring(scoring)
        try:
            T, nm, ni = backend.register_uniform(bd.src_des, bd.src_xyz, bd.tgt_des, bd.tgt_xyz, hypotheses=c["hypotheses"], dist_th=c["dist_th"],
                                                 similar_th=c["similar_th"], refine_thr=c["refine_thr"], seed=0)
        finally:
            backend.set_ransac_scoring(backend.RANSAC_TENSOR_FILTER)
        assert np.array_equal(nm.cpu().numpy(), nm_o) and np.array_equal(ni.cpu().numpy(), ni_o)
        assert np.array_equal(T.cpu().numpy(), To)
    recall, rte, rre = S.registration_recall(T.cpu(), b.T_gt)
    assert recall == 1.0
