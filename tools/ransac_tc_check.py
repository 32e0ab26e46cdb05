#!/usr/bin/env python3
"""Tensor-core RANSAC scoring filter vs the exact FP32 loop: equality of the packed bests / valid counts and timing.
python tools/ransac_tc_check.py [pairs]   (BFR_CFG selects the synthetic config, default 2)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buffer_b200 import _lib
if os.environ.get("BFR_SO"): _lib.SO_PATH = os.path.abspath(os.environ["BFR_SO"])
from buffer_b200 import backend as B, synthetic as S

dev = "cuda:0"


def run(corr, off, cnt, H, dist_th, sim_th, algo, splits=1, reps=3, seed=0):
    B.set_ransac_scoring(algo)
    ms = []
    for _ in range(reps):
        nv = torch.zeros(cnt.numel(), dtype=torch.int32, device=dev)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        bp = B.ransac_batched(corr, off, cnt, H, dist_th, sim_th, seed=seed, pair_id_base=3, splits=splits, valid_count=nv)
        eb.record(); eb.synchronize(); ms.append(ea.elapsed_time(eb))
    B.set_ransac_scoring(1)
    return bp.clone(), nv.clone(), min(ms)


def compare(name, corr, off, cnt, H, dist_th=0.1, sim_th=0.8, splits=1):
    b1, n1, t1 = run(corr, off, cnt, H, dist_th, sim_th, 1, splits)
    b0, n0, t0 = run(corr, off, cnt, H, dist_th, sim_th, 0, splits)
    same = bool(torch.equal(b0, b1)) and bool(torch.equal(n0, n1))
    print("%-44s tensor %.3f ms  fp32 %.3f ms  identical %s  (best count of pair 0: %d, valid %d)" % (name, t1, t0, same, int(b1[0].item()) >> 32, int(n1[0].item())), flush=True)
    if not same:
        bad = (b0 != b1).nonzero().flatten()[:8].tolist()
        print("   MISMATCH pairs", bad, [(int(b0[i].item()) >> 32, int(b1[i].item()) >> 32) for i in bad], flush=True)
    return same


def workload(P, cfg):
    c = S.CONFIGS[cfg]; N = c["gen"]["num_kpts"]
    parts = [S.make_pairs(min(128, P - p0), first_pair=p0, device=dev, **c["gen"]) for p0 in range(0, P, 128)]
    cat = lambda f: torch.cat([getattr(b, f) for b in parts], 0)
    off = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev)
    rm = B.mutual_matching_batched(cat("src_des").reshape(P * N, 32), cat("tgt_des").reshape(P * N, 32), off, off, N, N,
                                   cat("src_xyz").reshape(P * N, 3), cat("tgt_xyz").reshape(P * N, 3), want_nn=False, want_mids=False)
    return rm["corr"], off, rm["n_mutual"], c


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    cfg = int(os.environ.get("BFR_CFG", "2"))
    ok = True
    # small cases first (a hang or a fault shows up before the big launch)
    corr, off, cnt, c = workload(4, cfg)
    ok &= compare("4 pairs, 2000 hypotheses", corr, off, cnt, 2000, c["dist_th"], c["similar_th"])
    ok &= compare("4 pairs, 50000 hypotheses, 5 splits", corr, off, cnt, 50000, c["dist_th"], c["similar_th"], splits=5)
    g = torch.Generator().manual_seed(5)
    for K in (3, 100, 127, 128, 129, 1000, 5119, 5120, 5121):
        s = torch.rand(K, 3, generator=g) * 3 - 1.5
        q = s + 0.03 * torch.randn(K, 3, generator=g)
        out = torch.rand(K, generator=g) < 0.5
        q[out] = torch.rand(int(out.sum()), 3, generator=g) * 3 - 1.5
        rec = torch.zeros(K, 8); rec[:, :3] = s; rec[:, 4:7] = q
        o = torch.tensor([0, K], dtype=torch.int32, device=dev); n = torch.tensor([K], dtype=torch.int32, device=dev)
        ok &= compare("K = %d, identity motion, 50 %% outliers" % K, rec.to(dev), o, n, 6000)
    # residuals exactly on / next to the threshold: target = source + (thr, 0, 0) (+- 1 ulp) under the identity
    K = 2000
    s = torch.rand(K, 3, generator=g) * 2 - 1
    q = s.clone(); q[:, 0] += 0.1
    q[::3, 0] = torch.nextafter(q[::3, 0], torch.tensor(10.0)); q[1::3, 0] = torch.nextafter(q[1::3, 0], torch.tensor(-10.0))
    q[:600] = s[:600]                                                 # exact inliers so that good hypotheses exist
    rec = torch.zeros(K, 8); rec[:, :3] = s; rec[:, 4:7] = q
    o = torch.tensor([0, K], dtype=torch.int32, device=dev); n = torch.tensor([K], dtype=torch.int32, device=dev)
    ok &= compare("residuals on the threshold (+- 1 ulp)", rec.to(dev), o, n, 8000)
    # coordinates beyond the f16 split range -> the pair falls back to exact scoring
    rec2 = rec.clone(); rec2[:, :3] += 20000.0; rec2[:, 4:7] += 20000.0
    ok &= compare("coordinates ~ 2e4 (exact fallback)", rec2.to(dev), o, n, 8000)
    rec3 = rec.clone(); rec3[:, :3] *= 100.0; rec3[:, 4:7] *= 100.0
    ok &= compare("coordinates ~ 100, threshold 10", rec3.to(dev), o, n, 8000, dist_th=10.0)
    corr, off, cnt, c = workload(P, cfg)
    ok &= compare("%d pairs of config %d" % (P, cfg), corr, off, cnt, c["hypotheses"], c["dist_th"], c["similar_th"])
    print("ALL IDENTICAL" if ok else "MISMATCH", flush=True)
    sys.exit(0 if ok else 1)
