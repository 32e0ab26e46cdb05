#!/usr/bin/env python3
"""RANSAC-only timing on the bench workload: python tools/ransac_bench.py [pairs] [splits ...]  -> ms per launch for each split count"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buffer_b200 import _lib
if os.environ.get("BFR_SO"): _lib.SO_PATH = os.path.abspath(os.environ["BFR_SO"])      # experiment variants (tools/build_variant.sh)
from buffer_b200 import backend as B, synthetic as S
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1623
splits = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 4]
if os.environ.get("BFR_RANSAC_FP32"): B.set_ransac_scoring(0)
c = S.CONFIGS[int(os.environ.get("BFR_CFG", "2"))]; N = c["gen"]["num_kpts"]; dev = "cuda:0"
parts = [S.make_pairs(min(128, P - p0), first_pair=p0, device=dev, **c["gen"]) for p0 in range(0, P, 128)]
cat = lambda f: torch.cat([getattr(b, f) for b in parts], 0)
src_des, tgt_des = cat("src_des").reshape(P * N, 32), cat("tgt_des").reshape(P * N, 32)
src_xyz, tgt_xyz = cat("src_xyz").reshape(P * N, 3), cat("tgt_xyz").reshape(P * N, 3)
off = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev)
rm = B.mutual_matching_batched(src_des, tgt_des, off, off, N, N, src_xyz, tgt_xyz, want_nn=False, want_mids=False)
ref = None
for s in splits:
    best = None; ms = []
    for i in range(4):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        out = B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=0, splits=s)
        eb.record(); eb.synchronize(); ms.append(ea.elapsed_time(eb))
    T = out[0] if isinstance(out, (tuple, list)) else out
    same = True if ref is None else bool(torch.equal(T, ref))
    if ref is None: ref = T.clone()
    print("splits %d: %.3f ms (min of 3 warm runs), identical to splits[0]: %s" % (s, min(ms[1:]), same))
    L = _lib.lib()
    if hasattr(L, "bfr_dbg_ransac_counters"):       # -DRS_TIMING variant: cycles summed over CTAs (thread 0's clock), 4 launches
        import ctypes, numpy as np
        o = np.zeros(8, dtype=np.uint64); L.bfr_dbg_ransac_counters(ctypes.c_void_p(o.ctypes.data))
        tot, sc, fit, n, nsc = [float(x) for x in o[:5]]
        print("  per CTA: total %.0f cycles = stage1 %.0f + fit %.0f + score %.0f; scored hypotheses/CTA %.1f" % (tot / n, (tot - fit) / n, (fit - sc) / n, sc / n, nsc / n))
        if hasattr(L, "bfr_dbg_ransac_tc_counters"):
            o = np.zeros(8, dtype=np.uint64); L.bfr_dbg_ransac_tc_counters(ctypes.c_void_p(o.ctypes.data))
            b, loop, r, nf, nt = float(o[0]), float(o[1]), float(o[4]), float(o[5]), float(o[6])
            if nf: print("  tensor flush (thread 0): %.0f flushes/CTA, %.1f A tiles/flush; cycles per flush: B operands %.0f, tile loop %.0f (%.0f per tile), count reduction %.0f" % (nf / n, nt / nf, b / nf, loop / nf, loop / nt, r / nf))
