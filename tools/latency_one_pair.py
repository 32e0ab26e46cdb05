#!/usr/bin/env python3
"""Latency of ONE fragment pair through the whole back end (the reference's operating point: batch size 1), device-resident inputs.
usage: python tools/latency_one_pair.py [keypoints] [hypotheses]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buffer_b200 import backend as B, synthetic as S
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
b = S.make_pairs(1, N, cfg_id=2).to("cuda:0")
for _ in range(5): T, nm, ni = B.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=H, seed=0)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, c in ev:
    a.record(); T, nm, ni = B.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, hypotheses=H, seed=0); c.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(c) for a, c in ev)
rec, rte, rre = S.registration_recall(T.cpu(), b.T_gt.cpu())
print("one pair, %d x %d keypoints, %d hypotheses: median %.3f ms, min %.3f ms (device time, 7 kernels); mutual %d, inliers %d, recall %.0f" % (N, N, H, ms[len(ms) // 2], ms[0], int(nm), int(ni), rec))
