#!/usr/bin/env python3
"""CPU feasibility check for the round-2 plan: residual components x_i = u_(h,i) . v_c with hi/lo f16-split operands and FP32
accumulation, against the oracle's FP32 FMA chain.  Reports the error of d^2 and the fraction of (hypothesis, correspondence) pairs that would
fall inside the ambiguity band and need the exact re-check.  numpy only; synthetic pair of BASELINE config 2 / config 4 scale."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from buffer_b200 import synthetic as S


def split_f16(x):
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    return hi, lo


def run(name, gen, thr, nh=64):
    b = S.make_pairs(1, 5000, cfg_id=2, **gen) if gen else S.make_pairs(1, 5000, cfg_id=2)
    s = b.src_xyz[0].numpy().astype(np.float32); q = b.tgt_xyz[0][b.perm[0]].numpy().astype(np.float32) if hasattr(b, "perm") else None
    T = b.T_gt[0].numpy().astype(np.float64)
    # correspondences (s_c, q_c): planted matches i -> perm[i]
    perm = b.perm[0].numpy(); q = b.tgt_xyz[0].numpy().astype(np.float32)[perm]
    rng = np.random.RandomState(0)
    amb_frac, derr = [], []
    for h in range(nh):                                   # good hypotheses: ground truth perturbed like a 3-point fit (~1 cm / 0.3 deg)
        w = rng.randn(3) * 0.005; th = np.linalg.norm(w); K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        Rp = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        R = (Rp @ T[:3, :3]).astype(np.float32); t = (T[:3, 3] + rng.randn(3) * 0.01).astype(np.float32)
        # oracle chain (float32 fma emulated in float64 then rounded per op is overkill here: numpy float32 ops, no contraction)
        x32 = np.zeros((len(s), 3), np.float32)
        for i in range(3):
            acc = np.float32(t[i]) + R[i, 2] * s[:, 2]
            acc = acc + R[i, 1] * s[:, 1]
            acc = acc + R[i, 0] * s[:, 0]
            x32[:, i] = acc - q[:, i]
        d2_32 = (x32[:, 0] * x32[:, 0] + (x32[:, 1] * x32[:, 1] + x32[:, 2] * x32[:, 2])).astype(np.float32)
        # tensor-core emulation: u = (R_i0, R_i1, R_i2, t_i, -1), v = (s_x, s_y, s_z, 1, q_i): hi*hi + hi*lo + lo*hi, FP32 accumulation
        xtc = np.zeros((len(s), 3), np.float32)
        for i in range(3):
            u = np.array([R[i, 0], R[i, 1], R[i, 2], t[i], -1.0], np.float32)
            v = np.stack([s[:, 0], s[:, 1], s[:, 2], np.ones(len(s), np.float32), q[:, i]], 1)
            uh, ul = split_f16(u); vh, vl = split_f16(v)
            acc = np.zeros(len(s), np.float32)
            for k in range(5):
                acc = acc + uh[k] * vh[:, k]; acc = acc + uh[k] * vl[:, k]; acc = acc + ul[k] * vh[:, k]
            xtc[:, i] = acc
        d2_tc = (xtc ** 2).sum(1).astype(np.float32)
        err = np.abs(d2_tc.astype(np.float64) - d2_32.astype(np.float64))
        near = np.abs(d2_32 - thr * thr) < 4e-6                     # a band 16x the largest error seen
        derr.append(err[np.sqrt(d2_32) < 3 * thr].max()); amb_frac.append(near.mean())
    print("%-8s thr %.2f m: max |d2_tc - d2_fp32| near the threshold = %.2e (thr^2 = %.3g); pairs with |d2 - thr^2| < 4e-6 (would be re-checked exactly): %.2e of all (h,c)"
          % (name, thr, max(derr), thr * thr, float(np.mean(amb_frac))))


run("3DMatch", {}, 0.10)
