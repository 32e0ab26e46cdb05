#!/usr/bin/env python3
"""Randomised K1 cross-check on the GPU: the tensor-core filter path (algo 1) against the all-FP32 kernel (algo 0, bit-exact with the
oracle) on descriptor sets built to stress the re-check: clusters of near-duplicates (many in-band groups per row, list overflow -> warp
scan), planted matches, unrelated rows, mixed norms, ragged sizes, column splits.  usage: python tools/k1_stress.py [rounds] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buffer_b200 import backend as B

DEV = "cuda:0"
nrm = lambda x: torch.nn.functional.normalize(x, dim=-1)


def make(g, M, N, kind):
    r = lambda *s: torch.randn(*s, generator=g)
    if kind == "unrelated":
        return nrm(r(M, 32)), nrm(r(N, 32))
    if kind == "planted":
        src = nrm(r(M, 32)); k = min(M, N)
        tgt = nrm(r(N, 32)); perm = torch.randperm(N, generator=g)[:k]
        tgt[perm] = nrm(src[:k] + 0.05 * r(k, 32))
        return src, tgt
    if kind == "clusters":                      # targets in clusters of 2..24 near-duplicates: several groups inside every row's band
        nc = max(1, N // 9); centers = nrm(r(nc, 32))
        idx = torch.randint(0, nc, (N,), generator=g)
        tgt = nrm(centers[idx] + (10.0 ** torch.empty(N, 1).uniform_(-5, -2, generator=g)) * r(N, 32))
        src = nrm(centers[torch.randint(0, nc, (M,), generator=g)] + 0.02 * r(M, 32))
        return src, tgt
    if kind == "flood":                         # one dense blob: every row overflows its candidate list
        c = nrm(r(1, 32))
        return nrm(c + 1e-3 * r(M, 32)), torch.cat([nrm(c + 1e-4 * r(N // 2, 32)), nrm(r(N - N // 2, 32))])
    if kind == "scaled":                        # un-normalised, norms over four decades, some exact duplicates and zero rows
        src = r(M, 32) * 10.0 ** torch.empty(M, 1).uniform_(-2, 2, generator=g)
        tgt = r(N, 32) * 10.0 ** torch.empty(N, 1).uniform_(-2, 2, generator=g)
        nd = min(tgt[::11].shape[0], tgt[1::11].shape[0]); tgt[::11][:nd] = tgt[1::11][:nd]; src[::13] = 0.0
        return src, tgt
    raise ValueError(kind)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    g = torch.Generator().manual_seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
    kinds = ["unrelated", "planted", "clusters", "flood", "scaled"]
    bad = 0
    for it in range(rounds):
        kind = kinds[it % len(kinds)]
        M = int(torch.randint(1, 7000, (1,), generator=g)); N = int(torch.randint(1, 7000, (1,), generator=g))
        if it % 7 == 3: N = int(torch.randint(7000, 15000, (1,), generator=g))          # needs column splits in the tensor-core kernel
        src, tgt = make(g, M, N, kind)
        src, tgt = src.to(DEV), tgt.to(DEV)
        out = {}
        for algo in (B.K1_FP32, B.K1_TENSOR_FILTER):
            B.set_k1_algo(algo)
            out[algo] = B.mutual_matching_device(src, tgt, want_dist=True)
        a, b = out[B.K1_FP32], out[B.K1_TENSOR_FILTER]
        same = all(torch.equal(a[k], b[k]) for k in ("nn_s", "nn_t", "dist_s", "dist_t"))
        print("%-9s M=%5d N=%5d %s" % (kind, M, N, "ok" if same else "MISMATCH"), flush=True)
        bad += 0 if same else 1
    B.set_k1_algo(B.K1_TENSOR_FILTER)
    print("mismatches: %d of %d" % (bad, rounds))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
