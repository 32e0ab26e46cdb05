#!/usr/bin/env python3
"""Multi-GPU check (NCCL): run under torchrun with N >= 2 ranks on one node.
  1. pairs sharded by rank (buffer_b200.dist.register_sharded, no data-path collective, one all_gather of the results) == all pairs on one GPU
  2. one big pair split by hypothesis (ransac_split_hypotheses: local max, one all_reduce(MAX) of 8 bytes, winner regenerated locally)
     == the full hypothesis range on one GPU
Prints one line per check on rank 0; exit code 1 on any mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from buffer_b200 import backend as B, dist as D, synthetic as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True

# 1. sharded pairs
P, N, H = 37, 1500, 20000
b = S.make_pairs(P, N, cfg_id=7).to(dev)
kw = dict(hypotheses=H, seed=3)
T, nm, ni = D.register_sharded(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, **kw)
T1, nm1, ni1 = B.register_uniform(b.src_des, b.src_xyz, b.tgt_des, b.tgt_xyz, **kw)
same = bool(torch.equal(T, T1) and torch.equal(nm, nm1.int()) and torch.equal(ni, ni1.int()))
flag = torch.tensor([int(same)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("sharded pairs (%d pairs over %d ranks) == single GPU: %s" % (P, world, bool(flag.item())))
ok &= bool(flag.item())

# 2. one big pair, hypotheses split across ranks
N2, H2 = 20000, 200000
c = S.make_pairs(1, N2, cfg_id=8).to(dev)
off = torch.tensor([0, N2], dtype=torch.int32, device=dev)
rm = B.mutual_matching_batched(c.src_des.reshape(N2, 32), c.tgt_des.reshape(N2, 32), off, off, N2, N2, c.src_xyz.reshape(N2, 3), c.tgt_xyz.reshape(N2, 3),
                               want_nn=False, want_mids=False)
Ts, inl_s, bh_s = D.ransac_split_hypotheses(rm["corr"], off, rm["n_mutual"], H2, 0.1, 0.8, seed=5)
best = B.ransac_batched(rm["corr"], off, rm["n_mutual"], H2, 0.1, 0.8, seed=5)
Tf, inl_f, bh_f = B.ransac_finalize_batched(rm["corr"], off, rm["n_mutual"], best, 0.1, 0.8, seed=5)
same = bool(torch.equal(Ts, Tf) and torch.equal(inl_s, inl_f) and torch.equal(bh_s, bh_f))
flag = torch.tensor([int(same)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("one %d x %d pair, %d hypotheses split over %d ranks (all_reduce MAX of the packed best) == single GPU: %s (inliers %d, winning hypothesis %d)"
                    % (N2, N2, H2, world, bool(flag.item()), int(inl_f.item()), int(bh_f.item())))
ok &= bool(flag.item())
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
