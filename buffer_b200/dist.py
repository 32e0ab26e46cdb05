"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the little communication there is.

The path shards by fragment pair (the reference processes one pair per forward, batch size 1:
ThreeDMatch/config.py:21, dataloader.py:119), so the data path has NO collective: rank r owns a contiguous block of
pairs and results are collected with one all_gather of 18 numbers per pair (NCCL over NVLink; latency-bound).

One huge pair (BASELINE config 5, 100k x 100k keypoints) is split by HYPOTHESIS instead: every rank evaluates
h in [r*H/W, (r+1)*H/W) of the same Philox stream and the 8-byte packed best (count << 32 | ~h) is max-all-reduced;
because the RNG is counter-based, every rank then regenerates the winning fit locally — no broadcast of R, t.

The compute callables are injectable so the host logic is testable on CPU with the gloo backend (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_range(num_items, rank, world):
    """contiguous, balanced [start, end) of `num_items` for `rank` (first num_items % world ranks get one extra)"""
    base, extra = divmod(num_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def hypothesis_range(num_hypotheses, rank, world):
    return shard_range(num_hypotheses, rank, world)


def gather_pair_results(T_local, n_mutual_local, n_inliers_local, num_pairs, group=None):
    """all_gather the per-pair results of every rank's shard -> full [P,4,4], [P], [P] on every rank (pair order)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(num_pairs, r, world) for r in range(world)]
    cap = max(e - s for s, e in sizes)
    dev = T_local.device
    pack = torch.zeros(cap, 18, dtype=torch.float32, device=dev)
    n = T_local.shape[0]
    assert n == sizes[rank][1] - sizes[rank][0]
    pack[:n, :16] = T_local.reshape(n, 16)
    pack[:n, 16] = n_mutual_local.float()
    pack[:n, 17] = n_inliers_local.float()
    out = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(out, pack, group=group)
    parts = [o[: e - s] for o, (s, e) in zip(out, sizes)]
    full = torch.cat(parts, 0)
    return full[:, :16].reshape(-1, 4, 4), full[:, 16].round().int(), full[:, 17].round().int()


def register_sharded(src_des, src_xyz, tgt_des, tgt_xyz, register_fn=None, group=None, gather=True, **kw):
    """Uniform batch [P,N,...] present on every rank (or at least this rank's shard valid): each rank registers its
    contiguous shard with pair ids = global pair index, then (optionally) all ranks gather all poses."""
    if register_fn is None:
        from . import backend
        register_fn = backend.register_uniform
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    P = src_des.shape[0]
    s, e = shard_range(P, rank, world)
    T, nm, ni = register_fn(src_des[s:e], src_xyz[s:e], tgt_des[s:e], tgt_xyz[s:e], pair_id_base=s, **kw)
    if not gather or world == 1:
        return T, nm, ni
    return gather_pair_results(T, nm, ni, P, group)


def ransac_split_hypotheses(corr, corr_off, corr_cnt, hypotheses, dist_th, similar_th, seed=0, pair_id_base=0,
                            ransac_fn=None, finalize_fn=None, group=None):
    """Split-hypothesis RANSAC for pairs replicated on every rank: local max over this rank's hypothesis range, one
    all_reduce(MAX) of the packed int64 per pair, local regeneration of the winner.  -> T [P,4,4], inliers, best_h"""
    if ransac_fn is None or finalize_fn is None:
        from . import backend
        ransac_fn = ransac_fn or backend.ransac_batched
        finalize_fn = finalize_fn or backend.ransac_finalize_batched
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    h0, h1 = hypothesis_range(hypotheses, rank, world)
    best = ransac_fn(corr, corr_off, corr_cnt, hypotheses, dist_th, similar_th, seed=seed, pair_id_base=pair_id_base, h_begin=h0, h_end=h1)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MAX, group=group)     # packed value < 2^63: signed max == unsigned max
    return finalize_fn(corr, corr_off, corr_cnt, best, dist_th, similar_th, seed=seed, pair_id_base=pair_id_base)
