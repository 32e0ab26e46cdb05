"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the little communication there is.

The path shards by fragment pair (the reference processes one pair per forward, batch size 1:
ThreeDMatch/config.py:21, dataloader.py:119), so the data path has NO collective: rank r owns a contiguous block of
pairs and results are collected with one all_gather of 18 numbers per pair (NCCL over NVLink; latency-bound).

One huge pair (BASELINE config 5, 100k x 100k keypoints) replicated on every rank is split twice (HugePairSplit):
  * matching by ROW BLOCKS - 95 % of its work: rank r runs the fused distance + arg-max kernel on its 1/W share of the source row
    blocks (src->tgt direction) and of the target row blocks (tgt->src direction); the packed bests (key << 32 | ~index, 8 bytes per
    row of either side: 1.6 MB at 100k x 100k) are max-all-reduced over NVLink (NCCL), then every rank decodes / compacts them;
  * RANSAC by HYPOTHESIS: every rank evaluates h in [r*H/W, (r+1)*H/W) of the same Philox stream and the 8-byte packed best
    (count << 32 | ~h) is max-all-reduced; because the RNG is counter-based, every rank then regenerates the winning fit locally -
    no broadcast of R, t.  The refinement (<2 % of the work) is replicated.

The compute callables are injectable so the host logic is testable on CPU with the gloo backend (tests/test_host_cpu.py).
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


def all_reduce_max_u64(packed_i64, group=None):
    """unsigned 64-bit MAX of an int64-typed view: flipping bit 63 maps the unsigned order onto the signed one NCCL / gloo reduce"""
    packed_i64 ^= -0x8000000000000000
    dist.all_reduce(packed_i64, op=dist.ReduceOp.MAX, group=group)
    packed_i64 ^= -0x8000000000000000
    return packed_i64


class HugePairSplit:
    """One pair, replicated on every rank, registered cooperatively (see module docstring).  All device work is enqueued on the current
    stream; there is no host sync.  `timers` (optional dict) receives CUDA events around the two collectives."""

    def __init__(self, src_des, src_xyz, tgt_des, tgt_xyz, group=None):
        from . import backend
        self.B = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        M, N = src_des.shape[0], tgt_des.shape[0]
        dev = src_des.device
        self.M, self.N = M, N
        self.src_xyz, self.tgt_xyz = src_xyz, tgt_xyz
        self.src_off = torch.tensor([0, M], dtype=torch.int32).to(dev)
        self.tgt_off = torch.tensor([0, N], dtype=torch.int32).to(dev)
        self.k1 = backend.MutualNNSplit(src_des, tgt_des, self.src_off, self.tgt_off, M, N)

    def run(self, hypotheses=50000, dist_th=0.10, similar_th=0.8, refine_thr=0.10, refine_iters=20, seed=0, pair_id=0, events=None):
        B, world, rank = self.B, self.world, self.rank
        packed = self.k1.partial(rank, world)
        if world > 1:
            if events is not None:
                events["k1_ar0"].record()
            all_reduce_max_u64(packed, self.group)
            if events is not None:
                events["k1_ar1"].record()
        m = self.k1.select(self.src_xyz, self.tgt_xyz, want_nn=False, want_mids=False)
        corr, cnt = m["corr"], m["n_mutual"]
        h0, h1 = hypothesis_range(hypotheses, rank, world)
        best = B.ransac_batched(corr, self.src_off, cnt, hypotheses, dist_th, similar_th, seed=seed, pair_id_base=pair_id, h_begin=h0, h_end=h1)
        if world > 1:
            if events is not None:
                events["rs_ar0"].record()
            dist.all_reduce(best, op=dist.ReduceOp.MAX, group=self.group)        # packed value < 2^63: signed max == unsigned max
            if events is not None:
                events["rs_ar1"].record()
        T, inl, bh = B.ransac_finalize_batched(corr, self.src_off, cnt, best, dist_th, similar_th, seed=seed, pair_id_base=pair_id)
        if refine_iters > 0:
            T, _, _ = B.post_refinement_batched(T, corr, self.src_off, cnt, refine_thr, refine_iters, max_count=min(self.M, self.N))
        return T, cnt, inl, bh
