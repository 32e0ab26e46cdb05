#!/usr/bin/env python3
"""bench.py — fragment pairs/sec of the BUFFER correspondence-and-pose back end (match + RANSAC + SVD) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--config 2]

One "step" = one pass of the hot path (mutual-NN matching -> Philox RANSAC with 3-point Kabsch and inlier scoring ->
weighted-Kabsch post-refinement) over one batch of synthetic fragment pairs.  Default workload = BASELINE.json
configs[1]: 1,623 pairs x 5,000 keypoints x 32-d descriptors, 50,000 hypotheses per pair, 70 % outliers.
N > 1 (torchrun): STRONG scaling - the fixed workload is sharded by pair over the ranks, one all_gather of the poses inside the
timed region; time = max over ranks (the weak-scaling number is reported beside it as `weak_scaling`).  Every run also registers
BASELINE config 5's single 100k x 100k pair cooperatively on all ranks (`split_pair`: row-split matching + NCCL MAX all-reduce).
Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md §bench).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from buffer_b200 import synthetic as S  # noqa: E402

METRIC = "fragment_pairs_per_sec_match_ransac_svd"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config id (1-5); 2 is the one the metric is quoted on")
    ap.add_argument("--pairs", type=int, default=0, help="override the number of pairs per GPU (0 = the config's)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="pairs in the CPU baseline sample (0 = 4 per host thread)")
    ap.add_argument("--k1-algo", type=int, default=1, help="0 = FP32 FFMA2 mutual-NN kernel, 1 = tcgen05 f16 filter + exact FP32 re-check (bit-identical results)")
    ap.add_argument("--e2e-chunk", type=int, default=64, help="pairs per host->device chunk of the e2e leg (two chunks in flight on two streams)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-split-pair", action="store_true", help="skip the config-5 huge-pair leg (one 100k x 100k pair registered cooperatively by all ranks)")
    return ap.parse_args()


def workload(cfg_id, pairs_override):
    c = S.CONFIGS[cfg_id]
    P = pairs_override or c["num_pairs"]
    return c, P


def gen_pairs(c, P, first_pair, device, chunk=128):
    """generate P pairs in chunks (bounded temporary memory) -> PairBatch on `device`"""
    parts = []
    for p0 in range(0, P, chunk):
        n = min(chunk, P - p0)
        parts.append(S.make_pairs(n, first_pair=first_pair + p0, device=device, **c["gen"]))
    cat = lambda f: torch.cat([getattr(b, f) for b in parts], 0)
    return S.PairBatch(cat("src_des"), cat("tgt_des"), cat("src_xyz"), cat("tgt_xyz"), cat("T_gt"), cat("perm"), cat("inlier"))


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def fp32_peak_tflops(dev):
    """live FFMA2 issue peak on this GPU (bfr_fp32_probe), best of 5"""
    from buffer_b200 import _lib
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    grid, iters = sms * 4, 4000
    scratch = torch.full((grid * 256 + 128,), 1.0009765625, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.bfr_fp32_probe(grid, iters, scratch.data_ptr(), st), "bfr_fp32_probe")
        e1.record(); e1.synchronize()
        best = max(best, grid * 256.0 * iters * 64 * 4 / (e0.elapsed_time(e1) * 1e-3) * 1e-12)   # 64 FFMA2 = 256 flop / thread / iteration
    return best


def cpu_baseline(c, sample_pairs, threads, b=None):
    """the oracle port (oracle/bfr_oracle.c, OpenMP over pairs) on the host cores, on a bounded sample of the workload
    (`b`: the first pairs of the GPU workload copied to the host, else freshly generated ones)"""
    from oracle import oracle as O
    O.build()
    O.set_num_threads(threads)             # torchrun exports OMP_NUM_THREADS=1
    if b is None:
        b = S.make_pairs(sample_pairs, first_pair=0, **c["gen"])
    N = c["gen"]["num_kpts"]
    off = np.arange(sample_pairs + 1, dtype=np.int32) * N
    args = (b.src_des.reshape(-1, 32).numpy(), b.src_xyz.reshape(-1, 3).numpy(), off, b.tgt_des.reshape(-1, 32).numpy(),
            b.tgt_xyz.reshape(-1, 3).numpy(), off, c["hypotheses"], 0, 0, c["dist_th"], c["similar_th"], c["refine_thr"], 20)
    t0 = time.perf_counter()
    T, nm, ni = O.register_batched(*args)
    dt = time.perf_counter() - t0
    return sample_pairs / dt, dt, T, b


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python + third-party natives
    that cannot travel to the GPU box, so this arm times the oracle port with all host threads (tier rule)."""
    if rank != 0:
        return
    c, P = workload(args.config, args.pairs)
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_pairs or max(threads * 8, 32)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, _, _ = cpu_baseline(c, sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = sum(sample for _ in vals) / sum(dt for _, dt in vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, c, P, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d pairs of the workload per step (same generator, seeds 0..), whole back end" % sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args, c, P, world=1):
    return {"workload": "BASELINE.json configs[%d]: %d pairs x %d keypoints x 32-d, %d RANSAC hypotheses/pair, outlier ratio %s"
                        % (args.config - 1, P, c["gen"]["num_kpts"], c["hypotheses"], c["gen"].get("outlier_ratio")),
            "pairs_total": P, "pairs_per_gpu": (P + world - 1) // world, "keypoints": c["gen"]["num_kpts"], "desc_dim": 32, "hypotheses": c["hypotheses"],
            "dist_th": c["dist_th"], "similar_th": c["similar_th"], "refine_iters": 20, "confidence": 1.0,
            "cache": "inputs (%.2f GB in total) exceed the 126 MB L2; every step re-reads them from HBM" % (P * c["gen"]["num_kpts"] * 2 * (32 + 3) * 4 / 1e9),
            "parallelism": "the FIXED workload is sharded by pair over the ranks (contiguous blocks); no collective in the data path, one all_gather of "
                           "18 floats per pair (poses + counts) inside the timed region" if world > 1 else "single GPU"}


def gen_range(c, start, end, device, chunk=128):
    """pairs [start, end) of the workload, generated in the workload's global 128-pair chunks so that a pair's data does not depend on
    how the workload is sharded -> PairBatch on `device`"""
    parts = []
    for k in range(start // chunk, (end + chunk - 1) // chunk):
        p0 = k * chunk
        b = S.make_pairs(chunk, first_pair=p0, device=device, **c["gen"])
        lo, hi = max(start, p0) - p0, min(end, p0 + chunk) - p0
        parts.append([getattr(b, f)[lo:hi] for f in ("src_des", "tgt_des", "src_xyz", "tgt_xyz", "T_gt", "perm", "inlier")])
        del b
    return S.PairBatch(*[torch.cat([p[i] for p in parts], 0).contiguous() for i in range(7)])


def timed(fn, steps, barrier, dev, world, dist):
    """K calls of fn bracketed by barrier + synchronize, device-timed, max over ranks -> ms per step"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps, out


def split_pair_leg(args, dev, rank, world, dist, barrier, n_kpts=100000):
    """BASELINE config 5's single huge pair (100k x 100k keypoints, 50k hypotheses) registered cooperatively by all ranks: K1 row-split +
    NCCL MAX all-reduce of the packed bests, RANSAC split by hypothesis + 8-byte MAX all-reduce (buffer_b200.dist.HugePairSplit)"""
    from buffer_b200 import backend as B, dist as D
    c5 = S.CONFIGS[5]
    g = dict(c5["gen"], num_kpts=n_kpts)
    b = S.make_pairs(1, first_pair=10 ** 6, device=dev, **g)          # same seed on every rank: the pair is replicated
    hp = D.HugePairSplit(b.src_des[0], b.src_xyz[0], b.tgt_des[0], b.tgt_xyz[0])
    kw = dict(hypotheses=c5["hypotheses"], dist_th=c5["dist_th"], similar_th=c5["similar_th"], refine_thr=c5["refine_thr"], seed=0, pair_id=10 ** 6)
    for _ in range(3):
        hp.run(**kw)
    ms, (T, cnt, inl, bh) = timed(lambda: hp.run(**kw), max(args.steps, 5), barrier, dev, world, dist)
    out = {"keypoints": n_kpts, "hypotheses": c5["hypotheses"], "ms": ms, "mutual_matches": int(cnt.item()), "ransac_inliers": int(inl.item())}
    rec, rte, rre = S.registration_recall(T.cpu(), b.T_gt.cpu())
    out["recall"] = rec; out["rte_m"] = float(rte.max())
    if world > 1:
        ev = {k: torch.cuda.Event(enable_timing=True) for k in ("k1_ar0", "k1_ar1", "rs_ar0", "rs_ar1")}
        barrier()
        hp.run(events=ev, **kw)
        barrier()
        ar = torch.tensor([ev["k1_ar0"].elapsed_time(ev["k1_ar1"]), ev["rs_ar0"].elapsed_time(ev["rs_ar1"])], dtype=torch.float64, device=dev)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        out["allreduce_packed_bests_ms"] = ar[0].item(); out["allreduce_packed_bests_bytes"] = int(hp.k1.packed.numel() * 8)
        out["allreduce_ransac_best_ms"] = ar[1].item(); out["allreduce_share"] = (ar[0].item() + ar[1].item()) / ms
        # the same pair on ONE GPU (this rank alone, no collective): must give the same bits
        solo = D.HugePairSplit(b.src_des[0], b.src_xyz[0], b.tgt_des[0], b.tgt_xyz[0])
        solo.world, solo.rank = 1, 0
        T1, cnt1, inl1, bh1 = solo.run(**kw)
        same = torch.tensor([int(torch.equal(T, T1) and torch.equal(cnt, cnt1) and torch.equal(inl, inl1) and torch.equal(bh, bh1))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["bit_equal_to_one_gpu"] = bool(same.item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            solo.run(**kw)
        e1.record(); torch.cuda.synchronize()
        out["ms_one_gpu_same_run"] = e0.elapsed_time(e1) / 3
        out["design"] = ("K1 row blocks split over ranks, %d-byte MAX all-reduce of the packed bests (NCCL over NVLink), select replicated; RANSAC hypotheses split, "
                         "8-byte MAX all-reduce, winner regenerated locally; refinement replicated" % out["allreduce_packed_bests_bytes"])
    return out


def torch_baseline(c, batch, threads, timed_hypotheses=2000):
    """BASELINE config 1: the reference's CPU torch path (oracle/torch_ref.py) on the first pair of the workload, host cores"""
    from oracle import oracle as O, torch_ref as TR
    h = [getattr(batch, f)[0].cpu() for f in ("src_des", "tgt_des", "src_xyz", "tgt_xyz")]
    samples_fn = lambda K, n: np.stack([O.sample3(0, 0, i, K) for i in range(n)]).astype(np.int64)
    r = TR.time_pair(h[0], h[1], h[2], h[3], samples_fn, c["hypotheses"], timed_hypotheses, c["dist_th"], c["similar_th"], c["refine_thr"], threads)
    T = r.pop("T")
    ok, rte, rre = S.registration_recall(torch.from_numpy(T)[None], batch.T_gt[:1].cpu())
    r.update({"value": 1.0 / r["total_s"], "unit": UNIT, "cores": threads, "torch_threads": torch.get_num_threads(), "kind": "reference torch path, restated (oracle/torch_ref.py)",
              "sample": "pair 0 of the workload: torch.cdist + min matching (best of 3), Open3D-semantics Python loop on torch.svd Kabsch for the first %d of %d hypotheses "
                        "(time scaled linearly to %d; Open3D's own C++ loop is a third-party wheel that is absent), post_refinement with diag_embed" % (timed_hypotheses, c["hypotheses"], c["hypotheses"]),
              "recall_on_sample": ok})
    return r


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    if world > 1:
        bind_to_gpu_cpus(local)              # pinned host buffers of the e2e leg are then first-touched on the GPU's own NUMA node
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from buffer_b200 import _lib, backend as B, dist as D
    L = _lib.lib()
    B.set_k1_algo(args.k1_algo)

    c, P_total = workload(args.config, args.pairs)
    N = c["gen"]["num_kpts"]
    kw = dict(hypotheses=c["hypotheses"], dist_th=c["dist_th"], similar_th=c["similar_th"], refine_thr=c["refine_thr"], refine_iters=20, seed=0)
    p_lo, p_hi = D.shard_range(P_total, rank, world)              # STRONG scaling: the fixed workload is sharded by pair
    P = p_hi - p_lo
    batch = gen_range(c, p_lo, p_hi, dev)                         # resident in HBM before the timed region
    src_des = batch.src_des.reshape(P * N, 32); tgt_des = batch.tgt_des.reshape(P * N, 32)
    src_xyz = batch.src_xyz.reshape(P * N, 3); tgt_xyz = batch.tgt_xyz.reshape(P * N, 3)
    off = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        T, nm, ni = B.register_batched(src_des, src_xyz, off, tgt_des, tgt_xyz, off, N, N, pair_id_base=p_lo, **kw)
        if world > 1:
            return D.gather_pair_results(T, nm, ni, P_total) + (T,)     # every rank ends up with all poses
        return T, nm, ni, T

    peak_tf = fp32_peak_tflops(dev)
    for _ in range(max(args.warmup, 3)):
        out = step()
    # ---- timed region: exactly K steps, device-timed, K1 bracketed by its own events ----------------------------
    k1_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b_ in k1_ev:                      # torch creates events lazily: record once so the handles exist, the library re-records them
        a.record(); b_.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0.record()
    for i in range(args.steps):
        L.bfr_debug_set_k1_events(ctypes.c_void_p(k1_ev[i][0].cuda_event), ctypes.c_void_p(k1_ev[i][1].cuda_event))
        out = step()
    e1.record()
    L.bfr_debug_set_k1_events(None, None)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_ev) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = P_total / (ms_step * 1e-3)
    T_all, nm, ni, T = out
    ok_local, rte, rre = S.registration_recall(T.cpu(), batch.T_gt.cpu())
    q = torch.tensor([ok_local * P, float(P)], dtype=torch.float64, device=dev)
    qmax = torch.tensor([float(rte.max()) if P else 0.0, float(rre.max()) if P else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(q); dist.all_reduce(qmax, op=dist.ReduceOp.MAX)
    recall = (q[0] / q[1]).item()

    # ---- weak-scaling leg (N > 1 only; the headline above is STRONG scaling): every rank runs the full workload of its own pairs ----
    weak = None
    if world > 1:
        wb = gen_range(c, rank * P_total, (rank + 1) * P_total, dev)
        wargs = (wb.src_des.reshape(-1, 32), wb.src_xyz.reshape(-1, 3), (torch.arange(P_total + 1, dtype=torch.int32) * N).to(dev),
                 wb.tgt_des.reshape(-1, 32), wb.tgt_xyz.reshape(-1, 3))
        wstep = lambda: B.register_batched(wargs[0], wargs[1], wargs[2], wargs[3], wargs[4], wargs[2], N, N, pair_id_base=rank * P_total, **kw)
        for _ in range(2):
            wstep()
        wms, _ = timed(wstep, args.steps, barrier, dev, world, dist)
        weak = {"value": world * P_total / (wms * 1e-3), "unit": UNIT, "ms_per_step": wms, "pairs_per_gpu": P_total,
                "note": "every rank processes its own %d pairs, no collective" % P_total}
        del wb, wargs

    # ---- extra leg (outside the timed region): the all-FP32 mutual-NN kernel on the same data, for the FP32 roofline --------
    fp32_ms = None
    if rank == 0:
        B.set_k1_algo(B.K1_FP32)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a_, b_ in ev:
            a_.record(); b_.record()
        for a_, b_ in ev:
            L.bfr_debug_set_k1_events(ctypes.c_void_p(a_.cuda_event), ctypes.c_void_p(b_.cuda_event))
            r_fp = B.mutual_matching_batched(src_des, tgt_des, off, off, N, N, want_nn=True, want_mids=False, col_splits=1)
        L.bfr_debug_set_k1_events(None, None)
        torch.cuda.synchronize()
        fp32_ms = min(a_.elapsed_time(b_) for a_, b_ in ev)
        B.set_k1_algo(args.k1_algo)
        r_tc = B.mutual_matching_batched(src_des, tgt_des, off, off, N, N, want_nn=True, want_mids=False)
        k1_paths_identical = bool(torch.equal(r_fp["nn_s"], r_tc["nn_s"]) and torch.equal(r_fp["nn_t"], r_tc["nn_t"]))
        del r_fp, r_tc
        # ---- extra leg: the same K1 launch on UNRELATED descriptors (every pair matched against the next pair's targets: no true matches),
        # the filter's worst realistic case (workload sensitivity) --------
        k1_unrelated_ms = None
        if P > 1:
            tgt_roll = torch.roll(tgt_des.reshape(P, N, 32), 1, 0).reshape(P * N, 32).contiguous()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            for a_, b_ in ev:
                a_.record(); b_.record()
            for a_, b_ in ev:
                L.bfr_debug_set_k1_events(ctypes.c_void_p(a_.cuda_event), ctypes.c_void_p(b_.cuda_event))
                B.mutual_matching_batched(src_des, tgt_roll, off, off, N, N, want_nn=True, want_mids=False)
            L.bfr_debug_set_k1_events(None, None)
            torch.cuda.synchronize()
            k1_unrelated_ms = min(a_.elapsed_time(b_) for a_, b_ in ev)
            del tgt_roll
        # ---- extra leg: K2+K3 alone on the same data (events on the launching stream), H_valid for the 28*H_valid*C work model --------
        rm = B.mutual_matching_batched(src_des, tgt_des, off, off, N, N, src_xyz, tgt_xyz, want_nn=False, want_mids=False)
        nvalid = torch.zeros(P, dtype=torch.int32, device=dev)
        ransac_ms = []
        for i in range(3):
            nvalid.zero_()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=p_lo, valid_count=nvalid)
            eb.record(); eb.synchronize()
            ransac_ms.append(ea.elapsed_time(eb))
        ransac_ms = min(ransac_ms)
        hv_total = float(nvalid.sum().item()); c_mean = float(rm["n_mutual"].float().mean().item())
        # same launch with every residual in FP32 (BFR_CFG_RANSAC_TC = 0): the two scoring paths must return identical packed bests
        bp_tc = B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=p_lo).clone()
        B.set_ransac_scoring(B.RANSAC_FP32)
        ransac_fp32_ms = []
        for i in range(3):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            bp_fp = B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=p_lo)
            eb.record(); eb.synchronize()
            ransac_fp32_ms.append(ea.elapsed_time(eb))
        B.set_ransac_scoring(B.RANSAC_TENSOR_FILTER)
        ransac_fp32_ms = min(ransac_fp32_ms)
        ransac_same = bool(torch.equal(bp_tc, bp_fp))
        # Open3D's confidence early exit (the reference's 3DMatch setting, ThreeDMatch/config.py:65): same call with confidence = 0.999
        conf_ms = []
        for i in range(3):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            bp_c = B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=p_lo, confidence=0.999)
            eb.record(); eb.synchronize()
            conf_ms.append(ea.elapsed_time(eb))
        Tc_, inl_c, _ = B.ransac_finalize_batched(rm["corr"], off, rm["n_mutual"], bp_c, c["dist_th"], c["similar_th"], seed=0, pair_id_base=p_lo)
        conf_recall, _, _ = S.registration_recall(Tc_.cpu(), batch.T_gt.cpu())
        conf = {"confidence": 0.999, "ms_per_launch": min(conf_ms), "speedup_vs_all_hypotheses": ransac_ms / min(conf_ms), "recall_before_refinement": conf_recall,
                "inliers_mean": float(inl_c.float().mean())}
        del rm

    # ---- e2e: host (pinned) buffers -> poses on the host, copies inside the timed region ---------------------------
    e2e = None
    if not args.no_e2e:
        hb = batch.to("cpu")
        pin = lambda x: x.contiguous().pin_memory()
        h = [pin(hb.src_des), pin(hb.src_xyz), pin(hb.tgt_des), pin(hb.tgt_xyz)]
        Th = torch.empty(P, 4, 4).pin_memory(); nmh = torch.empty(P, dtype=torch.int32).pin_memory(); nih = torch.empty(P, dtype=torch.int32).pin_memory()
        chunk = max(1, min(P, args.e2e_chunk))
        reg = B.HostRegistrar(chunk, N, N, dev, ransac_splits=None, pair_id_base=p_lo, **kw)
        for _ in range(2):
            reg.run(*h, Th, nmh, nih)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            reg.run(*h, Th, nmh, nih)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(Th, T.cpu()))
        # the platform ceiling: the bare host->device copy of one step's inputs, ALL ranks copying at the same time (pinned, one stream each)
        dbuf = [torch.empty_like(x, device=dev) for x in h]
        h2d_ms = []
        for _ in range(3):
            ca, cb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ca.record()
            for d_, x in zip(dbuf, h):
                d_.copy_(x, non_blocking=True)
            cb_.record(); cb_.synchronize()
            tcp = torch.tensor([ca.elapsed_time(cb_)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tcp, op=dist.ReduceOp.MAX)
            h2d_ms.append(tcp.item())
        del dbuf
        h2d_ms = min(h2d_ms)
        step_ms = tt.item() / args.steps * 1e3
        bytes_in = int(sum(x.numel() * 4 for x in h)); bytes_out = int(Th.numel() * 4 + nmh.numel() * 4 + nih.numel() * 4)
        tot = torch.tensor([float(bytes_in), float(bytes_out)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": P_total * args.steps / tt.item(), "unit": UNIT, "ms_per_step": step_ms,
               "h2d_concurrent_copy_floor_ms": h2d_ms, "h2d_concurrent_copy_gbs_per_gpu": bytes_in / (h2d_ms * 1e-3) * 1e-9,
               "frac_of_concurrent_copy_floor": h2d_ms / step_ms,
               "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
               "api": "buffer_b200.backend.HostRegistrar.run -> ONE bfr_register_uniform_host_chunked call per rank and step (pinned host buffers, %d-pair chunks on 2 streams); "
                      "each rank's poses land in its own host buffer" % chunk,
               "poses_equal_device_path": same, "launches_per_step": 8 * ((P + chunk - 1) // chunk)}

    split_pair = None if args.no_split_pair else split_pair_leg(args, dev, rank, world, dist, barrier)

    if rank == 0:
        flops_k1 = 2.0 * N * N * 32 * P
        ach = flops_k1 / (k1_ms * 1e-3) * 1e-12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k1_tc_traffic.json" if args.k1_algo == 1 else "k1_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                if traffic is not None and tj.get("pairs"):
                    traffic = traffic * P / tj["pairs"]       # static: taken from the committed ncu capture, scaled to this launch's pairs
            except Exception:
                traffic = None
        mp = _measured_peaks()
        fp32_roof = {"kernel": "k1_mutual_nn_kernel (all products in FP32, FFMA2)", "bound": "fp32", "achieved": flops_k1 / (fp32_ms * 1e-3) * 1e-12,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": flops_k1 / (fp32_ms * 1e-3) * 1e-12 / peak_tf, "ms_per_launch": fp32_ms,
                     "peak_source": "live FFMA2 (fma.rn.f32x2) issue-rate probe on this GPU; MEASURED_PEAKS.json has no FP32 figure "
                                    "(theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s)",
                     "outputs_identical_to_tensor_path": k1_paths_identical}
        if args.k1_algo == 1:
            bf16_peak = mp.get("bf16_tflops") or 1590.0
            # tensor-pipe floor of this formulation: one accumulator tile = 128 own rows x 256 streamed rows x K 32 = two back-to-back f16 MMAs;
            # the pipe drains on every switch to another accumulator tile, ~345 cycles per tile whatever its N (profiles/r01_mma_pipeline_microbench.txt)
            tiles = 2.0 * P * ((N + 127) // 128) * ((N + 255) // 256)
            mma_floor_ms = tiles / 148.0 * 345.0 / ((clocks.get("sm_mhz") or 1965.0) * 1e3)
            roof = {"kernel": "k1_tc_kernel (one launch, both directions: tcgen05 f16 filter, 128x256 accumulator tiles in TMEM + exact FP32 re-check)", "bound": "tensor",
                    "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak, "traffic": traffic,
                    "traffic_source": "static: dram__bytes_read+write of the committed ncu --set full capture (profiles/k1_tc_traffic.json), scaled by pairs; not measured in this run",
                    "k1_ms_per_launch": k1_ms, "k1_share_of_step": k1_ms / ms_step, "algorithmic_flops_per_launch": flops_k1,
                    "executed_tensor_flops_per_launch": 2 * flops_k1,
                    "peak_source": "dense 16-bit tensor peak = measured cuBLAS bf16 burst in MEASURED_PEAKS.json (f16 and bf16 MMAs run at the same rate) (%s)" % ("measured" if mp.get("bf16_tflops") else "fallback 1.59 PF"),
                    "tensor_pipe_floor_ms": mma_floor_ms, "frac_of_tensor_pipe_floor": mma_floor_ms / k1_ms,
                    "k1_ms_unrelated_descriptors": k1_unrelated_ms,
                    "note": "K = 32 gives two MMAs per accumulator tile, so the tensor pipe is bound by its ~345-cycle drain per tile (tensor_pipe_floor_ms), "
                            "not by FLOPs, and the kernel as a whole by the TMEM -> register max-reduction epilogue and the MMA <-> epilogue hand-off; "
                            "the algorithmic FLOP rate exceeds the FP32 roofline because the products run on tensor cores and only near-best candidates are re-evaluated in FP32",
                    "vs_fp32_ffma2_peak": ach / peak_tf, "hbm_peak_gbs_measured": mp.get("hbm_gbs")}
        else:
            roof = dict(fp32_roof, traffic=traffic, k1_ms_per_launch=k1_ms, k1_share_of_step=k1_ms / ms_step, algorithmic_flops_per_launch=flops_k1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, c, P_total, world), "clocks": clocks, "gpu_launches": 8 * args.steps,
                "roofline": roof, "roofline_fp32_path": fp32_roof,
                "roofline_ransac": {"kernel": "ransac_kernel (Philox + Kabsch + checkers + inlier scoring: tcgen05 residual filter on 2-level f16 operand splits, exact FP32 re-check inside the error band)",
                                    "bound": "fp32", "ms_per_launch": ransac_ms, "ms_per_launch_fp32_scoring": ransac_fp32_ms, "outputs_identical_to_fp32_scoring": ransac_same,
                                    "valid_hypotheses_per_pair": hv_total / P, "hypotheses_per_pair": c["hypotheses"], "correspondences_per_pair": c_mean,
                                    "algorithmic_flops_per_launch": 28.0 * hv_total * c_mean, "achieved": 28.0 * hv_total * c_mean / (ransac_ms * 1e-3) * 1e-12,
                                    "peak": peak_tf, "unit": "TFLOP/s", "frac": 28.0 * hv_total * c_mean / (ransac_ms * 1e-3) * 1e-12 / peak_tf,
                                    "hypotheses_per_s": P * c["hypotheses"] / (ransac_ms * 1e-3),
                                    "streaming_model_gbs": 24.0 * c_mean * hv_total / 512.0 / (ransac_ms * 1e-3) * 1e-9,
                                    "streaming_model_note": "bytes = 24*C*H_valid/T_h with T_h = 512 hypotheses per CTA pass (the all-FP32 formulation); a pair's correspondences are loaded ONCE "
                                                            "into shared memory (compulsory HBM traffic is 32*C bytes per pair), so the kernel is issue-bound, not HBM-bound",
                                    "note": "achieved/frac count the 28 flop per (hypothesis, correspondence) of the FP32 formulation; with the tensor-core filter 19 of them run as tcgen05.mma "
                                            "and the kernel is bound by the accumulator hand-off per 128-correspondence tile (DESIGN.md K2+K3)",
                                    "open3d_confidence_exit": conf},
                "quality": {"registration_recall": recall, "rte_max_m": qmax[0].item(), "rre_max_deg": qmax[1].item(),
                            "mutual_matches_mean": float(nm.float().mean()), "ransac_inliers_mean": float(ni.float().mean())}}
        if weak is not None:
            line["weak_scaling"] = weak
        if e2e is not None:
            line["e2e"] = e2e
        if split_pair is not None:
            line["split_pair"] = split_pair
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            sample = min(P, args.cpu_sample_pairs or max(threads * 16, 64))
            hs = S.PairBatch(*[getattr(batch, f)[:sample].cpu() for f in ("src_des", "tgt_des", "src_xyz", "tgt_xyz", "T_gt", "perm", "inlier")])
            v, dt, Tc, _ = cpu_baseline(c, sample, threads, hs)
            same = bool(np.array_equal(Tc, T[:sample].cpu().numpy()))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
                                    "sample": "first %d pairs of the same workload, whole back end (oracle/bfr_oracle.c, OpenMP over pairs)" % sample,
                                    "poses_bit_identical_to_gpu": same}
            try:
                line["cpu_baseline_torch"] = torch_baseline(c, batch, threads)
            except Exception as exc:             # noqa: BLE001 - a baseline leg must not take the bench line down
                line["cpu_baseline_torch"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bind_to_gpu_cpus(index):
    """multi-GPU runs: restrict this rank to the CPUs NVML reports as local to its GPU (what `numactl --cpunodebind` would do), so that
    the pinned host memory of the end-to-end leg is allocated next to the PCIe root the GPU hangs off; silently skipped without NVML"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


if __name__ == "__main__":
    main()
