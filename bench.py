#!/usr/bin/env python3
"""bench.py — fragment pairs/sec of the BUFFER correspondence-and-pose back end (match + RANSAC + SVD) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--config 2]

One "step" = one pass of the hot path (mutual-NN matching -> Philox RANSAC with 3-point Kabsch and inlier scoring ->
weighted-Kabsch post-refinement) over one batch of synthetic fragment pairs.  Default workload = BASELINE.json
configs[1]: 1,623 pairs x 5,000 keypoints x 32-d descriptors, 50,000 hypotheses per pair, 70 % outliers.
N > 1 (torchrun): pairs are independent, every rank processes its own 1,623 pairs, no data-path collective (weak
scaling); time = max over ranks.  Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md §bench).
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
            "config": config_dict(args, c, P),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d pairs of the workload per step (same generator, seeds 0..), whole back end" % sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args, c, P):
    return {"workload": "BASELINE.json configs[%d]: %d pairs/GPU x %d keypoints x 32-d, %d RANSAC hypotheses/pair, outlier ratio %s"
                        % (args.config - 1, P, c["gen"]["num_kpts"], c["hypotheses"], c["gen"].get("outlier_ratio")),
            "pairs_per_gpu": P, "keypoints": c["gen"]["num_kpts"], "desc_dim": 32, "hypotheses": c["hypotheses"],
            "dist_th": c["dist_th"], "similar_th": c["similar_th"], "refine_iters": 20,
            "cache": "inputs (%.2f GB/GPU) exceed the 126 MB L2; every step re-reads them from HBM" % (P * c["gen"]["num_kpts"] * 2 * (32 + 3) * 4 / 1e9),
            "parallelism": "pairs sharded by rank, no collective in the data path"}


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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from buffer_b200 import _lib, backend as B
    L = _lib.lib()
    B.set_k1_algo(args.k1_algo)

    c, P = workload(args.config, args.pairs)
    N = c["gen"]["num_kpts"]
    kw = dict(hypotheses=c["hypotheses"], dist_th=c["dist_th"], similar_th=c["similar_th"], refine_thr=c["refine_thr"], refine_iters=20, seed=0)
    batch = gen_pairs(c, P, rank * P, dev)                       # resident in HBM before the timed region
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
        return B.register_batched(src_des, src_xyz, off, tgt_des, tgt_xyz, off, N, N, pair_id_base=rank * P, **kw)

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
    value = world * P / (ms_step * 1e-3)
    T, nm, ni = out
    recall, rte, rre = S.registration_recall(T.cpu(), batch.T_gt.cpu())

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
            B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], seed=0, pair_id_base=rank * P, valid_count=nvalid)
            eb.record(); eb.synchronize()
            ransac_ms.append(ea.elapsed_time(eb))
        ransac_ms = min(ransac_ms)
        hv_total = float(nvalid.sum().item()); c_mean = float(rm["n_mutual"].float().mean().item())
        del rm

    # ---- e2e: host (pinned) buffers -> poses on the host, copies inside the timed region ---------------------------
    e2e = None
    if not args.no_e2e:
        hb = batch.to("cpu")
        pin = lambda x: x.contiguous().pin_memory()
        h = [pin(hb.src_des), pin(hb.src_xyz), pin(hb.tgt_des), pin(hb.tgt_xyz)]
        Th = torch.empty(P, 4, 4).pin_memory(); nmh = torch.empty(P, dtype=torch.int32).pin_memory(); nih = torch.empty(P, dtype=torch.int32).pin_memory()
        chunk = min(P, args.e2e_chunk)
        reg = B.HostRegistrar(chunk, N, N, dev, ransac_splits=None, **kw)
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
        # context for the e2e number: the bare host->device copy of one step's inputs (pinned, one stream), outside the timed region
        dbuf = [torch.empty_like(x, device=dev) for x in h]
        ca, cb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d_ms = []
        for _ in range(3):
            ca.record()
            for d_, x in zip(dbuf, h):
                d_.copy_(x, non_blocking=True)
            cb_.record(); cb_.synchronize()
            h2d_ms.append(ca.elapsed_time(cb_))
        del dbuf
        h2d_ms = min(h2d_ms)
        e2e = {"value": world * P * args.steps / tt.item(), "unit": UNIT, "ms_per_step": tt.item() / args.steps * 1e3,
               "h2d_copy_alone_ms": h2d_ms, "h2d_copy_alone_gbs": sum(x.numel() * 4 for x in h) / (h2d_ms * 1e-3) * 1e-9,
               "h2d_bytes_per_step": int(sum(x.numel() * 4 for x in h)), "d2h_bytes_per_step": int(Th.numel() * 4 + nmh.numel() * 4 + nih.numel() * 4),
               "api": "buffer_b200.backend.HostRegistrar.run -> bfr_register_uniform_host (pinned host buffers, %d-pair chunks on 2 streams)" % chunk,
               "poses_equal_device_path": same, "launches_per_step": 8 * ((P + chunk - 1) // chunk)}

    if rank == 0:
        flops_k1 = 2.0 * N * N * 32 * P
        ach = flops_k1 / (k1_ms * 1e-3) * 1e-12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k1_tc_traffic.json" if args.k1_algo == 1 else "k1_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
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
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, c, P), "clocks": clocks, "gpu_launches": 7 * args.steps,
                "roofline": roof, "roofline_fp32_path": fp32_roof,
                "roofline_ransac": {"kernel": "ransac_kernel (Philox + Kabsch + checkers + inlier scoring)", "bound": "fp32", "ms_per_launch": ransac_ms,
                                    "valid_hypotheses_per_pair": hv_total / P, "hypotheses_per_pair": c["hypotheses"], "correspondences_per_pair": c_mean,
                                    "algorithmic_flops_per_launch": 28.0 * hv_total * c_mean, "achieved": 28.0 * hv_total * c_mean / (ransac_ms * 1e-3) * 1e-12,
                                    "peak": peak_tf, "unit": "TFLOP/s", "frac": 28.0 * hv_total * c_mean / (ransac_ms * 1e-3) * 1e-12 / peak_tf,
                                    "hypotheses_per_s": P * c["hypotheses"] / (ransac_ms * 1e-3),
                                    "streaming_model_gbs": 24.0 * c_mean * hv_total / 256.0 / (ransac_ms * 1e-3) * 1e-9,
                                    "streaming_model_note": "bytes = 24*C*H_valid/T_h with T_h = 256 hypotheses per CTA pass; the correspondences stream from L2/shared memory, "
                                                            "not HBM (compulsory HBM traffic is 32*C bytes per pair), so the kernel is FP32-issue-bound, not HBM-bound"},
                "quality": {"registration_recall": recall, "rte_max_m": float(rte.max()), "rre_max_deg": float(rre.max()),
                            "mutual_matches_mean": float(nm.float().mean()), "ransac_inliers_mean": float(ni.float().mean())}}
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = min(P, args.cpu_sample_pairs or max(threads * 16, 64))
            hs = S.PairBatch(*[getattr(batch, f)[:sample].cpu() for f in ("src_des", "tgt_des", "src_xyz", "tgt_xyz", "T_gt", "perm", "inlier")])
            v, dt, Tc, _ = cpu_baseline(c, sample, threads, hs)
            same = bool(np.array_equal(Tc, T[:sample].cpu().numpy()))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
                                    "sample": "first %d pairs of the same workload, whole back end (oracle/bfr_oracle.c, OpenMP over pairs)" % sample,
                                    "poses_bit_identical_to_gpu": same}
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
