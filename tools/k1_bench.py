#!/usr/bin/env python3
"""K1-only timing harness: python tools/k1_bench.py [so_path] [pairs] [N]  -> ms per launch, TFLOP/s (2*N*M*D flops)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
so = sys.argv[1] if len(sys.argv) > 1 else "buffer_b200/libbuffer_b200.so"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 296
N = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
L = C.CDLL(os.path.abspath(so))
L.bfr_mutual_nn_workspace_bytes.restype = C.c_size_t
dev = "cuda:0"
if len(sys.argv) > 4: L.bfr_config_set(1, int(sys.argv[4]))
g = torch.Generator(device=dev); g.manual_seed(1)
src = torch.nn.functional.normalize(torch.randn(P * N, 32, device=dev, generator=g), dim=-1)
tgt = torch.nn.functional.normalize(torch.randn(P * N, 32, device=dev, generator=g), dim=-1)
if len(sys.argv) > 5 and sys.argv[5] == "planted":      # every src row has a noisy copy at a permuted tgt row (the bench generator's structure)
    noisy = torch.nn.functional.normalize(src + 0.05 * torch.randn(P * N, 32, device=dev, generator=g), dim=-1).reshape(P, N, 32)
    perm = torch.argsort(torch.rand(P, N, device=dev, generator=g), dim=-1)
    tgt = torch.empty_like(noisy).scatter_(1, perm[:, :, None].expand(P, N, 32), noisy).reshape(P * N, 32)
off = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev)
nb = L.bfr_mutual_nn_workspace_bytes(P, N, N)
ws = torch.empty(nb + 1024, dtype=torch.uint8, device=dev)
nn_s = torch.empty(P * N, dtype=torch.int64, device=dev); nn_t = torch.empty_like(nn_s)
nm = torch.empty(P, dtype=torch.int32, device=dev)
vp = C.c_void_p
def run():
    rc = L.bfr_mutual_matching_batched(vp(src.data_ptr()), vp(tgt.data_ptr()), vp(off.data_ptr()), vp(off.data_ptr()), P, N, N, P * N, P * N, 32, 1,
                                       vp(nn_s.data_ptr()), vp(nn_t.data_ptr()), None, None, None, None, None, None, vp(nm.data_ptr()), None,
                                       vp(ws.data_ptr()), C.c_size_t(ws.numel()), vp(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for a, b in evs: a.record(); b.record()
L.bfr_debug_set_k1_events.argtypes = [vp, vp]
for _ in range(3): run()
torch.cuda.synchronize()
for a, b in evs:
    L.bfr_debug_set_k1_events(vp(a.cuda_event), vp(b.cuda_event)); run()
torch.cuda.synchronize()
ms = min(a.elapsed_time(b) for a, b in evs)
if hasattr(L, "bfr_dbg_counters"):
    import numpy as np
    out = np.zeros(8, dtype=np.uint64); L.bfr_dbg_counters(out.ctypes.data_as(vp))
    print("dbg counters (8 launches x 2 dirs): compactions %d entries_in %d entries_out %d overflow_valid %d overflow_invalid %d" % tuple(out[:5]))
print("%s P=%d N=%d: K1 %.3f ms  %.2f TFLOP/s  checksum %d" % (os.path.basename(so), P, N, ms, 2.0 * N * N * 32 * P / ms * 1e-9, int(nn_s.sum().item() % 1000003)))
