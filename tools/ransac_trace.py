#!/usr/bin/env python3
"""Timeline of one tensor-core flush of the RANSAC kernel (needs a -DRS_TRACE variant: tools/build_variant.sh tr -DRS_TRACE).
Per warp and A tile, cycles relative to the flush's first event: W = start waiting for the accumulator, F = accumulator full seen,
L = tcgen05.ld done, P = ready to hand back (A-tile wait, fence, warp sync), H = handed back (+ MMAs issued if this warp was last: *),
M = math done, T = after the (thread 0 only) TMA issue."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from buffer_b200 import _lib
_lib.SO_PATH = os.path.abspath(os.environ.get("BFR_SO", "variants/lib_tr.so"))
from buffer_b200 import backend as B, synthetic as S
NT = 12
c = S.CONFIGS[2]; N = c["gen"]["num_kpts"]; dev = "cuda:0"; P = 148
b = S.make_pairs(P, first_pair=0, device=dev, **c["gen"])
off = (torch.arange(P + 1, dtype=torch.int32) * N).to(dev)
rm = B.mutual_matching_batched(b.src_des.reshape(P * N, 32), b.tgt_des.reshape(P * N, 32), off, off, N, N, b.src_xyz.reshape(P * N, 3), b.tgt_xyz.reshape(P * N, 3), want_nn=False, want_mids=False)
B.ransac_batched(rm["corr"], off, rm["n_mutual"], c["hypotheses"], c["dist_th"], c["similar_th"], splits=1); torch.cuda.synchronize()
L = ctypes.CDLL(_lib.SO_PATH); out = np.zeros(20 * 8 * NT, dtype=np.uint32)
L.bfr_dbg_ransac_trace.restype = ctypes.c_uint
L.bfr_dbg_ransac_trace(ctypes.c_void_p(out.ctypes.data)); ev = out.reshape(20, NT, 8).astype(np.int64); iss = ev[16:]; ev = ev[:16]
t0 = ev[:, 0, 0].min()
rel = lambda x: int((x - t0) & 0xffffffff)
for w in (0, 1, 4, 7, 8, 9, 15):
    line = []
    for i in range(4, 10):
        e = ev[w, i]
        line.append("%d: W%d F%d L%d H%d M%d" % (i, rel(e[0]), rel(e[1]), rel(e[2]), rel(e[4]), rel(e[5])))
    print("warp %2d | " % w + " | ".join(line))
for g in range(4):
    print("issuer %d | " % g + " | ".join("%d: start %d, A tile there %d, buffer back %d, issued %d" % (i, rel(iss[g, i, 0]), rel(iss[g, i, 1]), rel(iss[g, i, 2]), rel(iss[g, i, 3])) for i in range(2, NT) if iss[g, i, 3]))
d = lambda a, b: float(np.mean((ev[:, 2:, a] - ev[:, 2:, b]) & 0xffffffff))
print("means over warps and tiles: wait %.0f, ld %.0f, hand back %.0f, math %.0f, tile period %.0f" % (d(1, 0), d(2, 1), d(4, 2), d(5, 4), float(np.mean((ev[:, 3:, 0] - ev[:, 2:-1, 0]) & 0xffffffff))))
if os.environ.get("BFR_TRACE_ALL"):
    for w in range(16):
        print("warp %2d | " % w + " | ".join("%d: W%d F%d L%d H%d M%d" % (i, rel(ev[w, i, 0]), rel(ev[w, i, 1]), rel(ev[w, i, 2]), rel(ev[w, i, 4]), rel(ev[w, i, 5])) for i in range(2, NT)))
