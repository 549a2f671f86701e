"""Time the persistent chain kernel alone for several batch sizes (clusters) and step counts.
    python tools/chain_sweep.py [--steps 200] [--batches 32,64,128,256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ldm_b200
from oracle import weights

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--batches", default="32,64,128,256")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
u = ldm_b200.ConditionalUNet(precision="bf16")
u.load_state_dict(weights.make_unet_state(42, "init"))
u = u.to(dev).eval()
d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, dev)
eng = u.engine(dev, 1000)
eng.set_schedule(*d._host_schedule)
print("chain", eng.info("chain"), "max co-resident clusters", eng.info("chain_max_clusters"),
      "peak bytes/step/CTA", eng.info("chain_peak_bytes_per_step"))
for B in [int(b) for b in a.batches.split(",")]:
    c = (torch.arange(B) % 102).to(dev)
    x = eng.randn(B, 256, 1, 0, 1000)
    eng.sample(x, 999, 1000 - a.steps, c, seed=3, use_graph=False)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(a.reps):
        x = eng.randn(B, 256, 1, 0, 1000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.sample(x, 999, 1000 - a.steps, c, seed=3, use_graph=False)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    eng.check_device_flags()
    print("B %4d  steps %d  %.3f ms  -> %.2f us/step  finite %s" % (B, a.steps, best, best * 1e3 / a.steps, bool(torch.isfinite(x).all())))

# ---- per-CTA timeline of one reverse step (cluster 0); needs an LDM_CHAIN_TRACE=1 build
import ctypes
import numpy as np
from ldm_b200 import _lib
L = _lib.lib()
B = int(os.environ.get("TRACE_B", "48"))
RANKS = [int(r) for r in os.environ.get("TRACE_RANKS", "0,1,4,12").split(",")]
TRACKS, LEN, CS = 5, 96, 16
TAG = {1: "iter", 2: "pre-acc-wait", 3: "acc-ready", 4: "acc-loaded(+partner)", 5: "dual-done", 6: "published-u", 7: "stats-u-in", 11: "stats-f-in",
       12: "traded", 13: "stats-done", 14: "group-bar", 15: "combined", 8: "operand-written", 9: "arrive-sent", 10: "handover-done", 20: "W-first-issued", 21: "W-last-issued", 30: "X-handover-seen", 31: "X-fenced",
       32: "X-first-issued", 33: "X-all-issued", 40: "MMA-first", 42: "MMA-half", 41: "MMA-all-issued"}
NAME = ["h", "u", "W", "X", "M"]
c = (torch.arange(B) % 102).to(dev)
x = eng.randn(B, 256, 1, 0, 1000)
_lib.check(L.ldm_debug_chain_trace(eng.ctx, 20, None, 0))
eng.sample(x, 999, 1000 - 40, c, seed=3, use_graph=False)
buf = np.zeros((CS, TRACKS, LEN), dtype=np.int64)
_lib.check(L.ldm_debug_chain_trace(eng.ctx, 0, buf.ctypes.data_as(ctypes.c_void_p), buf.size))
_lib.check(L.ldm_debug_chain_trace(eng.ctx, -1, None, 0))
MASK = (1 << 52) - 1
for r in RANKS:
    ev = []
    for w in range(TRACKS):
        for v in buf[r][w]:
            if v != 0:
                ev.append((int(v) & MASK, w, (int(v) >> 52) & 15, (int(v) >> 56) & 255))
    if not ev:
        continue
    ev.sort()
    t0 = min(e[0] for e in ev if e[1] == 0) if any(e[1] == 0 for e in ev) else ev[0][0]
    print("---- rank %d: %d events, step = %d cycles" % (r, len(ev), max(e[0] for e in ev if e[1] < 2) - t0))
    last = {}
    for t, w, p, tag in ev:
        print("  %7d  (+%5d)  %s  p%d  %s" % (t - t0, t - last.get(w, t), NAME[w], p, TAG.get(tag, str(tag))))
        last[w] = t
