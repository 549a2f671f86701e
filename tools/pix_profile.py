"""Timing / profiling driver of the v4 pixel-space path: `--steps N` reverse steps at `--batch B` images of 64 x 64
(CUDA-event timed, graph replay), optionally one eager forward for an ncu launch list (`--forward`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ldm_b200
from ldm_b200 import v4
from oracle import weights

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--forward", action="store_true")
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
torch.set_grad_enabled(False)
m = v4.SimpleUNet()
m.load_state_dict(weights.make_pix_state(45, "init"))
m = m.to("cuda").eval()
d = v4.DiffusionModel(m, 1000, device="cuda")
eng = d._engine("cuda")
x = eng.randn(a.batch, 3 * 64 * 64, 1, 0, 1000).view(a.batch, 3, 64, 64)
if a.forward:
    t = torch.full((a.batch,), 500.0, device="cuda")
    for _ in range(2):
        m(x, t)
    torch.cuda.synchronize()
for rep in range(a.reps):
    xs = x.clone()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    eng.pix_sample(xs, 999, 1000 - a.steps, seed=3, use_graph=not a.no_graph)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    flop = 5.63e9 * a.batch * a.steps
    print("rep %d: %d steps B=%d: %.3f ms total, %.1f us/step, %.1f TFLOP/s" % (rep, a.steps, a.batch, ms, ms * 1e3 / a.steps, flop / ms / 1e9))
