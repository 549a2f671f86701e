"""Run a few reverse steps + one decode WITHOUT graph replay so that ncu sees every kernel of a step.
    python tools/profile_step.py [--precision bf16] [--batch 256] [--steps 3] [--graph]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ldm_b200
from oracle import weights

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--no-decode", action="store_true")
ap.add_argument("--v3", action="store_true")
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
if a.v3:
    u = ldm_b200.v3.ConditionalUNet(precision=a.precision)
    u.load_state_dict(weights.make_unet3_state(44, "init"))
else:
    u = ldm_b200.ConditionalUNet(precision=a.precision)
    u.load_state_dict(weights.make_unet_state(42, "init"))
u = u.to(dev).eval()
ae = ldm_b200.SimpleAutoencoder(precision=a.precision)
ae.load_state_dict(weights.make_autoencoder_state(43, "init"))
ae = ae.to(dev).eval()
d = (ldm_b200.v3 if a.v3 else ldm_b200).ConditionalDenoiseDiffusion(u, 1000, dev)
eng = u.engine(dev, 1000)
eng.set_schedule(*d._host_schedule)
c = (torch.arange(a.batch) % 102).to(dev)
x = eng.randn(a.batch, 256, 1, 0, 1000)
torch.cuda.synchronize()
import time
for rep in range(2):
    t0 = time.perf_counter()
    if a.v3:
        eng.sample3(x, 999, 1000 - a.steps, c, c % 10, seed=3, use_graph=a.graph)
    else:
        eng.sample(x, 999, 1000 - a.steps, c, seed=3, use_graph=a.graph)
    torch.cuda.synchronize()
    print("rep", rep, "steps", a.steps, "ms", (time.perf_counter() - t0) * 1e3)
if not a.no_decode:
    img = ae.decode(x)
    torch.cuda.synchronize()
    print("decode ok", float(img.mean()))
