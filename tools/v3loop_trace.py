"""Per-phase timeline of the v3 persistent loop kernel (LDM_V3LOOP_TRACE=1): for a few CTAs, the time each phase of step 1
took in that CTA and how long the CTA then waited at the grid barrier (ns).  Run on a B200: python tools/v3loop_trace.py"""
import os
import sys
import time

os.environ.setdefault("LDM_V3LOOP_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import v3

B = int(os.environ.get("TRACE_B", "128"))
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = v3.ConditionalUNet(precision="bf16").to(dev).eval()
diff = v3.ConditionalDenoiseDiffusion(net, n_steps=1000, device=dev)
f = torch.randint(0, 102, (B,), device=dev)
c = torch.randint(0, 10, (B,), device=dev)
for n in (1000, 1000):
    torch.cuda.synchronize()
    t0 = time.time()
    out = diff.sample((B, 256), dev, f, c, seed=1, use_graph=False)
    torch.cuda.synchronize()
    print("B", B, "steps", n, "ms", round((time.time() - t0) * 1e3, 2), "finite", bool(torch.isfinite(out).all()))
