"""Time unet3_loop_kernel alone: B = 128, 500 reverse steps, best of 3 (CUDA events).  python tools/v3loop_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import v3

B, STEPS = int(os.environ.get("CASE_B", "128")), int(os.environ.get("CASE_STEPS", "500"))
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = v3.ConditionalUNet(precision="bf16").to(dev).eval()
diff = v3.ConditionalDenoiseDiffusion(net, n_steps=1000, device=dev)
eng = diff._engine(dev)
f = torch.randint(0, 102, (B,), device=dev)
c = torch.randint(0, 10, (B,), device=dev)
best, chk = 1e9, None
for rep in range(4):
    x = eng.randn(B, 256, 1, 0, 1000)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    eng.sample3(x, 999, 1000 - STEPS, f, c, seed=1, sample_offset=0, use_graph=False)
    e.record()
    torch.cuda.synchronize()
    if rep:
        best = min(best, s.elapsed_time(e))
    chk = float(x.double().sum())
print("B %d steps %d: %.3f ms -> %.2f us/step  checksum %.6f  tc_error %d" % (B, STEPS, best, best * 1000 / STEPS, chk, int(eng.info("tc_error"))))
