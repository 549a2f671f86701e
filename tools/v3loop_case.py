"""One short v3 chain (B = 128, 50 steps) through unet3_loop_kernel: the target of the ncu captures under profiles/.
   ncu --set full --clock-control none -k regex:unet3_loop -c 1 python tools/v3loop_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import v3

B, STEPS = int(os.environ.get("CASE_B", "128")), int(os.environ.get("CASE_STEPS", "50"))
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = v3.ConditionalUNet(precision="bf16").to(dev).eval()
diff = v3.ConditionalDenoiseDiffusion(net, n_steps=1000, device=dev)
eng = diff._engine(dev)
f = torch.randint(0, 102, (B,), device=dev)
c = torch.randint(0, 10, (B,), device=dev)
x = eng.randn(B, 256, 1, 0, 1000)
eng.sample3(x, 999, 1000 - STEPS, f, c, seed=1, sample_offset=0, use_graph=False)
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(x).all()), "tc_error", int(eng.info("tc_error")))
