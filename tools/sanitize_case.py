"""The tiny workload that runs under compute-sanitizer (one --tool per gpurun call, B200_PROFILING.md):
chain_kernel<2> (4 rows, 3 reverse steps + one forward), attn_tc_kernel (v3 forward, 8 rows), conv_tc_kernel /
conv_halo_kernel / pix_conv_in_tc_kernel (one v4 forward at 64 x 64) and the decoder (1 latent).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ldm_b200
from ldm_b200 import v3, v4
from oracle import weights

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
u = ldm_b200.ConditionalUNet(precision="bf16")
u.load_state_dict(weights.make_unet_state(42, "init"))
u = u.to(dev).eval()
d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, dev)
eng = u.engine(dev, 1000)
eng.set_schedule(*d._host_schedule)
c = torch.tensor([0, 33, 67, 101], device=dev)
x = eng.randn(4, 256, 1, 0, 1000)
eng.sample(x, 999, 997, c, seed=3, use_graph=False)
eps = u(x, torch.tensor([5], device=dev), c)
torch.cuda.synchronize()
print("chain ok", bool(torch.isfinite(x).all()), bool(torch.isfinite(eps).all()))

u3 = v3.ConditionalUNet(precision="bf16")
u3.load_state_dict(weights.make_unet3_state(44, "init"))
u3 = u3.to(dev).eval()
f, k = torch.arange(8, device=dev) % 102, torch.arange(8, device=dev) % 10
e3 = u3(torch.randn(8, 256, device=dev), torch.tensor([7], device=dev), f, k)
torch.cuda.synchronize()
print("v3 ok", bool(torch.isfinite(e3).all()))

m = v4.SimpleUNet()
m.load_state_dict(weights.make_pix_state(45, "init"))
m = m.to(dev).eval()
e4 = m(torch.randn(1, 3, 64, 64, device=dev), torch.tensor([9.0], device=dev))
torch.cuda.synchronize()
print("v4 ok", bool(torch.isfinite(e4).all()))

ae = ldm_b200.SimpleAutoencoder(precision="bf16")
ae.load_state_dict(weights.make_autoencoder_state(43, "init"))
ae = ae.to(dev).eval()
img = ae.decode(x[:1])
torch.cuda.synchronize()
eng.check_device_flags()
print("decode ok", bool(torch.isfinite(img).all()))
