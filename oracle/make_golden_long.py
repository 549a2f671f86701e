"""Generate the long-chain goldens (tests/golden/v2_tail.npz, v3_long.npz, v4_long.npz) by running the UNMODIFIED
reference scripts (build container only).  TEST INFRASTRUCTURE: never imported by the product package.

    python -m oracle.make_golden_long

* v2_tail: the reference's own `sample()` (v2:594-598) for GLOBAL samples 252..255 of the BASELINE configs[1] batch
  (B = 256, class b mod 102, noise stream of sample_offset 252), 1000 steps, both weight styles.  The reference is
  row-independent, so these four rows are what rows 252..255 of a B = 256 call must produce: they sit
  in the ragged last cluster of the chain kernel (rows 240..255 of 6 x 48 slots).
* v3_long: 1000 steps of the v3 reference `p_sample` (v3:874-887) at B = 128, the per-GPU call size of configs[3]
  (attention couples the rows of a call, so the whole call is the unit).
* v4_long: 100 steps of the v4 reference `p_sample` (v4:155-168) at 64 x 64, B = 2, from t = 99.
Noise comes from oracle/philox.py (the kernels' stream) and replaces the reference's torch.randn / randn_like."""
import os

import numpy as np
import torch

from . import philox, ref_loader, weights
from .make_golden import NOISE_SEED, UNET_SEED, run_reference_chain
from .make_golden_pix import image_noise

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TAIL_OFFSET = 252
V3_SEED, V3_NOISE, V3_B = 44, 4321, 128
V4_SEED, V4_NOISE, V4_B, V4_T = 45, 777, 2, 99


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)

    m = ref_loader.load()
    out = {"offset": np.asarray(TAIL_OFFSET)}
    for style in ("init", "perturbed"):
        unet = m.ConditionalUNet().eval()
        unet.load_state_dict(weights.make_unet_state(UNET_SEED, style), strict=True)
        diffusion = m.ConditionalDenoiseDiffusion(unet, n_steps=1000, device=None)
        c = (torch.arange(TAIL_OFFSET, TAIL_OFFSET + 4) % 102)
        x0, _ = run_reference_chain(m, diffusion, 4, c, NOISE_SEED, TAIL_OFFSET)
        out["c"] = c.numpy()
        out["x0_%s" % style] = x0.numpy()
        print("v2 tail", style, "x0 std", float(x0.std()))
    np.savez_compressed(os.path.join(OUT, "v2_tail.npz"), **out)

    m3 = ref_loader.load("v3")
    net = m3.ConditionalUNet().eval()
    net.load_state_dict(weights.make_unet3_state(V3_SEED, "init"), strict=True)
    diff = m3.ConditionalDenoiseDiffusion(net, n_steps=1000, device=torch.device("cpu"))
    flower, color = torch.arange(V3_B) % 102, (torch.arange(V3_B) * 7) % 10
    x = torch.from_numpy(philox.normal_rows(V3_NOISE, 0, V3_B, 1000))
    step = {"t": 999}
    real = m3.torch.randn_like
    m3.torch.randn_like = lambda v: torch.from_numpy(philox.normal_rows(V3_NOISE, 0, v.shape[0], step["t"], v.shape[1]))
    kept = {}
    try:
        for t in range(999, -1, -1):
            step["t"] = t
            x = diff.p_sample(x, t, flower, color)
            if t in (900, 500):
                kept[t] = x.clone()
    finally:
        m3.torch.randn_like = real
    np.savez_compressed(os.path.join(OUT, "v3_long.npz"), flower=flower.numpy(), color=color.numpy(), x0=x.numpy(),
                        x_after_t900=kept[900].numpy(), x_after_t500=kept[500].numpy())
    print("v3 long x0 std", float(x.std()))

    m4 = ref_loader.load("v4")
    net = m4.SimpleUNet().eval()
    net.load_state_dict(weights.make_pix_state(V4_SEED, "perturbed"), strict=True)
    diff = m4.DiffusionModel(net, n_steps=1000, device="cpu")
    x = image_noise(V4_NOISE + 5, 0, V4_B, 1000, (3, 64, 64))
    xs = x.clone()
    step = {"t": V4_T}
    real = m4.torch.randn_like
    m4.torch.randn_like = lambda v: image_noise(V4_NOISE + 6, 0, v.shape[0], step["t"], tuple(v.shape[1:]))
    try:
        for t in range(V4_T, -1, -1):
            step["t"] = t
            xs = diff.p_sample(xs, t)
    finally:
        m4.torch.randn_like = real
    np.savez_compressed(os.path.join(OUT, "v4_long.npz"), x=x.numpy(), x0=xs.numpy())
    print("v4 long x0 std", float(xs.std()))


if __name__ == "__main__":
    main()
