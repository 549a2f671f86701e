"""Shared helpers of the parity tests."""
import numpy as np
import torch

from oracle import philox, weights

UNET_SEED, AE_SEED, NOISE_SEED = 42, 43, 1234     # must match oracle/make_golden.py

# tolerances stated by BASELINE.json north_star (eps) and SURVEY.md 8c (latents)
EPS_TOL = {"fp32": 1e-3, "bf16": 2e-2}            # max|d| / max|ref| of eps per step
LATENT_TOL = {"fp32": 5e-3, "bf16": 5e-2}         # relative L2 of x_0 after the full 1000-step loop
IMAGE_TOL = {"fp32": 2e-4, "bf16": 3e-2}          # max abs error of decoded images (values in (0, 1))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def make_unet(style, precision, device="cuda"):
    import ldm_b200
    m = ldm_b200.ConditionalUNet(precision=precision)
    m.load_state_dict(weights.make_unet_state(UNET_SEED, style), strict=True)
    return m.to(device).eval()


def make_autoencoder(style, precision, device="cuda"):
    import ldm_b200
    ae = ldm_b200.SimpleAutoencoder(precision=precision)
    ae.load_state_dict(weights.make_autoencoder_state(AE_SEED, style), strict=True)
    return ae.to(device).eval()


def chain_noise(seed, offset, batch, t_start, t_end=0, dim=256):
    """(t_start - t_end + 1, batch, dim) explicit draws, slab j for t = t_start - j (oracle/philox.py spec)."""
    return np.stack([philox.normal_rows(seed, offset, batch, t, dim) for t in range(t_start, t_end - 1, -1)])
