"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Every array is an output of /root/reference/v2/model_train_test.py classes
(ConditionalUNet, ConditionalDenoiseDiffusion, SimpleAutoencoder) loaded with
the deterministic weights of oracle/weights.py, on CPU, eval mode, no_grad.
Noise (x_T and the per-step draws of v2:589) comes from oracle/philox.py and is
fed to the reference by temporarily replacing its torch.randn / randn_like, in
the reference's draw order (v2:595 first, then one draw per step after the
model forward).  The arrays are stored so that tests never need the reference.
"""
import os
import sys

import numpy as np
import torch

from . import philox, ref_loader, weights

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

UNET_SEED, AE_SEED, NOISE_SEED = 42, 43, 1234
EPS_T = (999, 500, 1, 0)
KEEP_T = (999, 900, 500, 100, 1, 0)
TE_ROWS = (0, 1, 2, 10, 100, 500, 998, 999)


class _Noise:
    """Stands in for the reference's torch.randn / randn_like (v2:589,595)."""

    def __init__(self, seed, offset, n_steps, start_step):
        self.seed, self.offset, self.step = seed, offset, start_step
        self.n_steps = n_steps

    def randn(self, shape, device=None):
        z = philox.normal_rows(self.seed, self.offset, shape[0], self.n_steps, shape[1])
        return torch.from_numpy(z)

    def randn_like(self, x):
        z = philox.normal_rows(self.seed, self.offset, x.shape[0], self.step, x.shape[1])
        self.step -= 1
        return torch.from_numpy(z)


def run_reference_chain(m, diffusion, B, c, seed, offset, t_start=None, x_start=None, keep=()):
    """Drive the reference's own sample()/p_sample() with Philox noise."""
    n = diffusion.n_steps
    noise = _Noise(seed, offset, n, (n - 1) if t_start is None else t_start)
    real_randn, real_like = m.torch.randn, m.torch.randn_like
    kept = {}
    try:
        m.torch.randn_like = noise.randn_like
        if t_start is None:
            # full chain through the reference's sample() (v2:594-598); record via p_sample wrapper
            m.torch.randn = noise.randn
            orig = diffusion.p_sample
            def spy(x, t, cc=None):
                y = orig(x, t, cc)
                if t in keep:
                    kept[t] = y.clone()
                return y
            diffusion.p_sample = spy
            x = diffusion.sample((B, 256), None, c)
            del diffusion.p_sample
        else:
            x = x_start
            for t in range(t_start, -1, -1):
                # explicit tensor t, as visualize_denoising_steps does (v2:689-690)
                x = diffusion.p_sample(x, torch.tensor([t]), c)
                if t in keep:
                    kept[t] = x.clone()
    finally:
        m.torch.randn, m.torch.randn_like = real_randn, real_like
    return x, kept


def main():
    if not ref_loader.available():
        sys.exit("reference not present; golden vectors can only be generated in the build container")
    m = ref_loader.load()
    torch.set_grad_enabled(False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)

    for style in ("init", "perturbed"):
        out = {}
        sd = weights.make_unet_state(UNET_SEED, style)
        unet = m.ConditionalUNet().eval()
        unet.load_state_dict(sd, strict=True)
        diffusion = m.ConditionalDenoiseDiffusion(unet, n_steps=1000, device=None)
        out["beta"], out["alpha"], out["alpha_bar"] = (v.numpy() for v in (diffusion.beta, diffusion.alpha, diffusion.alpha_bar))

        # --- single forwards (a4-a10)
        x = torch.from_numpy(philox.normal_rows(NOISE_SEED + 1, 0, 8, 7)) * 3.0
        c = (torch.arange(8) * 13 + 5) % 102
        out["fwd_x"], out["fwd_c"] = x.numpy(), c.numpy()
        for t in EPS_T:
            out[f"fwd_eps_t{t}"] = unet(x, torch.tensor([t]), c).numpy()
        tb = torch.tensor([3, 999, 0, 17, 500, 1, 2, 998])
        out["fwd_tb"] = tb.numpy()
        out["fwd_eps_tb"] = unet(x, tb, c).numpy()
        out["fwd_eps_noclass_t500"] = unet(x, torch.tensor([500]), None).numpy()
        out["time_emb_rows"] = np.asarray(TE_ROWS)
        out["time_emb"] = unet.time_emb(torch.tensor(TE_ROWS)).numpy()
        out["class_emb"] = unet.class_emb(torch.arange(102)).numpy()

        # --- one p_sample with noise and the t = 0 branch (a3)
        nz = torch.from_numpy(philox.normal_rows(NOISE_SEED + 2, 0, 8, 500))
        real_like = m.torch.randn_like
        m.torch.randn_like = lambda t_: nz
        out["ps_noise"] = nz.numpy()
        out["ps_t500"] = diffusion.p_sample(x, 500, c).numpy()
        m.torch.randn_like = real_like
        out["ps_t0"] = diffusion.p_sample(x, 0, c).numpy()

        # --- full 1000-step chain, B = 4 (BASELINE config 1), global samples 0..3 (a2)
        B = 4
        cc = torch.tensor([0, 33, 67, 101])
        out["chain_c"] = cc.numpy()
        out["chain_xT"] = philox.normal_rows(NOISE_SEED, 0, B, 1000)
        x0, kept = run_reference_chain(m, diffusion, B, cc, NOISE_SEED, 0, keep=KEEP_T)
        for t, v in kept.items():
            out[f"chain_x_after_t{t}"] = v.numpy()
        out["chain_x0"] = x0.numpy()
        # --- partial chain from t = 120 with tensor t (visualize_denoising_steps, v2:679-692)
        xs = torch.from_numpy(out["chain_xT"]) * 0.5 + 0.25
        out["partial_x_start"] = xs.numpy()
        xp, _ = run_reference_chain(m, diffusion, B, cc, NOISE_SEED + 3, 0, t_start=120, x_start=xs)
        out["partial_x0"] = xp.numpy()

        # --- decode (a11-a15)
        sda = weights.make_autoencoder_state(AE_SEED, style)
        ae = m.SimpleAutoencoder().eval()
        ae.load_state_dict(sda, strict=True)
        z = torch.from_numpy(philox.normal_rows(NOISE_SEED + 4, 0, 2, 0))
        out["dec_z"] = z.numpy()
        out["dec_img"] = ae.decode(z).numpy()
        out["dec_img_chain"] = ae.decode(x0).numpy()        # decode of the (huge-magnitude) chain latents

        path = os.path.join(OUT, f"v2_{style}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB", "x0 std", float(x0.std()))


if __name__ == "__main__":
    main()
