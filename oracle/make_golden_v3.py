"""Generate tests/golden/v3_*.npz by running the UNMODIFIED v3 reference (build container only).

    python -m oracle.make_golden_v3

Every array is an output of /root/reference/v3/model_train_test.py (ConditionalUNet.forward v3:804-853 and
ConditionalDenoiseDiffusion.p_sample v3:874-887) loaded with the deterministic weights of oracle/weights.py, on CPU,
eval mode, no_grad; the per-step noise comes from oracle/philox.py and replaces the reference's torch.randn_like."""
import os

import numpy as np
import torch

from . import philox, ref_loader, weights

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED, NOISE_SEED, B, T_START = 44, 4321, 6, 24


def main():
    torch.set_grad_enabled(False)
    m = ref_loader.load("v3")
    for style in ("init", "perturbed"):
        sd = weights.make_unet3_state(SEED, style)
        net = m.ConditionalUNet().eval()
        net.load_state_dict(sd, strict=True)
        diff = m.ConditionalDenoiseDiffusion(net, n_steps=1000, device=torch.device("cpu"))
        x = torch.from_numpy(philox.normal_rows(NOISE_SEED, 0, B, 1000))
        flower = torch.tensor([0, 101, 50, 7, 7, 33])
        color = torch.tensor([0, 9, 4, 1, 2, 1])
        out = {"x": x.numpy(), "flower": flower.numpy(), "color": color.numpy()}
        for t in (0, 1, 500, 999):
            out["eps_t%d" % t] = net(x, torch.tensor([t]), flower, color).numpy()
        tb = torch.tensor([999, 0, 500, 17, 17, 250])
        out["tb"] = tb.numpy()
        out["eps_tb"] = net(x, tb, flower, color).numpy()
        out["eps_t500_first3"] = net(x[:3], torch.tensor([500]), flower[:3], color[:3]).numpy()   # batch-coupled: differs from eps_t500[:3]
        # 25 reverse steps through the reference's own p_sample with Philox noise
        step = {"t": T_START}
        real = m.torch.randn_like
        m.torch.randn_like = lambda v: torch.from_numpy(philox.normal_rows(NOISE_SEED + 1, 0, v.shape[0], step["t"], v.shape[1]))
        try:
            xs = x.clone()
            for t in range(T_START, -1, -1):
                step["t"] = t
                xs = diff.p_sample(xs, t, flower, color)
        finally:
            m.torch.randn_like = real
        out["chain_x0"] = xs.numpy()
        os.makedirs(OUT, exist_ok=True)
        np.savez_compressed(os.path.join(OUT, "v3_%s.npz" % style), **out)
        print(style, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
