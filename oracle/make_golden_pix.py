"""Generate tests/golden/v4_*.npz and v5_*.npz by running the UNMODIFIED v4 / v5 reference (build container only).

    python -m oracle.make_golden_pix

Every array is an output of /root/reference/v{4,5}/model_train_test.py (SimpleUNet.forward v4:99-135 / v5:101-146 and
DiffusionModel.p_sample v4:155-168) loaded with the deterministic weights of oracle/weights.py, on CPU, eval mode,
no_grad; the per-step noise comes from oracle/philox.py and replaces the reference's torch.randn_like."""
import os

import numpy as np
import torch

from . import philox, ref_loader, weights

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED, NOISE_SEED, B, T_START = 45, 777, 2, 5


def image_noise(seed, offset, n, step, shape):
    return torch.from_numpy(philox.normal_rows(seed, offset, n, step, shape[0] * shape[1] * shape[2])).view(n, *shape)


def main():
    torch.set_grad_enabled(False)
    for ver, style in (("v4", "init"), ("v4", "perturbed"), ("v5", "perturbed")):
        m = ref_loader.load(ver)
        sd = weights.make_pix_state(SEED, style, v5=(ver == "v5"))
        net = m.SimpleUNet().eval()
        net.load_state_dict(sd, strict=True)
        diff = m.DiffusionModel(net, n_steps=1000, device="cpu")
        x = image_noise(NOISE_SEED, 0, B, 1000, (3, 64, 64))
        out = {"x": x.numpy()}
        ta, tb = torch.tensor([999, 3]), torch.tensor([0, 500])
        out["ta"], out["tb"] = ta.numpy(), tb.numpy()
        out["eps_ta"] = net(x, ta).numpy()
        out["eps_tb"] = net(x, tb).numpy()
        xs32 = image_noise(NOISE_SEED + 2, 0, 3, 1000, (3, 32, 32))       # another resolution, odd batch
        out["x32"] = xs32.numpy()
        out["eps32_t250"] = net(xs32, torch.full((3,), 250)).numpy()
        step = {"t": T_START}
        real = m.torch.randn_like
        m.torch.randn_like = lambda v: image_noise(NOISE_SEED + 1, 0, v.shape[0], step["t"], tuple(v.shape[1:]))
        try:
            xs = x.clone()
            for t in range(T_START, -1, -1):
                step["t"] = t
                xs = diff.p_sample(xs, t)
        finally:
            m.torch.randn_like = real
        out["chain_x0"] = xs.numpy()
        os.makedirs(OUT, exist_ok=True)
        np.savez_compressed(os.path.join(OUT, "%s_%s.npz" % (ver, style)), **out)
        print(ver, style, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
