"""Generate tests/golden/ublock.npz from the UNMODIFIED v2 reference classes UNetResidualBlock (v2:462-486) and
UNetAttentionBlock (v2:434-459) (build container only):    python -m oracle.make_golden_ublock"""
import os

import numpy as np
import torch

from . import ref_loader, weights

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ublock.npz")
CASES_RES = [(64, 64, 16), (64, 128, 8)]      # (in, out, side)
CASES_ATTN = [(128, 8), (64, 16)]             # (channels, side): head_dim 32 with 64 tokens, head_dim 16 with 256 tokens


def main():
    torch.set_grad_enabled(False)
    m = ref_loader.load("v2")
    g = torch.Generator().manual_seed(99)
    out = {}
    for cin, cout, side in CASES_RES:
        sd = weights.make_state(weights.ublock_res_spec(cin, cout), 46, "perturbed")
        blk = m.UNetResidualBlock(cin, cout).eval()
        blk.load_state_dict(sd, strict=True)
        x = torch.randn(3, cin, side, side, generator=g)
        t, c = torch.randn(3, 256, generator=g), torch.randn(3, 256, generator=g)
        k = "res_%d_%d_" % (cin, cout)
        out[k + "x"], out[k + "t"], out[k + "c"] = x.numpy(), t.numpy(), c.numpy()
        out[k + "y_tc"] = blk(x, t, c).numpy()
        out[k + "y_t"] = blk(x, t).numpy()
    for ch, side in CASES_ATTN:
        sd = weights.make_state(weights.ublock_attn_spec(ch), 47, "perturbed")
        blk = m.UNetAttentionBlock(ch).eval()
        blk.load_state_dict(sd, strict=True)
        x = torch.randn(2, ch, side, side, generator=g) * 1.5
        k = "attn_%d_" % ch
        out[k + "x"] = x.numpy()
        out[k + "y"] = blk(x).numpy()
    np.savez_compressed(OUT, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
