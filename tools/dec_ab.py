"""A/B of a decoder build switch: decode the same latents in two fresh processes (the switches are read once per process)
and compare the images, then print the traced per-kernel times of both.

    python tools/dec_ab.py LDM_DEC_FUSE_OUT 0 1 [B]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
from tests._util import make_autoencoder
from oracle import philox
torch.set_grad_enabled(False)
B = %(B)d
ae = make_autoencoder("perturbed", "bf16")
z = torch.from_numpy(philox.normal_rows(3, 0, B, 0)).cuda()
img = ae.decode(z)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3): ae.decode(z)
ev[0].record()
for _ in range(10): ae.decode(z)
ev[1].record(); torch.cuda.synchronize()
np.save(%(out)r, img.cpu().numpy())
print(json.dumps({"ms": ev[0].elapsed_time(ev[1]) / 10}))
"""


def run(var, val, B, out):
    env = dict(os.environ)
    env[var] = val
    src = CHILD % {"root": ROOT, "B": B, "out": out}
    r = subprocess.run([sys.executable, "-c", src], env=env, capture_output=True, text=True, cwd=ROOT)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-4000:])
        raise SystemExit("child failed (%s=%s)" % (var, val))
    return json.loads(r.stdout.strip().splitlines()[-1])


def main():
    import numpy as np
    var, a, b = sys.argv[1], sys.argv[2], sys.argv[3]
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    fa, fb = os.path.join(ROOT, "gpurun_out", "ab_a.npy"), os.path.join(ROOT, "gpurun_out", "ab_b.npy")
    ra, rb = run(var, a, B, fa), run(var, b, B, fb)
    xa, xb = np.load(fa), np.load(fb)
    print(json.dumps({"var": var, "B": B, a: ra, b: rb, "max_abs_diff": float(np.abs(xa - xb).max()),
                      "finite": bool(np.isfinite(xb).all()), "range_b": [float(xb.min()), float(xb.max())]}))
    os.remove(fa); os.remove(fb)


if __name__ == "__main__":
    main()
