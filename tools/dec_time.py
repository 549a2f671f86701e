"""Time the bf16 decode of B latents (CUDA events, 20 repetitions after 5 warm-ups) and print a checksum of the images.
    python tools/dec_time.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import philox  # noqa: E402
from tests._util import make_autoencoder  # noqa: E402

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ae = make_autoencoder("perturbed", "bf16")
z = torch.from_numpy(philox.normal_rows(3, 0, B, 0)).cuda()
img = ae.decode(z)
for _ in range(5):
    ae.decode(z)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ae.decode(z)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"B": B, "ms": e0.elapsed_time(e1) / 20, "sum": float(img.double().sum())}))
