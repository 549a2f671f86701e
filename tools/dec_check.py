"""Decoder check on the GPU: bf16 tensor-core decode vs the CPU restatement (B = 5, perturbed weights), then timing at B = 256."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ldm_b200
from oracle import philox, restate as R, weights
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for style in ("perturbed", "init"):
    sd = weights.make_autoencoder_state(43, style)
    ae = ldm_b200.SimpleAutoencoder(precision=prec)
    ae.load_state_dict(sd, strict=True)
    ae = ae.to(dev).eval()
    z = torch.from_numpy(philox.normal_rows(3, 0, 5, 0))
    want = R.decode(sd, z)
    got = ae.decode(z.to(dev)).cpu()
    print(style, prec, "image max-abs err %.3e  mean-abs %.3e  finite %s" % (float((got - want).abs().max()), float((got - want).abs().mean()), bool(torch.isfinite(got).all())))
    eng = ldm_b200.get_engine(dev, prec)
    eng.check_device_flags()
z = torch.randn(256, 256, device=dev)
for _ in range(3):
    ae.decode(z)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ae.decode(z)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print("decode B=256: %.3f ms  -> %.1f TFLOP/s" % (ms, 2809570560 * 256 / ms / 1e9))
