"""The CPU restatement (oracle/restate.py) against the golden vectors produced by the reference itself
(oracle/make_golden.py).  Runs wherever the tests run; /root/reference is not needed."""
import numpy as np
import torch

from oracle import philox, restate as R, weights
from tests._util import AE_SEED, NOISE_SEED, T, UNET_SEED

torch.set_grad_enabled(False)
TOL = 2e-5   # BLAS blocking differs between hosts; the same host reproduces the goldens bit for bit


def test_schedule_bit_exact(golden, style):
    g = golden(style)
    beta, alpha, alpha_bar = R.schedule(1000)
    assert np.array_equal(beta.numpy(), g["beta"])
    assert np.array_equal(alpha.numpy(), g["alpha"])
    assert np.array_equal(alpha_bar.numpy(), g["alpha_bar"])


def test_embeddings(golden, style):
    g = golden(style)
    sd = weights.make_unet_state(UNET_SEED, style)
    te = R.time_embedding(sd, T(g["time_emb_rows"]))
    assert R.max_rel(te, g["time_emb"]) < TOL
    ce = R.class_embedding(sd, torch.arange(102))
    assert R.max_rel(ce, g["class_emb"]) < TOL


def test_unet_forward(golden, style):
    g = golden(style)
    sd = weights.make_unet_state(UNET_SEED, style)
    x, c = T(g["fwd_x"]), T(g["fwd_c"])
    for t in (999, 500, 1, 0):
        assert R.max_rel(R.unet_forward(sd, x, torch.tensor([t]), c), g["fwd_eps_t%d" % t]) < TOL
    assert R.max_rel(R.unet_forward(sd, x, T(g["fwd_tb"]), c), g["fwd_eps_tb"]) < TOL
    assert R.max_rel(R.unet_forward(sd, x, torch.tensor([500]), None), g["fwd_eps_noclass_t500"]) < TOL


def test_p_sample(golden, style):
    g = golden(style)
    sd = weights.make_unet_state(UNET_SEED, style)
    sched = R.schedule(1000)
    x, c = T(g["fwd_x"]), T(g["fwd_c"])
    assert R.max_rel(R.p_sample(sd, sched, x, 500, c, noise=T(g["ps_noise"])), g["ps_t500"]) < TOL
    assert R.max_rel(R.p_sample(sd, sched, x, 0, c), g["ps_t0"]) < TOL


def test_noise_stream_matches_golden_xT(golden, style):
    g = golden(style)
    assert np.array_equal(philox.normal_rows(NOISE_SEED, 0, 4, 1000), g["chain_xT"])


def test_full_chain(golden, style):
    """BASELINE config 1: B = 4, 1000 steps on the CPU."""
    g = golden(style)
    sd = weights.make_unet_state(UNET_SEED, style)
    sched = R.schedule(1000)
    keep = (999, 900, 500, 100, 1, 0)
    x0, kept = R.sample(sd, sched, T(g["chain_xT"]), T(g["chain_c"]),
                        noise_fn=lambda t: T(philox.normal_rows(NOISE_SEED, 0, 4, t)), keep=keep)
    for t in keep:
        assert R.rel_l2(kept[t], g["chain_x_after_t%d" % t]) < 1e-3, t
    assert R.rel_l2(x0, g["chain_x0"]) < 1e-3


def test_partial_chain(golden, style):
    g = golden(style)
    sd = weights.make_unet_state(UNET_SEED, style)
    x0, _ = R.sample(sd, R.schedule(1000), T(g["partial_x_start"]), T(g["chain_c"]),
                     noise_fn=lambda t: T(philox.normal_rows(NOISE_SEED + 3, 0, 4, t)), t_start=120)
    assert R.rel_l2(x0, g["partial_x0"]) < 1e-3


def test_decode(golden, style):
    g = golden(style)
    sd = weights.make_autoencoder_state(AE_SEED, style)
    img = R.decode(sd, T(g["dec_z"]))
    assert img.shape == (2, 3, 64, 64)
    assert float((img - T(g["dec_img"])).abs().max()) < 1e-5
    img2 = R.decode(sd, T(g["chain_x0"]))
    assert float((img2 - T(g["dec_img_chain"])).abs().max()) < 1e-4
