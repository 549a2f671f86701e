"""v4 / v5 pixel-space diffusion (BASELINE config 5): the oracle (oracle/restate_pix.py) pinned to the
reference's own outputs (tests/golden/v4_*.npz, v5_*.npz from the unmodified scripts; live reference when present), and
parity of the sm_100a path through the module mirror / C ABI in both modes: eps max|d|/max|ref| <= 1e-3 (strict fp32) and
<= 2e-2 (bf16, tcgen05) as north_star states; images after a chain within relative L2 5e-3 / 5e-2; Philox indexing exact."""
import os

import numpy as np
import pytest
import torch

from oracle import philox, ref_loader, restate as R, restate_pix as P, weights
from tests._util import EPS_TOL, LATENT_TOL, T

torch.set_grad_enabled(False)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED, NOISE_SEED, T_START = 45, 777, 5
CASES = [("v4", "init"), ("v4", "perturbed"), ("v5", "perturbed")]


def gold(ver, style):
    return np.load(os.path.join(GOLD, "%s_%s.npz" % (ver, style)))


def image_noise(seed, offset, n, step, shape=(3, 64, 64)):
    return torch.from_numpy(philox.normal_rows(seed, offset, n, step, shape[0] * shape[1] * shape[2])).view(n, *shape)


# ----------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("ver,style", CASES)
def test_pix_restatement_reproduces_reference_goldens(ver, style):
    g = gold(ver, style)
    sd = weights.make_pix_state(SEED, style, v5=(ver == "v5"))
    x = T(g["x"])
    assert torch.equal(x, image_noise(NOISE_SEED, 0, 2, 1000))
    assert R.max_rel(P.unet_forward(sd, x, T(g["ta"])), T(g["eps_ta"])) < 1e-5
    assert R.max_rel(P.unet_forward(sd, x, T(g["tb"])), T(g["eps_tb"])) < 1e-5
    assert R.max_rel(P.unet_forward(sd, T(g["x32"]), torch.full((3,), 250)), T(g["eps32_t250"])) < 1e-5
    x0 = P.sample(sd, R.schedule(1000), x, noise_fn=lambda t: image_noise(NOISE_SEED + 1, 0, 2, t), t_start=T_START)
    assert R.rel_l2(x0, T(g["chain_x0"])) < 1e-5


@pytest.mark.skipif(not ref_loader.available("v4"), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("ver", ["v4", "v5"])
def test_pix_restatement_against_the_live_reference(ver):
    m = ref_loader.load(ver)
    sd = weights.make_pix_state(SEED, "perturbed", v5=(ver == "v5"))
    net = m.SimpleUNet().eval()
    net.load_state_dict(sd, strict=True)
    torch.manual_seed(9)
    x, t = torch.randn(2, 3, 32, 32), torch.tensor([17, 803])
    assert torch.equal(net(x, t), P.unet_forward(sd, x, t))
    d = m.DiffusionModel(net, n_steps=1000, device="cpu")
    sched = R.schedule(1000)
    assert all(torch.equal(a, b) for a, b in zip((d.beta, d.alpha, d.alpha_bar), sched))
    z = torch.randn(2, 3, 32, 32)
    real = m.torch.randn_like
    m.torch.randn_like = lambda v: z
    try:
        assert torch.equal(d.p_sample(x, 400), P.p_sample(sd, sched, x, 400, z))
        assert torch.equal(d.p_sample(x, 0), P.p_sample(sd, sched, x, 0))          # t = 0 adds no noise (v4:162-167)
    finally:
        m.torch.randn_like = real
    t = torch.tensor([5, 900])
    assert torch.equal(d.q_sample(x, t, z), P.q_sample(sched, x, t, z))


def test_pix_mirror_state_dict_layout():
    from ldm_b200 import v4
    for v5 in (False, True):
        sd = weights.make_pix_state(SEED, "init", v5=v5)
        m = v4.SimpleUNet(res_ratio=v5)
        assert list(m.state_dict().keys()) == list(sd.keys())
        assert all(tuple(m.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
        m.load_state_dict(sd, strict=True)
    m = v4.SimpleUNet()                        # a v5 checkpoint switches the residual ratio on
    m.load_state_dict(weights.make_pix_state(SEED, "perturbed", v5=True), strict=True)
    assert "res_ratio" in m.state_dict()
    with pytest.raises(RuntimeError):          # no CPU fallback
        m.eval()(torch.zeros(1, 3, 64, 64), torch.zeros(1))
    d = v4.DiffusionModel(m, 1000, device="cpu")
    sched = R.schedule(1000)
    assert all(torch.equal(a, b) for a, b in zip((d.beta, d.alpha, d.alpha_bar), sched))
    x, z, t = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8), torch.tensor([1, 999])
    assert torch.equal(d.q_sample(x, t, z), P.q_sample(sched, x, t, z))


# ----------------------------------------------------------------------------- GPU: parity through the C ABI
def _model(ver, style, precision="bf16"):
    from ldm_b200 import v4
    m = v4.SimpleUNet(res_ratio=(ver == "v5"), precision=precision)
    m.load_state_dict(weights.make_pix_state(SEED, style, v5=(ver == "v5")), strict=True)
    return m.to("cuda").eval()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ver,style", CASES)
def test_pix_forward_against_reference_goldens(ver, style, precision):
    g = gold(ver, style)
    m = _model(ver, style, precision)
    x = T(g["x"]).cuda()
    for tk, ek in (("ta", "eps_ta"), ("tb", "eps_tb")):
        e = R.max_rel(m(x, T(g[tk]).cuda()).cpu(), T(g[ek]))
        assert e < EPS_TOL[precision], (tk, e)
    e = R.max_rel(m(T(g["x32"]).cuda(), torch.full((3,), 250, device="cuda")).cpu(), T(g["eps32_t250"]))
    assert e < EPS_TOL[precision], e
    # (B, 1) float timesteps are what the reference's own view(B, 1).float() produces (v4:104)
    e = R.max_rel(m(x, T(g["ta"]).cuda().view(2, 1).float()).cpu(), T(g["eps_ta"]))
    assert e < EPS_TOL[precision], e


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ver,style", CASES)
def test_pix_chain_against_reference_goldens(ver, style, precision):
    from ldm_b200 import v4
    g = gold(ver, style)
    m = _model(ver, style, precision)
    d = v4.DiffusionModel(m, 1000, device="cuda")
    x = T(g["x"]).cuda()
    noise = torch.stack([image_noise(NOISE_SEED + 1, 0, 2, t) for t in range(T_START, -1, -1)]).cuda()
    eng = d._engine("cuda")
    for use_graph in (False, True):
        xs = x.clone()
        eng.pix_sample(xs, T_START, 0, noise=noise, use_graph=use_graph)
        assert R.rel_l2(xs.cpu(), T(g["chain_x0"])) < LATENT_TOL[precision], (use_graph, R.rel_l2(xs.cpu(), T(g["chain_x0"])))
    # in-kernel Philox == the spec's stream (seed NOISE_SEED + 1, global sample index, step t)
    xs = x.clone()
    eng.pix_sample(xs, T_START, 0, seed=NOISE_SEED + 1, sample_offset=0, use_graph=True)
    assert R.rel_l2(xs.cpu(), T(g["chain_x0"])) < LATENT_TOL[precision]
    # p_sample with explicit noise = one step of the reference (v4:155-168); t = 0 adds none
    sd = weights.make_pix_state(SEED, style, v5=(ver == "v5"))
    sched = R.schedule(1000)
    one = d.p_sample(x, 700, noise=noise[0])
    step_tol = 5e-3 if precision == "bf16" else 1e-5                                          # bf16 eps error x c2(t)
    assert R.rel_l2(one.cpu(), P.p_sample(sd, sched, T(g["x"]), 700, noise[0].cpu())) < step_tol
    zero = d.p_sample(x, 0, noise=noise[0])
    assert R.rel_l2(zero.cpu(), P.p_sample(sd, sched, T(g["x"]), 0)) < step_tol


@pytest.mark.gpu
def test_pix_sharding_and_sample_entry():
    """Samples are independent: a shard with sample_offset reproduces its rows of the full batch bit for bit; the public
    sample() draws x_T from the Philox stream at step n_steps."""
    from ldm_b200 import v4
    m = _model("v4", "init")
    d = v4.DiffusionModel(m, 20, device="cuda")          # a short schedule keeps the test quick
    full = d.sample((5, 3, 32, 32), seed=31)
    part = d.sample((3, 3, 32, 32), seed=31, sample_offset=2)
    assert torch.equal(full[2:], part)
    x_T = image_noise(31, 0, 5, 20, (3, 32, 32))
    sd = weights.make_pix_state(SEED, "init")
    sched = R.schedule(20)
    want = P.sample(sd, sched, x_T, noise_fn=lambda t: image_noise(31, 0, 5, t, (3, 32, 32)))
    assert R.rel_l2(full.cpu(), want) < LATENT_TOL["bf16"], R.rel_l2(full.cpu(), want)
    frames = d.sample_with_intermediates((1, 3, 32, 32), [15, 3, 0], seed=4)
    assert len(frames) == 3 and frames[0].shape == (32, 32, 3) and float(frames[0].min()) >= 0.0
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 30, 30, device="cuda"), torch.zeros(1, device="cuda"))      # H, W must tile


@pytest.mark.gpu
def test_pix_full_size_batch_properties():
    """BASELINE config 5's per-GPU share (64 images of 64 x 64): finite outputs, and rows of the big batch equal the same
    rows computed in a small batch (size-independent property: no cross-sample coupling)."""
    m = _model("v4", "init")
    x = image_noise(5, 0, 64, 1000).cuda()
    t = torch.arange(64, device="cuda") * 15
    big = m(x, t)
    assert torch.isfinite(big).all()
    small = m(x[40:44], t[40:44])
    assert torch.equal(big[40:44], small)


@pytest.mark.gpu
def test_pix_boundary_errors():
    """The pixel path fails loudly instead of degrading: channel counts off the 64 grid, training mode, bad arguments."""
    import ldm_b200
    from ldm_b200 import v4
    small = v4.SimpleUNet(base_channels=32).cuda().eval()
    with pytest.raises(ldm_b200.LdmError):
        small(torch.zeros(1, 3, 64, 64, device="cuda"), torch.zeros(1, device="cuda"))
    ok = _model("v4", "init")
    with pytest.raises(RuntimeError):
        ok.train()(torch.zeros(1, 3, 64, 64, device="cuda"), torch.zeros(1, device="cuda"))
    with pytest.raises(RuntimeError):        # t must have one entry per sample, as the reference's view(B, 1) demands
        ok.eval()(torch.zeros(2, 3, 64, 64, device="cuda"), torch.zeros(3, device="cuda"))
    d = v4.DiffusionModel(ok, 1000, device="cuda")
    with pytest.raises(IndexError):
        d.p_sample(torch.zeros(1, 3, 64, 64, device="cuda"), 1000)


@pytest.mark.gpu
def test_pix_repacks_for_every_new_module():
    """Short-lived modules with alternating weights: a new module allocated on the address / storage of a collected one
    must be re-packed (the engine keys its pack cache on a per-object serial, not on id())."""
    import gc
    for style in ["init", "perturbed"] * 4:
        g = gold("v4", style)
        m = _model("v4", style)
        e = R.max_rel(m(T(g["x"]).cuda(), T(g["ta"]).cuda()).cpu(), T(g["eps_ta"]))
        assert e < EPS_TOL["bf16"], (style, e)
        del m
        gc.collect()
        torch.cuda.empty_cache() if style == "init" else None
