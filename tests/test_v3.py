"""v3 multi-conditional denoiser (BASELINE config 4): oracle pinned to the reference's own outputs
(tests/golden/v3_*.npz, made by oracle/make_golden_v3.py from the unmodified v3 script; live reference when present),
and parity of the sm_100a path through the module mirror / C ABI.  Tolerances as for v2: eps max|d|/max|ref| <= 1e-3
(fp32 mode), 2e-2 (bf16); latents relative L2."""
import os

import numpy as np
import pytest
import torch

from oracle import philox, ref_loader, restate as R, weights
from tests._util import EPS_TOL, LATENT_TOL, T

torch.set_grad_enabled(False)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED, NOISE_SEED, T_START = 44, 4321, 24


def gold(style):
    return np.load(os.path.join(GOLD, "v3_%s.npz" % style))


# ----------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_v3_restatement_reproduces_reference_goldens(style):
    g = gold(style)
    sd = weights.make_unet3_state(SEED, style)
    x, f, k = T(g["x"]), T(g["flower"]), T(g["color"])
    for t in (0, 1, 500, 999):
        assert R.max_rel(R.unet3_forward(sd, x, torch.tensor([t]), f, k), T(g["eps_t%d" % t])) < 1e-5
    assert R.max_rel(R.unet3_forward(sd, x, T(g["tb"]), f, k), T(g["eps_tb"])) < 1e-5
    # attention runs ACROSS the batch (v3:832): the first three rows alone give a different answer
    sub = R.unet3_forward(sd, x[:3], torch.tensor([500]), f[:3], k[:3])
    assert R.max_rel(sub, T(g["eps_t500_first3"])) < 1e-5
    assert R.max_rel(sub, T(g["eps_t500"])[:3]) > 1e-3
    noise = lambda t: torch.from_numpy(philox.normal_rows(NOISE_SEED + 1, 0, 6, t))
    x0 = R.sample3(sd, R.schedule(1000), x, f, k, noise_fn=noise, t_start=T_START)
    assert R.rel_l2(x0, T(g["chain_x0"])) < 1e-5


@pytest.mark.skipif(not ref_loader.available("v3"), reason="the reference tree is only present in the build container")
def test_v3_restatement_against_the_live_reference():
    m = ref_loader.load("v3")
    sd = weights.make_unet3_state(SEED, "perturbed")
    net = m.ConditionalUNet().eval()
    net.load_state_dict(sd, strict=True)
    torch.manual_seed(3)
    x, f, k = torch.randn(9, 256) * 4, torch.randint(0, 102, (9,)), torch.randint(0, 10, (9,))
    for t in (3, 777):
        assert torch.allclose(net(x, torch.tensor([t]), f, k), R.unet3_forward(sd, x, torch.tensor([t]), f, k), rtol=0, atol=1e-5)


def test_v3_mirror_state_dict_layout():
    import ldm_b200
    m = ldm_b200.v3.ConditionalUNet()
    sd = weights.make_unet3_state(SEED, "init")
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(tuple(m.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):      # no CPU fallback
        m.eval()(torch.zeros(2, 256), torch.tensor([1]), torch.zeros(2, dtype=torch.long), torch.zeros(2, dtype=torch.long))


# ----------------------------------------------------------------------------- GPU: parity through the C ABI
def _unet(style, precision):
    import ldm_b200
    m = ldm_b200.v3.ConditionalUNet(precision=precision)
    m.load_state_dict(weights.make_unet3_state(SEED, style), strict=True)
    return m.to("cuda").eval()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_v3_forward_against_reference_goldens(style, precision):
    g = gold(style)
    u = _unet(style, precision)
    x, f, k = T(g["x"]).cuda(), T(g["flower"]).cuda(), T(g["color"]).cuda()
    for t in (0, 1, 500, 999):
        eps = u(x, torch.tensor([t], device="cuda"), f, k).cpu()
        assert R.max_rel(eps, T(g["eps_t%d" % t])) < EPS_TOL[precision], (t, R.max_rel(eps, T(g["eps_t%d" % t])))
    assert R.max_rel(u(x, T(g["tb"]).cuda(), f, k).cpu(), T(g["eps_tb"])) < EPS_TOL[precision]
    assert R.max_rel(u(x[:3], torch.tensor([500], device="cuda"), f[:3], k[:3]).cpu(), T(g["eps_t500_first3"])) < EPS_TOL[precision]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_v3_chain_and_larger_batches(precision):
    import ldm_b200
    g = gold("perturbed")
    u = _unet("perturbed", precision)
    d = ldm_b200.v3.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
    x, f, k = T(g["x"]).cuda(), T(g["flower"]).cuda(), T(g["color"]).cuda()
    noise = np.stack([philox.normal_rows(NOISE_SEED + 1, 0, 6, t) for t in range(T_START, -1, -1)])
    eng = d._engine("cuda")
    xs = x.clone()
    eng.sample3(xs, T_START, 0, f, k, noise=torch.from_numpy(noise).cuda())
    assert R.rel_l2(xs.cpu(), T(g["chain_x0"])) < LATENT_TOL[precision]
    # p_sample with explicit noise = one step of the reference (v3:874-887)
    one = d.p_sample(x, T_START, f, k, noise=torch.from_numpy(noise[0]).cuda())
    sd = weights.make_unet3_state(SEED, "perturbed")
    want = R.ddpm_update(R.schedule(1000), T(g["x"]), R.unet3_forward(sd, T(g["x"]), torch.tensor([T_START]), T(g["flower"]), T(g["color"])),
                         T_START, torch.from_numpy(noise[0]))
    assert R.max_rel(one.cpu(), want) < EPS_TOL[precision] * 2
    # a batch that spans several key tiles of the attention kernel, against the restatement
    torch.manual_seed(5)
    B = 77
    xb, fb, kb = torch.randn(B, 256) * 2, torch.randint(0, 102, (B,)), torch.randint(0, 10, (B,))
    want = R.unet3_forward(sd, xb, torch.tensor([321]), fb, kb)
    got = u(xb.cuda(), torch.tensor([321], device="cuda"), fb.cuda(), kb.cuda()).cpu()
    assert R.max_rel(got, want) < EPS_TOL[precision], R.max_rel(got, want)
    # several query / key tiles of the tensor-core attention kernel (128 rows each): online-softmax rescale across tiles
    B = 300
    xb2, fb2, kb2 = torch.randn(B, 256) * 2, torch.randint(0, 102, (B,)), torch.randint(0, 10, (B,))
    want2 = R.unet3_forward(sd, xb2, torch.tensor([77]), fb2, kb2)
    got2 = u(xb2.cuda(), torch.tensor([77], device="cuda"), fb2.cuda(), kb2.cuda()).cpu()
    assert R.max_rel(got2, want2) < EPS_TOL[precision], R.max_rel(got2, want2)
    # out-of-range labels raise like nn.Embedding would
    with pytest.raises(IndexError):
        u(xb.cuda(), torch.tensor([1], device="cuda"), fb.cuda(), (kb + 10).cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_v3_full_size_call(precision):
    """BASELINE configs[3]: 1024 samples over 8 GPUs = one reference call of 128 rows per GPU.  Parity of that call's eps
    against the restatement, and of 10 reverse steps from x_T with Philox noise."""
    import ldm_b200
    sd = weights.make_unet3_state(SEED, "init")
    u = _unet("init", precision)
    B = 128
    x = torch.from_numpy(philox.normal_rows(9, 0, B, 1000))
    f, k = torch.arange(B) % 102, (torch.arange(B) * 7) % 10
    want = R.unet3_forward(sd, x, torch.tensor([640]), f, k)
    got = u(x.cuda(), torch.tensor([640], device="cuda"), f.cuda(), k.cuda()).cpu()
    assert R.max_rel(got, want) < EPS_TOL[precision], R.max_rel(got, want)
    d = ldm_b200.v3.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
    eng = d._engine("cuda")
    xs = x.cuda().clone()
    eng.sample3(xs, 999, 990, f.cuda(), k.cuda(), seed=21, sample_offset=0)
    ref = R.sample3(sd, R.schedule(1000), x, f, k, noise_fn=lambda t: torch.from_numpy(philox.normal_rows(21, 0, B, t)), t_start=999, t_end=990)
    assert R.rel_l2(xs.cpu(), ref) < LATENT_TOL[precision], R.rel_l2(xs.cpu(), ref)
