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


# ----------------------------------------------------------------------------- GPU: the persistent loop kernel (v3loop.cu)
@pytest.mark.gpu
def test_v3_loop_kernel_is_the_bf16_path_and_agrees_with_the_per_layer_sequence():
    """bf16 calls of up to 128 rows run unet3_loop_kernel (ONE launch for the whole chain, no graph); LDM_V3LOOP=0 selects the
    per-layer sequence (27 launches per step).  Both must reproduce the restatement's chain, and agree with each other."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, torch
import ldm_b200
from oracle import philox, weights
torch.set_grad_enabled(False)
u = ldm_b200.v3.ConditionalUNet(precision="bf16")
u.load_state_dict(weights.make_unet3_state(44, "perturbed"), strict=True)
u = u.to("cuda").eval()
d = ldm_b200.v3.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
eng = d._engine("cuda")
B = 32
x = torch.from_numpy(philox.normal_rows(3, 0, B, 1000)).cuda()
f, k = (torch.arange(B) * 5 %% 102).cuda(), (torch.arange(B) %% 10).cuda()
eng.ktrace_start()
eng.sample3(x, 999, 960, f, k, seed=17, sample_offset=0, use_graph=False)
names = [n for n, _ in eng.ktrace_stop()]
print("RESULT", int(eng.info("launches_per_step")), len(names), names[-1])
np.save("%s", x.cpu().numpy())
"""
    outs = {}
    for name, env_extra in (("loop", {}), ("layers", {"LDM_V3LOOP": "0"})):
        path = os.path.join(root, "gpurun_out", "v3_%s.npy" % name) if os.path.isdir(os.path.join(root, "gpurun_out")) else "/tmp/v3_%s.npy" % name
        env = dict(os.environ, PYTHONPATH=root, **env_extra)
        r = subprocess.run([sys.executable, "-c", code % path], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "RESULT" in r.stdout, r.stdout + r.stderr
        per_step, n_launches, last = r.stdout.split("RESULT")[1].split()[:3]
        if name == "loop":
            assert int(per_step) == 0 and int(n_launches) <= 8 and last == "unet3_loop", r.stdout
        else:
            assert int(per_step) == 27 and int(n_launches) >= 40 * 27, r.stdout
        outs[name] = torch.from_numpy(np.load(path))
    sd = weights.make_unet3_state(SEED, "perturbed")
    B = 32
    x = torch.from_numpy(philox.normal_rows(3, 0, B, 1000))
    f, k = torch.arange(B) * 5 % 102, torch.arange(B) % 10
    ref = R.sample3(sd, R.schedule(1000), x, f, k, noise_fn=lambda t: torch.from_numpy(philox.normal_rows(17, 0, B, t)), t_start=999, t_end=960)
    for name in outs:
        assert R.rel_l2(outs[name], ref) < LATENT_TOL["bf16"], (name, R.rel_l2(outs[name], ref))
    assert R.rel_l2(outs["loop"], outs["layers"]) < LATENT_TOL["bf16"]


@pytest.mark.gpu
def test_v3_loop_kernel_other_architecture_and_ragged_calls():
    """A non-default architecture the loop kernel covers (head widths 16 and 32, 128-wide latent, three stages) at ragged call
    sizes, single and per-row timesteps, against the restatement; one the kernel does not cover falls back to the layers."""
    import ldm_b200
    kw = dict(latent_dim=128, hidden_dims=[128, 256, 128, 128], time_emb_dim=256, num_classes=7, num_colors=3)
    sd = weights.make_state(weights.unet3_spec(latent_dim=128, hidden=kw["hidden_dims"], temb=256, num_classes=7, num_colors=3), 5, "perturbed")
    u = ldm_b200.v3.ConditionalUNet(precision="bf16", **kw)
    u.load_state_dict(sd, strict=True)
    u = u.to("cuda").eval()
    torch.manual_seed(11)
    for B in (1, 5, 127, 128):
        x, f, k = torch.randn(B, 128) * 1.5, torch.randint(0, 7, (B,)), torch.randint(0, 3, (B,))
        t1 = torch.tensor([613])
        got = u(x.cuda(), t1.cuda(), f.cuda(), k.cuda()).cpu()
        assert R.max_rel(got, R.unet3_forward(sd, x, t1, f, k)) < EPS_TOL["bf16"], B
        tb = torch.randint(0, 1000, (B,))
        got = u(x.cuda(), tb.cuda(), f.cuda(), k.cuda()).cpu()
        assert R.max_rel(got, R.unet3_forward(sd, x, tb, f, k)) < EPS_TOL["bf16"], B
    eng = ldm_b200.get_engine(torch.device("cuda"), "bf16")
    assert int(eng.info("launches_per_step")) == 0          # the loop kernel took these calls
    # head width 48 is outside the tensor-core attention: per-layer sequence, same answers
    kw2 = dict(latent_dim=128, hidden_dims=[128, 384, 128], time_emb_dim=256, num_classes=7, num_colors=3)
    sd2 = weights.make_state(weights.unet3_spec(latent_dim=128, hidden=kw2["hidden_dims"], temb=256, num_classes=7, num_colors=3), 6, "perturbed")
    u2 = ldm_b200.v3.ConditionalUNet(precision="bf16", **kw2)
    u2.load_state_dict(sd2, strict=True)
    u2 = u2.to("cuda").eval()
    x, f, k = torch.randn(9, 128), torch.randint(0, 7, (9,)), torch.randint(0, 3, (9,))
    got = u2(x.cuda(), torch.tensor([5], device="cuda"), f.cuda(), k.cuda()).cpu()
    assert R.max_rel(got, R.unet3_forward(sd2, x, torch.tensor([5]), f, k)) < EPS_TOL["bf16"]
    assert int(eng.info("launches_per_step")) > 0
