"""Full-length chains and full-size calls against outputs of the UNMODIFIED reference (tests/golden/v2_tail.npz,
v3_long.npz, v4_long.npz from oracle/make_golden_long.py), and the plumbing either side of the path:

* v2: 1000 steps for GLOBAL rows 252..255 of a B = 256 call (BASELINE configs[1]); on the GPU these rows sit in the
  ragged last cluster of the chain kernel (rows 240..255 of 6 x 48 slots, the rest zero-filled by TMA).
* v3: 1000 steps at B = 128 (the per-GPU call of configs[3]; attention couples the rows of a call).
* v4: 100 steps at 64 x 64, B = 2, and eps of a B = 64 call (the per-GPU share of configs[4]) against the oracle.
* Encoder.forward (v2:181-239) against the live reference; generate_class_samples / generate_sharded on the GPU.
Tolerances: tests/_util.py (north_star: eps 1e-3 fp32 mode / 2e-2 bf16; latents relative L2 5e-3 / 5e-2)."""
import os

import numpy as np
import pytest
import torch

from oracle import philox, ref_loader, restate as R, restate_pix as P, weights
from tests._util import AE_SEED, EPS_TOL, IMAGE_TOL, LATENT_TOL, NOISE_SEED, T, UNET_SEED, make_autoencoder, make_unet

torch.set_grad_enabled(False)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V3_SEED, V3_NOISE, V3_B = 44, 4321, 128
V4_SEED, V4_NOISE, V4_T = 45, 777, 99


def _g(name):
    return np.load(os.path.join(GOLD, name))


def _img_noise(seed, offset, n, step, shape=(3, 64, 64)):
    return torch.from_numpy(philox.normal_rows(seed, offset, n, step, shape[0] * shape[1] * shape[2])).view(n, *shape)


# ----------------------------------------------------------------------------- CPU: the oracle stays pinned over long chains
@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_v2_tail_rows_restatement(style):
    g = _g("v2_tail.npz")
    off = int(g["offset"])
    sd = weights.make_unet_state(UNET_SEED, style)
    x_T = T(philox.normal_rows(NOISE_SEED, off, 4, 1000))
    x0, _ = R.sample(sd, R.schedule(1000), x_T, T(g["c"]), noise_fn=lambda t: T(philox.normal_rows(NOISE_SEED, off, 4, t)))
    assert R.rel_l2(x0, g["x0_%s" % style]) < 1e-3
    assert list(g["c"]) == [b % 102 for b in range(off, off + 4)]


def test_v3_long_chain_restatement():
    g = _g("v3_long.npz")
    sd = weights.make_unet3_state(V3_SEED, "init")
    f, k = T(g["flower"]), T(g["color"])
    x = T(philox.normal_rows(V3_NOISE, 0, V3_B, 1000))
    # the first 100 steps on the CPU (the full 1000 take a minute); the GPU tests below run the whole chain
    x900 = R.sample3(sd, R.schedule(1000), x, f, k, noise_fn=lambda t: T(philox.normal_rows(V3_NOISE, 0, V3_B, t)), t_start=999, t_end=900)
    assert R.rel_l2(x900, g["x_after_t900"]) < 1e-3


def test_v4_long_chain_restatement():
    g = _g("v4_long.npz")
    sd = weights.make_pix_state(V4_SEED, "perturbed")
    x = T(g["x"])
    assert torch.equal(x, _img_noise(V4_NOISE + 5, 0, 2, 1000))
    x0 = P.sample(sd, R.schedule(1000), x, noise_fn=lambda t: _img_noise(V4_NOISE + 6, 0, 2, t), t_start=V4_T)
    assert R.rel_l2(x0, g["x0"]) < 1e-4


@pytest.mark.skipif(not ref_loader.available(), reason="the reference tree is only present in the build container")
def test_encoder_forward_against_the_live_reference():
    """Encoder.forward (v2:181-239) and the encode helpers of SimpleAutoencoder (v2:345-353): plain torch, bit-equal."""
    import ldm_b200
    m = ref_loader.load()
    sd = weights.make_autoencoder_state(AE_SEED, "perturbed")
    ref = m.SimpleAutoencoder().eval()
    ref.load_state_dict(sd, strict=True)
    ours = ldm_b200.SimpleAutoencoder().eval()
    ours.load_state_dict(sd, strict=True)
    torch.manual_seed(2)
    x = torch.rand(3, 3, 64, 64)
    mu_r, lv_r = ref.encoder(x)
    mu_o, lv_o = ours.encoder(x)
    assert torch.equal(mu_r, mu_o) and torch.equal(lv_r, lv_o)
    assert len(ours.encoder.skip_features) == len(ref.encoder.skip_features) == 4
    assert all(torch.equal(a, b) for a, b in zip(ref.encoder.skip_features, ours.encoder.skip_features))
    a, b = ref.encode_with_params(x), ours.encode_with_params(x)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    torch.manual_seed(7)
    z_r = ref.encode(x)
    torch.manual_seed(7)
    z_o = ours.encode(x)
    assert torch.equal(z_r, z_o)
    assert torch.equal(ref.classify(z_r), ours.classify(z_o))


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_v2_full_batch_tail_rows_against_reference(style):
    """B = 256, 1000 steps, in-kernel Philox: rows 252..255 against the reference's own run of those global samples."""
    import ldm_b200
    g = _g("v2_tail.npz")
    off = int(g["offset"])
    for precision in ("bf16", "fp32"):
        u = make_unet(style, precision)
        d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
        c = (torch.arange(256) % 102).cuda()
        x0 = d.sample((256, 256), torch.device("cuda"), c, seed=NOISE_SEED, sample_offset=0).cpu()
        err = R.rel_l2(x0[off:off + 4], g["x0_%s" % style])
        assert err < LATENT_TOL[precision], (precision, err)
        # and the same rows of the neighbouring full cluster slots (rows 236..239) agree with a small call
        sub = d.sample((4, 256), torch.device("cuda"), c[236:240].contiguous(), seed=NOISE_SEED, sample_offset=236).cpu()
        assert torch.equal(sub, x0[236:240])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_v3_thousand_steps_at_call_size(precision):
    import ldm_b200
    g = _g("v3_long.npz")
    u = ldm_b200.v3.ConditionalUNet(precision=precision)
    u.load_state_dict(weights.make_unet3_state(V3_SEED, "init"), strict=True)
    u = u.to("cuda").eval()
    d = ldm_b200.v3.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
    f, k = T(g["flower"]).cuda(), T(g["color"]).cuda()
    eng = d._engine("cuda")
    x = T(philox.normal_rows(V3_NOISE, 0, V3_B, 1000)).cuda()
    eng.sample3(x, 999, 900, f, k, seed=V3_NOISE, sample_offset=0)
    e900 = R.rel_l2(x.cpu(), g["x_after_t900"])
    eng.sample3(x, 899, 500, f, k, seed=V3_NOISE, sample_offset=0)
    e500 = R.rel_l2(x.cpu(), g["x_after_t500"])
    eng.sample3(x, 499, 0, f, k, seed=V3_NOISE, sample_offset=0)
    e0 = R.rel_l2(x.cpu(), g["x0"])
    assert max(e900, e500, e0) < LATENT_TOL[precision], (e900, e500, e0)
    # the public entry draws x_T in the kernel from the same stream (libm ulps away from the host-side normals above)
    x0 = d.sample((V3_B, 256), torch.device("cuda"), f, k, seed=V3_NOISE, sample_offset=0)
    e_pub = R.rel_l2(x0.cpu(), g["x0"])
    assert e_pub < LATENT_TOL[precision], e_pub


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_v4_hundred_steps_and_full_size_eps(precision):
    from ldm_b200 import v4
    g = _g("v4_long.npz")
    sd = weights.make_pix_state(V4_SEED, "perturbed")
    m = v4.SimpleUNet(precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda").eval()
    d = v4.DiffusionModel(m, 1000, device="cuda")
    eng = d._engine("cuda")
    xs = T(g["x"]).cuda().clone()
    eng.pix_sample(xs, V4_T, 0, seed=V4_NOISE + 6, sample_offset=0, use_graph=True)
    err = R.rel_l2(xs.cpu(), g["x0"])
    assert err < LATENT_TOL[precision], err
    if precision == "bf16":
        # eps of the per-GPU share of BASELINE configs[4] (64 images of 64 x 64) against the oracle, per-sample timesteps
        x = _img_noise(5, 0, 64, 1000)
        t = torch.arange(64) * 15
        want = P.unet_forward(sd, x, t)
        got = m(x.cuda(), t.cuda()).cpu()
        e = R.max_rel(got, want)
        assert e < EPS_TOL[precision], e


@pytest.mark.gpu
def test_generate_class_samples_and_sharded_entry(tmp_path):
    """The reference's generate_class_samples (v2:856-882) and the multi-GPU entry at world size 1, against the oracle."""
    import ldm_b200
    from ldm_b200 import sharding
    style, n_steps, n = "perturbed", 40, 5
    u = make_unet(style, "bf16")
    ae = make_autoencoder(style, "bf16")
    d = ldm_b200.ConditionalDenoiseDiffusion(u, n_steps, torch.device("cuda"))
    path = str(tmp_path / "row.png")
    imgs, lat = ldm_b200.generate_class_samples(ae, d, 33, num_samples=n, save_path=path, seed=11)
    assert imgs.shape == (n, 3, 64, 64) and lat.shape == (n, 256) and imgs.is_cuda
    assert open(path, "rb").read()[:8] == b"\x89PNG\r\n\x1a\n"
    sd_u, sd_a = weights.make_unet_state(UNET_SEED, style), weights.make_autoencoder_state(AE_SEED, style)
    c = torch.full((n,), 33)
    x_T = T(philox.normal_rows(11, 0, n, n_steps))
    want, _ = R.sample(sd_u, R.schedule(n_steps), x_T, c, noise_fn=lambda t: T(philox.normal_rows(11, 0, n, t)))
    assert R.rel_l2(lat.cpu(), want) < LATENT_TOL["bf16"]
    assert float((imgs.cpu() - R.decode(sd_a, lat.cpu())).abs().max()) < IMAGE_TOL["bf16"]
    # by name, as the reference looks the class up in its list (v2:860-864)
    names = ["class %d" % i for i in range(102)]
    imgs2, lat2 = ldm_b200.generate_class_samples(ae, d, "class 33", num_samples=n, class_names=names, seed=11)
    assert torch.equal(lat2, lat) and torch.equal(imgs2, imgs)
    with pytest.raises(ValueError):
        ldm_b200.generate_class_samples(ae, d, "no such flower", class_names=names)
    # world size 1: the sharded entry is the same computation
    classes = torch.full((n,), 33)
    imgs3, lat3, (lo, hi) = sharding.generate_sharded(ae, d, classes, seed=11)
    assert (lo, hi) == (0, n) and torch.equal(lat3, lat) and torch.equal(imgs3, imgs)


@pytest.mark.gpu
def test_v3_host_entry_and_kernel_trace():
    """ldm_generate3_host (host labels in, host images out) equals sample + decode; ldm_debug_ktrace lists the launches."""
    import ldm_b200
    u = ldm_b200.v3.ConditionalUNet(precision="bf16")
    u.load_state_dict(weights.make_unet3_state(V3_SEED, "init"), strict=True)
    u = u.to("cuda").eval()
    ae = make_autoencoder("init", "bf16")
    d = ldm_b200.v3.ConditionalDenoiseDiffusion(u, 60, torch.device("cuda"))
    eng = d._engine("cuda")
    eng.pack_decoder(ae.decoder)
    B = 10
    f, k = (torch.arange(B) * 11) % 102, torch.arange(B) % 10
    img = torch.empty(B, 3, 64, 64).pin_memory()
    lat = torch.empty(B, 256).pin_memory()
    eng.generate3_host(f.pin_memory(), k.pin_memory(), img, lat, seed=9, sample_offset=0)
    x0 = d.sample((B, 256), torch.device("cuda"), f.cuda(), k.cuda(), seed=9, sample_offset=0)
    assert torch.equal(lat, x0.cpu())
    eng.ktrace_start()
    ref = ae.decode(x0)
    trace = eng.ktrace_stop()
    assert torch.equal(img, ref.cpu())
    names = [n for n, _ in trace]
    assert len(trace) == 41 and names[0] == "launch_load_x" and names[-1] == "final_gn_conv3" and names.count("conv_tc") == 9 and names.count("conv_halo") == 1
    assert all(0.0 < ms < 50.0 for _, ms in trace)
    with pytest.raises(ldm_b200.LdmError):
        eng.generate3_host((f + 200).pin_memory(), k.pin_memory(), img, None, seed=9)      # flower label out of range


@pytest.mark.gpu
def test_strict_cuda_core_paths_in_a_fresh_process():
    """The strict mode runs the three-term bf16 split on the tensor cores by default; LDM_CHAIN=0 and LDM_DEC_F32=1 select the
    fp32 CUDA-core kernels (gemm_f32 / conv_f32), kept as the independent cross-check: both must meet the strict tolerances,
    and agree with each other far inside them."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, torch
from oracle import restate as R
from tests._util import EPS_TOL, IMAGE_TOL, T, make_autoencoder, make_unet
torch.set_grad_enabled(False)
g = np.load("tests/golden/v2_perturbed.npz")
u = make_unet("perturbed", "fp32")
eng = u.engine("cuda", 1000)
x, c = T(g["fwd_x"]).cuda(), T(g["fwd_c"]).cuda()
eps = u(x, torch.tensor([500], device="cuda"), c).cpu()
img = make_autoencoder("perturbed", "fp32").decode(T(g["dec_z"]).cuda()).cpu()
print("RESULT", int(eng.info("chain")), R.max_rel(eps, T(g["fwd_eps_t500"])), float((img - T(g["dec_img"])).abs().max()))
np.save("%s", np.concatenate([eps.numpy().ravel(), img.numpy().ravel()]))
"""
    outs = {}
    for name, env_extra in (("tc", {}), ("cuda_cores", {"LDM_CHAIN": "0", "LDM_DEC_F32": "1"})):
        path = os.path.join(root, "gpurun_out", "strict_%s.npy" % name) if os.path.isdir(os.path.join(root, "gpurun_out")) else "/tmp/strict_%s.npy" % name
        env = dict(os.environ, PYTHONPATH=root, **env_extra)
        r = subprocess.run([sys.executable, "-c", code % path], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "RESULT" in r.stdout, r.stdout + r.stderr
        chain, e_eps, e_img = r.stdout.split("RESULT")[1].split()[:3]
        assert int(chain) == (1 if name == "tc" else 0)
        assert float(e_eps) < EPS_TOL["fp32"] and float(e_img) < IMAGE_TOL["fp32"], (name, e_eps, e_img)
        outs[name] = np.load(path)
    assert float(np.abs(outs["tc"] - outs["cuda_cores"]).max() / np.abs(outs["cuda_cores"]).max()) < 1e-4
