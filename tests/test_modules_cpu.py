"""Host-side mirror of the reference's module API: state_dict layout, loud failures, bit-exact host tables."""
import numpy as np
import pytest
import torch

import ldm_b200
from oracle import restate as R, weights
from tests._util import UNET_SEED


def test_state_dict_layout_is_the_references():
    u, ae = ldm_b200.ConditionalUNet(), ldm_b200.SimpleAutoencoder()
    assert [(k, tuple(v.shape)) for k, v in u.state_dict().items()] == [(k, tuple(s)) for k, s, _ in weights.unet_spec()]
    assert [(k, tuple(v.shape)) for k, v in ae.state_dict().items()] == [(k, tuple(s)) for k, s, _ in weights.autoencoder_spec()]
    assert sum(p.numel() for p in u.parameters()) == 11131137
    assert sum(p.numel() for p in ae.parameters()) == 69218997
    assert sum(p.numel() for p in ae.decoder.parameters()) == 26066217
    u.load_state_dict(weights.make_unet_state(UNET_SEED, "perturbed"), strict=True)
    ae.load_state_dict(weights.make_autoencoder_state(1, "init"), strict=True)


def test_checkpoint_loader_accepts_both_reference_formats():
    ae = ldm_b200.SimpleAutoencoder()
    sd = weights.make_autoencoder_state(2, "perturbed")
    ldm_b200.load_autoencoder_checkpoint(ae, {"autoencoder": sd, "discriminator": {}})      # v2:1179-1191
    assert torch.equal(ae.decoder.fc[0].weight, sd["decoder.fc.0.weight"])
    ae2 = ldm_b200.SimpleAutoencoder()
    ldm_b200.load_autoencoder_checkpoint(ae2, sd)                                            # v2:1326
    assert torch.equal(ae2.decoder.fc[3].bias, sd["decoder.fc.3.bias"])


def test_schedule_and_sinusoid_tables_are_bit_exact(golden):
    g = golden("init")
    u = ldm_b200.ConditionalUNet().eval()
    d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, None)
    assert d.n_steps == 1000 and d.eps_model is u and d.device is None
    assert np.array_equal(d.beta.numpy(), g["beta"])
    assert np.array_equal(d.alpha.numpy(), g["alpha"])
    assert np.array_equal(d.alpha_bar.numpy(), g["alpha_bar"])
    tab = u.time_emb.sinusoid_table(1000)
    assert torch.equal(tab, R.sinusoid(torch.arange(1000)))
    assert torch.equal(tab[[0, 17, 999]], R.sinusoid(torch.tensor([0, 17, 999])))   # row t IS the reference's value for t


def test_q_sample_matches_restatement():
    u = ldm_b200.ConditionalUNet().eval()
    d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, None)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(3, 256, generator=g), torch.randn(3, 256, generator=g)
    t = torch.tensor([0, 500, 999])
    assert torch.equal(d.q_sample(x0, t, eps), R.q_sample(R.schedule(1000), x0, t, eps))


def test_cpu_tensors_and_training_mode_fail_loudly():
    u = ldm_b200.ConditionalUNet()
    x, t = torch.zeros(2, 256), torch.tensor([5])
    with pytest.raises(RuntimeError, match="eval"):
        u(x, t)                                      # training mode
    u.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        u(x, t)                                      # no CPU fallback
    d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        d.sample((2, 256), "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        ldm_b200.SimpleAutoencoder().eval().decode(torch.zeros(1, 256))
    with pytest.raises(IndexError):
        d.p_sample(x, 1000)
    with pytest.raises(ValueError):
        d.p_sample(x, torch.tensor([1, 2]))


def test_product_package_never_imports_the_oracle():
    import os
    pkg = os.path.dirname(ldm_b200.LIB_PATH)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle." not in src.replace("oracle/", ""), f


def test_checkpoint_conventions_and_png_writer(tmp_path):
    """epoch-in-filename resume (v2:1352-1364), both autoencoder wrappings (v2:1179-1191,1326), PNG grid."""
    import struct
    import zlib
    import ldm_b200
    from oracle import weights
    sd = weights.make_unet_state(42, "perturbed")
    p = tmp_path / "conditional_diffusion_epoch_340.pt"
    torch.save(sd, p)
    u = ldm_b200.ConditionalUNet()
    assert ldm_b200.load_unet_checkpoint(u, str(p)) == 340
    assert all(torch.equal(u.state_dict()[k], v) for k, v in sd.items())
    q = tmp_path / "conditional_diffusion_final.pt"
    torch.save(sd, q)
    assert ldm_b200.load_unet_checkpoint(ldm_b200.ConditionalUNet(), str(q)) == 0          # no epoch in the name: start from 0
    assert ldm_b200.load_unet_checkpoint(u, str(tmp_path / "missing_epoch_3.pt")) == 0
    sd_a = weights.make_autoencoder_state(43, "perturbed")
    for obj in (sd_a, {"autoencoder": sd_a, "discriminator": {}}):
        ae = ldm_b200.SimpleAutoencoder()
        ldm_b200.load_autoencoder_checkpoint(ae, obj)
        assert torch.equal(ae.state_dict()["decoder.fc.0.weight"], sd_a["decoder.fc.0.weight"])
    imgs = torch.rand(5, 3, 8, 8)
    out = ldm_b200.save_image_grid(imgs, str(tmp_path / "row.png"))
    raw = open(out, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (5 * 8 + 6 * 2, 8 + 2 * 2)
    idat = raw[raw.index(b"IDAT") + 4: raw.index(b"IEND") - 8]
    rows = zlib.decompress(idat)
    assert len(rows) == h * (1 + 3 * w)
    px = torch.frombuffer(bytearray(rows), dtype=torch.uint8).view(h, 1 + 3 * w)[:, 1:].view(h, w, 3)
    assert torch.equal(px[2:10, 2:10], ldm_b200.to_uint8(imgs)[0])


def test_empty_batches_return_empty_tensors():
    """Edge case: a batch of zero samples launches nothing and returns an empty tensor of the right shape."""
    import ldm_b200
    from ldm_b200 import v3, v4
    u = ldm_b200.ConditionalUNet().eval()
    assert tuple(u(torch.zeros(0, 256), torch.tensor([3]), torch.zeros(0, dtype=torch.long)).shape) == (0, 256)
    d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, torch.device("cpu"))
    assert tuple(d.sample((0, 256), torch.device("cpu"), torch.zeros(0, dtype=torch.long)).shape) == (0, 256)
    ae = ldm_b200.SimpleAutoencoder().eval()
    assert tuple(ae.decode(torch.zeros(0, 256)).shape) == (0, 3, 64, 64)
    u3 = v3.ConditionalUNet().eval()
    z = torch.zeros(0, dtype=torch.long)
    assert tuple(u3(torch.zeros(0, 256), torch.tensor([3]), z, z).shape) == (0, 256)
    m = v4.SimpleUNet().eval()
    assert tuple(m(torch.zeros(0, 3, 64, 64), torch.zeros(0)).shape) == (0, 3, 64, 64)
    assert tuple(v4.DiffusionModel(m, 10, device="cpu").sample((0, 3, 32, 32)).shape) == (0, 3, 32, 32)


def test_pack_cache_key_survives_object_address_reuse():
    """A new module that lands on the address (and parameter storage) of a collected one must not look 'already packed'."""
    import gc
    from ldm_b200 import engine
    keys = set()
    for _ in range(50):
        m = torch.nn.Linear(4, 4)
        keys.add(engine._state_key(m))
        del m
        gc.collect()
    assert len(keys) == 50
    m = torch.nn.Linear(4, 4)
    k0 = engine._state_key(m)
    assert engine._state_key(m) == k0                      # stable for an untouched module
    with torch.no_grad():
        m.weight.add_(1.0)
    assert engine._state_key(m) != k0                      # in-place updates re-pack


def test_default_precision_reaches_every_mirror(monkeypatch):
    """set_default_precision() / LDM_B200_PRECISION select the engine of EVERY mirror whose `precision` was left unset
    (v4.SimpleUNet and the conv U-Net blocks used to fall back to bf16 on their own)."""
    import ldm_b200
    from ldm_b200 import engine, modules, v4
    seen = []

    class Stop(Exception):
        pass

    def fake(device, precision=None):
        seen.append(precision or engine.default_precision())
        raise Stop

    for mod in (modules, v4):
        monkeypatch.setattr(mod, "get_engine", fake)
    engine.set_default_precision("fp32")
    try:
        for call in (lambda: v4.SimpleUNet().eval().engine(torch.device("cuda", 0)),
                     lambda: ldm_b200.UNetAttentionBlock(64).eval()(torch.zeros(1, 64, 4, 4, device="meta")),
                     lambda: ldm_b200.UNetResidualBlock(64, 64).eval()(torch.zeros(1, 64, 4, 4, device="meta"), torch.zeros(1, 256)),
                     lambda: v4.SimpleUNet(precision="bf16").eval().engine(torch.device("cuda", 0))):
            with pytest.raises(Stop):
                call()
    finally:
        engine.set_default_precision(None)
    assert seen == ["fp32", "fp32", "fp32", "bf16"]


def test_pack_serial_is_not_inherited_by_copies_and_invalidate_forgets():
    import copy
    from ldm_b200 import engine
    m = torch.nn.Linear(4, 4)
    k = engine._state_key(m)
    assert "_ldm_b200_serial" not in m.__dict__                    # kept outside the module: deepcopy / pickle cannot carry it
    m2 = copy.deepcopy(m)
    assert engine._state_key(m2)[0] != k[0]
    with torch.no_grad():
        m.weight.data.mul_(2.0)                                    # an edit through .data does not bump the version counter ...
    assert engine._state_key(m) == k

    class E(engine.Engine):                                        # ... which is what invalidate() is for (no context needed here)
        def __init__(self):
            self.ctx = None
            self._unet_key, self._dec_key, self._pix_key, self._cls_key, self._ublocks = k, ("other",), None, 1, {}
    e = E()
    e.invalidate(m)
    assert e._unet_key is None and e._cls_key is None and e._dec_key == ("other",)
    e.invalidate()
    assert e._dec_key is None
