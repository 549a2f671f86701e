"""The conv U-Net blocks of the v2 script: UNetResidualBlock v2:462-486, UNetAttentionBlock v2:434-459,
SwitchSequential v2:489-498.  The reference defines them but never runs them, so parity is module-level: the oracle
restatement is pinned to the reference classes (tests/golden/ublock.npz from the unmodified script; live reference when
present) and the sm_100a operators are compared with it in both modes (max|d|/max|ref| <= 1e-3 strict fp32, 2e-2 bf16)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, restate as R, weights
from tests._util import EPS_TOL, T

torch.set_grad_enabled(False)
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ublock.npz"))
CASES_RES = [(64, 64), (64, 128)]
CASES_ATTN = [128, 64]


@pytest.mark.parametrize("cin,cout", CASES_RES)
def test_residual_block_restatement_reproduces_reference_goldens(cin, cout):
    sd = weights.make_state(weights.ublock_res_spec(cin, cout), 46, "perturbed")
    k = "res_%d_%d_" % (cin, cout)
    x, t, c = T(G[k + "x"]), T(G[k + "t"]), T(G[k + "c"])
    assert R.max_rel(R.unet_residual_block(sd, "", x, t, c), T(G[k + "y_tc"])) < 1e-5
    assert R.max_rel(R.unet_residual_block(sd, "", x, t), T(G[k + "y_t"])) < 1e-5


@pytest.mark.parametrize("ch", CASES_ATTN)
def test_attention_block_restatement_reproduces_reference_goldens(ch):
    sd = weights.make_state(weights.ublock_attn_spec(ch), 47, "perturbed")
    assert R.max_rel(R.unet_attention_block(sd, "", T(G["attn_%d_x" % ch])), T(G["attn_%d_y" % ch])) < 1e-5


@pytest.mark.skipif(not ref_loader.available("v2"), reason="the reference tree is only present in the build container")
def test_block_restatements_against_the_live_reference():
    m = ref_loader.load("v2")
    torch.manual_seed(1)
    sd = weights.make_state(weights.ublock_res_spec(128, 64), 46, "perturbed")
    blk = m.UNetResidualBlock(128, 64).eval()
    blk.load_state_dict(sd, strict=True)
    x, t, c = torch.randn(2, 128, 8, 8), torch.randn(2, 256), torch.randn(2, 256)
    assert torch.equal(blk(x, t, c), R.unet_residual_block(sd, "", x, t, c))
    sd = weights.make_state(weights.ublock_attn_spec(256), 47, "perturbed")
    att = m.UNetAttentionBlock(256).eval()
    att.load_state_dict(sd, strict=True)
    x = torch.randn(2, 256, 4, 4)
    assert torch.equal(att(x), R.unet_attention_block(sd, "", x))


def test_block_mirrors_state_dict_layout():
    import ldm_b200
    for cin, cout in ((64, 64), (64, 128)):
        m = ldm_b200.UNetResidualBlock(cin, cout)
        sd = weights.make_state(weights.ublock_res_spec(cin, cout), 46, "init")
        assert list(m.state_dict().keys()) == list(sd.keys())
        m.load_state_dict(sd, strict=True)
    a = ldm_b200.UNetAttentionBlock(128)
    sd = weights.make_state(weights.ublock_attn_spec(128), 47, "init")
    assert list(a.state_dict().keys()) == list(sd.keys())
    a.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):          # no CPU fallback
        a.eval()(torch.zeros(1, 128, 8, 8))


# ----------------------------------------------------------------------------- GPU: parity through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,cout", CASES_RES)
def test_residual_block_against_reference_goldens(cin, cout, precision):
    import ldm_b200
    m = ldm_b200.UNetResidualBlock(cin, cout, precision=precision)
    m.load_state_dict(weights.make_state(weights.ublock_res_spec(cin, cout), 46, "perturbed"), strict=True)
    m = m.cuda().eval()
    k = "res_%d_%d_" % (cin, cout)
    x, t, c = T(G[k + "x"]).cuda(), T(G[k + "t"]).cuda(), T(G[k + "c"]).cuda()
    e = R.max_rel(m(x, t, c).cpu(), T(G[k + "y_tc"]))
    assert e < EPS_TOL[precision], e
    e = R.max_rel(m(x, t).cpu(), T(G[k + "y_t"]))
    assert e < EPS_TOL[precision], e


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ch", CASES_ATTN)
def test_attention_block_against_reference_goldens(ch, precision):
    import ldm_b200
    m = ldm_b200.UNetAttentionBlock(ch, precision=precision)
    m.load_state_dict(weights.make_state(weights.ublock_attn_spec(ch), 47, "perturbed"), strict=True)
    m = m.cuda().eval()
    e = R.max_rel(m(T(G["attn_%d_x" % ch]).cuda()).cpu(), T(G["attn_%d_y" % ch]))
    assert e < EPS_TOL[precision], e


@pytest.mark.gpu
def test_switch_sequential_and_long_sequences():
    """SwitchSequential (v2:489-498) routes (x, t, c) to residual blocks and x alone to attention blocks; a 32 x 32 map
    gives 1024 tokens = 8 query tiles x 8 key tiles of the attention kernel (online softmax across tiles)."""
    import ldm_b200
    sd_r = weights.make_state(weights.ublock_res_spec(64, 128), 46, "perturbed")
    sd_a = weights.make_state(weights.ublock_attn_spec(128), 47, "perturbed")
    seq = ldm_b200.SwitchSequential(ldm_b200.UNetResidualBlock(64, 128), ldm_b200.UNetAttentionBlock(128))
    seq[0].load_state_dict(sd_r, strict=True)
    seq[1].load_state_dict(sd_a, strict=True)
    seq = seq.cuda().eval()
    torch.manual_seed(7)
    x, t, c = torch.randn(2, 64, 32, 32), torch.randn(2, 256), torch.randn(2, 256)
    want = R.unet_attention_block(sd_a, "", R.unet_residual_block(sd_r, "", x, t, c))
    got = seq(x.cuda(), t.cuda(), c.cuda()).cpu()
    assert R.max_rel(got, want) < EPS_TOL["bf16"], R.max_rel(got, want)


@pytest.mark.gpu
def test_block_handles_are_released_and_reused():
    """Re-packs (changed weights, .data edits + invalidate, collected modules) release the previous device copy: the
    context's handle table stays at the number of LIVE blocks instead of growing with every pack."""
    import gc
    import ldm_b200
    from ldm_b200 import engine
    sd = weights.make_state(weights.ublock_attn_spec(128), 47, "perturbed")
    x = T(G["attn_128_x"]).cuda()
    eng = engine.get_engine(torch.device("cuda", 0), "bf16")
    eng.invalidate()
    gc.collect()
    base = len(eng._ublocks)
    handles = set()
    for i in range(6):
        m = ldm_b200.UNetAttentionBlock(128, precision="bf16")
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        y0 = m(x)
        assert R.max_rel(y0.cpu(), T(G["attn_128_y"])) < EPS_TOL["bf16"]
        with torch.no_grad():
            m.proj.weight.mul_(0.5)                                   # version bump: re-pack in place of the old handle
        y1 = m(x)
        assert not torch.equal(y0, y1)
        m.proj.weight.data.mul_(2.0)                                  # .data edit: invisible until invalidate()
        assert torch.equal(m(x), y1)
        ldm_b200.invalidate(m)
        assert R.max_rel(m(x).cpu(), T(G["attn_128_y"])) < EPS_TOL["bf16"]
        handles.add(eng._ublocks[engine._module_serial(m)][1])
        assert len(eng._ublocks) == base + 1
        del m
        gc.collect()
        assert len(eng._ublocks) == base                              # the finalizer released the collected module's block
    assert len(handles) <= 2, handles                                 # freed slots are handed out again


@pytest.mark.gpu
def test_ddpm_step_validates_its_operands():
    import ldm_b200
    from ldm_b200 import engine
    eng = engine.get_engine(torch.device("cuda", 0), "fp32")
    d = ldm_b200.ConditionalDenoiseDiffusion(ldm_b200.ConditionalUNet(precision="fp32").cuda().eval(), 1000, torch.device("cuda"))
    eng.set_schedule(*d._host_schedule)
    x, eps, z = torch.randn(5, 192, device="cuda"), torch.randn(5, 192, device="cuda"), torch.randn(5, 192, device="cuda")
    want = R.ddpm_update(R.schedule(1000), x.cpu(), eps.cpu(), 321, z.cpu())
    got = eng.ddpm_step(x.clone(), eps, 321, noise=z)                 # any row width (a multiple of 4), not only latent_dim
    assert torch.equal(got.cpu(), want)
    for bad in (eps[:, :100], eps.double(), eps.cpu(), eps.t().contiguous().t()):
        with pytest.raises(ValueError):
            eng.ddpm_step(x.clone(), bad, 321, noise=z)
