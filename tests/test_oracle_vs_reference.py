"""The restatement against the LIVE reference (only where /root/reference exists: the build container)."""
import pytest
import torch

from oracle import ref_loader, restate as R, weights

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_state_dict_layout_matches_reference(ref):
    for mod, spec in ((ref.ConditionalUNet(), weights.unet_spec()), (ref.SimpleAutoencoder(), weights.autoencoder_spec())):
        assert [(k, tuple(v.shape)) for k, v in mod.state_dict().items()] == [(k, tuple(s)) for k, s, _ in spec]


@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_unet_and_update_bit_equal(ref, style):
    sd = weights.make_unet_state(7, style)
    u = ref.ConditionalUNet().eval()
    u.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 256, generator=g) * 2
    c = torch.tensor([0, 101, 50, 7, 7])
    for t in (999, 321, 0):
        assert torch.equal(R.unet_forward(sd, x, torch.tensor([t]), c), u(x, torch.tensor([t]), c))
    tb = torch.tensor([0, 999, 5, 500, 77])
    assert torch.equal(R.unet_forward(sd, x, tb, c), u(x, tb, c))
    assert torch.equal(R.unet_forward(sd, x, tb, None), u(x, tb, None))
    d = ref.ConditionalDenoiseDiffusion(u, 1000, None)
    sched = R.schedule(1000)
    assert all(torch.equal(a, b) for a, b in zip(sched, (d.beta, d.alpha, d.alpha_bar)))
    nz = torch.randn(5, 256, generator=g)
    old = ref.torch.randn_like
    ref.torch.randn_like = lambda t_: nz
    try:
        want = d.p_sample(x, 400, c)
    finally:
        ref.torch.randn_like = old
    assert torch.equal(R.p_sample(sd, sched, x, 400, c, noise=nz), want)
    assert torch.equal(R.p_sample(sd, sched, x, 0, c), d.p_sample(x, torch.tensor([0]), c))


@pytest.mark.parametrize("style", ["init", "perturbed"])
def test_decode_bit_equal(ref, style):
    sd = weights.make_autoencoder_state(9, style)
    ae = ref.SimpleAutoencoder().eval()
    ae.load_state_dict(sd, strict=True)
    z = torch.randn(2, 256, generator=torch.Generator().manual_seed(5))
    assert torch.equal(R.decode(sd, z), ae.decode(z))


def test_mirror_modules_match_reference_layout(ref):
    import ldm_b200
    for mine, theirs in ((ldm_b200.ConditionalUNet(), ref.ConditionalUNet()), (ldm_b200.SimpleAutoencoder(), ref.SimpleAutoencoder())):
        a = [(k, tuple(v.shape), v.dtype) for k, v in mine.state_dict().items()]
        b = [(k, tuple(v.shape), v.dtype) for k, v in theirs.state_dict().items()]
        assert a == b
        mine.load_state_dict(theirs.state_dict(), strict=True)
        theirs.load_state_dict(mine.state_dict(), strict=True)
