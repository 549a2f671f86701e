"""Race proxy for the hand-rolled mbarrier / st.async / DSMEM / TMA protocols (compute-sanitizer is closed on the GPU pool,
profiles/r02_sanitizer_closed.txt): identical inputs must give bit-identical outputs on every repetition, and no bounded
wait may have timed out."""
import pytest
import torch

from oracle import weights
from tests._util import make_unet

torch.set_grad_enabled(False)
pytestmark = pytest.mark.gpu


def test_chain_kernel_is_bit_stable_over_repetitions():
    import ldm_b200
    u = make_unet("init", "bf16")
    d = ldm_b200.ConditionalDenoiseDiffusion(u, 1000, torch.device("cuda"))
    eng = u.engine(torch.device("cuda"), 1000)
    eng.set_schedule(*d._host_schedule)
    for B in (256, 33, 100):                      # 6 x 48 rows with a ragged last cluster; one partial cluster; 32-row clusters
        c = (torch.arange(B) % 102).cuda()
        x_T = eng.randn(B, 256, 5, 0, 1000)
        ref = None
        for rep in range(8):
            x = x_T.clone()
            eng.sample(x, 999, 800, c, seed=5, use_graph=False)
            eng.check_device_flags()
            assert torch.isfinite(x).all()
            ref = x if ref is None else ref
            assert torch.equal(x, ref), (B, rep)


def test_attention_and_halo_kernels_are_bit_stable_over_repetitions():
    from ldm_b200 import v3, v4
    u3 = v3.ConditionalUNet(precision="bf16")
    u3.load_state_dict(weights.make_unet3_state(44, "init"))
    u3 = u3.cuda().eval()
    B = 300                                        # three query / key tiles of attn_tc_kernel
    x = torch.randn(B, 256, device="cuda")
    f, k = torch.arange(B, device="cuda") % 102, torch.arange(B, device="cuda") % 10
    ref = u3(x, torch.tensor([77], device="cuda"), f, k)
    for rep in range(10):
        assert torch.equal(u3(x, torch.tensor([77], device="cuda"), f, k), ref), rep
    m = v4.SimpleUNet()
    m.load_state_dict(weights.make_pix_state(45, "init"))
    m = m.cuda().eval()
    xi, t = torch.randn(8, 3, 64, 64, device="cuda"), torch.arange(8, device="cuda") * 100.0
    ref = m(xi, t)
    for rep in range(10):
        assert torch.equal(m(xi, t), ref), rep
    u3.engine(torch.device("cuda"), 1000).check_device_flags()


def test_v3_loop_kernel_is_bit_stable_over_repetitions():
    """unet3_loop_kernel: 19 grid-wide barriers per step between generic-proxy stores and TMA reads; any missing release /
    acquire / proxy fence would show up as run-to-run differences."""
    from ldm_b200 import v3
    u3 = v3.ConditionalUNet(precision="bf16")
    u3.load_state_dict(weights.make_unet3_state(44, "perturbed"))
    u3 = u3.cuda().eval()
    d = v3.ConditionalDenoiseDiffusion(u3, 1000, torch.device("cuda"))
    eng = d._engine("cuda")
    for B in (128, 37):
        f, k = torch.arange(B, device="cuda") % 102, torch.arange(B, device="cuda") % 10
        x_T = eng.randn(B, 256, 7, 0, 1000)
        ref = None
        for rep in range(8):
            x = x_T.clone()
            eng.sample3(x, 999, 900, f, k, seed=7, sample_offset=0, use_graph=False)
            assert torch.isfinite(x).all()
            ref = x if ref is None else ref
            assert torch.equal(x, ref), (B, rep)
        assert int(eng.info("launches_per_step")) == 0
        eng.check_device_flags(u3.num_classes)
    assert int(eng.info("tc_error")) == 0


def test_full_batch_decoder_kernels_against_the_oracle_and_over_repetitions():
    """The decoder kernels that only full batches reach (sa_map_gate_kernel: one CTA per sample, B >= 96) next to the ones every
    batch uses (convt_halo_kernel, final_gn_conv3_kernel, programmatic dependent launches between all of them): rows of a B = 128
    decode against the CPU restatement, against a small call of the same latents, and bit-stable over repetitions (an early
    start of a dependent kernel that read its predecessor's output before griddepcontrol.wait would show up here)."""
    from oracle import philox, restate as R
    from tests._util import AE_SEED, IMAGE_TOL, make_autoencoder
    sd = weights.make_autoencoder_state(AE_SEED, "perturbed")
    ae = make_autoencoder("perturbed", "bf16")
    z = torch.from_numpy(philox.normal_rows(11, 0, 128, 0))
    ref = ae.decode(z.cuda())
    for rep in range(6):
        assert torch.equal(ae.decode(z.cuda()), ref), rep
    rows = [0, 63, 64, 127]
    want = R.decode(sd, z[rows])
    got = ref[rows].cpu()
    assert float((got - want).abs().max()) < IMAGE_TOL["bf16"]
    small = ae.decode(z[rows].cuda()).cpu()                 # two-kernel map / gate sequence, other statistics splits
    assert float((small - got).abs().max()) < IMAGE_TOL["bf16"]
    assert float(ref.min()) > 0 and float(ref.max()) < 1
