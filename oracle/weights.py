"""Deterministic synthetic weights in the reference's state_dict layout.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference ships no
checkpoints (.gitignore drops *.pt), so parity and the benchmark use
random-init weights.  The generator is independent of module construction
order: every tensor is drawn from its own torch.Generator seeded by
(seed, crc32(key)), so the reference (in the build container) and the product
(on the GPU box) receive bit-identical tensors.

Layouts follow v2/model_train_test.py: ConditionalUNet v2:501-533 (82 tensors),
SimpleAutoencoder v2:305-324 with Encoder v2:181-222 and Decoder v2:242-278
(132 tensors incl. the two buffers).  `style`:
  "init"      kaiming_normal(a=0.2) weights / zero biases / unit norms, i.e. the
              statistics of v2:326-337 and v2:1346-1350 (MHA in_proj keeps
              xavier-uniform, embedding N(0,1), residual_weight 0.1)
  "perturbed" init + 0.05*randn on EVERY tensor, so biases, norm affines and
              the unused parameters are all non-trivial.
"""
import math
import zlib

import torch

HIDDEN = [256, 512, 1024, 512, 256]


def unet_spec(latent_dim=256, hidden=HIDDEN, temb=256, num_classes=102):
    """[(key, shape, kind)] in the reference's state_dict order (v2:501-533)."""
    s = [("residual_weight", (), "rw")]
    def lin(name, o, i):
        s.append((name + ".weight", (o, i), "w")); s.append((name + ".bias", (o,), "b"))
    def ln(name, d):
        s.append((name + ".weight", (d,), "g")); s.append((name + ".bias", (d,), "b"))
    lin("time_emb.lin1", 2 * temb, temb); lin("time_emb.lin2", temb, 2 * temb)
    s.append(("class_emb.embedding.weight", (num_classes, temb), "emb"))
    lin("class_emb.lin1", temb, temb); lin("class_emb.lin2", temb, temb)
    lin("latent_proj", hidden[0], latent_dim)
    for i, d in enumerate(hidden):
        lin(f"time_projections.{i}", d, temb)
    for i, d in enumerate(hidden):
        s.append((f"attention_layers.{i}.in_proj_weight", (3 * d, d), "xavier"))
        s.append((f"attention_layers.{i}.in_proj_bias", (3 * d,), "b"))
        lin(f"attention_layers.{i}.out_proj", d, d)
    for i in range(len(hidden) - 1):
        lin(f"layers.{i}.0.0", hidden[i], hidden[i]); ln(f"layers.{i}.0.1", hidden[i])
        ln(f"layers.{i}.1", hidden[i]); lin(f"layers.{i}.2", hidden[i + 1], hidden[i])
    lin("final_time_proj", hidden[-1], temb); lin("final_class_proj", hidden[-1], temb)
    ln("final_norm", hidden[-1]); lin("final", latent_dim, hidden[-1])
    return s


def unet3_spec(latent_dim=256, hidden=HIDDEN, temb=256, num_classes=102, num_colors=10):
    """v3 ConditionalUNet (v3:769-803), state_dict order: 91 tensors."""
    s = [("residual_weight", (), "rw")]
    def lin(name, o, i):
        s.append((name + ".weight", (o, i), "w")); s.append((name + ".bias", (o,), "b"))
    def ln(name, d):
        s.append((name + ".weight", (d,), "g")); s.append((name + ".bias", (d,), "b"))
    lin("time_emb.lin1", 2 * temb, temb); lin("time_emb.lin2", temb, 2 * temb)
    s.append(("multi_cond_emb.flower_emb.weight", (num_classes, temb), "emb"))
    s.append(("multi_cond_emb.color_emb.weight", (num_colors, temb), "emb"))
    lin("multi_cond_emb.fc", temb, 2 * temb)
    lin("latent_proj", hidden[0], latent_dim)
    for i, d in enumerate(hidden):
        lin(f"time_projections.{i}", d, temb)
    for i, d in enumerate(hidden):
        lin(f"cond_projections.{i}", d, temb)
    for i, d in enumerate(hidden):
        s.append((f"attention_layers.{i}.in_proj_weight", (3 * d, d), "xavier"))
        s.append((f"attention_layers.{i}.in_proj_bias", (3 * d,), "b"))
        lin(f"attention_layers.{i}.out_proj", d, d)
    for i in range(len(hidden) - 1):
        lin(f"layers.{i}.0.0", hidden[i], hidden[i]); ln(f"layers.{i}.0.1", hidden[i])
        ln(f"layers.{i}.1", hidden[i]); lin(f"layers.{i}.2", hidden[i + 1], hidden[i])
    lin("final_time_proj", hidden[-1], temb); lin("final_class_proj", hidden[-1], temb)
    ln("final_norm", hidden[-1]); lin("final", latent_dim, hidden[-1])
    return s


def _resblock(s, name, c):
    for j in (1, 2):
        s.append((f"{name}.conv{j}.weight", (c, c, 3, 3), "w")); s.append((f"{name}.conv{j}.bias", (c,), "b"))
        s.append((f"{name}.ln{j}.weight", (c,), "g")); s.append((f"{name}.ln{j}.bias", (c,), "b"))
    s.append((f"{name}.ca.conv_du.0.weight", (c // 8, c, 1, 1), "w"))
    s.append((f"{name}.ca.conv_du.2.weight", (c, c // 8, 1, 1), "w"))
    s.append((f"{name}.sa.conv.weight", (1, 2, 7, 7), "w"))


def decoder_spec(latent_dim=256, out_channels=3, prefix="decoder."):
    """Decoder tensors (v2:242-278), 59 entries."""
    s = []
    def wb(name, shape, fan_kind="w"):
        s.append((name + ".weight", shape, fan_kind)); s.append((name + ".bias", (shape[0],), "b"))
    def norm(name, d):
        s.append((name + ".weight", (d,), "g")); s.append((name + ".bias", (d,), "b"))
    wb("fc.0", (512, latent_dim)); norm("fc.1", 512)
    wb("fc.3", (512 * 8 * 8, 512)); norm("fc.4", 512 * 8 * 8)
    for name, c in (("3", 512), ("2", 256), ("1", 128)):
        _resblock(s, "res" + name, c)
        # ConvTranspose2d weight is (Cin, Cout, 4, 4); bias is Cout (v2:256)
        s.append((f"up{name}.0.weight", (c, c // 2, 4, 4), "wT")); s.append((f"up{name}.0.bias", (c // 2,), "b"))
        norm(f"up{name}.1", c // 2)
    wb("final_conv.0", (32, 64, 3, 3)); norm("final_conv.1", 32)
    wb("final_conv.3", (out_channels, 32, 3, 3))
    # reorder into the module's registration order: res3, up3, res2, up2, res1, up1 already holds
    return [(prefix + k, sh, kd) for k, sh, kd in s]


def encoder_spec(in_channels=3, latent_dim=256, prefix="encoder."):
    """Encoder tensors (v2:181-222). Not on the hot path; needed so a full AE
    state_dict round-trips."""
    s = []
    def wb(name, shape):
        s.append((name + ".weight", shape, "w")); s.append((name + ".bias", (shape[0],), "b"))
    def norm(name, d):
        s.append((name + ".weight", (d,), "g")); s.append((name + ".bias", (d,), "b"))
    wb("initial_conv.0", (64, in_channels, 3, 3)); norm("initial_conv.1", 64)
    for i, (ci, co) in enumerate(((64, 128), (128, 256), (256, 512)), start=1):
        wb(f"down{i}.0", (co, ci, 4, 4)); norm(f"down{i}.1", co)
        _resblock(s, f"res{i}", co)
    for head in ("fc_mu", "fc_logvar"):
        wb(head + ".0", (512, 512 * 8 * 8)); norm(head + ".1", 512); wb(head + ".3", (latent_dim, 512))
    return [(prefix + k, sh, kd) for k, sh, kd in s]


def classifier_spec(latent_dim=256, num_classes=102, prefix="classifier."):
    s = []
    def wb(name, shape):
        s.append((name + ".weight", shape, "w")); s.append((name + ".bias", (shape[0],), "b"))
    def norm(name, d):
        s.append((name + ".weight", (d,), "g")); s.append((name + ".bias", (d,), "b"))
    wb("0", (512, latent_dim)); norm("1", 512); wb("4", (256, 512)); norm("5", 256); wb("8", (num_classes, 256))
    return [(prefix + k, sh, kd) for k, sh, kd in s]


def autoencoder_spec():
    """Full SimpleAutoencoder state_dict (v2:305-324): buffers first, then
    encoder, decoder, classifier."""
    return ([("class_centers", (102, 256), "buf"), ("center_counts", (102,), "buf")]
            + encoder_spec() + decoder_spec() + classifier_spec())


def pix_unet_spec(in_channels=3, base=64, temb=128, v5=False):
    """v4 SimpleUNet (v4:37-97), state_dict order: 44 tensors; v5 adds the scalar `res_ratio` first (v5:54).
    The reference never re-initialises this model, so "init" is torch's default: U(+-1/sqrt(fan_in)) for weights
    AND biases (kinds "u:<fan_in>")."""
    s = [("res_ratio", (), "rw")] if v5 else []
    def lin(name, o, i):
        s.append((name + ".weight", (o, i), "u:%d" % i)); s.append((name + ".bias", (o,), "u:%d" % i))
    def conv(name, o, i, k):
        s.append((name + ".weight", (o, i, k, k), "u:%d" % (i * k * k))); s.append((name + ".bias", (o,), "u:%d" % (i * k * k)))
    def convT(name, i, o, k):   # weight (Cin, Cout, k, k); torch's fan_in of it is Cout * k * k
        s.append((name + ".weight", (i, o, k, k), "u:%d" % (o * k * k))); s.append((name + ".bias", (o,), "u:%d" % (o * k * k)))
    lin("time_embed.0", temb, 1); lin("time_embed.2", temb, temb)
    lin("time_fc1", base, temb); lin("time_fc2", 2 * base, temb); lin("time_fc3", 4 * base, temb)
    conv("conv1.0", base, in_channels, 3); conv("conv1.2", base, base, 3)
    conv("down1", 2 * base, base, 4)
    conv("conv2.0", 2 * base, 2 * base, 3); conv("conv2.2", 2 * base, 2 * base, 3)
    conv("down2", 4 * base, 2 * base, 4)
    conv("conv3.0", 4 * base, 4 * base, 3); conv("conv3.2", 4 * base, 4 * base, 3)
    conv("bottleneck.0", 8 * base, 4 * base, 3); conv("bottleneck.2", 4 * base, 8 * base, 3)
    convT("up1", 4 * base, 2 * base, 4)
    conv("conv4.0", 2 * base, 4 * base, 3); conv("conv4.2", 2 * base, 2 * base, 3)
    convT("up2", 2 * base, base, 4)
    conv("conv5.0", base, 2 * base, 3); conv("conv5.2", base, base, 3)
    conv("out_conv", in_channels, base, 3)
    return s


def ublock_res_spec(cin, cout, d_time=256, prefix=""):
    """UNetResidualBlock (v2:462-473) state_dict order; torch default init ("u:<fan_in>")."""
    s = [(prefix + "norm1.weight", (cin,), "g"), (prefix + "norm1.bias", (cin,), "b"),
         (prefix + "conv1.weight", (cout, cin, 3, 3), "u:%d" % (cin * 9)), (prefix + "conv1.bias", (cout,), "u:%d" % (cin * 9)),
         (prefix + "time_emb.weight", (cout, d_time), "u:%d" % d_time), (prefix + "time_emb.bias", (cout,), "u:%d" % d_time),
         (prefix + "class_emb.weight", (cout, d_time), "u:%d" % d_time), (prefix + "class_emb.bias", (cout,), "u:%d" % d_time),
         (prefix + "norm2.weight", (cout,), "g"), (prefix + "norm2.bias", (cout,), "b"),
         (prefix + "conv2.weight", (cout, cout, 3, 3), "u:%d" % (cout * 9)), (prefix + "conv2.bias", (cout,), "u:%d" % (cout * 9))]
    if cin != cout:
        s += [(prefix + "residual.weight", (cout, cin, 1, 1), "u:%d" % cin), (prefix + "residual.bias", (cout,), "u:%d" % cin)]
    return s


def ublock_attn_spec(c, prefix=""):
    """UNetAttentionBlock (v2:435-442) state_dict order."""
    return [(prefix + "norm.weight", (c,), "g"), (prefix + "norm.bias", (c,), "b"),
            (prefix + "qkv.weight", (3 * c, c, 1, 1), "u:%d" % c), (prefix + "qkv.bias", (3 * c,), "u:%d" % c),
            (prefix + "proj.weight", (c, c, 1, 1), "u:%d" % c), (prefix + "proj.bias", (c,), "u:%d" % c)]


def _draw(key, shape, kind, seed, style):
    g = torch.Generator().manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    n = lambda: torch.randn(shape, generator=g, dtype=torch.float32)
    if kind in ("w", "wT"):
        if len(shape) == 4:
            # torch fan_in = size(1) * receptive field (also for ConvTranspose2d, whose dim 1 is Cout)
            fan_in = shape[1] * shape[2] * shape[3]
        else:
            fan_in = shape[1]
        gain = math.sqrt(2.0 / (1 + 0.2 ** 2))
        t = n() * (gain / math.sqrt(fan_in))
    elif kind.startswith("u:"):
        bound = 1.0 / math.sqrt(int(kind[2:]))
        t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    elif kind == "xavier":
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    elif kind == "emb":
        t = n()
    elif kind == "g":
        t = torch.ones(shape)
    elif kind == "rw":
        t = torch.tensor(0.1)
    else:  # "b", "buf"
        t = torch.zeros(shape)
    if style == "perturbed" and kind.startswith("u:"):
        t = t + 0.5 / math.sqrt(int(kind[2:])) * n()     # relative to the init scale, so that 17 stacked convolutions keep a finite gain
    elif style == "perturbed":
        t = t + 0.05 * n()
    elif style != "init":
        raise ValueError(style)
    return t.contiguous()


def make_state(spec, seed=42, style="init"):
    return {k: _draw(k, sh, kd, seed, style) for k, sh, kd in spec}


def make_unet_state(seed=42, style="init"):
    return make_state(unet_spec(), seed, style)


def make_unet3_state(seed=44, style="init"):
    return make_state(unet3_spec(), seed, style)


def make_pix_state(seed=45, style="init", v5=False):
    """"perturbed" adds 0.5 / sqrt(fan_in) * randn to every tensor (half the init bound), res_ratio 0.1 + 0.05 randn."""
    return make_state(pix_unet_spec(v5=v5), seed, style)


def make_decoder_state(seed=43, style="init"):
    return make_state(decoder_spec(), seed, style)


def make_autoencoder_state(seed=43, style="init"):
    """Decoder tensors are identical to make_decoder_state(seed, style)."""
    return make_state(autoencoder_spec(), seed, style)
