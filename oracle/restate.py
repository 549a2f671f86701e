"""CPU restatement of the reference's v2 sampling hot path (torch fp32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker the CUDA path is
compared against on the GPU box, where /root/reference does not exist.  Every
function follows the reference line by line in the SAME operation order (no
algebraic folding), takes the reference's state_dict layout, and cites
v2/model_train_test.py.  All arithmetic is torch fp32 because the reference's
arithmetic IS torch fp32 (SURVEY.md 8c "third-party arithmetic").

Pinned: tests/test_oracle_vs_reference.py compares every function here with the
live reference in the build container; tests/test_oracle_golden.py compares it
with tests/golden/*.npz (outputs of the reference itself, made by
oracle/make_golden.py) wherever the tests run.
"""
import math

import torch
import torch.nn.functional as F


def swish(x):
    """v2:48-50 / v2:396-398."""
    return x * torch.sigmoid(x)


# --------------------------------------------------------------------------
# schedule (a1)
# --------------------------------------------------------------------------
def schedule(n_steps=1000):
    """v2:569-571: beta = linspace(1e-4, 0.02, n), alpha = 1 - beta,
    alpha_bar = cumprod(alpha).  Built with torch on CPU exactly as the
    reference does when device is cpu/None."""
    beta = torch.linspace(0.0001, 0.02, n_steps)
    alpha = 1 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return beta, alpha, alpha_bar


# --------------------------------------------------------------------------
# denoiser (a4-a10)
# --------------------------------------------------------------------------
def sinusoid(t, n_channels=256):
    """v2:410-414: the sinusoidal features BEFORE the MLP. t: int64 (n,)."""
    half_dim = n_channels // 2
    emb = math.log(10000) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, device=t.device) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=1)


def time_embedding(sd, t, n_channels=256):
    """TimeEmbedding.forward v2:409-418 (256 == 2*128 so the padding branch
    v2:415-417 never runs)."""
    e = sinusoid(t, n_channels)
    h = F.linear(e, sd["time_emb.lin1.weight"], sd["time_emb.lin1.bias"])
    return F.linear(swish(h), sd["time_emb.lin2.weight"], sd["time_emb.lin2.bias"])


def class_embedding(sd, c):
    """ClassEmbedding.forward v2:429-431."""
    e = sd["class_emb.embedding.weight"][c]
    h = F.linear(e, sd["class_emb.lin1.weight"], sd["class_emb.lin1.bias"])
    return F.linear(swish(h), sd["class_emb.lin2.weight"], sd["class_emb.lin2.bias"])


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def unet_forward(sd, x, t, c=None, n_stages=None, v1=False, literal_attention=False):
    """ConditionalUNet.forward v2:535-561, eval mode (Dropout v2:521 is the
    identity).  The L=1 attention (v2:550-552) is softmax over ONE key, i.e.
    out_proj(V(h_norm)) with V = in_proj rows [2d:3d].
    literal_attention=True runs torch's multi_head_attention_forward exactly as
    nn.MultiheadAttention.forward does for the reference's call (need_weights
    defaults to True, so the fused fast path is not taken): same result, and
    the reference's real CPU cost (full 3d in_proj, bmm, softmax) -- used when
    the oracle is TIMED as the CPU baseline."""
    if n_stages is None:   # len(hidden_dims) - 1 (v2:517): one layers.{i} entry per stage
        n_stages = sum(1 for k in sd if k.startswith("layers.") and k.endswith(".2.weight"))
    residual = x
    te = time_embedding(sd, t)                                   # v2:537
    ce = class_embedding(sd, c) if c is not None else None       # v2:538
    h = F.linear(x, sd["latent_proj.weight"], sd["latent_proj.bias"])  # v2:539
    for i in range(n_stages):
        tw, tb = sd[f"time_projections.{i}.weight"], sd[f"time_projections.{i}.bias"]
        h = h + F.linear(te, tw, tb)                             # v2:541-542
        if ce is not None:
            h = h + F.linear(ce, tw, tb)                         # v2:543-545 (same Linear: bias twice)
        h_res = h
        u = F.linear(h, sd[f"layers.{i}.0.0.weight"], sd[f"layers.{i}.0.0.bias"])
        h = swish(_ln(u, sd, f"layers.{i}.0.1")) + h_res         # v2:546-548
        n = _ln(h, sd, f"layers.{i}.1")                          # v2:549
        d = n.shape[-1]
        wv = sd[f"attention_layers.{i}.in_proj_weight"][2 * d:3 * d]
        bv = sd[f"attention_layers.{i}.in_proj_bias"][2 * d:3 * d]
        if literal_attention:
            q = n.unsqueeze(0)                                   # v2:550: (L=1, N=B, E)
            a, _ = F.multi_head_attention_forward(
                q, q, q, d, 8, sd[f"attention_layers.{i}.in_proj_weight"], sd[f"attention_layers.{i}.in_proj_bias"],
                None, None, False, 0.3, sd[f"attention_layers.{i}.out_proj.weight"],
                sd[f"attention_layers.{i}.out_proj.bias"], training=False, need_weights=True)
            a = a.squeeze(0)
        else:
            a = F.linear(F.linear(n, wv, bv), sd[f"attention_layers.{i}.out_proj.weight"],
                         sd[f"attention_layers.{i}.out_proj.bias"])  # v2:550-551
        h = h + a                                                # v2:552
        h = F.linear(h, sd[f"layers.{i}.2.weight"], sd[f"layers.{i}.2.bias"])  # v2:553
    h = h + F.linear(te, sd["final_time_proj.weight"], sd["final_time_proj.bias"])      # v2:554-555
    if ce is not None:
        h = h + F.linear(ce, sd["final_class_proj.weight"], sd["final_class_proj.bias"])  # v2:556-558
    h = _ln(h, sd, "final_norm")                                 # v2:559
    out = F.linear(h, sd["final.weight"], sd["final.bias"])      # v2:560
    if v1:
        return out                                               # v1:561
    return out + torch.sigmoid(sd["residual_weight"]) * F.linear(residual, sd["final.weight"], sd["final.bias"])  # v2:561


# --------------------------------------------------------------------------
# v3 multi-conditional denoiser: flower + colour condition, attention ACROSS the batch
# --------------------------------------------------------------------------
def multi_cond_embedding(sd, flower, color):
    """MultiConditionEmbedding.forward v3:745-749."""
    e = torch.cat((sd["multi_cond_emb.flower_emb.weight"][flower], sd["multi_cond_emb.color_emb.weight"][color]), dim=-1)
    return F.linear(e, sd["multi_cond_emb.fc.weight"], sd["multi_cond_emb.fc.bias"])


def batch_attention(n, w_in, b_in, w_out, b_out, heads=8):
    """nn.MultiheadAttention on h_norm.unsqueeze(1) (v3:832-835): with batch_first=False that is (L=B, N=1, E), so the
    B samples of a call attend to EACH OTHER.  Eval mode: no dropout."""
    B, d = n.shape
    hd = d // heads
    qkv = F.linear(n, w_in, b_in)
    q, k, v = (t.reshape(B, heads, hd).transpose(0, 1) for t in qkv.split(d, dim=1))     # (heads, B, hd)
    p = torch.softmax((q * (hd ** -0.5)) @ k.transpose(1, 2), dim=-1)                      # (heads, B, B)
    o = (p @ v).transpose(0, 1).reshape(B, d)
    return F.linear(o, w_out, b_out)


def unet3_forward(sd, x, t, flower, color, n_stages=None):
    """v3 ConditionalUNet.forward v3:804-853 (eval mode): separate cond_projections (no double bias), real
    cross-batch attention, and NO final residual (`return out`, v3:853; residual_weight is unused)."""
    if n_stages is None:
        n_stages = sum(1 for k in sd if k.startswith("layers.") and k.endswith(".2.weight"))
    te = time_embedding(sd, t)                                   # v3:809
    ce = multi_cond_embedding(sd, flower, color)                 # v3:810
    h = F.linear(x, sd["latent_proj.weight"], sd["latent_proj.bias"])
    for i in range(n_stages):
        h = h + F.linear(te, sd[f"time_projections.{i}.weight"], sd[f"time_projections.{i}.bias"]) \
              + F.linear(ce, sd[f"cond_projections.{i}.weight"], sd[f"cond_projections.{i}.bias"])   # v3:818-822
        u = F.linear(h, sd[f"layers.{i}.0.0.weight"], sd[f"layers.{i}.0.0.bias"])
        h = swish(_ln(u, sd, f"layers.{i}.0.1")) + h             # v3:825-827
        n = _ln(h, sd, f"layers.{i}.1")                          # v3:830
        h = h + batch_attention(n, sd[f"attention_layers.{i}.in_proj_weight"], sd[f"attention_layers.{i}.in_proj_bias"],
                                sd[f"attention_layers.{i}.out_proj.weight"], sd[f"attention_layers.{i}.out_proj.bias"])
        h = F.linear(h, sd[f"layers.{i}.2.weight"], sd[f"layers.{i}.2.bias"])   # v3:841
    h = h + F.linear(te, sd["final_time_proj.weight"], sd["final_time_proj.bias"]) \
          + F.linear(ce, sd["final_class_proj.weight"], sd["final_class_proj.bias"])   # v3:844-846
    return F.linear(_ln(h, sd, "final_norm"), sd["final.weight"], sd["final.bias"])   # v3:849-853


def sample3(sd, sched, x_T, flower, color, noise_fn=None, t_start=None, t_end=0):
    """v3 ConditionalDenoiseDiffusion.sample / p_sample (v3:876-893) from a given x_T with supplied noise."""
    n_steps = sched[0].shape[0]
    x = x_T
    t_start = n_steps - 1 if t_start is None else t_start
    for t in range(t_start, t_end - 1, -1):
        eps = unet3_forward(sd, x, torch.tensor([t]), flower, color)
        x = ddpm_update(sched, x, eps, t, noise_fn(t) if (noise_fn is not None and t > 0) else None)
    return x


# --------------------------------------------------------------------------
# DDPM reverse process (a2, a3)
# --------------------------------------------------------------------------
def p_sample(sd, sched, xt, t, c=None, noise=None, literal_attention=False):
    """ConditionalDenoiseDiffusion.p_sample v2:580-592.  `t` is a python int
    or an int64 tensor of shape (1,); `noise` replaces randn_like (v2:589)."""
    beta, alpha, alpha_bar = sched
    if not isinstance(t, torch.Tensor):
        t = torch.tensor([t])
    eps_theta = unet_forward(sd, xt, t, c, literal_attention=literal_attention)
    alpha_t = alpha[t].reshape(-1, 1)
    alpha_bar_t = alpha_bar[t].reshape(-1, 1)
    mean = (xt - ((1 - alpha_t) / torch.sqrt(1 - alpha_bar_t)) * eps_theta) / torch.sqrt(alpha_t)
    var = beta[t].reshape(-1, 1)
    if t[0] > 0:
        if noise is None:
            noise = torch.randn_like(xt)
        return mean + torch.sqrt(var) * noise
    return mean


def ddpm_update(sched, xt, eps_theta, t, noise=None):
    """The posterior update of v2:584-592 alone, given eps."""
    beta, alpha, alpha_bar = sched
    t = torch.tensor([int(t)])
    alpha_t = alpha[t].reshape(-1, 1)
    alpha_bar_t = alpha_bar[t].reshape(-1, 1)
    mean = (xt - ((1 - alpha_t) / torch.sqrt(1 - alpha_bar_t)) * eps_theta) / torch.sqrt(alpha_t)
    if int(t[0]) > 0:
        return mean + torch.sqrt(beta[t].reshape(-1, 1)) * noise
    return mean


def sample(sd, sched, x_T, c=None, noise_fn=None, t_start=None, t_end=0, keep=()):
    """ConditionalDenoiseDiffusion.sample v2:594-598 from a given x_T.
    noise_fn(t) -> (B, D) tensor supplies the draw of v2:589 for step t
    (t = t_start .. 1).  Returns x_{t_end} and {t: x after step t} for t in keep."""
    n_steps = sched[0].shape[0]
    x = x_T
    kept = {}
    t_start = n_steps - 1 if t_start is None else t_start
    for t in range(t_start, t_end - 1, -1):
        nz = noise_fn(t) if (noise_fn is not None and t > 0) else None
        x = p_sample(sd, sched, x, t, c, noise=nz)
        if t in keep:
            kept[t] = x.clone()
    return x, kept


def q_sample(sched, x0, t, eps):
    """v2:574-578."""
    alpha_bar_t = sched[2][t].reshape(-1, 1)
    return torch.sqrt(alpha_bar_t) * x0 + torch.sqrt(1 - alpha_bar_t) * eps


# --------------------------------------------------------------------------
# VAE decoder (a11-a15)
# --------------------------------------------------------------------------
def layernorm2d(x, w, b, eps=1e-5):
    """LayerNorm2d.forward v2:151-156: per-(n, c) statistics over H x W."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), keepdim=True, unbiased=False)
    x = (x - mean) / torch.sqrt(var + eps)
    return x * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def ca_layer(x, w0, w2):
    """CALayer.forward v2:64-67 (1x1 convs without bias, v2:58-61)."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = torch.sigmoid(F.conv2d(swish(F.conv2d(y, w0)), w2))
    return x * y


def spatial_attention(x, w):
    """SpatialAttention.forward v2:75-81."""
    avg_out = torch.mean(x, dim=1, keepdim=True)
    max_out, _ = torch.max(x, dim=1, keepdim=True)
    a = torch.sigmoid(F.conv2d(torch.cat([avg_out, max_out], dim=1), w, padding=w.shape[-1] // 2))
    return x * a


def residual_block(sd, p, x):
    """ResidualBlock.forward v2:170-178."""
    out = swish(layernorm2d(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1),
                            sd[p + "ln1.weight"], sd[p + "ln1.bias"]))
    out = layernorm2d(F.conv2d(out, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1),
                      sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    out = ca_layer(out, sd[p + "ca.conv_du.0.weight"], sd[p + "ca.conv_du.2.weight"])
    out = spatial_attention(out, sd[p + "sa.conv.weight"])
    return swish(out + x)


def up_block(sd, p, x, groups):
    """up3/up2/up1 v2:255-271: ConvTranspose2d(4, 2, 1) -> GroupNorm -> Swish."""
    y = F.conv_transpose2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1)
    return swish(F.group_norm(y, groups, sd[p + "1.weight"], sd[p + "1.bias"], 1e-5))


def decoder_fc(sd, z, p="decoder."):
    """Decoder.fc v2:246-253 then view v2:282. Returns (B, 512, 8, 8)."""
    x = F.linear(z, sd[p + "fc.0.weight"], sd[p + "fc.0.bias"])
    x = swish(F.layer_norm(x, (512,), sd[p + "fc.1.weight"], sd[p + "fc.1.bias"], 1e-5))
    x = F.linear(x, sd[p + "fc.3.weight"], sd[p + "fc.3.bias"])
    x = swish(F.layer_norm(x, (512 * 8 * 8,), sd[p + "fc.4.weight"], sd[p + "fc.4.bias"], 1e-5))
    return x.view(-1, 512, 8, 8)


def decode(sd, z, p="decoder.", stages=None):
    """SimpleAutoencoder.decode v2:355-357 -> Decoder.forward v2:280-290.
    `stages`, if a dict, receives the intermediate activations (NCHW)."""
    x = decoder_fc(sd, z, p)
    rec = (lambda k, v: stages.__setitem__(k, v.clone())) if stages is not None else (lambda k, v: None)
    rec("fc", x)
    x = residual_block(sd, p + "res3.", x); rec("res3", x)
    x = up_block(sd, p + "up3.", x, 32); rec("up3", x)
    x = residual_block(sd, p + "res2.", x); rec("res2", x)
    x = up_block(sd, p + "up2.", x, 16); rec("up2", x)
    x = residual_block(sd, p + "res1.", x); rec("res1", x)
    x = up_block(sd, p + "up1.", x, 8); rec("up1", x)
    x = F.conv2d(x, sd[p + "final_conv.0.weight"], sd[p + "final_conv.0.bias"], padding=1)   # v2:273
    x = swish(F.group_norm(x, 8, sd[p + "final_conv.1.weight"], sd[p + "final_conv.1.bias"], 1e-5))
    rec("final0", x)
    x = torch.sigmoid(F.conv2d(x, sd[p + "final_conv.3.weight"], sd[p + "final_conv.3.bias"], padding=1))
    return x


# --------------------------------------------------------------------------
# the conv U-Net blocks the v2 script defines but never instantiates
# --------------------------------------------------------------------------
def unet_residual_block(sd, p, x, t, c=None):
    """UNetResidualBlock.forward v2:475-486, eval mode (Dropout v2:484 = identity). x (B, Cin, H, W); t, c (B, d_time)
    embedding vectors.  `residual` is nn.Identity (no parameters) when in_channels == out_channels (v2:473)."""
    h = swish(layernorm2d(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"]))
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    t_emb = swish(F.linear(t, sd[p + "time_emb.weight"], sd[p + "time_emb.bias"]))
    h = h + t_emb.view(-1, t_emb.shape[1], 1, 1)
    if c is not None:
        c_emb = swish(F.linear(c, sd[p + "class_emb.weight"], sd[p + "class_emb.bias"]))
        h = h + c_emb.view(-1, c_emb.shape[1], 1, 1)
    h = swish(layernorm2d(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"]))
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    res = F.conv2d(x, sd[p + "residual.weight"], sd[p + "residual.bias"]) if p + "residual.weight" in sd else x
    return h + res


def unet_attention_block(sd, p, x, num_heads=4):
    """UNetAttentionBlock.forward v2:444-459, line by line (including the (head_dim, head) channel order that
    out.permute(0, 3, 1, 2).reshape(b, c, h, w) produces, v2:456-457)."""
    b, c, h, w = x.shape
    residual = x
    x = F.group_norm(x, 1, sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-5)
    qkv = F.conv2d(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(b, 3, num_heads, c // num_heads, h * w)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q = q.permute(0, 1, 3, 2)
    k = k.permute(0, 1, 2, 3)
    v = v.permute(0, 1, 3, 2)
    scale = (c // num_heads) ** -0.5
    attn = torch.matmul(q, k) * scale
    attn = F.softmax(attn, dim=-1)
    out = torch.matmul(attn, v)
    out = out.permute(0, 3, 1, 2)
    out = out.reshape(b, c, h, w)
    return F.conv2d(out, sd[p + "proj.weight"], sd[p + "proj.bias"]) + residual


# --------------------------------------------------------------------------
# error measures used by every parity test
# --------------------------------------------------------------------------
def max_rel(a, ref):
    """max|a - ref| / max|ref|  (north_star's 'max relative error'; element-wise
    relative error is unbounded near zeros)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def rel_l2(a, ref):
    a = torch.as_tensor(a, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    return float((a - ref).norm() / ref.norm().clamp_min(1e-30))
