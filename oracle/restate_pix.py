"""CPU restatement of the reference's v4 / v5 pixel-space sampling path (torch fp32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows v4/model_train_test.py operation for operation
(SimpleUNet.forward v4:99-135, DiffusionModel v4:140-183); v5 differs only by `out + res_ratio * x_input`
(v5:54,144), applied when the state_dict carries `res_ratio`.  Takes the reference's state_dict layout.

Pinned: tests/test_pix.py compares every function with the live reference in the build container and with
tests/golden/v4_*.npz / v5_*.npz (outputs of the reference itself, oracle/make_golden_pix.py) wherever the tests run.
"""
import torch
import torch.nn.functional as F


def time_embedding(sd, t):
    """v4:103-105: the RAW timestep as a float feature -> Linear(1,128) -> ReLU -> Linear(128,128). t: (B,) int64/float."""
    t = t.view(-1, 1).float()
    h = F.relu(F.linear(t, sd["time_embed.0.weight"], sd["time_embed.0.bias"]))
    return F.linear(h, sd["time_embed.2.weight"], sd["time_embed.2.bias"])


def stage_terms(sd, t):
    """v4:108-110: per-stage additive terms (B, 64), (B, 128), (B, 256)."""
    te = time_embedding(sd, t)
    return [F.linear(te, sd["time_fc%d.weight" % i], sd["time_fc%d.bias" % i]) for i in (1, 2, 3)]


def _pair(sd, name, x):
    """nn.Sequential(Conv3x3, ReLU, Conv3x3, ReLU) (v4:54-59 etc.)."""
    x = F.relu(F.conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], padding=1))
    return F.relu(F.conv2d(x, sd[name + ".2.weight"], sd[name + ".2.bias"], padding=1))


def unet_forward(sd, x, t, keep=None):
    """SimpleUNet.forward v4:99-135 (v5:101-146 when `res_ratio` is present). x (B,3,H,W) fp32, t (B,)."""
    B = x.size(0)
    e1, e2, e3 = [e.view(B, -1, 1, 1) for e in stage_terms(sd, t)]
    x1 = _pair(sd, "conv1", x) + e1
    x2 = F.conv2d(x1, sd["down1.weight"], sd["down1.bias"], stride=2, padding=1)
    x2 = _pair(sd, "conv2", x2) + e2
    x3 = F.conv2d(x2, sd["down2.weight"], sd["down2.bias"], stride=2, padding=1)
    x3 = _pair(sd, "conv3", x3) + e3
    x4 = _pair(sd, "bottleneck", x3)
    x5 = F.conv_transpose2d(x4, sd["up1.weight"], sd["up1.bias"], stride=2, padding=1)
    x5 = _pair(sd, "conv4", torch.cat([x5, x2], dim=1))
    x6 = F.conv_transpose2d(x5, sd["up2.weight"], sd["up2.bias"], stride=2, padding=1)
    x6 = _pair(sd, "conv5", torch.cat([x6, x1], dim=1))
    out = F.conv2d(x6, sd["out_conv.weight"], sd["out_conv.bias"], padding=1)
    if keep is not None:
        keep.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, x6=x6)
    if "res_ratio" in sd:
        out = out + sd["res_ratio"] * x
    return out


def p_sample(sd, sched, xt, t, noise=None):
    """DiffusionModel.p_sample v4:155-168; t is a python int, `noise` replaces randn_like (v4:163)."""
    beta, alpha, alpha_bar = sched
    B = xt.size(0)
    t_tensor = torch.full((B,), t, dtype=torch.long)
    eps_pred = unet_forward(sd, xt, t_tensor)
    alpha_t = alpha[t]
    alpha_bar_t = alpha_bar[t]
    mean = (xt - ((1 - alpha_t) / torch.sqrt(1 - alpha_bar_t)) * eps_pred) / torch.sqrt(alpha_t)
    if t > 0:
        if noise is None:
            noise = torch.randn_like(xt)
        return mean + torch.sqrt(beta[t]) * noise
    return mean


def sample(sd, sched, x_T, noise_fn=None, t_start=None, t_end=0):
    """DiffusionModel.sample v4:170-175 from a given x_T; noise_fn(t) supplies the draw of v4:163."""
    x = x_T
    t_start = sched[0].shape[0] - 1 if t_start is None else t_start
    for t in range(t_start, t_end - 1, -1):
        x = p_sample(sd, sched, x, t, noise_fn(t) if (noise_fn is not None and t > 0) else None)
    return x


def q_sample(sched, x0, t, noise):
    """v4:148-153."""
    ab = sched[2][t].view(-1, 1, 1, 1)
    return torch.sqrt(ab) * x0 + torch.sqrt(1 - ab) * noise
