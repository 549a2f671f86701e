"""Drop-in mirror of the v4 / v5 pixel-space diffusion model (v4/model_train_test.py:37-183, v5/model_train_test.py:
38-196): same class names, constructor arguments, method signatures and state_dict keys; the arithmetic runs in
libldm_b200.so (ldm_pix_pack / ldm_pix_forward / ldm_pix_sample: implicit-GEMM convolutions on tcgen05 in bf16 - the
default - or the strict fp32 CUDA-core path with `precision="fp32"`).

    from ldm_b200 import v4
    model = v4.SimpleUNet().to("cuda").eval()                 # v5: v4.SimpleUNet(res_ratio=True)
    model.load_state_dict(torch.load("diffusion_model.pt"))   # a v4 (44 tensors) or v5 (45 tensors) checkpoint, unchanged
    diffusion = v4.DiffusionModel(model, n_steps=1000, device="cuda")
    images = diffusion.sample((B, 3, 64, 64))

Samples are independent (plain convolutions and ReLU, no normalisation), so a batch shards over GPUs by sample like
v2: pass `sample_offset` = the first global sample index of the shard and the Philox draws do not depend on the
sharding."""
import torch
from torch import nn

from .engine import get_engine
from .modules import _fresh_seed, _require_eval


class SimpleUNet(nn.Module):
    """v4:37-135.  `res_ratio=True` adds v5's learnable residual ratio (v5:54,144) - it is also switched on
    automatically by load_state_dict when the checkpoint carries the `res_ratio` key."""

    def __init__(self, in_channels=3, base_channels=64, time_emb_dim=128, *, res_ratio=False, precision=None, max_timesteps=1000):
        super().__init__()
        self.in_channels, self.base_channels, self.time_emb_dim = in_channels, base_channels, time_emb_dim
        self.precision = precision
        self.max_timesteps = max_timesteps
        c = base_channels
        self.time_embed = nn.Sequential(nn.Linear(1, time_emb_dim), nn.ReLU(), nn.Linear(time_emb_dim, time_emb_dim))
        self.time_fc1 = nn.Linear(time_emb_dim, c)
        self.time_fc2 = nn.Linear(time_emb_dim, c * 2)
        self.time_fc3 = nn.Linear(time_emb_dim, c * 4)
        if res_ratio:
            self.res_ratio = nn.Parameter(torch.tensor(0.1))

        def pair(i, m, o):
            return nn.Sequential(nn.Conv2d(i, m, 3, padding=1), nn.ReLU(), nn.Conv2d(m, o, 3, padding=1), nn.ReLU())

        self.conv1 = pair(in_channels, c, c)
        self.down1 = nn.Conv2d(c, c * 2, 4, stride=2, padding=1)
        self.conv2 = pair(c * 2, c * 2, c * 2)
        self.down2 = nn.Conv2d(c * 2, c * 4, 4, stride=2, padding=1)
        self.conv3 = pair(c * 4, c * 4, c * 4)
        self.bottleneck = pair(c * 4, c * 8, c * 4)
        self.up1 = nn.ConvTranspose2d(c * 4, c * 2, 4, stride=2, padding=1)
        self.conv4 = pair(c * 4, c * 2, c * 2)
        self.up2 = nn.ConvTranspose2d(c * 2, c, 4, stride=2, padding=1)
        self.conv5 = pair(c * 2, c, c)
        self.out_conv = nn.Conv2d(c, in_channels, 3, padding=1)

    def load_state_dict(self, state_dict, *args, **kwargs):
        if "res_ratio" in state_dict and not hasattr(self, "res_ratio"):
            self.res_ratio = nn.Parameter(torch.tensor(0.1, device=self.out_conv.weight.device))
        return super().load_state_dict(state_dict, *args, **kwargs)

    def engine(self, device=None, n_t=None):
        device = device if device is not None else self.out_conv.weight.device
        eng = get_engine(device, self.precision)
        eng.pack_pix(self, max(self.max_timesteps, n_t or 0))
        return eng

    def forward(self, x, t):
        """eps(x, t) (v4:99-135). x (B, 3, H, W); t (B,) or (B, 1), any numeric dtype (the reference casts to float)."""
        _require_eval(self, "v4.SimpleUNet.forward")
        if x.shape[0] == 0:
            return x.new_empty(tuple(x.shape), dtype=torch.float32)
        return self.engine(x.device).pix_forward(x, t)


class DiffusionModel:
    """v4:140-183 (plain class, as in the reference)."""

    def __init__(self, model, n_steps=1000, beta_start=0.0001, beta_end=0.02, device="cpu"):
        self.model = model
        self.n_steps = n_steps
        self.device = device
        beta = torch.linspace(beta_start, beta_end, n_steps)      # v4:144-146; built on the CPU so that the tables are
        alpha = 1 - beta                                           # bit-equal to the reference's CPU construction
        alpha_bar = torch.cumprod(alpha, dim=0)
        self._host_schedule = (beta, alpha, alpha_bar)
        self.beta, self.alpha, self.alpha_bar = beta.to(device), alpha.to(device), alpha_bar.to(device)

    def _engine(self, device):
        eng = self.model.engine(device, n_t=self.n_steps)
        eng.set_schedule(*self._host_schedule)
        return eng

    def q_sample(self, x0, t, noise=None):
        """v4:148-153."""
        if noise is None:
            noise = torch.randn_like(x0)
        alpha_bar_t = self.alpha_bar.to(x0.device)[t].view(-1, 1, 1, 1)
        return torch.sqrt(alpha_bar_t) * x0 + torch.sqrt(1 - alpha_bar_t) * noise

    def p_sample(self, xt, t, *, noise=None, seed=None, sample_offset=0):
        """v4:155-168; t is a python int."""
        _require_eval(self.model, "v4.DiffusionModel.p_sample")
        t = int(t)
        if not 0 <= t < self.n_steps:
            raise IndexError("timestep %d outside [0, %d)" % (t, self.n_steps))
        eng = self._engine(xt.device)
        x = xt.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        if noise is not None:
            noise = noise.reshape(1, *x.shape)
        eng.pix_sample(x, t, t, noise=noise, seed=_fresh_seed() if seed is None else int(seed), sample_offset=int(sample_offset),
                       use_graph=False)
        return x

    def sample(self, shape, *, seed=None, sample_offset=0, x_T=None, noise=None, use_graph=True):
        """v4:170-175: x_T ~ N(0, I) on self.device, then n_steps reverse steps as one CUDA-graph launch."""
        _require_eval(self.model, "v4.DiffusionModel.sample")
        B, C, H, W = (int(s) for s in shape)
        if B == 0:
            return torch.empty((0, C, H, W), device=self.device, dtype=torch.float32)
        eng = self._engine(self.device)
        seed = _fresh_seed() if seed is None else int(seed)
        if x_T is None:
            x = eng.randn(B, C * H * W, seed, int(sample_offset), self.n_steps).view(B, C, H, W)
        else:
            x = x_T.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        eng.pix_sample(x, self.n_steps - 1, 0, noise=noise, seed=seed, sample_offset=int(sample_offset), use_graph=use_graph)
        eng.check_device_flags()
        return x

    def loss(self, x0):
        """v4:177-183, evaluation only (training is out of scope)."""
        B = x0.size(0)
        t = torch.randint(0, self.n_steps, (B,), device=x0.device).long()
        noise = torch.randn_like(x0)
        return torch.nn.functional.mse_loss(self.model(self.q_sample(x0, t, noise), t), noise)

    def sample_with_intermediates(self, shape, capture_steps, *, seed=None):
        """v4:185-199: frames (H, W, 3 numpy, clamped to [0, 1]) after the steps listed in capture_steps."""
        _require_eval(self.model, "v4.DiffusionModel.sample_with_intermediates")
        eng = self._engine(self.device)
        B, C, H, W = (int(s) for s in shape)
        seed = _fresh_seed() if seed is None else int(seed)
        x = eng.randn(B, C * H * W, seed, 0, self.n_steps).view(B, C, H, W)
        stops = sorted({int(t) for t in capture_steps if 0 <= int(t) < self.n_steps}, reverse=True)
        frames, t_hi = [], self.n_steps - 1
        for t in stops + ([] if (stops and stops[-1] == 0) else [None]):
            t_lo = 0 if t is None else t
            if t_hi >= t_lo:
                eng.pix_sample(x, t_hi, t_lo, seed=seed, use_graph=False)
            if t is not None:
                frames.append(x.clamp(0, 1).squeeze().cpu().permute(1, 2, 0).numpy())
            t_hi = t_lo - 1
        return frames
