"""Host-side mirror of the reference's PyTorch module API for the sampling path.

Same class names, constructor arguments, method signatures and state_dict keys as
/root/reference/v2/model_train_test.py (cited per class), so reference checkpoints load with
strict=True and the reference's own callers (generate_class_samples v2:856-869,
visualize_denoising_steps v2:657-692, create_diffusion_animation v2:884-936) run unchanged.  The
modules hold parameters only: forward / p_sample / sample / decode hand raw device pointers to
libldm_b200.so through `engine.Engine`.  They are inference-only (eval mode); training is out of
scope and raises instead of silently running something else.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import get_engine


def _require_eval(m, what):
    if m.training:
        raise RuntimeError("%s: the B200 path implements eval-mode inference only (call .eval()); "
                           "training (dropout, autograd) is out of scope" % what)


def _fresh_seed():
    """A 62-bit seed drawn from torch's default CPU generator, so torch.manual_seed() controls the
    in-kernel Philox stream the way it controls torch.randn in the reference (v2:589,595)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class Swish(nn.Module):
    """v2:48-50. Parameter-free; present because it occupies Sequential slots in the state_dict numbering."""

    def forward(self, x):
        return x * torch.sigmoid(x)


# ------------------------------------------------------------------------------------------------
# denoiser
# ------------------------------------------------------------------------------------------------
class TimeEmbedding(nn.Module):
    """v2:401-418: sinusoid -> Linear -> Swish -> Linear."""

    def __init__(self, n_channels=256):
        super().__init__()
        self.n_channels = n_channels
        self.lin1 = nn.Linear(n_channels, 2 * n_channels)
        self.act = Swish()
        self.lin2 = nn.Linear(2 * n_channels, n_channels)

    def sinusoid_table(self, n_t):
        """Rows t = 0..n_t-1 of the sinusoidal features (v2:410-417), computed on the CPU with the
        reference's own torch expression so the table is bit-identical to what its forward builds."""
        t = torch.arange(n_t, dtype=torch.int64)
        half_dim = self.n_channels // 2
        emb = math.log(10000) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim) * -emb)
        emb = t[:, None] * emb[None, :]
        emb = torch.cat((emb.sin(), emb.cos()), dim=1)
        if emb.shape[1] < self.n_channels:
            emb = torch.cat([emb, torch.zeros(emb.shape[0], self.n_channels - emb.shape[1])], dim=1)
        return emb


class ClassEmbedding(nn.Module):
    """v2:421-431: Embedding -> Linear -> Swish -> Linear."""

    def __init__(self, num_classes=102, n_channels=256):
        super().__init__()
        self.embedding = nn.Embedding(num_classes, n_channels)
        self.lin1 = nn.Linear(n_channels, n_channels)
        self.act = Swish()
        self.lin2 = nn.Linear(n_channels, n_channels)


class ConditionalUNet(nn.Module):
    """v2:501-561.  Parameter tree identical to the reference (82 tensors, including the
    time_projections[-1] / attention_layers[-1] that forward never uses)."""

    def __init__(self, latent_dim=256, hidden_dims=[256, 512, 1024, 512, 256], time_emb_dim=256, num_classes=102,
                 dropout_rate=0.3, *, precision=None, max_timesteps=1000):
        super().__init__()
        self.latent_dim, self.time_emb_dim, self.num_classes = latent_dim, time_emb_dim, num_classes
        self.hidden_dims = list(hidden_dims)
        self.precision = precision            # None -> engine.default_precision()
        self.max_timesteps = max_timesteps    # rows of the hoisted time-embedding table
        self.time_emb = TimeEmbedding(n_channels=time_emb_dim)
        self.class_emb = ClassEmbedding(num_classes=num_classes, n_channels=time_emb_dim)
        self.latent_proj = nn.Linear(latent_dim, hidden_dims[0])
        self.time_projections = nn.ModuleList(nn.Linear(time_emb_dim, d) for d in hidden_dims)
        self.attention_layers = nn.ModuleList(nn.MultiheadAttention(embed_dim=d, num_heads=8, dropout=dropout_rate)
                                              for d in hidden_dims)
        self.layers = nn.ModuleList()
        for d_in, d_out in zip(hidden_dims[:-1], hidden_dims[1:]):
            block = nn.Sequential(nn.Linear(d_in, d_in), nn.LayerNorm(d_in), nn.Dropout(dropout_rate), Swish())
            self.layers.append(nn.ModuleList([block, nn.LayerNorm(d_in), nn.Linear(d_in, d_out)]))
        self.final_time_proj = nn.Linear(time_emb_dim, hidden_dims[-1])
        self.final_class_proj = nn.Linear(time_emb_dim, hidden_dims[-1])
        self.final_norm = nn.LayerNorm(hidden_dims[-1])
        self.final = nn.Linear(hidden_dims[-1], latent_dim)
        self.residual_weight = nn.Parameter(torch.tensor(0.1))

    def engine(self, device=None, n_t=None):
        device = device if device is not None else self.residual_weight.device
        eng = get_engine(device, self.precision)
        eng.pack_unet(self, max(self.max_timesteps, n_t or 0))
        return eng

    def forward(self, x, t, c=None):
        """eps_theta(x_t, t, c) (v2:535-561). x (B, latent) fp32; t int64 (1,) or (B,); c int64 (B,) or None."""
        _require_eval(self, "ConditionalUNet.forward")
        if x.shape[0] == 0:                       # empty batch: nothing to launch (torch's own modules return empty tensors too)
            return x.new_empty((0, self.latent_dim), dtype=torch.float32)
        eng = self.engine(x.device)
        out = eng.unet_forward(x, t, c)
        eng.check_device_flags(self.num_classes)      # IndexError on a bad label / timestep, like the reference
        return out


class ConditionalDenoiseDiffusion:
    """v2:564-607 (a plain class, not an nn.Module).  Attributes beta / alpha / alpha_bar / n_steps /
    eps_model / device as in the reference."""

    def __init__(self, eps_model, n_steps=1000, device=None):
        self.eps_model = eps_model
        self.device = device
        # v2:569-571.  Built on the CPU and moved, like beta in the reference; alpha_bar is ALSO built on the
        # CPU (the reference runs cumprod on `device`; a CUDA scan differs from the CPU's sequential product in
        # the last bits, and the CPU run is the oracle).
        beta = torch.linspace(0.0001, 0.02, n_steps)
        alpha = 1 - beta
        alpha_bar = torch.cumprod(alpha, dim=0)
        self._host_schedule = (beta, alpha, alpha_bar)
        self.beta, self.alpha, self.alpha_bar = beta.to(device), alpha.to(device), alpha_bar.to(device)
        self.n_steps = n_steps

    # -- helpers -------------------------------------------------------------------------------
    def _engine(self, device):
        eng = self.eps_model.engine(device, n_t=self.n_steps)
        eng.set_schedule(*self._host_schedule)
        return eng

    @staticmethod
    def _t_int(t):
        if isinstance(t, torch.Tensor):
            if t.numel() != 1:
                raise ValueError("p_sample takes a scalar timestep (int or tensor of shape (1,)), as the reference's callers do")
            return int(t.reshape(-1)[0].item())
        return int(t)

    # -- reference API --------------------------------------------------------------------------
    def q_sample(self, x0, t, eps=None):
        """v2:574-578 (forward noising; used by create_diffusion_animation v2:933-934). Plain torch: not on the hot path."""
        if eps is None:
            eps = torch.randn_like(x0)
        alpha_bar_t = self.alpha_bar.to(x0.device)[t].reshape(-1, 1)
        return torch.sqrt(alpha_bar_t) * x0 + torch.sqrt(1 - alpha_bar_t) * eps

    def p_sample(self, xt, t, c=None, *, noise=None, seed=None, sample_offset=0):
        """One reverse step x_t -> x_{t-1} (v2:580-592). `t`: python int or int64 tensor of shape (1,).
        Keyword-only extensions: `noise` (B, latent) replaces the draw of v2:589; `seed` / `sample_offset`
        select the in-kernel Philox stream (default: a fresh seed from torch's generator)."""
        _require_eval(self.eps_model, "ConditionalDenoiseDiffusion.p_sample")
        ti = self._t_int(t)
        if not 0 <= ti < self.n_steps:
            raise IndexError("timestep %d outside [0, %d)" % (ti, self.n_steps))
        eng = self._engine(xt.device)
        x = xt.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        if noise is not None:
            noise = noise.reshape(1, *x.shape)
        if c is not None and c.numel() and (int(c.min()) < 0 or int(c.max()) >= self.eps_model.num_classes):
            raise IndexError("class label out of range [0, %d)" % self.eps_model.num_classes)     # what nn.Embedding raises
        eng.sample(x, ti, ti, c, noise=noise, seed=_fresh_seed() if seed is None else int(seed),
                   sample_offset=int(sample_offset), use_graph=False)
        return x

    def sample(self, shape, device, c=None, *, seed=None, sample_offset=0, x_T=None, noise=None, use_graph=True):
        """x_T ~ N(0, I), then n_steps reverse steps (v2:594-598); returns x_0 of `shape` on `device`.
        The whole loop is ONE CUDA-graph launch.  Keyword-only extensions: `seed` / `sample_offset` (global
        index of row 0: the draw of every sample is independent of how the batch is sharded), `x_T` and
        `noise` (n_steps, B, latent) to replay given draws."""
        _require_eval(self.eps_model, "ConditionalDenoiseDiffusion.sample")
        B, D = int(shape[0]), int(shape[1])
        if D != self.eps_model.latent_dim:
            raise ValueError("shape[1] must be latent_dim=%d" % self.eps_model.latent_dim)
        if B == 0:
            return torch.empty((0, D), device=device, dtype=torch.float32)
        eng = self._engine(device)
        seed = _fresh_seed() if seed is None else int(seed)
        if c is not None and c.numel() and (int(c.min()) < 0 or int(c.max()) >= self.eps_model.num_classes):
            raise IndexError("class label out of range [0, %d)" % self.eps_model.num_classes)
        if x_T is None:
            x = eng.randn(B, D, seed, int(sample_offset), self.n_steps)          # v2:595
        else:
            x = x_T.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        eng.sample(x, self.n_steps - 1, 0, c, noise=noise, seed=seed, sample_offset=int(sample_offset), use_graph=use_graph)
        return x

    def loss(self, x0, labels=None):
        """v2:600-607, evaluation only (no autograd graph is built: training is out of scope)."""
        batch_size = x0.shape[0]
        t = torch.randint(0, self.n_steps, (batch_size,), device=x0.device, dtype=torch.long)
        eps = torch.randn_like(x0)
        xt = self.q_sample(x0, t, eps)
        eps_theta = self.eps_model(xt, t, labels)
        return euclidean_distance_loss(eps, eps_theta)


def euclidean_distance_loss(x, y, reduction="mean"):
    """v2:293-302."""
    dist = torch.sqrt(((x - y) ** 2).view(x.size(0), -1).sum(dim=1) + 1e-8)
    if reduction == "mean":
        return dist.mean()
    if reduction == "sum":
        return dist.sum()
    return dist


# ------------------------------------------------------------------------------------------------
# autoencoder (decode is the hot path; everything else exists for state_dict / caller compatibility)
# ------------------------------------------------------------------------------------------------
class CALayer(nn.Module):
    """v2:53-67."""

    def __init__(self, channel, reduction=8):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_du = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1, padding=0, bias=False), Swish(),
                                     nn.Conv2d(channel // reduction, channel, 1, padding=0, bias=False), nn.Sigmoid())

    def forward(self, x):
        return x * self.conv_du(self.avg_pool(x))


class SpatialAttention(nn.Module):
    """v2:69-81."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        pooled = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
        return x * self.sigmoid(self.conv(pooled))


class LayerNorm2d(nn.Module):
    """v2:144-156: per-(n, c) statistics over H x W (an affine instance norm)."""

    def __init__(self, num_channels, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(dim=(2, 3), keepdim=True)
        var = x.var(dim=(2, 3), keepdim=True, unbiased=False)
        return (x - mean) / torch.sqrt(var + self.eps) * self.weight.view(1, -1, 1, 1) + self.bias.view(1, -1, 1, 1)


class ResidualBlock(nn.Module):
    """v2:159-178. The torch forward serves the Encoder only; the Decoder's blocks run in the library."""

    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1)
        self.ln1 = LayerNorm2d(channels)
        self.swish = Swish()
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1)
        self.ln2 = LayerNorm2d(channels)
        self.ca = CALayer(channels)
        self.sa = SpatialAttention()

    def forward(self, x):
        out = self.ln2(self.conv2(self.swish(self.ln1(self.conv1(x)))))
        return self.swish(self.sa(self.ca(out)) + x)


class Encoder(nn.Module):
    """v2:181-239.  NOT on the sampling path: kept so that a full autoencoder state_dict
    round-trips and so the reference's visualisation callers can encode; plain torch ops."""

    def __init__(self, in_channels=3, latent_dim=256):
        super().__init__()
        self.latent_dim = latent_dim
        self.initial_conv = nn.Sequential(nn.Conv2d(in_channels, 64, 3, padding=1), LayerNorm2d(64), Swish())
        self.skip_features = []
        for i, (ci, co) in enumerate(((64, 128), (128, 256), (256, 512)), start=1):
            setattr(self, "down%d" % i, nn.Sequential(nn.Conv2d(ci, co, 4, stride=2, padding=1), LayerNorm2d(co), Swish()))
            setattr(self, "res%d" % i, ResidualBlock(co))
        head = lambda: nn.Sequential(nn.Linear(512 * 8 * 8, 512), nn.LayerNorm(512), Swish(), nn.Linear(512, latent_dim))
        self.fc_mu = head()
        self.fc_logvar = head()

    def forward(self, x):
        self.skip_features = []
        x = self.initial_conv(x)
        self.skip_features.append(x)
        for i in (1, 2, 3):
            x = getattr(self, "res%d" % i)(getattr(self, "down%d" % i)(x))
            self.skip_features.append(x)
        flat = x.reshape(x.size(0), -1)
        return self.fc_mu(flat), self.fc_logvar(flat)


class Decoder(nn.Module):
    """v2:242-290: 256-d latent -> (B, 3, 64, 64) in (0, 1).  59 tensors, names as in the reference."""

    def __init__(self, latent_dim=256, out_channels=3, *, precision=None):
        super().__init__()
        if out_channels != 3:
            raise ValueError("the B200 decoder is built for out_channels=3")
        self.latent_dim = latent_dim
        self.precision = precision
        self.fc = nn.Sequential(nn.Linear(latent_dim, 512), nn.LayerNorm(512), Swish(),
                                nn.Linear(512, 512 * 8 * 8), nn.LayerNorm(512 * 8 * 8), Swish())
        for name, c, groups in (("3", 512, 32), ("2", 256, 16), ("1", 128, 8)):
            setattr(self, "res" + name, ResidualBlock(c))
            setattr(self, "up" + name, nn.Sequential(nn.ConvTranspose2d(c, c // 2, 4, stride=2, padding=1),
                                                     nn.GroupNorm(groups, c // 2), Swish()))
        self.final_conv = nn.Sequential(nn.Conv2d(64, 32, 3, padding=1), nn.GroupNorm(8, 32), Swish(),
                                        nn.Conv2d(32, out_channels, 3, padding=1), nn.Sigmoid())

    def forward(self, z, encoder_features=None):
        """v2:280-290; `encoder_features` is accepted and ignored, as in the reference."""
        _require_eval(self, "Decoder.forward")
        if z.numel() == 0:
            return z.new_empty((0, 3, 64, 64), dtype=torch.float32)
        eng = get_engine(z.device, self.precision)
        eng.pack_decoder(self)
        out = eng.decode(z.reshape(-1, self.latent_dim))
        return out


class SimpleAutoencoder(nn.Module):
    """v2:305-393: encoder + decoder + classifier head + class-centre buffers (132 state_dict entries)."""

    def __init__(self, in_channels=3, latent_dim=256, num_classes=102, *, precision=None):
        super().__init__()
        self.latent_dim = latent_dim
        self.encoder = Encoder(in_channels, latent_dim)
        self.decoder = Decoder(latent_dim, in_channels, precision=precision)
        self.classifier = nn.Sequential(nn.Linear(latent_dim, 512), nn.LayerNorm(512), Swish(), nn.Dropout(0.3),
                                        nn.Linear(512, 256), nn.LayerNorm(256), Swish(), nn.Dropout(0.2),
                                        nn.Linear(256, num_classes))
        self.register_buffer("class_centers", torch.zeros(num_classes, latent_dim))
        self.register_buffer("center_counts", torch.zeros(num_classes))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        """v2:326-337."""
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(m.weight, a=0.2)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def reparameterize(self, mu, logvar):
        """v2:339-343."""
        std = torch.exp(0.5 * torch.clamp(logvar, min=-2.0, max=10.0))
        return mu + torch.randn_like(std) * std

    def encode(self, x):
        """v2:345-348 (not on the sampling path)."""
        mu, logvar = self.encoder(x)
        return self.reparameterize(mu, logvar)

    def encode_with_params(self, x):
        """v2:350-353."""
        mu, logvar = self.encoder(x)
        return mu, torch.clamp(logvar, min=-2.0, max=10.0)

    def decode(self, z):
        """v2:355-357: the hot-path entry (B, latent) -> (B, 3, 64, 64)."""
        return self.decoder(z, getattr(self, "stored_encoder_features", None))

    def classify(self, z):
        """v2:359-360."""
        return self.classifier(z)

    def forward(self, x):
        """v2:362-366 (eval use only here)."""
        z = self.encode(x)
        return self.decode(z), z


class UNetAttentionBlock(nn.Module):
    """v2:434-459 (defined by the reference, never instantiated there).  forward runs in the library: GroupNorm(1, C),
    the 1x1 qkv / proj convolutions as tcgen05 GEMMs over pixels, the 4-head attention over the H W tokens as the
    tcgen05 attention kernel."""

    def __init__(self, channels, num_heads=4, *, precision=None):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads
        self.precision = precision
        self.norm = nn.GroupNorm(1, channels)
        self.qkv = nn.Conv2d(channels, channels * 3, 1)
        self.proj = nn.Conv2d(channels, channels, 1)

    def forward(self, x):
        _require_eval(self, "UNetAttentionBlock.forward")
        return get_engine(x.device, self.precision).ublock_attn_forward(self, x)


class UNetResidualBlock(nn.Module):
    """v2:462-486 (defined by the reference, never instantiated there).  forward(x, t, c=None): t and c are the
    (B, d_time) embedding vectors.  Runs in the library: LayerNorm2d + Swish passes, 3x3 implicit-GEMM convolutions on
    tcgen05 with the time / class term as the epilogue's per-sample add, identity or 1x1 residual."""

    def __init__(self, in_channels, out_channels, d_time=256, dropout_rate=0.2, *, precision=None):
        super().__init__()
        self.precision = precision
        self.norm1 = LayerNorm2d(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.time_emb = nn.Linear(d_time, out_channels)
        self.class_emb = nn.Linear(d_time, out_channels)
        self.act = Swish()
        self.dropout = nn.Dropout(dropout_rate)
        self.norm2 = LayerNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        self.residual = nn.Identity() if in_channels == out_channels else nn.Conv2d(in_channels, out_channels, 1)

    def forward(self, x, t, c=None):
        _require_eval(self, "UNetResidualBlock.forward")
        return get_engine(x.device, self.precision).ublock_res_forward(self, x, t, c)


class SwitchSequential(nn.Sequential):
    """v2:489-498."""

    def forward(self, x, t=None, c=None):
        for layer in self:
            if isinstance(layer, UNetResidualBlock):
                x = layer(x, t, c)
            else:
                x = layer(x)
        return x


def init_weights(m):
    """The denoiser initialisation of main() (v2:1346-1350), used for the benchmark's random-init weights."""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
        nn.init.kaiming_normal_(m.weight, a=0.2)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def load_autoencoder_checkpoint(autoencoder, obj, strict=False):
    """Accept both on-disk forms the reference writes: the wrapped
    {'autoencoder': sd, 'discriminator': sd} of v2:1179-1191 and the raw state_dict of v2:1326."""
    sd = obj["autoencoder"] if isinstance(obj, dict) and "autoencoder" in obj and isinstance(obj["autoencoder"], dict) else obj
    return autoencoder.load_state_dict(sd, strict=strict)


def generate_class_samples(autoencoder, diffusion, target_class, num_samples=5, *, save_path=None, class_names=None, seed=None,
                           sample_offset=0):
    """The reference's generate_class_samples (v2:856-882): returns (images (n, 3, 64, 64), latents (n, latent)) on the
    model's device and, with `save_path`, writes the row of samples as a PNG (io_utils.save_image_grid in place of the
    matplotlib figure of v2:870-881).  `target_class` is an index, or a name looked up in `class_names` (the reference
    reads its global list, v2:860-864)."""
    if isinstance(target_class, str):
        if class_names is None or target_class not in class_names:
            raise ValueError("Invalid class name: %s. Must be one of %s" % (target_class, class_names))
        target_class = class_names.index(target_class)
    device = next(autoencoder.parameters()).device
    autoencoder.eval()
    diffusion.eps_model.eval()
    class_tensor = torch.tensor([int(target_class)] * num_samples, device=device)
    with torch.no_grad():
        latents = diffusion.sample((num_samples, autoencoder.latent_dim), device, class_tensor, seed=seed,
                                   sample_offset=sample_offset)
        samples = autoencoder.decode(latents)
    if save_path:
        from .io_utils import save_image_grid
        save_image_grid(samples, save_path)
    return samples, latents
