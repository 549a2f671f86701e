"""Drop-in mirror of the v3 multi-conditional denoiser (v3/model_train_test.py:739-893): same class names,
constructor arguments, method signatures and state_dict keys as the reference's v3 script; the arithmetic runs in
libldm_b200.so (ldm_unet3_pack / ldm_unet3_set_conditions / ldm_unet_forward / ldm_sample).

    from ldm_b200 import v3
    unet = v3.ConditionalUNet(num_classes=102, num_colors=10).to("cuda").eval()
    unet.load_state_dict(torch.load("conditional_diffusion_final.pt"))        # a v3 checkpoint, unchanged
    diffusion = v3.ConditionalDenoiseDiffusion(unet, n_steps=1000, device="cuda")
    z = diffusion.sample((B, 256), "cuda", flower_label, color_label)

v3's nn.MultiheadAttention receives h_norm.unsqueeze(1) (v3:832), i.e. (L = B, N = 1, E): the B samples of a call
attend to each other.  A call here is one reference call: results depend on the batch composition exactly as the
reference's do, so multi-GPU use replicates calls (each GPU = one reference call) instead of sharding one."""
import torch
from torch import nn

from .engine import get_engine
from .modules import Swish, TimeEmbedding, _fresh_seed, _require_eval, euclidean_distance_loss


class MultiConditionEmbedding(nn.Module):
    """v3:739-749."""

    def __init__(self, num_flower_types=102, num_colors=10, n_channels=256):
        super().__init__()
        self.flower_emb = nn.Embedding(num_flower_types, n_channels)
        self.color_emb = nn.Embedding(num_colors, n_channels)
        self.fc = nn.Linear(n_channels * 2, n_channels)


class ConditionalUNet(nn.Module):
    """v3:769-853.  Parameter tree identical to the reference (91 tensors, including residual_weight and the
    [-1] projections / attention layer that forward never uses)."""

    def __init__(self, latent_dim=256, hidden_dims=[256, 512, 1024, 512, 256], time_emb_dim=256, num_classes=102,
                 num_colors=10, dropout_rate=0.3, *, precision=None, max_timesteps=1000):
        super().__init__()
        self.latent_dim, self.time_emb_dim = latent_dim, time_emb_dim
        self.num_classes, self.num_colors = num_classes, num_colors
        self.hidden_dims = list(hidden_dims)
        self.precision = precision
        self.max_timesteps = max_timesteps
        self.time_emb = TimeEmbedding(n_channels=time_emb_dim)
        self.multi_cond_emb = MultiConditionEmbedding(num_flower_types=num_classes, num_colors=num_colors, n_channels=time_emb_dim)
        self.latent_proj = nn.Linear(latent_dim, hidden_dims[0])
        self.time_projections = nn.ModuleList(nn.Linear(time_emb_dim, d) for d in hidden_dims)
        self.cond_projections = nn.ModuleList(nn.Linear(time_emb_dim, d) for d in hidden_dims)
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
        eng.pack_unet3(self, max(self.max_timesteps, n_t or 0))
        return eng

    def forward(self, x, t, flower_label, color_label):
        """eps_theta(x_t, t, flower, color) (v3:804-853). x (B, latent); t int64 (1,) or (B,); labels int64 (B,)."""
        _require_eval(self, "v3.ConditionalUNet.forward")
        if x.shape[0] == 0:
            return x.new_empty((0, self.latent_dim), dtype=torch.float32)
        eng = self.engine(x.device)
        out = eng.unet3_forward(x, t, flower_label, color_label)
        eng.check_device_flags(self.num_classes)
        return out


class ConditionalDenoiseDiffusion:
    """v3:860-893."""

    def __init__(self, eps_model, n_steps=1000, device=None):
        self.eps_model = eps_model
        self.device = device
        beta = torch.linspace(0.0001, 0.02, n_steps)          # v3:865-867, built on the CPU (see modules.py)
        alpha = 1 - beta
        alpha_bar = torch.cumprod(alpha, dim=0)
        self._host_schedule = (beta, alpha, alpha_bar)
        self.beta, self.alpha, self.alpha_bar = beta.to(device), alpha.to(device), alpha_bar.to(device)
        self.n_steps = n_steps

    def _engine(self, device):
        eng = self.eps_model.engine(device, n_t=self.n_steps)
        eng.set_schedule(*self._host_schedule)
        return eng

    def q_sample(self, x0, t, eps=None):
        """v3:869-873."""
        if eps is None:
            eps = torch.randn_like(x0)
        alpha_bar_t = self.alpha_bar.to(x0.device)[t].reshape(-1, 1)
        return torch.sqrt(alpha_bar_t) * x0 + torch.sqrt(1 - alpha_bar_t) * eps

    def p_sample(self, xt, t, flower_label, color_label, *, noise=None, seed=None, sample_offset=0):
        """v3:874-887.  `t`: python int or int64 tensor of shape (1,)."""
        _require_eval(self.eps_model, "v3.ConditionalDenoiseDiffusion.p_sample")
        ti = int(t.reshape(-1)[0].item()) if isinstance(t, torch.Tensor) else int(t)
        if not 0 <= ti < self.n_steps:
            raise IndexError("timestep %d outside [0, %d)" % (ti, self.n_steps))
        eng = self._engine(xt.device)
        x = xt.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        if noise is not None:
            noise = noise.reshape(1, *x.shape)
        eng.sample3(x, ti, ti, flower_label, color_label, noise=noise, seed=_fresh_seed() if seed is None else int(seed),
                    sample_offset=int(sample_offset), use_graph=False)
        eng.check_device_flags(self.eps_model.num_classes)      # out-of-range labels raise here, not in a later unrelated call
        return x

    def sample(self, shape, device, flower_label, color_label, *, seed=None, sample_offset=0, x_T=None, noise=None,
               use_graph=True):
        """v3:888-892: x_T ~ N(0, I), then n_steps reverse steps as one CUDA-graph launch."""
        _require_eval(self.eps_model, "v3.ConditionalDenoiseDiffusion.sample")
        B, D = int(shape[0]), int(shape[1])
        if D != self.eps_model.latent_dim:
            raise ValueError("shape[1] must be latent_dim=%d" % self.eps_model.latent_dim)
        if B == 0:
            return torch.empty((0, D), device=device, dtype=torch.float32)
        eng = self._engine(device)
        seed = _fresh_seed() if seed is None else int(seed)
        if x_T is None:
            x = eng.randn(B, D, seed, int(sample_offset), self.n_steps)
        else:
            x = x_T.detach().to(device=eng.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        eng.sample3(x, self.n_steps - 1, 0, flower_label, color_label, noise=noise, seed=seed, sample_offset=int(sample_offset),
                    use_graph=use_graph)
        eng.check_device_flags(self.eps_model.num_classes)
        return x

    def loss(self, x0, flower_label, color_label):
        """v3:894-900, evaluation only."""
        t = torch.randint(0, self.n_steps, (x0.shape[0],), device=x0.device, dtype=torch.long)
        eps = torch.randn_like(x0)
        return euclidean_distance_loss(eps, self.eps_model(self.q_sample(x0, t, eps), t, flower_label, color_label))
