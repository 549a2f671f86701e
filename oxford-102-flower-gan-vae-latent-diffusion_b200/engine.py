"""Per-(device, precision) handle on an ldm_ctx: packs module weights into it and issues the calls.

PyTorch is used here for device memory and streams only; all arithmetic of the hot path happens inside
libldm_b200.so.  One denoiser and one decoder are resident per engine; packing is redone when a
different module, or a module whose parameters changed (tensor._version / data_ptr), is used."""
import ctypes
import itertools
import os
import threading
import weakref

import torch

from . import _lib
from ._lib import check, lib

_engines = {}
_lock = threading.Lock()
_default_precision = None


def default_precision():
    """'bf16' (tcgen05 tensor cores, eps within 2e-2) unless LDM_B200_PRECISION=fp32 or set_default_precision()."""
    if _default_precision is not None:
        return _default_precision
    p = os.environ.get("LDM_B200_PRECISION", "bf16").lower()
    if p not in _lib.PRECISION:
        raise ValueError("LDM_B200_PRECISION must be one of %s" % sorted(_lib.PRECISION))
    return p


def set_default_precision(p):
    global _default_precision
    if p is not None and p not in _lib.PRECISION:
        raise ValueError("precision must be one of %s" % sorted(_lib.PRECISION))
    _default_precision = p


def _dev_index(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the B200 sampling path runs on CUDA devices only (got %s); there is no CPU fallback" % device)
    return device.index if device.index is not None else torch.cuda.current_device()


def get_engine(device, precision=None):
    precision = precision or default_precision()
    key = (_dev_index(device), precision)
    with _lock:
        eng = _engines.get(key)
        if eng is None:
            eng = _engines[key] = Engine(key[0], precision)
    return eng


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t, device):
    """fp32 contiguous view/copy of a parameter on `device` (kept alive by the caller during packing)."""
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


_serials = itertools.count(1)
_serial_of = weakref.WeakKeyDictionary()      # module object -> serial; not an attribute, so copy.deepcopy / pickle do not carry it


def _module_serial(module):
    """A process-unique number per module OBJECT.  id(module) is not enough: CPython reuses the address of a collected
    module for the next one, and torch's caching allocator hands the same blocks to its parameters, so (id, versions,
    data_ptrs) of a NEW module can equal those of a dead one whose packed weights are still resident."""
    s = _serial_of.get(module)
    if s is None:
        s = _serial_of[module] = next(_serials)
    return s


def _state_key(module, extra=()):
    """What a pack is valid for: the module object, and the version counter and storage of every parameter.  In-place
    edits through autograd-visible ops (`p.add_()`, `p.copy_()`, optimizer steps, load_state_dict) bump `_version` and
    re-pack on the next call.  Edits made through `p.data` (`p.data.mul_()`, EMA updates written that way) and edits of
    BUFFERS do NOT change this key: call `Engine.invalidate(module)` (or `ldm_b200.invalidate(module)`) after them."""
    ps = list(module.parameters())
    return (_module_serial(module), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps)) + tuple(extra)


def invalidate(module=None):
    """Forget the packed copy of `module` (or of everything) in every engine: the next call re-packs from the module's
    current tensors.  Needed only after weight edits that bypass the version counter (see _state_key)."""
    with _lock:
        engines = list(_engines.values())
    for eng in engines:
        eng.invalidate(module)


class Engine:
    def __init__(self, device_index, precision):
        self.device = torch.device("cuda", device_index)
        self.precision = precision
        self.ctx = ctypes.c_void_p()
        check(lib().ldm_ctx_create(ctypes.byref(self.ctx), device_index, _lib.PRECISION[precision]), "ldm_ctx_create")
        self._unet_key = None
        self._dec_key = None
        self._sched_key = None
        self._cls_key = None
        self._pix_key = None
        self._ublocks = {}      # module serial -> (state key, handle)

    def __del__(self):
        try:
            if self.ctx:
                lib().ldm_ctx_destroy(self.ctx)
        except Exception:
            pass

    def invalidate(self, module=None):
        """Drop the packed state of `module` (None: of every module) so that its next use re-packs."""
        ser = None if module is None else _module_serial(module)
        for attr in ("_unet_key", "_dec_key", "_pix_key"):
            k = getattr(self, attr)
            if k is not None and (ser is None or k[0] == ser):
                setattr(self, attr, None)
                if attr == "_unet_key":
                    self._cls_key = None
        for s in [s for s in self._ublocks if ser is None or s == ser]:
            self._ublock_release(s)

    # ------------------------------------------------------------------ helpers
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def info(self, key):
        out = ctypes.c_double()
        check(lib().ldm_get_info(self.ctx, key.encode(), ctypes.byref(out)), "ldm_get_info")
        return out.value

    def launches(self):
        out = ctypes.c_uint64()
        check(lib().ldm_kernel_launch_count(self.ctx, ctypes.byref(out)), "ldm_kernel_launch_count")
        return out.value

    def ktrace_start(self):
        """Start recording a CUDA event after every kernel this context launches eagerly on the current stream."""
        check(lib().ldm_debug_ktrace(self.ctx, 1, self.stream(), None, 0, None, 0, None), "ldm_debug_ktrace")

    def ktrace_stop(self, max_kernels=8192):
        """-> [(kernel name, milliseconds)] in launch order since ktrace_start()."""
        names = ctypes.create_string_buffer(64 * max_kernels)
        ms = (ctypes.c_float * max_kernels)()
        n = ctypes.c_int()
        check(lib().ldm_debug_ktrace(self.ctx, 0, None, names, len(names), ms, max_kernels, ctypes.byref(n)), "ldm_debug_ktrace")
        return list(zip(names.value.decode().split("\n")[:n.value], [float(ms[i]) for i in range(n.value)]))

    def check_device_flags(self, num_classes=None):
        """Raise like the reference's nn.Embedding / tensor indexing would on a bad index."""
        f = int(self.info("device_flags"))
        if f & 1:
            raise IndexError("class label out of range [0, %s)" % (num_classes if num_classes is not None else "num_classes"))
        if f & 2:
            raise IndexError("timestep out of range of the time-embedding table")
        e = int(self.info("tc_error"))
        if e:
            raise _lib.LdmError("tensor-core kernel barrier timeout (code %d, block %d)" % (e >> 16, e & 0xFFFF))

    # ------------------------------------------------------------------ schedule
    def set_schedule(self, beta, alpha, alpha_bar):
        b, a, ab = (t.detach().to("cpu", torch.float32).contiguous() for t in (beta, alpha, alpha_bar))
        key = (b.numpy().tobytes(), a.numpy().tobytes(), ab.numpy().tobytes())
        if key == self._sched_key:
            return
        check(lib().ldm_set_schedule(self.ctx, _ptr(b), _ptr(a), _ptr(ab), b.numel()), "ldm_set_schedule")
        self._sched_key = key

    # ------------------------------------------------------------------ denoiser
    def pack_unet(self, m, n_t):
        key = _state_key(m, (n_t,))
        if key == self._unet_key:
            return
        dev, keep = self.device, []
        def P(t):
            v = _f32(t, dev); keep.append(v); return v.data_ptr()
        w = _lib.UnetWeights()
        hidden = list(m.hidden_dims)
        nst = len(hidden) - 1
        if nst > _lib.MAX_STAGES:
            raise ValueError("at most %d stages are supported" % _lib.MAX_STAGES)
        w.latent_dim, w.time_dim, w.num_classes, w.n_stages, w.n_t = m.latent_dim, m.time_emb_dim, m.num_classes, nst, n_t
        for i, d in enumerate(hidden):
            w.hidden[i] = d
        w.sinusoid = P(m.time_emb.sinusoid_table(n_t))
        w.residual_weight = P(m.residual_weight.reshape(1))
        w.time_lin1_w, w.time_lin1_b = P(m.time_emb.lin1.weight), P(m.time_emb.lin1.bias)
        w.time_lin2_w, w.time_lin2_b = P(m.time_emb.lin2.weight), P(m.time_emb.lin2.bias)
        w.class_embedding = P(m.class_emb.embedding.weight)
        w.class_lin1_w, w.class_lin1_b = P(m.class_emb.lin1.weight), P(m.class_emb.lin1.bias)
        w.class_lin2_w, w.class_lin2_b = P(m.class_emb.lin2.weight), P(m.class_emb.lin2.bias)
        w.latent_proj_w, w.latent_proj_b = P(m.latent_proj.weight), P(m.latent_proj.bias)
        for i in range(nst):
            w.time_proj_w[i], w.time_proj_b[i] = P(m.time_projections[i].weight), P(m.time_projections[i].bias)
            att = m.attention_layers[i]
            w.attn_in_proj_w[i], w.attn_in_proj_b[i] = P(att.in_proj_weight), P(att.in_proj_bias)
            w.attn_out_w[i], w.attn_out_b[i] = P(att.out_proj.weight), P(att.out_proj.bias)
            block, ln, down = m.layers[i]
            w.block_lin_w[i], w.block_lin_b[i] = P(block[0].weight), P(block[0].bias)
            w.block_ln_w[i], w.block_ln_b[i] = P(block[1].weight), P(block[1].bias)
            w.stage_ln_w[i], w.stage_ln_b[i] = P(ln.weight), P(ln.bias)
            w.down_w[i], w.down_b[i] = P(down.weight), P(down.bias)
        w.final_time_w, w.final_time_b = P(m.final_time_proj.weight), P(m.final_time_proj.bias)
        w.final_class_w, w.final_class_b = P(m.final_class_proj.weight), P(m.final_class_proj.bias)
        w.final_norm_w, w.final_norm_b = P(m.final_norm.weight), P(m.final_norm.bias)
        w.final_w, w.final_b = P(m.final.weight), P(m.final.bias)
        check(lib().ldm_unet_pack(self.ctx, ctypes.byref(w), self.stream()), "ldm_unet_pack")
        self._unet_key = key
        self._cls_key = None

    def pack_unet3(self, m, n_t):
        """v3 multi-conditional denoiser (ldm_unet3_pack)."""
        key = _state_key(m, (n_t, "v3"))
        if key == self._unet_key:
            return
        dev, keep = self.device, []
        def P(t):
            v = _f32(t, dev); keep.append(v); return v.data_ptr()
        w = _lib.Unet3Weights()
        hidden = list(m.hidden_dims)
        nst = len(hidden) - 1
        if nst > _lib.MAX_STAGES:
            raise ValueError("at most %d stages are supported" % _lib.MAX_STAGES)
        w.latent_dim, w.time_dim, w.num_classes, w.num_colors = m.latent_dim, m.time_emb_dim, m.num_classes, m.num_colors
        w.n_stages, w.n_t = nst, n_t
        for i, d in enumerate(hidden):
            w.hidden[i] = d
        w.sinusoid = P(m.time_emb.sinusoid_table(n_t))
        w.time_lin1_w, w.time_lin1_b = P(m.time_emb.lin1.weight), P(m.time_emb.lin1.bias)
        w.time_lin2_w, w.time_lin2_b = P(m.time_emb.lin2.weight), P(m.time_emb.lin2.bias)
        w.flower_emb, w.color_emb = P(m.multi_cond_emb.flower_emb.weight), P(m.multi_cond_emb.color_emb.weight)
        w.cond_fc_w, w.cond_fc_b = P(m.multi_cond_emb.fc.weight), P(m.multi_cond_emb.fc.bias)
        w.latent_proj_w, w.latent_proj_b = P(m.latent_proj.weight), P(m.latent_proj.bias)
        for i in range(nst):
            w.time_proj_w[i], w.time_proj_b[i] = P(m.time_projections[i].weight), P(m.time_projections[i].bias)
            w.cond_proj_w[i], w.cond_proj_b[i] = P(m.cond_projections[i].weight), P(m.cond_projections[i].bias)
            att = m.attention_layers[i]
            w.attn_in_proj_w[i], w.attn_in_proj_b[i] = P(att.in_proj_weight), P(att.in_proj_bias)
            w.attn_out_w[i], w.attn_out_b[i] = P(att.out_proj.weight), P(att.out_proj.bias)
            block, ln, down = m.layers[i]
            w.block_lin_w[i], w.block_lin_b[i] = P(block[0].weight), P(block[0].bias)
            w.block_ln_w[i], w.block_ln_b[i] = P(block[1].weight), P(block[1].bias)
            w.stage_ln_w[i], w.stage_ln_b[i] = P(ln.weight), P(ln.bias)
            w.down_w[i], w.down_b[i] = P(down.weight), P(down.bias)
        w.final_time_w, w.final_time_b = P(m.final_time_proj.weight), P(m.final_time_proj.bias)
        w.final_class_w, w.final_class_b = P(m.final_class_proj.weight), P(m.final_class_proj.bias)
        w.final_norm_w, w.final_norm_b = P(m.final_norm.weight), P(m.final_norm.bias)
        w.final_w, w.final_b = P(m.final.weight), P(m.final.bias)
        check(lib().ldm_unet3_pack(self.ctx, ctypes.byref(w), self.stream()), "ldm_unet3_pack")
        self._unet_key = key
        self._cls_key = None

    def set_conditions(self, flower, color, batch):
        f = flower.detach().to(device=self.device, dtype=torch.int64).contiguous()
        k = color.detach().to(device=self.device, dtype=torch.int64).contiguous()
        if f.dim() != 1 or f.numel() != batch or k.shape != f.shape:
            raise ValueError("flower / color labels must both have shape (%d,)" % batch)
        check(lib().ldm_unet3_set_conditions(self.ctx, _ptr(f), _ptr(k), batch, self.stream()), "ldm_unet3_set_conditions")
        return f, k

    def unet3_forward(self, x, t, flower, color):
        x = x.detach().to(device=self.device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        t = t.detach().to(device=self.device, dtype=torch.int64).contiguous().reshape(-1)
        if t.numel() not in (1, B):
            raise ValueError("t must have 1 or %d entries, got %d" % (B, t.numel()))
        keep = self.set_conditions(flower, color, B)
        out = torch.empty_like(x)
        check(lib().ldm_unet_forward(self.ctx, _ptr(x), _ptr(t), t.numel(), _ptr(out), B, self.stream()), "ldm_unet_forward")
        del keep
        return out

    def sample3(self, x, t_start, t_end, flower, color, noise=None, seed=0, sample_offset=0, use_graph=True):
        """In-place v3 chain on x (B, latent) for t = t_start .. t_end."""
        B = x.shape[0]
        keep = self.set_conditions(flower, color, B)
        if noise is not None:
            noise = noise.detach().to(device=self.device, dtype=torch.float32).contiguous()
            if noise.shape != (t_start - t_end + 1, B, x.shape[1]):
                raise ValueError("noise must have shape (%d, %d, %d)" % (t_start - t_end + 1, B, x.shape[1]))
        check(lib().ldm_sample(self.ctx, _ptr(x), int(t_start), int(t_end), _ptr(noise) if noise is not None else None,
                               seed, sample_offset, B, 1 if use_graph else 0, self.stream()), "ldm_sample")
        del keep
        return x

    def set_classes(self, c, batch):
        if c is None:
            check(lib().ldm_unet_set_classes(self.ctx, None, batch, self.stream()), "ldm_unet_set_classes")
            return None
        c = c.detach().to(device=self.device, dtype=torch.int64).contiguous()
        if c.dim() != 1 or c.numel() != batch:
            raise ValueError("class tensor must have shape (%d,), got %s" % (batch, tuple(c.shape)))
        check(lib().ldm_unet_set_classes(self.ctx, _ptr(c), batch, self.stream()), "ldm_unet_set_classes")
        return c

    def unet_forward(self, x, t, c):
        x = x.detach().to(device=self.device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        t = t.detach().to(device=self.device, dtype=torch.int64).contiguous().reshape(-1)
        if t.numel() not in (1, B):
            raise ValueError("t must have 1 or %d entries, got %d" % (B, t.numel()))
        keep = self.set_classes(c, B)
        out = torch.empty_like(x)
        check(lib().ldm_unet_forward(self.ctx, _ptr(x), _ptr(t), t.numel(), _ptr(out), B, self.stream()), "ldm_unet_forward")
        del keep
        return out

    def ddpm_step(self, x, eps, t, noise=None, seed=0, sample_offset=0):
        """In-place posterior update of x (B, D) for a given eps (v2:584-592)."""
        for name, v in (("x", x), ("eps", eps), ("noise", noise)):
            if v is None:
                continue
            if not (v.is_cuda and v.device == self.device and v.dtype == torch.float32 and v.is_contiguous() and v.dim() == 2
                    and tuple(v.shape) == tuple(x.shape)):
                raise ValueError("ddpm_step: %s must be a contiguous fp32 (B, D) tensor on %s with the shape of x" % (name, self.device))
        check(lib().ldm_ddpm_step(self.ctx, _ptr(x), _ptr(eps), int(t), _ptr(noise) if noise is not None else None,
                                  seed, sample_offset, x.shape[0], x.shape[1], self.stream()), "ldm_ddpm_step")
        return x

    def randn(self, batch, dim, seed, sample_offset, step):
        out = torch.empty(batch, dim, device=self.device, dtype=torch.float32)
        check(lib().ldm_randn(self.ctx, _ptr(out), seed, sample_offset, int(step), batch, dim, self.stream()), "ldm_randn")
        return out

    def sample(self, x, t_start, t_end, c, noise=None, seed=0, sample_offset=0, use_graph=True):
        """In-place chain on x (B, latent) for t = t_start .. t_end."""
        B = x.shape[0]
        keep = self.set_classes(c, B)
        if noise is not None:
            noise = noise.detach().to(device=self.device, dtype=torch.float32).contiguous()
            if noise.shape != (t_start - t_end + 1, B, x.shape[1]):
                raise ValueError("noise must have shape (%d, %d, %d)" % (t_start - t_end + 1, B, x.shape[1]))
        check(lib().ldm_sample(self.ctx, _ptr(x), int(t_start), int(t_end), _ptr(noise) if noise is not None else None,
                               seed, sample_offset, B, 1 if use_graph else 0, self.stream()), "ldm_sample")
        del keep
        return x

    # ------------------------------------------------------------------ v4 / v5 pixel-space denoiser
    def pack_pix(self, m, n_t):
        key = _state_key(m, (n_t,))
        if key == self._pix_key:
            return
        dev, keep = self.device, []
        def P(t):
            v = _f32(t, dev); keep.append(v); return v.data_ptr()
        def C(dst, conv):
            dst.w, dst.b = P(conv.weight), P(conv.bias)
        w = _lib.PixWeights()
        w.in_channels, w.base_channels, w.time_emb_dim, w.n_t = m.in_channels, m.base_channels, m.time_emb_dim, int(n_t)
        rr = getattr(m, "res_ratio", None)
        w.res_ratio = P(rr.reshape(1)) if rr is not None else None
        w.time_embed0_w, w.time_embed0_b = P(m.time_embed[0].weight), P(m.time_embed[0].bias)
        w.time_embed2_w, w.time_embed2_b = P(m.time_embed[2].weight), P(m.time_embed[2].bias)
        for i, fc in enumerate((m.time_fc1, m.time_fc2, m.time_fc3)):
            w.time_fc_w[i], w.time_fc_b[i] = P(fc.weight), P(fc.bias)
        for name in ("conv1", "conv2", "conv3", "bottleneck", "conv4", "conv5"):
            seq = getattr(m, name)
            C(getattr(w, name)[0], seq[0]); C(getattr(w, name)[1], seq[2])
        for name in ("down1", "down2", "up1", "up2", "out_conv"):
            C(getattr(w, name), getattr(m, name))
        check(lib().ldm_pix_pack(self.ctx, ctypes.byref(w), self.stream()), "ldm_pix_pack")
        self._pix_key = key

    def pix_forward(self, x, t):
        x = x.detach().to(device=self.device, dtype=torch.float32).contiguous()
        B, C, H, W = x.shape
        t = t.detach().to(device=self.device).reshape(-1).float().contiguous()      # v4:104: t.view(B, 1).float()
        if t.numel() != B:
            raise RuntimeError("shape '[%d, 1]' is invalid for input of size %d" % (B, t.numel()))   # what .view(B, 1) raises
        out = torch.empty_like(x)
        check(lib().ldm_pix_forward(self.ctx, _ptr(x), _ptr(t), _ptr(out), B, H, W, self.stream()), "ldm_pix_forward")
        return out

    def pix_sample(self, x, t_start, t_end, noise=None, seed=0, sample_offset=0, use_graph=True):
        """In-place chain on x (B, 3, H, W) for t = t_start .. t_end."""
        B, C, H, W = x.shape
        if noise is not None:
            noise = noise.detach().to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(noise.shape) != (t_start - t_end + 1, B, C, H, W):
                raise ValueError("noise must have shape %s" % ((t_start - t_end + 1, B, C, H, W),))
        check(lib().ldm_pix_sample(self.ctx, _ptr(x), int(t_start), int(t_end), _ptr(noise) if noise is not None else None,
                                   seed, sample_offset, B, H, W, 1 if use_graph else 0, self.stream()), "ldm_pix_sample")
        return x

    # ------------------------------------------------------------------ conv U-Net blocks (v2:434-486)
    def _ublock_release(self, serial):
        hit = self._ublocks.pop(serial, None)
        if hit is not None and self.ctx:
            lib().ldm_ublock_free(self.ctx, hit[1])       # best effort: the context frees whatever is left when it dies

    def _ublock(self, m, build):
        """Handle of the packed block of module `m`: re-packed (the old device copy released) when its weights changed,
        released when the module is collected."""
        key, ser = _state_key(m), _module_serial(m)
        hit = self._ublocks.get(ser)
        if hit is None or hit[0] != key:
            if hit is not None:
                self._ublock_release(ser)
            else:
                me = weakref.ref(self)
                weakref.finalize(m, lambda: me() is not None and me()._ublock_release(ser))
            hit = (key, build())
            self._ublocks[ser] = hit
        return hit[1]

    def ublock_res_forward(self, m, x, t, c=None):
        dev = self.device

        def build():
            keep = []
            def P(v):
                v = _f32(v, dev); keep.append(v); return v.data_ptr()
            w = _lib.UBlockResWeights()
            w.in_channels, w.out_channels, w.d_time = m.conv1.in_channels, m.conv1.out_channels, m.time_emb.in_features
            w.norm1_w, w.norm1_b, w.conv1_w, w.conv1_b = P(m.norm1.weight), P(m.norm1.bias), P(m.conv1.weight), P(m.conv1.bias)
            w.time_w, w.time_b, w.class_w, w.class_b = P(m.time_emb.weight), P(m.time_emb.bias), P(m.class_emb.weight), P(m.class_emb.bias)
            w.norm2_w, w.norm2_b, w.conv2_w, w.conv2_b = P(m.norm2.weight), P(m.norm2.bias), P(m.conv2.weight), P(m.conv2.bias)
            if isinstance(m.residual, torch.nn.Conv2d):
                w.res_w, w.res_b = P(m.residual.weight), P(m.residual.bias)
            h = ctypes.c_int()
            check(lib().ldm_ublock_res_pack(self.ctx, ctypes.byref(w), ctypes.byref(h), self.stream()), "ldm_ublock_res_pack")
            return h.value

        handle = self._ublock(m, build)
        x = _f32(x, dev)
        B, _, H, W = x.shape
        t = _f32(t, dev).reshape(B, -1)
        cc = _f32(c, dev).reshape(B, -1) if c is not None else None
        out = torch.empty(B, m.conv1.out_channels, H, W, device=dev, dtype=torch.float32)
        check(lib().ldm_ublock_res_forward(self.ctx, handle, _ptr(x), _ptr(t), _ptr(cc) if cc is not None else None, _ptr(out),
                                           B, H, W, self.stream()), "ldm_ublock_res_forward")
        return out

    def ublock_attn_forward(self, m, x):
        dev = self.device

        def build():
            keep = []
            def P(v):
                v = _f32(v, dev); keep.append(v); return v.data_ptr()
            w = _lib.UBlockAttnWeights()
            w.channels, w.num_heads = m.channels, m.num_heads
            w.norm_w, w.norm_b, w.qkv_w, w.qkv_b = P(m.norm.weight), P(m.norm.bias), P(m.qkv.weight), P(m.qkv.bias)
            w.proj_w, w.proj_b = P(m.proj.weight), P(m.proj.bias)
            h = ctypes.c_int()
            check(lib().ldm_ublock_attn_pack(self.ctx, ctypes.byref(w), ctypes.byref(h), self.stream()), "ldm_ublock_attn_pack")
            return h.value

        handle = self._ublock(m, build)
        x = _f32(x, dev)
        B, _, H, W = x.shape
        out = torch.empty_like(x)
        check(lib().ldm_ublock_attn_forward(self.ctx, handle, _ptr(x), _ptr(out), B, H, W, self.stream()), "ldm_ublock_attn_forward")
        return out

    # ------------------------------------------------------------------ decoder
    def pack_decoder(self, dec):
        key = _state_key(dec)
        if key == self._dec_key:
            return
        dev, keep = self.device, []
        def P(t):
            v = _f32(t, dev); keep.append(v); return v.data_ptr()
        w = _lib.DecoderWeights()
        w.latent_dim = dec.latent_dim
        w.fc0_w, w.fc0_b, w.fc1_w, w.fc1_b = P(dec.fc[0].weight), P(dec.fc[0].bias), P(dec.fc[1].weight), P(dec.fc[1].bias)
        w.fc3_w, w.fc3_b, w.fc4_w, w.fc4_b = P(dec.fc[3].weight), P(dec.fc[3].bias), P(dec.fc[4].weight), P(dec.fc[4].bias)
        for i, (res, up) in enumerate(((dec.res3, dec.up3), (dec.res2, dec.up2), (dec.res1, dec.up1))):
            r = w.res[i]
            r.conv1_w, r.conv1_b, r.ln1_w, r.ln1_b = P(res.conv1.weight), P(res.conv1.bias), P(res.ln1.weight), P(res.ln1.bias)
            r.conv2_w, r.conv2_b, r.ln2_w, r.ln2_b = P(res.conv2.weight), P(res.conv2.bias), P(res.ln2.weight), P(res.ln2.bias)
            r.ca_w0, r.ca_w2, r.sa_w = P(res.ca.conv_du[0].weight), P(res.ca.conv_du[2].weight), P(res.sa.conv.weight)
            w.up_w[i], w.up_b[i], w.up_gn_w[i], w.up_gn_b[i] = P(up[0].weight), P(up[0].bias), P(up[1].weight), P(up[1].bias)
        fc = dec.final_conv
        w.fin0_w, w.fin0_b, w.fin_gn_w, w.fin_gn_b = P(fc[0].weight), P(fc[0].bias), P(fc[1].weight), P(fc[1].bias)
        w.fin3_w, w.fin3_b = P(fc[3].weight), P(fc[3].bias)
        check(lib().ldm_decoder_pack(self.ctx, ctypes.byref(w), self.stream()), "ldm_decoder_pack")
        self._dec_key = key

    def decode(self, z):
        z = z.detach().to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty(z.shape[0], 3, 64, 64, device=self.device, dtype=torch.float32)
        check(lib().ldm_decode(self.ctx, _ptr(z), _ptr(out), z.shape[0], self.stream()), "ldm_decode")
        return out

    def generate_host(self, c_host, img_host, latents_host=None, seed=0, sample_offset=0):
        """Host buffers in, host buffers out (pinned recommended): labels -> images, one call."""
        B = img_host.shape[0]
        cp = _ptr(c_host) if c_host is not None else None
        lp = _ptr(latents_host) if latents_host is not None else None
        check(lib().ldm_generate_host(self.ctx, cp, B, seed, sample_offset, _ptr(img_host), lp, self.stream()),
              "ldm_generate_host")
        return img_host

    def generate3_host(self, flower_host, color_host, img_host, latents_host=None, seed=0, sample_offset=0):
        """v3: host (flower, color) labels in, host images out, one call (ldm_generate3_host)."""
        B = img_host.shape[0]
        lp = _ptr(latents_host) if latents_host is not None else None
        check(lib().ldm_generate3_host(self.ctx, _ptr(flower_host), _ptr(color_host), B, seed, sample_offset, _ptr(img_host), lp,
                                       self.stream()), "ldm_generate3_host")
        return img_host
