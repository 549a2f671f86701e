"""ctypes binding of libldm_b200.so (include/ldm_b200.h).  Nothing else in this package touches the
library, and nothing here falls back to another implementation: a missing or unloadable library,
or a missing GPU, raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libldm_b200.so")

MAX_STAGES = 8
PRECISION = {"fp32": 0, "bf16": 2}

_f = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p


class UnetWeights(ctypes.Structure):
    _fields_ = (
        [("latent_dim", ctypes.c_int32), ("time_dim", ctypes.c_int32), ("num_classes", ctypes.c_int32),
         ("n_stages", ctypes.c_int32), ("hidden", ctypes.c_int32 * (MAX_STAGES + 1)), ("n_t", ctypes.c_int32),
         ("sinusoid", _vp), ("residual_weight", _vp),
         ("time_lin1_w", _vp), ("time_lin1_b", _vp), ("time_lin2_w", _vp), ("time_lin2_b", _vp),
         ("class_embedding", _vp),
         ("class_lin1_w", _vp), ("class_lin1_b", _vp), ("class_lin2_w", _vp), ("class_lin2_b", _vp),
         ("latent_proj_w", _vp), ("latent_proj_b", _vp)]
        + [(n, _vp * MAX_STAGES) for n in (
            "time_proj_w", "time_proj_b", "attn_in_proj_w", "attn_in_proj_b", "attn_out_w", "attn_out_b",
            "block_lin_w", "block_lin_b", "block_ln_w", "block_ln_b", "stage_ln_w", "stage_ln_b", "down_w", "down_b")]
        + [(n, _vp) for n in ("final_time_w", "final_time_b", "final_class_w", "final_class_b",
                              "final_norm_w", "final_norm_b", "final_w", "final_b")]
    )


class Unet3Weights(ctypes.Structure):
    """ldm_unet3_weights (v3 multi-conditional denoiser)."""
    _fields_ = (
        [("latent_dim", ctypes.c_int32), ("time_dim", ctypes.c_int32), ("num_classes", ctypes.c_int32),
         ("num_colors", ctypes.c_int32), ("n_stages", ctypes.c_int32), ("hidden", ctypes.c_int32 * (MAX_STAGES + 1)),
         ("n_t", ctypes.c_int32), ("sinusoid", _vp),
         ("time_lin1_w", _vp), ("time_lin1_b", _vp), ("time_lin2_w", _vp), ("time_lin2_b", _vp),
         ("flower_emb", _vp), ("color_emb", _vp), ("cond_fc_w", _vp), ("cond_fc_b", _vp),
         ("latent_proj_w", _vp), ("latent_proj_b", _vp)]
        + [(n, _vp * MAX_STAGES) for n in (
            "time_proj_w", "time_proj_b", "cond_proj_w", "cond_proj_b", "attn_in_proj_w", "attn_in_proj_b", "attn_out_w",
            "attn_out_b", "block_lin_w", "block_lin_b", "block_ln_w", "block_ln_b", "stage_ln_w", "stage_ln_b", "down_w",
            "down_b")]
        + [(n, _vp) for n in ("final_time_w", "final_time_b", "final_class_w", "final_class_b",
                              "final_norm_w", "final_norm_b", "final_w", "final_b")]
    )


class ResBlockWeights(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("conv1_w", "conv1_b", "ln1_w", "ln1_b", "conv2_w", "conv2_b", "ln2_w", "ln2_b",
                                   "ca_w0", "ca_w2", "sa_w")]


class DecoderWeights(ctypes.Structure):
    _fields_ = (
        [("latent_dim", ctypes.c_int32)]
        + [(n, _vp) for n in ("fc0_w", "fc0_b", "fc1_w", "fc1_b", "fc3_w", "fc3_b", "fc4_w", "fc4_b")]
        + [("res", ResBlockWeights * 3)]
        + [(n, _vp * 3) for n in ("up_w", "up_b", "up_gn_w", "up_gn_b")]
        + [(n, _vp) for n in ("fin0_w", "fin0_b", "fin_gn_w", "fin_gn_b", "fin3_w", "fin3_b")]
    )


class PixConv(ctypes.Structure):
    _fields_ = [("w", _vp), ("b", _vp)]


class PixWeights(ctypes.Structure):
    """ldm_pix_weights (v4 / v5 pixel-space SimpleUNet)."""
    _fields_ = (
        [("in_channels", ctypes.c_int32), ("base_channels", ctypes.c_int32), ("time_emb_dim", ctypes.c_int32),
         ("n_t", ctypes.c_int32), ("res_ratio", _vp),
         ("time_embed0_w", _vp), ("time_embed0_b", _vp), ("time_embed2_w", _vp), ("time_embed2_b", _vp),
         ("time_fc_w", _vp * 3), ("time_fc_b", _vp * 3),
         ("conv1", PixConv * 2), ("down1", PixConv), ("conv2", PixConv * 2), ("down2", PixConv), ("conv3", PixConv * 2),
         ("bottleneck", PixConv * 2), ("up1", PixConv), ("conv4", PixConv * 2), ("up2", PixConv), ("conv5", PixConv * 2),
         ("out_conv", PixConv)]
    )


class UBlockResWeights(ctypes.Structure):
    _fields_ = ([("in_channels", ctypes.c_int32), ("out_channels", ctypes.c_int32), ("d_time", ctypes.c_int32)]
                + [(n, _vp) for n in ("norm1_w", "norm1_b", "conv1_w", "conv1_b", "time_w", "time_b", "class_w", "class_b",
                                      "norm2_w", "norm2_b", "conv2_w", "conv2_b", "res_w", "res_b")])


class UBlockAttnWeights(ctypes.Structure):
    _fields_ = ([("channels", ctypes.c_int32), ("num_heads", ctypes.c_int32)]
                + [(n, _vp) for n in ("norm_w", "norm_b", "qkv_w", "qkv_b", "proj_w", "proj_b")])


# name -> (restype, argtypes); exactly the prototypes of include/ldm_b200.h
PROTOTYPES = {
    "ldm_version": (ctypes.c_int, []),
    "ldm_last_error": (ctypes.c_char_p, []),
    "ldm_ctx_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int]),
    "ldm_ctx_destroy": (ctypes.c_int, [_vp]),
    "ldm_set_schedule": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int]),
    "ldm_unet_pack": (ctypes.c_int, [_vp, ctypes.POINTER(UnetWeights), _vp]),
    "ldm_unet_set_classes": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "ldm_unet3_pack": (ctypes.c_int, [_vp, ctypes.POINTER(Unet3Weights), _vp]),
    "ldm_unet3_set_conditions": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp]),
    "ldm_unet_forward": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp]),
    "ldm_ddpm_step": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_randn": (ctypes.c_int, [_vp, _vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_sample": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_uint64, ctypes.c_uint64,
                                  ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_decoder_pack": (ctypes.c_int, [_vp, ctypes.POINTER(DecoderWeights), _vp]),
    "ldm_decode": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp]),
    "ldm_generate_host": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _vp, _vp, _vp]),
    "ldm_generate3_host": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _vp, _vp, _vp]),
    "ldm_pix_pack": (ctypes.c_int, [_vp, ctypes.POINTER(PixWeights), _vp]),
    "ldm_pix_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_pix_sample": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_ublock_res_pack": (ctypes.c_int, [_vp, ctypes.POINTER(UBlockResWeights), ctypes.POINTER(ctypes.c_int), _vp]),
    "ldm_ublock_res_forward": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_ublock_attn_pack": (ctypes.c_int, [_vp, ctypes.POINTER(UBlockAttnWeights), ctypes.POINTER(ctypes.c_int), _vp]),
    "ldm_ublock_free": (ctypes.c_int, [_vp, ctypes.c_int]),
    "ldm_ublock_attn_forward": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "ldm_kernel_launch_count": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "ldm_get_info": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]),
    "ldm_debug_chain_trace": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int]),
    "ldm_debug_ktrace": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_char_p, ctypes.c_int, _vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
}

_lib = None


class LdmError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise LdmError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no fallback implementation" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)      # AttributeError if the ABI and this table ever drift apart
            fn.restype, fn.argtypes = res, args
        if l.ldm_version() != 1:
            raise LdmError("ABI version mismatch: library %d, binding 1" % l.ldm_version())
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ldm_last_error().decode("utf-8", "replace")
        raise LdmError("%s failed (code %d): %s" % (what or "ldm call", rc, msg))
