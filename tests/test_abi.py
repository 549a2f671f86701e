"""The C-ABI library: it loads on a CPU-only host and exports every symbol include/ldm_b200.h declares
(no compute calls here).  Also: context creation without a GPU fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ldm_b200.h")


@pytest.fixture(scope="module")
def built():
    entry.build()
    import ldm_b200
    return ldm_b200.LIB_PATH


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"LDM_API\s+[\w\s\*]+?\b(ldm_\w+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = _declared()
    for must in ("ldm_ctx_create", "ldm_ctx_destroy", "ldm_set_schedule", "ldm_unet_pack", "ldm_unet_set_classes",
                 "ldm_unet_forward", "ldm_ddpm_step", "ldm_randn", "ldm_sample", "ldm_decoder_pack", "ldm_decode",
                 "ldm_generate_host", "ldm_version", "ldm_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (ldm_\w+)", out)))
    assert exported == _declared()          # nothing undeclared leaks out either


def test_binding_table_matches_header(built):
    from importlib import import_module
    _lib = import_module("ldm_b200._lib")
    assert sorted(_lib.PROTOTYPES) == _declared()
    lib = _lib.lib()
    assert lib.ldm_version() == 1
    # struct sizes are what the C side expects: 4+9+1 ints (padded) then pointers
    assert ctypes.sizeof(_lib.ResBlockWeights) == 11 * 8
    assert ctypes.sizeof(_lib.UnetWeights) % 8 == 0 and ctypes.sizeof(_lib.DecoderWeights) % 8 == 0


def test_is_sm100a_only_and_uses_tcgen05_and_tma(built):
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    assert "UTCHMMA" in sass or "UTCMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass                           # TMA tensor loads
    assert "LDTM" in sass                              # tcgen05.ld


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(built):
    from importlib import import_module
    _lib = import_module("ldm_b200._lib")
    ctx = ctypes.c_void_p()
    rc = _lib.lib().ldm_ctx_create(ctypes.byref(ctx), 0, 0)
    assert rc != 0 and not ctx
    assert b"no CUDA device" in _lib.lib().ldm_last_error()
    with pytest.raises(_lib.LdmError):
        _lib.check(rc, "ldm_ctx_create")
    rc = _lib.lib().ldm_ctx_create(ctypes.byref(ctx), 0, 7)
    assert rc < 0 and b"precision" in _lib.lib().ldm_last_error()
