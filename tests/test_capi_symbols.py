"""CPU test: the C-ABI library builds for sm_100a, loads, and exports every symbol include/reflax_c.h declares.
No compute call is made (there is no GPU here); rfx_create must fail loudly instead of falling back."""
import ctypes as C
import os
import re

from reflaxman_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "reflax_c.h")).read()
    return sorted(set(re.findall(r"RFX_API[^;]*?\b(rfx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(n for n, _, _ in capi.SYMBOLS)


def test_library_exports_every_declared_symbol(rfx_lib):
    for name in declared_symbols():
        assert hasattr(rfx_lib, name), name
    assert b"sm_100a" in rfx_lib.rfx_version()


def test_sass_is_sm100_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_create_fails_loudly_without_gpu(rfx_lib):
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = rfx_lib.rfx_create(C.byref(h), 0)
    assert rc < 0 and not h.value
    assert b"no CPU fallback" in rfx_lib.rfx_last_error(None)


def test_product_never_references_oracle():
    """the product path must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "reflaxman_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "rfxo_" not in txt and "librfx_oracle" not in txt, os.path.join(dp, f)
