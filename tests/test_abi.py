"""The C-ABI library builds, loads without a GPU and exports every symbol include/marie_b200.h declares; on a
machine without a CUDA device mb_init refuses to run (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from marie_icr_b200 import _lib
    return _lib.load_library()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "marie_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/marie_b200.h but not exported by libmarie_b200.so"


def test_version_and_no_cpu_fallback(lib):
    import torch
    assert b"sm_100a" in lib.mb_version()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    h = ctypes.c_void_p()
    rc = lib.mb_init(0, ctypes.byref(h))
    assert rc == -3 and not h.value          # MB_ERR_NO_DEVICE
    from marie_icr_b200._lib import Context, MarieB200Error
    with pytest.raises(MarieB200Error):
        Context(0)
