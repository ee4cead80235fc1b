"""CPU: libq100.so loads, exports every symbol include/q100.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import q100
from q100 import _lib
from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "q100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(q100_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared_symbols()
    assert len(names) >= 10
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/q100.h but not exported by libq100.so"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_abi_version():
    assert _lib.lib.q100_abi_version() == 2


def test_packed_shape():
    assert q100.packed_shape(4096, 4096, 4) == (32, 64, 1024)
    assert q100.packed_shape(4096, 4096, 2) == (32, 64, 512)
    assert q100.packed_shape(4096, 4096, 8) == (32, 64, 2048)
    assert q100.packed_shape(4096, 4096, 3) == (32, 64, 768)
    assert q100.packed_shape(130, 128, 4) == (2, 2, 1024)          # rows padded to 128
    with pytest.raises(ValueError):
        q100.packed_shape(128, 96, 4)
    with pytest.raises(ValueError):
        q100.packed_shape(128, 4096, 5)


def test_workspace_bytes_is_pure():
    assert _lib.lib.q100_gemm_workspace_bytes(16, 14336, 4096, 4, 0) > 0
    assert _lib.lib.q100_gemm_workspace_bytes(0, 14336, 4096, 4, 0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    assert _lib.lib.q100_device_ok(0) == _lib.ENODEVICE
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.addressof(buf)
    rc = _lib.lib.q100_quantize_groupwise(p, 0, 4, 128, 128, 4, 0, 0, None, None, p, None, p, p, None, 0, None)
    assert rc == _lib.ENODEVICE and "no CPU path" in _lib.last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q100.quantize_groupwise(torch.randn(4, 128), 128, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q100.fakequant_rows(torch.randn(4, 128), 8)


def test_argument_validation():
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.addressof(buf)
    L = _lib.lib
    assert L.q100_quantize_groupwise(None, 0, 4, 128, 128, 4, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.EINVAL
    assert L.q100_quantize_groupwise(p, 7, 4, 128, 128, 4, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.EINVAL
    assert L.q100_quantize_groupwise(p, 0, 4, 100, 128, 4, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.EINVAL
    assert L.q100_quantize_groupwise(p, 0, 4, 128, 128, 16, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.EINVAL
    assert L.q100_quantize_groupwise(p, 0, 4, 128, 128, 9, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.EUNSUPPORTED
    assert L.q100_quantize_groupwise(p, 0, 4, 128, 128, 5, 0, 0, None, None, None, p, p, p, None, 0, None) == _lib.EUNSUPPORTED
    # empty tensors are a no-op, like the reference on an empty weight
    assert L.q100_quantize_groupwise(p, 0, 0, 128, 128, 4, 0, 0, None, None, p, None, p, p, None, 0, None) == _lib.OK
    assert L.q100_gemm_wna16(p, p, p, None, p, 4, 128, 128, 128, 4, 0, 2, None, 0, 0, None) == _lib.EUNSUPPORTED
    assert L.q100_gemm_wna16(p, p, p, None, p, 4, 128, 128, 128, 5, 0, 0, None, 0, 0, None) == _lib.EUNSUPPORTED
    assert L.q100_gemm_wna16(None, p, p, None, p, 4, 128, 128, 128, 4, 0, 0, None, 0, 0, None) == _lib.EINVAL
    assert L.q100_gemm_wna16(p, p, p, None, p, 0, 128, 128, 128, 4, 0, 0, None, 0, 0, None) == _lib.OK
