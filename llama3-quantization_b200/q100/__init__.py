"""q100: B200 (sm_100a) kernels for the group-wise quantized-linear hot path.

Importing this package loads ``libq100.so`` through ctypes and fails loudly if it is
missing.  See include/q100.h for the C ABI and DESIGN.md for the kernels.
"""
from . import _lib  # noqa: F401  (loads the shared library, raises ImportError if absent)
from ._lib import Q100Error, LIB_PATH  # noqa: F401
from .ops import (  # noqa: F401
    PackedWeight, QuantResult, dequantize, fakequant_rows, gemm_wna16, pack_weight,
    packed_shape, quantize_groupwise, selftest_div,
)

__all__ = ["PackedWeight", "QuantResult", "Q100Error", "dequantize", "fakequant_rows", "gemm_wna16",
           "pack_weight", "packed_shape", "quantize_groupwise", "selftest_div", "LIB_PATH"]
