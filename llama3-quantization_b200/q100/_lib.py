"""ctypes binding of libq100.so (C ABI declared in include/q100.h).

Loading is eager and loud: if the shared library has not been built
(``python tools/build_ext.py`` / ``__graft_entry__.build()``) importing this module
raises ``ImportError``.  There is no Python or CPU fallback for any kernel.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# Q100_LIB selects an experiment build (tools/build_ext.py, Q100_VARIANT); the default is the in-tree library
LIB_PATH = os.environ.get("Q100_LIB") or os.path.join(_HERE, "libq100.so")

OK, EINVAL, EUNSUPPORTED, ECUDA, ENODEVICE, EWORKSPACE = range(6)
F16, BF16, F32 = 0, 1, 2
FLAG_IEEE_DIV = 0x1
FLAG_SCALAR_MATH = 0x2
FLAG_K1_GENERIC = 0x4
FLAG_CUDA_SCALE = 0x8
GEMM_FORCE_SKINNY = 0x1
GEMM_FORCE_TC = 0x2
GEMM_EXACT_WEIGHTS = 0x4
GEMM_PDL = 0x8
GEMM_FORCE_MMA_SYNC = 0x10
GEMM_FORCE_STREAM_TC = 0x20
GEMM_FORCE_TC_SMEM = 0x40
GEMM_NO_IMMA = 0x80
GEMM_X_GATE_UP = 0x100
GEMM_NO_ROWS = 0x200

_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

PH_GEMM, PH_ATTN, PH_COMBINE = 0, 1, 2
X_PLAIN, X_GATE_UP, X_NORM, X_ATTN = 0, 1, 2, 3
EPI_STORE, EPI_ADD_SSQ = 0, 1


class Phase(C.Structure):
    """struct q100_phase of include/q100.h (one phase of a decode chain)."""
    _fields_ = [("type", _i),
                ("x", _vp), ("packed", _vp), ("meta", _vp), ("bias", _vp), ("y", _vp),
                ("N", _i64), ("K", _i64),
                ("x_mode", _i), ("norm_weight", _vp), ("ssq_in", _vp), ("ssq_in_parts", _i), ("eps", C.c_float),
                ("epi_mode", _i), ("residual", _vp), ("ssq_out", _vp),
                ("qkv", _vp), ("cos_cache", _vp), ("sin_cache", _vp), ("position", _vp), ("k_cache", _vp),
                ("v_cache", _vp), ("att_partials", _vp), ("att_out", _vp),
                ("batch", _i), ("heads", _i), ("kv_heads", _i), ("head_dim", _i),
                ("max_len", _i64), ("rope_len", _i64)]


# name -> (restype, argtypes); must list every symbol include/q100.h declares
SIGNATURES = {
    "q100_last_error": (C.c_char_p, []),
    "q100_abi_version": (_i, []),
    "q100_device_ok": (_i, [_i]),
    "q100_quantize_groupwise": (_i, [_vp, _i, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "q100_quantize_backward": (_i, [_vp, _vp, _i, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "q100_fakequant_rows": (_i, [_vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "q100_rows_minmax": (_i, [_vp, _i, _i64, _i64, _vp, _vp]),
    "q100_quantize_rows_stats": (_i, [_vp, _i, _i64, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "q100_packed_words": (_i64, [_i64, _i64, _i]),
    "q100_dequantize": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "q100_gemm_wna16": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "q100_gemm_wna8": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "q100_gemm_w8a8": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _vp]),
    "q100_code_rowsums": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "q100_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64, _i, _i]),
    "q100_gemm_wna16_block": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, C.c_float, _vp, _i, _vp]),
    "q100_gemm_block_fusable": (_i, [_i64, _i64, _i64, _i, _i]),
    "q100_chain_desc_bytes": (_sz, [_i]),
    "q100_chain_ssq_parts": (_i64, [_i64]),
    "q100_chain_attn_partial_floats": (_i64, [_i, _i, _i64]),
    "q100_chain_build": (_i, [C.POINTER(Phase), _i, _i, _i, _i, _i, _i, _vp, _sz, _vp, _sz, C.POINTER(_sz), _vp]),
    "q100_chain_run": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "q100_chain_error": (_i, [_vp, _vp]),
    "q100_add_rmsnorm": (_i, [_vp, _vp, _vp, C.c_float, _vp, _vp, _i64, _i64, _i, _vp]),
    "q100_rope_kv_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _vp, _i, _vp]),
    "q100_attn_decode_workspace_bytes": (_sz, [_i, _i, _i, _i64]),
    "q100_attn_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i64, _i64, _vp, _i, _vp]),
    "q100_rope_qkv_prefill": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _i64, _vp, _i, _vp]),
    "q100_silu_mul": (_i, [_vp, _vp, _i64, _i64, _i, _vp]),
    "q100_argmax_workspace_bytes": (_sz, [_i64]),
    "q100_argmax_rows": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _sz, _vp]),
    "q100_allreduce_oneshot": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    "q100_gemm_wna16_scatter": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "q100_reduce_bcast": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "q100_allreduce_multimem": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "q100_allreduce_add_rmsnorm": (_i, [_vp, _vp, _vp, C.c_float, _vp, _vp, _i64, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "q100_selftest_div": (_i, [_i, _vp, _vp]),
    "q100_debug_gemm_dense_tc": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i, _vp]),
}


class Q100Error(RuntimeError):
    """A libq100 entry point returned a non-zero status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libq100 error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python tools/build_ext.py or __graft_entry__.build()). q100 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError == ABI mismatch, let it propagate
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return lib.q100_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise Q100Error(rc, last_error())
