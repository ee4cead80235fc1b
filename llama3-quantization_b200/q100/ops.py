"""Torch-facing wrappers of the libq100 kernels.

PyTorch is plumbing only here: it owns device memory and the stream; every
computation below is one call through the C ABI (include/q100.h).  All functions
require CUDA tensors and raise otherwise -- there is no CPU path.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check

_DT = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"q100: unsupported dtype {t.dtype} (fp16, bf16, fp32 only)") from None


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("q100: CUDA tensor required -- the quantized hot path has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _on:
    """Device guard that only switches when the tensor is not on the current device (torch.cuda.device() costs
    several microseconds per call even when it has nothing to do; the eager decode path is host-bound)."""
    __slots__ = ("ctx",)

    def __init__(self, dev: torch.device):
        self.ctx = None if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


@functools.lru_cache(maxsize=4096)
def _gemm_workspace_bytes(M: int, N: int, K: int, n_bits: int, flags: int) -> int:
    return int(lib.q100_gemm_workspace_bytes(M, N, K, n_bits, flags))


TILE_ROWS, TILE_K = 128, 64


def packed_shape(rows: int, cols: int, n_bits: int):
    """[n_tiles, k_tiles, tile_words] of the tile-major packed layout (DESIGN.md "Packed layout")."""
    w = lib.q100_packed_words(rows, cols, n_bits)
    if w < 0:
        raise ValueError(f"q100: cannot pack [{rows}, {cols}] at {n_bits} bits (cols % 64 == 0, bits in 2/3/4/8)")
    nt, kt = (rows + TILE_ROWS - 1) // TILE_ROWS, cols // TILE_K
    return nt, kt, (int(w) // (nt * kt) if nt * kt else 0)


@dataclass
class QuantResult:
    """Outputs of one quantize pass (any subset)."""
    dequant: Optional[torch.Tensor] = None     # forward() value, same shape/dtype as the input
    packed: Optional[torch.Tensor] = None      # int32 [n_tiles, k_tiles, tile_words] (tile-major)
    scales: Optional[torch.Tensor] = None      # [n_groups, 1]  == quantizer.scale
    zeros: Optional[torch.Tensor] = None       # [n_groups, 1]  == quantizer.round_zero_point (None if disabled)
    meta: Optional[torch.Tensor] = None        # [cols/group, rows, 2] (scale, zero) side band for the GEMMs


def quantize_groupwise(x: torch.Tensor, group: int, n_bits: int, symmetric: bool = False,
                       disable_zero_point: bool = False, lwc_up: Optional[torch.Tensor] = None,
                       lwc_low: Optional[torch.Tensor] = None, want_dequant: bool = True,
                       want_packed: bool = False, want_meta: bool = False, flags: int = 0) -> QuantResult:
    """One fused pass of UniformAffineQuantizer.forward over a 2-D tensor [rows, cols].

    ``group == 0`` quantises one group per row.  ``lwc_up`` / ``lwc_low`` are the
    post-sigmoid LWC multipliers (quant/quantizer.py:143-145), shape [n_groups, 1].
    """
    _require_cuda(x, lwc_up, lwc_low)
    if x.dim() != 2:
        raise ValueError("q100.quantize_groupwise expects a 2-D tensor")
    x = x.contiguous()
    rows, cols = x.shape
    g = cols if group in (0, None) else int(group)
    if cols % g != 0:
        raise ValueError(f"cols={cols} not divisible by group={g}")
    n_groups = rows * (cols // g)
    dev, dt = x.device, x.dtype
    res = QuantResult()
    if want_dequant:
        res.dequant = torch.empty_like(x)
    if want_packed:
        shp = packed_shape(rows, cols, n_bits)
        res.packed = (torch.zeros if rows % TILE_ROWS else torch.empty)(shp, dtype=torch.int32, device=dev)
    res.scales = torch.empty((n_groups, 1), dtype=dt, device=dev)
    if not disable_zero_point:
        res.zeros = torch.empty((n_groups, 1), dtype=dt, device=dev)
    if want_meta:
        res.meta = torch.zeros((cols // g, rows, 2), dtype=dt, device=dev)
    if lwc_up is not None:
        lwc_up = lwc_up.to(dt).contiguous()
        lwc_low = lwc_low.to(dt).contiguous()
        if lwc_up.numel() != n_groups or lwc_low.numel() != n_groups:
            raise ValueError("LWC multipliers must have one entry per group")
    with _on(dev):
        check(lib.q100_quantize_groupwise(
            x.data_ptr(), _dtype_code(x), rows, cols, 0 if group in (0, None) else g, n_bits,
            int(bool(symmetric)), int(bool(disable_zero_point)), _ptr(lwc_up), _ptr(lwc_low),
            _ptr(res.dequant), _ptr(res.packed), _ptr(res.scales), _ptr(res.zeros), _ptr(res.meta),
            flags, _stream(x)))
    return res


BACKWARD_GROUPS = (32, 64, 128, 256)


def quantize_backward(x: torch.Tensor, grad_out: torch.Tensor, group: int, n_bits: int, symmetric: bool = False,
                      disable_zero_point: bool = False, lwc_up: Optional[torch.Tensor] = None,
                      lwc_low: Optional[torch.Tensor] = None, need_x: bool = True, need_lwc: bool = False,
                      flags: int = 0):
    """Fused backward of :func:`quantize_groupwise` (K1b).  Returns (grad_x, grad_up, grad_low); entries that were
    not asked for (or have no meaning without LWC) are None.  ``group`` must be in BACKWARD_GROUPS."""
    _require_cuda(x, grad_out, lwc_up, lwc_low)
    if x.dim() != 2 or grad_out.shape != x.shape:
        raise ValueError("q100.quantize_backward expects 2-D x and grad_out of the same shape")
    x = x.contiguous()
    grad_out = grad_out.to(x.dtype).contiguous()
    rows, cols = x.shape
    n_groups = rows * (cols // group)
    gx = torch.empty_like(x) if need_x else None
    gup = glow = None
    if lwc_up is not None:
        lwc_up = lwc_up.to(x.dtype).contiguous()
        lwc_low = lwc_low.to(x.dtype).contiguous()
        if need_lwc:
            gup = torch.empty((n_groups, 1), dtype=x.dtype, device=x.device)
            glow = torch.empty((n_groups, 1), dtype=x.dtype, device=x.device)
    with _on(x.device):
        check(lib.q100_quantize_backward(
            x.data_ptr(), grad_out.data_ptr(), _dtype_code(x), rows, cols, int(group), n_bits,
            int(bool(symmetric)), int(bool(disable_zero_point)), _ptr(lwc_up), _ptr(lwc_low),
            _ptr(gx), _ptr(gup), _ptr(glow), flags, _stream(x)))
    return gx, gup, glow


def fakequant_rows(x: torch.Tensor, n_bits: int, flags: int = 0) -> QuantResult:
    """Per-token activation fake-quant over the last dim (any leading shape)."""
    _require_cuda(x)
    shp = x.shape
    x2 = x.contiguous().view(-1, shp[-1])
    rows, cols = x2.shape
    res = QuantResult(dequant=torch.empty_like(x2),
                      scales=torch.empty((rows, 1), dtype=x.dtype, device=x.device),
                      zeros=torch.empty((rows, 1), dtype=x.dtype, device=x.device))
    with _on(x.device):
        check(lib.q100_fakequant_rows(x2.data_ptr(), _dtype_code(x2), rows, cols, n_bits,
                                      res.dequant.data_ptr(), res.scales.data_ptr(), res.zeros.data_ptr(),
                                      flags, _stream(x2)))
    res.dequant = res.dequant.view(shp)
    res.scales = res.scales.view(*shp[:-1], 1)
    res.zeros = res.zeros.view(*shp[:-1], 1)
    return res


def rows_minmax(x: torch.Tensor) -> torch.Tensor:
    """[rows, 2] (min, max) over the last dim of ``x`` (any leading shape), in x's dtype -- the statistics pass of K2.
    With the feature dim sharded over tensor-parallel ranks the pairs are MIN / MAX all-reduced and handed to
    :func:`quantize_rows_stats` (q100/tp.py ``allreduce_minmax``)."""
    _require_cuda(x)
    x2 = x.contiguous().view(-1, x.shape[-1])
    out = torch.empty((x2.shape[0], 2), dtype=x.dtype, device=x.device)
    with _on(x.device):
        check(lib.q100_rows_minmax(x2.data_ptr(), _dtype_code(x2), x2.shape[0], x2.shape[1], out.data_ptr(), _stream(x2)))
    return out


def quantize_rows_stats(x: torch.Tensor, minmax: torch.Tensor, n_bits: int, symmetric: bool = False,
                        disable_zero_point: bool = False, lwc_up: Optional[torch.Tensor] = None,
                        lwc_low: Optional[torch.Tensor] = None, want_dequant: bool = True,
                        want_packed: bool = False, want_meta: bool = False, flags: int = 0) -> QuantResult:
    """One group per row of the 2-D ``x``, quantised with externally supplied row statistics ``minmax`` [rows, 2]
    (UniformAffineQuantizer.forward when the row is sharded over ranks: quantizer.py:140-159 on the global
    amin / amax)."""
    _require_cuda(x, minmax, lwc_up, lwc_low)
    if x.dim() != 2:
        raise ValueError("q100.quantize_rows_stats expects a 2-D tensor")
    x = x.contiguous()
    rows, cols = x.shape
    if minmax.shape != (rows, 2) or minmax.dtype != x.dtype:
        raise ValueError("minmax must be [rows, 2] in x's dtype")
    minmax = minmax.contiguous()
    dev, dt = x.device, x.dtype
    res = QuantResult()
    if want_dequant:
        res.dequant = torch.empty_like(x)
    if want_packed:
        shp = packed_shape(rows, cols, n_bits)
        res.packed = (torch.zeros if rows % TILE_ROWS else torch.empty)(shp, dtype=torch.int32, device=dev)
    res.scales = torch.empty((rows, 1), dtype=dt, device=dev)
    if not disable_zero_point:
        res.zeros = torch.empty((rows, 1), dtype=dt, device=dev)
    if want_meta:
        res.meta = torch.zeros((1, rows, 2), dtype=dt, device=dev)
    if lwc_up is not None:
        lwc_up = lwc_up.to(dt).contiguous()
        lwc_low = lwc_low.to(dt).contiguous()
    with _on(dev):
        check(lib.q100_quantize_rows_stats(
            x.data_ptr(), _dtype_code(x), rows, cols, n_bits, int(bool(symmetric)), int(bool(disable_zero_point)),
            minmax.data_ptr(), _ptr(lwc_up), _ptr(lwc_low), _ptr(res.dequant), _ptr(res.packed), _ptr(res.scales),
            _ptr(res.zeros), _ptr(res.meta), flags, _stream(x)))
    return res


def dequantize(packed: torch.Tensor, scales: Optional[torch.Tensor], zeros: Optional[torch.Tensor],
               rows: int, cols: int, group: int, n_bits: int, dtype: torch.dtype,
               disable_zero_point: bool = False) -> torch.Tensor:
    """Packed codes -> dense [rows, cols] tensor (raw integer codes when ``scales`` is None)."""
    _require_cuda(packed, scales, zeros)
    out = torch.empty((rows, cols), dtype=dtype, device=packed.device)
    with _on(packed.device):
        check(lib.q100_dequantize(packed.data_ptr(), _ptr(scales), _ptr(zeros), _DT[dtype], rows, cols,
                                  0 if group in (0, None) else group, n_bits, int(bool(disable_zero_point)),
                                  out.data_ptr(), _stream(packed)))
    return out


@dataclass
class PackedWeight:
    """A quantised nn.Linear weight [N, K] in the q100 v1 layout."""
    packed: torch.Tensor       # int32 [ceil(N/128), K/64, tile_words] (tile-major)
    meta: torch.Tensor         # [K/group, N, 2] (scale, zero) in the compute dtype
    scales: torch.Tensor       # [N*K/group, 1] reference layout
    zeros: Optional[torch.Tensor]
    n_bits: int
    group: int                 # elements per group (== K for per-channel)
    N: int
    K: int
    disable_zero_point: bool = False
    code_sums: Optional[torch.Tensor] = None   # int32 [N]: sum over k of the row's 8-bit codes (K3d, filled on first use)

    @property
    def dtype(self) -> torch.dtype:
        return self.meta.dtype

    def nbytes(self) -> int:
        return self.packed.numel() * 4 + self.meta.numel() * self.meta.element_size()


def pack_weight(w: torch.Tensor, n_bits: int, group: Optional[int], symmetric: bool = False,
                disable_zero_point: bool = False, lwc_up=None, lwc_low=None,
                want_dequant: bool = False):
    """Quantise + pack an [N, K] weight.  Returns (PackedWeight, QuantResult)."""
    g = w.shape[1] if group in (0, None) else int(group)
    r = quantize_groupwise(w, g, n_bits, symmetric, disable_zero_point, lwc_up, lwc_low,
                           want_dequant=want_dequant, want_packed=True, want_meta=True)
    pw = PackedWeight(r.packed, r.meta, r.scales, r.zeros, n_bits, g, w.shape[0], w.shape[1],
                      bool(disable_zero_point))
    return pw, r


_WORKSPACES: dict = {}
_DEFAULT_GEMM_FLAGS = 0


def set_default_gemm_flags(flags: int) -> int:
    """Flags OR-ed into every gemm_wna16 call that does not pass its own (e.g. ``_lib.GEMM_PDL`` while a
    decode step is captured: weights are static there, which is what the flag requires).  Returns the
    previous value."""
    global _DEFAULT_GEMM_FLAGS
    old, _DEFAULT_GEMM_FLAGS = _DEFAULT_GEMM_FLAGS, int(flags)
    return old


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) zero-initialised scratch; kernels leave it zeroed."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 22), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def gemm_wna16(x: torch.Tensor, w: PackedWeight, bias: Optional[torch.Tensor] = None,
               flags: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ dequant(w)^T (+ bias) straight from packed codes (K3).  ``out``: a preallocated contiguous
    [rows, N] destination (e.g. a symmetric-memory buffer a collective reads next)."""
    _require_cuda(x, w.packed, bias)
    if x.dtype != w.dtype:
        raise TypeError(f"q100.gemm_wna16: x is {x.dtype} but the packed weight was built for {w.dtype}")
    shp = x.shape
    flags |= _DEFAULT_GEMM_FLAGS
    kx = 2 * w.K if flags & _lib.GEMM_X_GATE_UP else w.K      # gate | up in, act_fn(gate) * up staged by the kernel
    if shp[-1] != kx:
        raise ValueError(f"q100.gemm_wna16: x has K={shp[-1]}, expected {kx}")
    x2 = x.contiguous().view(-1, kx)
    M = x2.shape[0]
    if out is None:
        y = torch.empty((M, w.N), dtype=x.dtype, device=x.device)
    else:
        if out.numel() != M * w.N or out.dtype != x.dtype or not out.is_contiguous():
            raise ValueError("q100.gemm_wna16: out must be a contiguous [rows, N] tensor in the dtype of x")
        y = out
    if M == 0:
        return y.view(*shp[:-1], w.N)
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    need = _gemm_workspace_bytes(M, w.N, w.K, w.n_bits, flags)
    ws = _workspace(x.device, need)
    with _on(x.device):
        check(lib.q100_gemm_wna16(x2.data_ptr(), w.packed.data_ptr(), w.meta.data_ptr(), _ptr(bias),
                                  y.data_ptr(), M, w.N, w.K, w.group, w.n_bits, int(w.disable_zero_point),
                                  _DT[x.dtype], ws.data_ptr(), ws.numel(), flags, _stream(x2)))
    return y.view(*shp[:-1], w.N)


def gemm_block_fusable(rows: int, w: PackedWeight) -> bool:
    """True when ``gemm_wna16_block`` takes ``rows`` x this weight (the row-owner decode kernel K3r applies)."""
    return bool(lib.q100_gemm_block_fusable(rows, w.N, w.K, w.group, w.n_bits))


def gemm_wna16_block(x: torch.Tensor, w: PackedWeight, bias: Optional[torch.Tensor] = None, flags: int = 0,
                     norm_weight: Optional[torch.Tensor] = None, eps: float = 0.0,
                     residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode GEMM with the glue of a decoder block folded in (q100_gemm_wna16_block): ``norm_weight`` -- x is the
    residual stream and OmniLlamaRMSNorm(x) is what gets multiplied; ``residual`` -- returns residual + x W^T.
    Raises Q100Error(EUNSUPPORTED) unless ``gemm_block_fusable``."""
    _require_cuda(x, w.packed, bias, norm_weight, residual)
    if x.dtype != w.dtype:
        raise TypeError(f"q100.gemm_wna16_block: x is {x.dtype} but the packed weight was built for {w.dtype}")
    flags |= _DEFAULT_GEMM_FLAGS & _lib.GEMM_PDL
    kx = 2 * w.K if flags & _lib.GEMM_X_GATE_UP else w.K
    if x.shape[-1] != kx:
        raise ValueError(f"q100.gemm_wna16_block: x has K={x.shape[-1]}, expected {kx}")
    shp = x.shape
    x2 = x.contiguous().view(-1, kx)
    M = x2.shape[0]
    y = torch.empty((M, w.N), dtype=x.dtype, device=x.device) if out is None else out
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    if norm_weight is not None:
        norm_weight = norm_weight.to(x.dtype).contiguous()
        if norm_weight.numel() != w.K:
            raise ValueError("q100.gemm_wna16_block: norm_weight must have K elements")
    if residual is not None:
        residual = residual.contiguous()
        if residual.numel() != M * w.N or residual.dtype != x.dtype:
            raise ValueError("q100.gemm_wna16_block: residual must be [rows, N] in the dtype of x")
    with _on(x.device):
        check(lib.q100_gemm_wna16_block(x2.data_ptr(), w.packed.data_ptr(), w.meta.data_ptr(), _ptr(bias), y.data_ptr(),
                                        M, w.N, w.K, w.group, w.n_bits, int(w.disable_zero_point), _DT[x.dtype],
                                        _ptr(norm_weight), float(eps), _ptr(residual), flags, _stream(x2)))
    return y.view(*shp[:-1], w.N)


def gemm_wna8(x: torch.Tensor, w: PackedWeight, a_bits: int, minmax: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """y = fakequant_rows(x, a_bits) @ dequant(w)^T with the activation quantisation fused into the GEMM and the
    contraction on the integer tensor cores (K3c, decode shapes: at most 15 rows).  ``minmax`` [rows, 2]: row
    statistics over the full K (default: computed here with rows_minmax; pass the all-reduced ones under TP)."""
    _require_cuda(x, w.packed, bias, minmax)
    if x.dtype != w.dtype:
        raise TypeError(f"q100.gemm_wna8: x is {x.dtype} but the packed weight was built for {w.dtype}")
    shp = x.shape
    x2 = x.contiguous().view(-1, w.K)
    M = x2.shape[0]
    if minmax is None:
        minmax = rows_minmax(x2)
    y = torch.empty((M, w.N), dtype=x.dtype, device=x.device)
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    need = _gemm_workspace_bytes(M, w.N, w.K, w.n_bits, 0)
    ws = _workspace(x.device, need)
    with _on(x.device):
        check(lib.q100_gemm_wna8(x2.data_ptr(), minmax.data_ptr(), w.packed.data_ptr(), w.meta.data_ptr(), _ptr(bias),
                                 y.data_ptr(), M, w.N, w.K, w.group, w.n_bits, int(a_bits), int(w.disable_zero_point),
                                 _DT[x.dtype], ws.data_ptr(), ws.numel(), flags, _stream(x2)))
    return y.view(*shp[:-1], w.N)


def code_rowsums(packed: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """int32 [rows]: the sum over k of every row's codes of a packed 8-bit [rows, cols] matrix (the zero-point
    correction terms of K3d)."""
    _require_cuda(packed)
    out = torch.empty((rows,), dtype=torch.int32, device=packed.device)
    with _on(packed.device):
        check(lib.q100_code_rowsums(packed.data_ptr(), rows, cols, out.data_ptr(), _stream(packed)))
    return out


def gemm_w8a8_supported(rows: int, w: PackedWeight) -> bool:
    """True when :func:`gemm_w8a8` takes ``rows`` x this weight: 8-bit per-channel codes with zero-points."""
    return (w.n_bits == 8 and w.group == w.K and not w.disable_zero_point and w.K % 128 == 0 and w.K <= 32768
            and w.N % 8 == 0 and rows > 0)


def gemm_w8a8(x: torch.Tensor, w: PackedWeight, minmax: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """y = fakequant_rows(x, 8) @ dequant(w)^T for a per-channel 8-bit weight at prefill sizes (K3d): the activations are
    quantised and packed once (K2 in pack mode; with ``minmax`` [rows, 2] the externally supplied -- all-reduced -- row
    statistics are used), then codes x codes on the integer tensor cores (tcgen05.mma kind::i8) with the zero-point
    terms applied in the epilogue.  ``flags``: FLAG_CUDA_SCALE for the activation quantiser."""
    _require_cuda(x, w.packed, bias, minmax)
    if x.dtype != w.dtype:
        raise TypeError(f"q100.gemm_w8a8: x is {x.dtype} but the packed weight was built for {w.dtype}")
    if x.shape[-1] != w.K:
        raise ValueError(f"q100.gemm_w8a8: x has K={x.shape[-1]}, expected {w.K}")
    if not gemm_w8a8_supported(1, w):
        raise ValueError("q100.gemm_w8a8: per-channel 8-bit weights with zero-points, K % 128 == 0, N % 8 == 0")
    shp = x.shape
    x2 = x.contiguous().view(-1, w.K)
    M = x2.shape[0]
    y = torch.empty((M, w.N), dtype=x.dtype, device=x.device)
    if M == 0:
        return y.view(*shp[:-1], w.N)
    if minmax is None:
        r = quantize_groupwise(x2, 0, 8, want_dequant=False, want_packed=True, want_meta=True, flags=flags)
    else:
        r = quantize_rows_stats(x2, minmax, 8, want_dequant=False, want_packed=True, want_meta=True, flags=flags)
    xs = code_rowsums(r.packed, M, w.K)
    if w.code_sums is None:
        w.code_sums = code_rowsums(w.packed, w.N, w.K)
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    with _on(x.device):
        check(lib.q100_gemm_w8a8(r.packed.data_ptr(), r.meta.data_ptr(), xs.data_ptr(), w.packed.data_ptr(),
                                 w.meta.data_ptr(), w.code_sums.data_ptr(), _ptr(bias), y.data_ptr(), M, w.N, w.K,
                                 _DT[x.dtype], _stream(x2)))
    return y.view(*shp[:-1], w.N)


def add_rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, residual: Optional[torch.Tensor] = None,
                out_norm: Optional[torch.Tensor] = None, out_sum: Optional[torch.Tensor] = None):
    """(h, n) with h = residual + x (or x) and n = OmniLlamaRMSNorm(h); one launch (K4).  ``out_norm`` / ``out_sum``:
    preallocated destinations (fixed addresses for a CUDA graph / a decode chain)."""
    _require_cuda(x, weight, residual)
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    norm = torch.empty_like(x) if out_norm is None else out_norm
    h = x
    if residual is not None:
        residual = residual.contiguous()
        h = torch.empty_like(x) if out_sum is None else out_sum
    with _on(x.device):
        check(lib.q100_add_rmsnorm(x.data_ptr(), _ptr(residual), weight.data_ptr(), float(eps),
                                   h.data_ptr() if residual is not None else None, norm.data_ptr(), rows, cols,
                                   _DT[x.dtype], _stream(x)))
    return h, norm


def rope_kv_decode(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, position: torch.Tensor,
                   k_cache: torch.Tensor, v_cache: torch.Tensor, heads: int, kv_heads: int,
                   err_flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rotary on q/k of one new token per sequence + append of k, v to the static cache (K5).
    Returns q [batch, heads, 1, head_dim].  A position outside the cache / the rotary table writes nothing
    and sets ``err_flag`` (int32 [1] on the device) to 1."""
    _require_cuda(qkv, cos, sin, position, k_cache, v_cache)
    bsz = k_cache.shape[0]
    hd, max_len = k_cache.shape[-1], k_cache.shape[-2]
    qkv = qkv.contiguous()
    if qkv.numel() != bsz * (heads + 2 * kv_heads) * hd or position.dtype != torch.int64:
        raise ValueError("q100.rope_kv_decode: qkv must be [batch, (heads + 2 kv_heads) * head_dim], position int64")
    if cos.dtype != qkv.dtype or cos.shape[-1] != hd or not cos.is_contiguous():
        raise ValueError("q100.rope_kv_decode: cos/sin cache must be contiguous [pos, head_dim] in the model dtype")
    q = torch.empty((bsz, heads, 1, hd), dtype=qkv.dtype, device=qkv.device)
    with _on(qkv.device):
        check(lib.q100_rope_kv_decode(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), position.data_ptr(),
                                      q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), bsz, heads, kv_heads,
                                      hd, max_len, cos.shape[0], _ptr(err_flag), _DT[qkv.dtype], _stream(qkv)))
    return q


_ATTN_WS: dict = {}


def attn_decode(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, position: torch.Tensor,
                k_cache: torch.Tensor, v_cache: torch.Tensor, heads: int, kv_heads: int,
                err_flag: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The attention core of a decode step in one launch (K8): rotary on q / k of one new token per sequence, append of
    k, v to the static cache at ``position`` and softmax(q k^T / sqrt(d)) v over the slots 0..position.  Returns
    [batch, heads * head_dim] (what o_proj reads).  ``position``: int64 scalar on the device, shared by the batch."""
    _require_cuda(qkv, cos, sin, position, k_cache, v_cache)
    bsz, hd, max_len = k_cache.shape[0], k_cache.shape[-1], k_cache.shape[-2]
    qkv = qkv.contiguous()
    if qkv.numel() != bsz * (heads + 2 * kv_heads) * hd or position.dtype != torch.int64:
        raise ValueError("q100.attn_decode: qkv must be [batch, (heads + 2 kv_heads) * head_dim], position int64")
    if cos.dtype != qkv.dtype or cos.shape[-1] != hd or not cos.is_contiguous() or not sin.is_contiguous():
        raise ValueError("q100.attn_decode: cos/sin cache must be contiguous [pos, head_dim] in the model dtype")
    if not (k_cache.is_contiguous() and v_cache.is_contiguous()) or k_cache.dtype != qkv.dtype:
        raise ValueError("q100.attn_decode: contiguous [batch, kv_heads, max_len, head_dim] caches in the model dtype")
    dev = qkv.device
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, bsz, heads, kv_heads, max_len)
    ws = _ATTN_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(lib.q100_attn_decode_workspace_bytes(bsz, heads, kv_heads, max_len)), dtype=torch.uint8,
                         device=dev)
        _ATTN_WS[key] = ws
    if out is None:
        out = torch.empty((bsz, heads * hd), dtype=qkv.dtype, device=dev)
    with _on(dev):
        check(lib.q100_attn_decode(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), position.data_ptr(),
                                   k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                   bsz, heads, kv_heads, hd, max_len, cos.shape[0], _ptr(err_flag), _DT[qkv.dtype],
                                   _stream(qkv)))
    return out


def attn_decode_supported(heads: int, kv_heads: int, head_dim: int) -> bool:
    rep = heads // max(kv_heads, 1)
    return head_dim == 128 and kv_heads > 0 and heads % kv_heads == 0 and 1 <= rep <= 8


def rope_qkv_prefill(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, position_ids: torch.Tensor,
                     heads: int, kv_heads: int, err_flag: Optional[torch.Tensor] = None):
    """Rotary on q / k of ``q_len`` new tokens per sequence and the head-major layouts the attention core reads (K5p).
    ``qkv`` [batch, q_len, (heads + 2 kv_heads) * head_dim] (the fused projection); ``position_ids`` int64 [1, q_len]
    or [batch, q_len].  Returns q [batch, heads, q_len, hd], k, v [batch, kv_heads, q_len, hd]."""
    _require_cuda(qkv, cos, sin, position_ids)
    bsz, q_len, width = qkv.shape
    hd = width // (heads + 2 * kv_heads)
    qkv = qkv.contiguous()
    if hd * (heads + 2 * kv_heads) != width or position_ids.dtype != torch.int64:
        raise ValueError("q100.rope_qkv_prefill: qkv must be [batch, q_len, (heads + 2 kv_heads) * head_dim], positions int64")
    if cos.dtype != qkv.dtype or cos.shape[-1] != hd or not cos.is_contiguous() or not sin.is_contiguous():
        raise ValueError("q100.rope_qkv_prefill: cos/sin cache must be contiguous [pos, head_dim] in the model dtype")
    pos = position_ids.reshape(-1, q_len).contiguous()
    if pos.shape[0] not in (1, bsz):
        raise ValueError("q100.rope_qkv_prefill: position_ids must be [1, q_len] or [batch, q_len]")
    q = torch.empty((bsz, heads, q_len, hd), dtype=qkv.dtype, device=qkv.device)
    k = torch.empty((bsz, kv_heads, q_len, hd), dtype=qkv.dtype, device=qkv.device)
    v = torch.empty_like(k)
    with _on(qkv.device):
        check(lib.q100_rope_qkv_prefill(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr(),
                                        0 if pos.shape[0] == 1 else q_len, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                        bsz, q_len, heads, kv_heads, hd, cos.shape[0], _ptr(err_flag),
                                        _DT[qkv.dtype], _stream(qkv)))
    return q, k, v


def silu_mul(gate_up: torch.Tensor) -> torch.Tensor:
    """silu(gate) * up on the fused gate|up GEMM output [..., 2I] -> [..., I] (K6)."""
    _require_cuda(gate_up)
    gate_up = gate_up.contiguous()
    inter = gate_up.shape[-1] // 2
    rows = gate_up.numel() // (2 * inter)
    out = torch.empty((*gate_up.shape[:-1], inter), dtype=gate_up.dtype, device=gate_up.device)
    with _on(gate_up.device):
        check(lib.q100_silu_mul(gate_up.data_ptr(), out.data_ptr(), rows, inter, _DT[gate_up.dtype],
                                _stream(gate_up)))
    return out


_ARGMAX_WS: dict = {}


def argmax_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int64 argmax over the last dim of a 16-bit tensor (K10; the greedy token choice of a decode step).  Ties go to
    the lowest index, NaN counts as the maximum.  Returns shape x.shape[:-1]."""
    _require_cuda(x)
    x2 = x.contiguous().view(-1, x.shape[-1])
    rows, cols = x2.shape
    dev = x.device
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, rows)
    ws = _ARGMAX_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(lib.q100_argmax_workspace_bytes(rows)), dtype=torch.uint8, device=dev)
        _ARGMAX_WS[key] = ws
    res = torch.empty(rows, dtype=torch.int64, device=dev) if out is None else out
    if res.numel() != rows or res.dtype != torch.int64 or not res.is_contiguous():
        raise ValueError("q100.argmax_rows: out must be a contiguous int64 tensor with one entry per row")
    with _on(dev):
        check(lib.q100_argmax_rows(x2.data_ptr(), rows, cols, _dtype_code(x2), res.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream(x2)))
    return res.view(x.shape[:-1]) if out is None else out


def selftest_div(dtype: torch.dtype, device="cuda") -> int:
    """Exhaustive check of the fast 16-bit division; returns the mismatch count."""
    out = torch.zeros(1, dtype=torch.int64, device=device)
    with _on(out.device):
        check(lib.q100_selftest_div(_DT[dtype], out.data_ptr(), _stream(out)))
    return int(out.item())


def debug_gemm_dense_tc(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x[M,K] @ w[N,K]^T through the tcgen05 plumbing with a dense 16-bit weight (regression hook)."""
    _require_cuda(x, w)
    x = x.contiguous()
    w = w.contiguous()
    y = torch.empty((x.shape[0], w.shape[0]), dtype=x.dtype, device=x.device)
    with _on(x.device):
        check(lib.q100_debug_gemm_dense_tc(x.data_ptr(), w.data_ptr(), y.data_ptr(), x.shape[0], w.shape[0],
                                           x.shape[1], _DT[x.dtype], _stream(x)))
    return y
