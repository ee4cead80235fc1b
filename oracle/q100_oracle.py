"""CPU oracle for the group-wise quantized-linear hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``llama3-quantization_b200/q100`` + ``quant``) never imports it and has no CPU
fallback.

What it restates (all citations relative to the reference checkout):

* ``quant/quantizer.py:15-19``    round_ste            -> :func:`round_ste` (round-half-even)
* ``quant/quantizer.py:132-159``  per_token_dynamic_calibration -> :func:`calibrate`
* ``quant/quantizer.py:94-115``   fake_quant           -> :func:`fake_quant`
* ``quant/quantizer.py:118-130``  forward              -> :func:`quantize_groupwise`
* ``quant/int_linear.py:48-65``   QuantLinear.forward  -> :func:`quant_linear`
* ``quant/int_matmul.py:31-43``   QuantMatMul          -> :func:`fakequant_rows` (+ numpy matmul)

Arithmetic model (numpy only, independent of torch): the reference runs every
elementwise op in the tensor's storage dtype, and torch's CPU kernels for
fp16/bf16 compute each op in fp32 and round the result to the storage dtype.
So every op below is done in fp32 and followed by :func:`rnd` to the storage
dtype.  Division is IEEE (numpy fp32 true division), rounding is
round-half-to-even (``np.rint``).

Parity pin: ``oracle/make_golden.py`` imports the unmodified reference modules
from ``/root/reference`` (possible in the build container only), runs them on
seeded inputs and edge cases, checks this oracle against them bit-for-bit and
writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures everywhere (no reference needed).
"""
from __future__ import annotations

import numpy as np

CLIPMIN = np.float32(1e-5)   # quant/quantizer.py:10
CLIPMAX = np.float32(1e4)    # quant/quantizer.py:149,154
ZP_LIM = np.float32(1e4)     # quant/quantizer.py:159

DTYPES = ("f16", "bf16", "f32")


# --------------------------------------------------------------------------- dtypes
def rnd(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round an fp32 array to ``dtype`` (RNE) and return it widened back to fp32."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == "f32":
        return x
    if dtype == "f16":
        with np.errstate(over="ignore"):
            return x.astype(np.float16).astype(np.float32)
    if dtype == "bf16":
        u = x.view(np.uint32).astype(np.uint64)
        nan = np.isnan(x)
        lsb = (u >> np.uint64(16)) & np.uint64(1)
        u = (u + np.uint64(0x7FFF) + lsb) & np.uint64(0xFFFF0000)
        out = u.astype(np.uint32).view(np.float32).copy()
        out[nan] = np.nan
        return out.reshape(x.shape)
    raise ValueError(dtype)


def to_bits(x: np.ndarray, dtype: str) -> np.ndarray:
    """fp32 array holding values exactly representable in ``dtype`` -> raw storage bits."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if dtype == "f32":
        return x.view(np.uint32).copy()
    if dtype == "f16":
        return x.astype(np.float16).view(np.uint16).copy()
    if dtype == "bf16":
        return (x.view(np.uint32) >> np.uint32(16)).astype(np.uint16)
    raise ValueError(dtype)


def from_bits(b: np.ndarray, dtype: str) -> np.ndarray:
    """Raw storage bits -> fp32 values."""
    if dtype == "f32":
        return np.ascontiguousarray(b, dtype=np.uint32).view(np.float32).copy()
    if dtype == "f16":
        return np.ascontiguousarray(b, dtype=np.uint16).view(np.float16).astype(np.float32)
    if dtype == "bf16":
        return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32).copy()
    raise ValueError(dtype)


def _rint(x):
    """torch.round == round-half-to-even (quant/quantizer.py:19,159)."""
    return np.rint(x)


def round_ste(v, dtype):
    """quant/quantizer.py:15-19 forward value: ``(v.round() - v) + v`` with per-op rounding.

    Equals rint(v) except that a zero result is always +0.0 and +-inf becomes NaN
    (inf - inf); both matter for bit-exactness (signed zeros without a zero-point,
    fp16 overflow of x/scale when scale sits at CLIPMIN).
    """
    return rnd(rnd(_rint(v) - v, dtype) + v, dtype)


def _clamp(x, lo, hi):
    """torch.clamp semantics: NaN propagates, otherwise min(max(x, lo), hi)."""
    out = np.minimum(np.maximum(x, np.float32(lo)), np.float32(hi))
    out = np.where(np.isnan(x), x, out)
    return out.astype(np.float32)


def qrange(n_bits: int, disable_zero_point: bool):
    """quant/quantizer.py:46-51."""
    if disable_zero_point:
        return -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
    return 0, 2 ** n_bits - 1


# --------------------------------------------------------------------------- calibration
def calibrate(xg: np.ndarray, dtype: str, n_bits: int, symmetric: bool = False,
              disable_zero_point: bool = False, lwc_up=None, lwc_low=None, ext_minmax=None):
    """Per-row min/max -> (scale, round_zero_point).  quant/quantizer.py:140-159.
    ``ext_minmax`` [groups, 2]: amin / amax of the WHOLE row when ``xg`` is only a tensor-parallel slice of it
    (the reference has no sharding; this states what its quantizer.py:140-141 sees on the unsharded row).

    ``xg`` is ``[groups, g]`` fp32 holding ``dtype``-representable values.
    ``lwc_up`` / ``lwc_low`` are the *post-sigmoid* multipliers ``[groups, 1]``
    (quantizer.py:143-145); the sigmoid itself is left to torch, see SURVEY §7.
    Returns fp32 arrays ``[groups, 1]``; zp is ``None`` when disabled.
    """
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        xmin = np.min(xg, axis=-1, keepdims=True)          # amin/amax propagate NaN like torch
        xmax = np.max(xg, axis=-1, keepdims=True)
        if ext_minmax is not None:
            em = np.asarray(ext_minmax, np.float32).reshape(-1, 2)
            xmin, xmax = em[:, :1].copy(), em[:, 1:].copy()
        if lwc_up is not None:
            xmax = rnd(np.asarray(lwc_up, np.float32).reshape(-1, 1) * xmax, dtype)
            xmin = rnd(np.asarray(lwc_low, np.float32).reshape(-1, 1) * xmin, dtype)
        if symmetric:
            abs_max = np.maximum(np.abs(xmax), np.abs(xmin))
            abs_max = np.where(np.isnan(xmax) | np.isnan(xmin), np.float32(np.nan), abs_max)
            scale = rnd(abs_max / np.float32(2 ** (n_bits - 1) - 1), dtype)
            scale = rnd(_clamp(scale, CLIPMIN, CLIPMAX), dtype)
            zp = rnd(np.full_like(scale, np.float32(2 ** (n_bits - 1) - 1)), dtype)
        else:
            rng = rnd(xmax - xmin, dtype)
            scale = rnd(rng / np.float32(2 ** n_bits - 1), dtype)
            scale = rnd(_clamp(scale, CLIPMIN, CLIPMAX), dtype)
            zp = rnd(rnd(-xmin, dtype) / scale, dtype)
        if disable_zero_point:
            return scale.astype(np.float32), None
        rzp = _rint(rnd(_clamp(zp, -ZP_LIM, ZP_LIM), dtype))
        return scale.astype(np.float32), rzp.astype(np.float32)


def fake_quant(xg: np.ndarray, scale: np.ndarray, rzp, dtype: str, n_bits: int,
               disable_zero_point: bool = False):
    """quant/quantizer.py:103-110 on ``[groups, g]``; returns (dequant fp32, codes fp32).

    ``codes`` are the clamped integer-valued ``x_int`` the reference never
    materialises (extracted with the oracle's own scale/zp, never
    as round(deq/scale)+zp).
    """
    qmin, qmax = qrange(n_bits, disable_zero_point)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        x_int = round_ste(rnd(xg / scale, dtype), dtype)
        if rzp is not None:
            x_int = rnd(x_int + rzp, dtype)
        codes = _clamp(x_int, qmin, qmax)
        deq = codes
        if rzp is not None:
            deq = rnd(deq - rzp, dtype)
        deq = rnd(deq * scale, dtype)
    return deq.astype(np.float32), codes.astype(np.float32)


def quantize_groupwise(w: np.ndarray, dtype: str, group: int, n_bits: int,
                       symmetric: bool = False, disable_zero_point: bool = False,
                       lwc_up=None, lwc_low=None, ext_minmax=None):
    """UniformAffineQuantizer.forward on a 2-D tensor (quant/quantizer.py:118-130).

    ``group == 0`` means one group per row (``group_size=None``).  ``w`` is fp32
    holding ``dtype``-representable values, ``[rows, cols]`` with
    ``cols % group == 0``.  Returns dict(dequant, codes, scales, zeros); scales /
    zeros are ``[rows*cols/group, 1]`` exactly like the reference's
    ``self.scale`` / ``self.round_zero_point``.
    """
    w = np.asarray(w, dtype=np.float32)
    rows, cols = w.shape
    g = cols if group in (0, None) else group
    assert cols % g == 0
    xg = w.reshape(-1, g)
    scale, rzp = calibrate(xg, dtype, n_bits, symmetric, disable_zero_point, lwc_up, lwc_low, ext_minmax)
    deq, codes = fake_quant(xg, scale, rzp, dtype, n_bits, disable_zero_point)
    return {
        "dequant": deq.reshape(rows, cols),
        "codes": codes.reshape(rows, cols),
        "scales": scale,
        "zeros": rzp,
    }


def fakequant_rows(x: np.ndarray, dtype: str, n_bits: int):
    """Per-token activation fake-quant: asymmetric, one group per last-dim row.

    quant/int_linear.py:59-60 and quant/int_matmul.py:31-39 with the
    ``act_quant_params`` of main.py:326-349 (symmetric False, group_size None).
    """
    shp = x.shape
    r = quantize_groupwise(np.asarray(x, np.float32).reshape(-1, shp[-1]), dtype, 0, n_bits)
    r["dequant"] = r["dequant"].reshape(shp)
    r["codes"] = r["codes"].reshape(shp)
    return r


# --------------------------------------------------------------------------- dequant / linear
def dequant_codes(codes: np.ndarray, scales: np.ndarray, zeros, dtype: str, group: int):
    """(codes - zp) * scale with per-op rounding (quant/quantizer.py:107-110)."""
    rows, cols = codes.shape
    g = cols if group in (0, None) else group
    c = codes.reshape(-1, g).astype(np.float32)
    s = np.asarray(scales, np.float32).reshape(-1, 1)
    d = c if zeros is None else rnd(c - np.asarray(zeros, np.float32).reshape(-1, 1), dtype)
    return rnd(d * s, dtype).reshape(rows, cols)


def linear(x: np.ndarray, w_deq: np.ndarray, bias, dtype: str) -> np.ndarray:
    """F.linear on already fake-quantised weights (quant/int_linear.py:62).

    Accumulates in float64 and rounds once to ``dtype``: the ideal result any
    fp32-accumulating GEMM (cuBLAS, MKL, ours) is compared against with a
    tolerance, never bit-exactly.
    """
    y = np.asarray(x, np.float64).reshape(-1, x.shape[-1]) @ np.asarray(w_deq, np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    return rnd(y.astype(np.float32), dtype).reshape(*x.shape[:-1], w_deq.shape[0])


def quant_linear(x, w, bias, dtype, group, w_bits, a_bits=16, symmetric=False,
                 disable_zero_point=False):
    """QuantLinear.forward with use_weight_quant (and use_act_quant if a_bits < 16).

    quant/int_linear.py:48-65: weight re-quantised, input per-token fake-quantised
    when ``a_bits < 16``, then F.linear.
    """
    wq = quantize_groupwise(w, dtype, group, w_bits, symmetric, disable_zero_point)["dequant"] \
        if w_bits < 16 else np.asarray(w, np.float32)
    xq = fakequant_rows(x, dtype, a_bits)["dequant"] if a_bits < 16 else np.asarray(x, np.float32)
    return linear(xq, wq, bias, dtype)


# --------------------------------------------------------------------------- packed layout (q100 v2)
# The reference never materialises integer codes; its --real_quant packing lives in
# AutoGPTQ (not vendored).  The packed layout is therefore this project's own and is
# pinned by the round trip unpack(pack(codes)) == codes plus the CUDA kernel writing
# byte-identical words.  See DESIGN.md "Packed layout".
#
#   Tile-major: the [N, K] code matrix is cut into tiles of 128 rows x 64 k; tiles are stored
#   contiguously, ordered [n_tile][k_tile] (so a CTA streaming along K for one row tile reads one
#   contiguous run of HBM); N is padded up to a multiple of 128 with zero codes.
#   Inside a tile, rows are stored one after the other.  A row's 64 codes are packed as
#     n_bits in {2,4,8}: F = 32/n_bits codes per uint32, consecutive k; code e of a word sits in
#       field (e >> 1) + (e & 1) * F/2, i.e. bit n_bits*(e>>1) + 16*(e&1)
#       (so ``(word >> n_bits*j) & pair_mask`` yields the half2 pair (e[2j], e[2j+1])).
#     n_bits == 3: the tile holds two planes: all rows' low 2 bits packed as n_bits=2 words
#       (128 rows x 4 words) followed by all rows' high bit packed as 1-bit fields
#       (128 rows x 2 words, same even/odd split).
#   Stored code = x_int - qmin (unsigned).
TILE_ROWS = 128
TILE_K = 64


def tile_words(n_bits: int) -> int:
    if n_bits == 3:
        return TILE_ROWS * (4 + 2)
    return TILE_ROWS * (TILE_K * n_bits // 32)


def packed_shape(rows: int, cols: int, n_bits: int):
    assert cols % TILE_K == 0
    return ((rows + TILE_ROWS - 1) // TILE_ROWS, cols // TILE_K, tile_words(n_bits))


def _pack_plane(u: np.ndarray, bits: int) -> np.ndarray:
    """u: [..., 64] unsigned codes -> [..., 64*bits/32] words."""
    F = 32 // bits
    e = u.reshape(*u.shape[:-1], u.shape[-1] // F, F).astype(np.uint32)
    word = np.zeros(e.shape[:-1], dtype=np.uint32)
    for j in range(F // 2):
        word |= e[..., 2 * j] << np.uint32(bits * j)
        word |= e[..., 2 * j + 1] << np.uint32(bits * (F // 2 + j))
    return word


def _unpack_plane(word: np.ndarray, bits: int) -> np.ndarray:
    F = 32 // bits
    mask = np.uint32((1 << bits) - 1)
    e = np.zeros((*word.shape, F), dtype=np.uint32)
    for j in range(F // 2):
        e[..., 2 * j] = (word >> np.uint32(bits * j)) & mask
        e[..., 2 * j + 1] = (word >> np.uint32(bits * (F // 2 + j))) & mask
    return e.reshape(*word.shape[:-1], word.shape[-1] * F)


def pack_codes(codes: np.ndarray, n_bits: int, disable_zero_point: bool = False) -> np.ndarray:
    """Integer-valued codes ``[rows, cols]`` -> uint32 ``[n_tiles, k_tiles, tile_words]``."""
    qmin, _ = qrange(n_bits, disable_zero_point)
    rows, cols = codes.shape
    NT, KT, TW = packed_shape(rows, cols, n_bits)
    u = (np.nan_to_num(np.asarray(codes, np.float64), nan=float(qmin)) - qmin).astype(np.int64)
    assert u.min() >= 0 and u.max() < (1 << n_bits)
    up = np.zeros((NT * TILE_ROWS, cols), dtype=np.uint32)
    up[:rows] = u
    t = up.reshape(NT, TILE_ROWS, KT, TILE_K).transpose(0, 2, 1, 3)          # [NT, KT, 128, 64]
    if n_bits == 3:
        lo = _pack_plane(t & np.uint32(3), 2).reshape(NT, KT, -1)
        hi = _pack_plane(t >> np.uint32(2), 1).reshape(NT, KT, -1)
        return np.ascontiguousarray(np.concatenate([lo, hi], axis=2))
    return np.ascontiguousarray(_pack_plane(t, n_bits).reshape(NT, KT, TW))


def unpack_codes(words: np.ndarray, rows: int, cols: int, n_bits: int,
                 disable_zero_point: bool = False) -> np.ndarray:
    qmin, _ = qrange(n_bits, disable_zero_point)
    NT, KT, TW = packed_shape(rows, cols, n_bits)
    words = np.asarray(words, dtype=np.uint32).reshape(NT, KT, TW)
    if n_bits == 3:
        lo = _unpack_plane(words[:, :, : TILE_ROWS * 4].reshape(NT, KT, TILE_ROWS, 4), 2)
        hi = _unpack_plane(words[:, :, TILE_ROWS * 4:].reshape(NT, KT, TILE_ROWS, 2), 1)
        t = lo | (hi << np.uint32(2))
    else:
        t = _unpack_plane(words.reshape(NT, KT, TILE_ROWS, -1), n_bits)
    u = t.transpose(0, 2, 1, 3).reshape(NT * TILE_ROWS, cols)[:rows]
    return u.astype(np.float32) + np.float32(qmin)
