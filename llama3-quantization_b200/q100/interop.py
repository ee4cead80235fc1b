"""Packed checkpoint format and interop: the data formats either side of the path.

* q100 v2 packed weights <-> state dicts: ``packed_state`` / ``packed_from_state`` for one ``PackedWeight``,
  ``save_packed`` / ``load_packed`` for every ``QuantLinear`` of a module tree -- the real-quantised counterpart of
  what the reference leaves in a checkpoint after ``register_scales_and_zeros`` (quant/quantizer.py:161-165, buffers
  ``scales`` / ``zeros``) and ``--save_dir`` (main.py:399-414).
* q100 v2 <-> AutoGPTQ v1 tensors (``qweight`` / ``qzeros`` / ``scales`` / ``g_idx``): what the reference's
  ``--real_quant`` path produces by handing ``scales`` / ``zeros`` to AutoGPTQ's ``QuantLinear.pack``
  (quant/omniquant.py:315-336) and what its evaluation scripts load.

AutoGPTQ (auto-gptq 0.7.1, environment.yml:24) is not vendored in the reference and not installed here: its layout
is restated from its published ``qlinear_cuda_old`` / ``qlinear_triton`` ``pack()``:

    intweight[k, n] in [0, 2^b)                          the code of W[n, k]
    qweight  int32 [K / 32 * b, N]     b in {2, 4, 8}: 32 / b consecutive k per word, element i at bit b * i
                                       b == 3: 32 consecutive k per three words (10 | 1 split 2+1 | 10 | 1 split 1+2 | 10)
    qzeros   int32 [K / g, N / 32 * b] (zero - 1), packed along n the same way
    scales   fp16  [K / g, N]
    g_idx    int32 [K]                 k // g   (no activation reordering)
    dequantised weight = scales * (code - (stored_zero + 1))

No test of the reference pins results at this boundary ("parity unpinned"); the converter is pinned by
round trips and by ``dequant(AutoGPTQ tensors) == dequant(q100)`` in tests/test_interop.py.  Zero-points the v1
format cannot hold (the reference's ``round_zero_point`` is only confined to [-1e4, 1e4]; v1 stores ``zero - 1`` in
b bits, i.e. zero in [1, 2^b]) raise ``ValueError``.

Format conversion is not the hot path: plain torch integer ops on whatever device the tensors live on.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops as _ops
from .ops import PackedWeight, TILE_K, TILE_ROWS

FORMAT_VERSION = 2


# ------------------------------------------------------------------------------------ q100 v2 layout in torch ops
def _pack_plane(u: torch.Tensor, bits: int) -> torch.Tensor:
    """[..., 64] unsigned codes (int64) -> [..., 64 * bits / 32] words (int64 holding uint32 values).  Element e of a
    word sits at bit bits * (e >> 1) + 16 * (e & 1) (DESIGN.md "Packed layout")."""
    F = 32 // bits
    e = u.reshape(*u.shape[:-1], u.shape[-1] // F, F)
    word = torch.zeros(e.shape[:-1], dtype=torch.int64, device=u.device)
    for j in range(F // 2):
        word |= e[..., 2 * j] << (bits * j)
        word |= e[..., 2 * j + 1] << (bits * (F // 2 + j))
    return word


def _unpack_plane(word: torch.Tensor, bits: int) -> torch.Tensor:
    F = 32 // bits
    mask = (1 << bits) - 1
    e = torch.empty((*word.shape, F), dtype=torch.int64, device=word.device)
    for j in range(F // 2):
        e[..., 2 * j] = (word >> (bits * j)) & mask
        e[..., 2 * j + 1] = (word >> (bits * (F // 2 + j))) & mask
    return e.reshape(*word.shape[:-1], word.shape[-1] * F)


def _as_i32(words: torch.Tensor) -> torch.Tensor:
    """int64 holding uint32 values -> int32 with the same bits."""
    return torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)


def _as_u32(words: torch.Tensor) -> torch.Tensor:
    return words.to(torch.int64) & 0xFFFFFFFF


def pack_codes(codes: torch.Tensor, n_bits: int) -> torch.Tensor:
    """Stored (unsigned, ``x_int - qmin``) codes ``[N, K]`` -> int32 ``[ceil(N/128), K/64, tile_words]``, the layout the
    quantize kernel writes (tests compare the two bit for bit)."""
    if n_bits not in (2, 3, 4, 8):
        raise ValueError("n_bits in {2, 3, 4, 8}")
    N, K = codes.shape
    if K % TILE_K:
        raise ValueError("K must be a multiple of 64")
    u = codes.to(torch.int64)
    if int(u.min()) < 0 or int(u.max()) >= (1 << n_bits):
        raise ValueError("codes outside [0, 2^n_bits)")
    NT, KT = (N + TILE_ROWS - 1) // TILE_ROWS, K // TILE_K
    up = torch.zeros((NT * TILE_ROWS, K), dtype=torch.int64, device=codes.device)
    up[:N] = u
    t = up.reshape(NT, TILE_ROWS, KT, TILE_K).permute(0, 2, 1, 3)           # [NT, KT, 128, 64]
    if n_bits == 3:
        lo = _pack_plane(t & 3, 2).reshape(NT, KT, -1)
        hi = _pack_plane(t >> 2, 1).reshape(NT, KT, -1)
        return _as_i32(torch.cat([lo, hi], dim=2)).contiguous()
    return _as_i32(_pack_plane(t, n_bits).reshape(NT, KT, -1)).contiguous()


def unpack_codes(packed: torch.Tensor, N: int, K: int, n_bits: int) -> torch.Tensor:
    """Inverse of :func:`pack_codes`: int64 ``[N, K]`` stored codes."""
    NT, KT = (N + TILE_ROWS - 1) // TILE_ROWS, K // TILE_K
    w = _as_u32(packed).reshape(NT, KT, -1)
    if n_bits == 3:
        lo = _unpack_plane(w[:, :, :TILE_ROWS * 4].reshape(NT, KT, TILE_ROWS, 4), 2)
        hi = _unpack_plane(w[:, :, TILE_ROWS * 4:].reshape(NT, KT, TILE_ROWS, 2), 1)
        t = lo | (hi << 2)
    else:
        t = _unpack_plane(w.reshape(NT, KT, TILE_ROWS, -1), n_bits)
    return t.permute(0, 2, 1, 3).reshape(NT * TILE_ROWS, K)[:N].contiguous()


def make_meta(scales: torch.Tensor, zeros: Optional[torch.Tensor], N: int, K: int, group: int) -> torch.Tensor:
    """Reference-layout ``scales`` / ``zeros`` ``[N*K/g, 1]`` -> the GEMM side band ``[K/g, N, 2]``."""
    kg = K // group
    s = scales.reshape(N, kg).t()
    z = torch.zeros_like(s) if zeros is None else zeros.reshape(N, kg).t()
    return torch.stack((s, z), dim=-1).contiguous()


# ------------------------------------------------------------------------------------ state dicts
def packed_state(pw: PackedWeight, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Tensors only (safetensors-friendly): codes, reference-layout scales / zeros and a small int64 header
    ``[version, n_bits, group, N, K, disable_zero_point]``.  The side band is rebuilt on load."""
    d = {prefix + "q100_qweight": pw.packed, prefix + "q100_scales": pw.scales,
         prefix + "q100_header": torch.tensor([FORMAT_VERSION, pw.n_bits, pw.group, pw.N, pw.K,
                                               int(pw.disable_zero_point)], dtype=torch.int64)}
    if pw.zeros is not None:
        d[prefix + "q100_zeros"] = pw.zeros
    return d


def packed_from_state(state: Dict[str, torch.Tensor], prefix: str = "", device=None) -> PackedWeight:
    hdr = [int(v) for v in state[prefix + "q100_header"]]
    if hdr[0] != FORMAT_VERSION:
        raise ValueError(f"q100 packed format version {hdr[0]} (this build reads {FORMAT_VERSION})")
    _, n_bits, group, N, K, nozp = hdr
    mv = (lambda t: t if device is None else t.to(device))
    packed, scales = mv(state[prefix + "q100_qweight"]), mv(state[prefix + "q100_scales"])
    zeros = state.get(prefix + "q100_zeros")
    zeros = None if zeros is None else mv(zeros)
    want = _ops.packed_shape(N, K, n_bits)
    if tuple(packed.shape) != tuple(want) or packed.dtype != torch.int32:
        raise ValueError(f"q100_qweight has shape {tuple(packed.shape)}, expected {tuple(want)} int32")
    if scales.numel() != N * (K // group):
        raise ValueError("q100_scales does not have one entry per group")
    return PackedWeight(packed.contiguous(), make_meta(scales, zeros, N, K, group), scales, zeros, n_bits, group,
                        N, K, bool(nozp))


def _quant_linears(module):
    for name, m in module.named_modules():
        if type(m).__name__ == "QuantLinear" and hasattr(m, "weight_quantizer"):
            yield name, m


def save_packed(module, path: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """State dict of every QuantLinear's packed twin under ``module`` (plus biases); written with ``torch.save`` when
    ``path`` is given.  Linears without a packed twin (never quantised, unsupported shape) are skipped."""
    out: Dict[str, torch.Tensor] = {}
    for name, m in _quant_linears(module):
        pw = m.weight_quantizer.packed_for(m.weight)
        if pw is None:
            continue
        out.update({k: v.detach().cpu() for k, v in packed_state(pw, name + ".").items()})
        if m.bias is not None:
            out[name + ".bias"] = m.bias.detach().cpu()
    if path is not None:
        torch.save(out, path)
    return out


def load_packed(module, state, device=None, release_dense: bool = True) -> int:
    """Attach the packed twins in ``state`` (a dict or a path) to the QuantLinears of ``module``: each becomes an
    inference-only layer that runs the fused dequant-GEMM straight from the codes (its dense ``weight`` is released
    unless ``release_dense=False``, in which case it is rebuilt by the dequantize kernel).  Returns the number of
    layers restored."""
    if isinstance(state, str):
        state = torch.load(state, map_location="cpu")
    n = 0
    for name, m in _quant_linears(module):
        if name + ".q100_header" not in state:
            continue
        dev = device if device is not None else m.weight.device
        pw = packed_from_state(state, name + ".", dev)
        if (pw.N, pw.K) != (m.out_features, m.in_features):
            raise ValueError(f"{name}: packed weight is [{pw.N}, {pw.K}], layer is [{m.out_features}, {m.in_features}]")
        wq = m.weight_quantizer
        wq._packed = pw
        if release_dense:
            m.weight = torch.empty(0, dtype=pw.dtype, device=dev)
        else:
            m.weight = _ops.dequantize(pw.packed, pw.scales, pw.zeros, pw.N, pw.K, pw.group, pw.n_bits, pw.dtype,
                                       pw.disable_zero_point)
        wq._packed_of = (m.weight.data_ptr(), m.weight._version, tuple(m.weight.shape), m.weight.dtype)
        wq.scale, wq.round_zero_point = pw.scales, pw.zeros
        if name + ".bias" in state:
            m.bias = state[name + ".bias"].to(dev)
        m.use_weight_quant, m.use_temporary_parameter = False, False
        n += 1
    return n


# ------------------------------------------------------------------------------------ AutoGPTQ v1
def _gptq_pack_rows(iw: torch.Tensor, bits: int) -> torch.Tensor:
    """[R, C] values in [0, 2^bits) -> int32 [R / 32 * bits, C], packing along dim 0 like AutoGPTQ's pack()."""
    R, C = iw.shape
    iw = iw.to(torch.int64)
    if bits in (2, 4, 8):
        per = 32 // bits
        if R % per:
            raise ValueError(f"dimension {R} is not a multiple of {per}")
        e = iw.reshape(R // per, per, C)
        word = torch.zeros((R // per, C), dtype=torch.int64, device=iw.device)
        for i in range(per):
            word |= e[:, i] << (bits * i)
        return _as_i32(word)
    if bits == 3:
        if R % 32:
            raise ValueError(f"dimension {R} is not a multiple of 32")
        e = iw.reshape(R // 32, 32, C)
        w0 = torch.zeros((R // 32, C), dtype=torch.int64, device=iw.device)
        w1, w2 = torch.zeros_like(w0), torch.zeros_like(w0)
        for j in range(10):
            w0 |= e[:, j] << (3 * j)
        w0 |= (e[:, 10] << 30) & 0xFFFFFFFF
        w1 |= (e[:, 10] >> 2) & 1
        for j in range(10):
            w1 |= e[:, 11 + j] << (3 * j + 1)
        w1 |= (e[:, 21] << 31) & 0xFFFFFFFF
        w2 |= (e[:, 21] >> 1) & 3
        for j in range(10):
            w2 |= e[:, 22 + j] << (3 * j + 2)
        return _as_i32(torch.stack((w0, w1, w2), dim=1).reshape(R // 32 * 3, C))
    raise ValueError("bits in {2, 3, 4, 8}")


def _gptq_unpack_rows(qw: torch.Tensor, bits: int) -> torch.Tensor:
    w = _as_u32(qw)
    Rw, C = w.shape
    if bits in (2, 4, 8):
        per, mask = 32 // bits, (1 << bits) - 1
        e = torch.stack([(w >> (bits * i)) & mask for i in range(per)], dim=1)
        return e.reshape(Rw * per, C)
    if bits == 3:
        w3 = w.reshape(Rw // 3, 3, C)
        w0, w1, w2 = w3[:, 0], w3[:, 1], w3[:, 2]
        out = []
        for j in range(10):
            out.append((w0 >> (3 * j)) & 7)
        out.append(((w0 >> 30) & 3) | ((w1 & 1) << 2))
        for j in range(10):
            out.append((w1 >> (3 * j + 1)) & 7)
        out.append(((w1 >> 31) & 1) | ((w2 & 3) << 1))
        for j in range(10):
            out.append((w2 >> (3 * j + 2)) & 7)
        return torch.stack(out, dim=1).reshape(Rw // 3 * 32, C)
    raise ValueError("bits in {2, 3, 4, 8}")


def to_autogptq(pw: PackedWeight) -> Dict[str, torch.Tensor]:
    """q100 v2 -> the tensors of an AutoGPTQ v1 ``QuantLinear`` (what ``pack(linear, scales, zeros)`` leaves,
    quant/omniquant.py:326-334).  Raises ``ValueError`` for zero-points outside [1, 2^bits] (see module docstring)."""
    b, g, N, K = pw.n_bits, pw.group, pw.N, pw.K
    codes = unpack_codes(pw.packed, N, K, b)                                 # stored codes, [N, K]
    kg = K // g
    if pw.zeros is None:
        zero = torch.full((N, kg), float(1 << (b - 1)), device=pw.scales.device)   # stored code = x_int + 2^(b-1)
    else:
        zero = pw.zeros.reshape(N, kg).float()
    if bool((zero != zero.round()).any()) or float(zero.min()) < 1 or float(zero.max()) > (1 << b):
        raise ValueError(
            f"zero-points in [{float(zero.min()):g}, {float(zero.max()):g}] do not fit AutoGPTQ v1, which stores "
            f"(zero - 1) in {b} bits (zero in [1, {1 << b}]); the reference only confines round_zero_point to "
            "[-1e4, 1e4] (quantizer.py:159)")
    if N % 32:
        raise ValueError("AutoGPTQ v1 packs qzeros along N in 32-value units: N must be a multiple of 32")
    return {"qweight": _gptq_pack_rows(codes.t().contiguous(), b),
            "qzeros": _gptq_pack_rows((zero - 1).to(torch.int64), b).t().contiguous(),     # packed along N
            "scales": pw.scales.reshape(N, kg).t().contiguous().to(torch.float16),
            "g_idx": (torch.arange(K, device=pw.packed.device) // g).to(torch.int32),
            "bits": torch.tensor(b), "group_size": torch.tensor(g)}


def from_autogptq(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                  g_idx: Optional[torch.Tensor], bits: int, group_size: int,
                  dtype: torch.dtype = torch.float16) -> PackedWeight:
    """AutoGPTQ v1 tensors -> a PackedWeight the fused dequant-GEMM runs on.  The dequantised weight is
    ``scales * (code - (stored_zero + 1))`` rounded like the reference's dequant (r(r(q - zp) * scale))."""
    K = qweight.shape[0] * 32 // bits
    N = qweight.shape[1]
    g = K if group_size in (-1, 0, None) else int(group_size)
    kg = K // g
    if g_idx is not None:
        want = torch.arange(K, device=g_idx.device) // g
        if not torch.equal(g_idx.to(torch.int64), want):
            raise ValueError("activation-reordered checkpoints (g_idx != k // group_size) are not supported")
    codes = _gptq_unpack_rows(qweight, bits)[:K].t().contiguous()            # [N, K]
    zero = (_gptq_unpack_rows(qzeros.t().contiguous(), bits)[:N].t().contiguous() + 1)   # [K/g, N]
    if tuple(scales.shape) != (kg, N) or tuple(zero.shape) != (kg, N):
        raise ValueError(f"scales / qzeros do not match K={K}, N={N}, group={g}")
    s_ref = scales.t().contiguous().reshape(-1, 1).to(dtype)
    z_ref = zero.t().contiguous().reshape(-1, 1).to(dtype)
    return PackedWeight(pack_codes(codes, bits), make_meta(s_ref, z_ref, N, K, g), s_ref, z_ref, bits, g, N, K, False)


def autogptq_dequant(t: Dict[str, torch.Tensor]) -> torch.Tensor:
    """What AutoGPTQ's v1 kernels compute from their tensors, in fp32: ``[N, K]`` weight (test helper and reference
    for importers)."""
    b, g = int(t["bits"]), int(t["group_size"])
    iw = _gptq_unpack_rows(t["qweight"], b).float()                           # [K, N]
    z = (_gptq_unpack_rows(t["qzeros"].t().contiguous(), b).t().contiguous() + 1).float()   # [K/g, N]
    gi = t["g_idx"].to(torch.int64)
    return (t["scales"].float()[gi] * (iw - z[gi])).t().contiguous()
