"""GPU: the dequant-GEMM kernels at EVERY linear shape BASELINE.json names -- the seven LLaMA3-8B and LLaMA3-70B
projections plus the two decode-time fusions (q|k|v, gate|up) -- at M in {1, 4, 8, 9, 16, 2048} and W2 / W3 / W4,
g128, through the C ABI (q100_gemm_wna16), with the default kernel choice (what QuantLinear.forward runs) and, at
decode sizes, every forced kernel that applies.

Reference contraction: a float64 matmul ON THE GPU of x with the fake-quantised weight the quantize kernel returned
(the CPU float64 product of a 2048 x 28672 x 8192 case would take minutes); that weight is itself pinned against the
numpy oracle bit for bit, here for the 8B shapes and in tests/test_parity_quant.py for everything small.  Tolerances
are the ones stated in tests/test_parity_gemm.py with one addition for the long contractions that only appear here:
the fp32-accumulator term is max(8, K / 1024) * 2^-24 * sum_k |x_k w_k| -- a K = 28672 chain (70B down_proj) makes
1792 accumulator updates per output, and the tensor core's accumulator rounding error grows with that count (first
measured here: 8 of 16.7 M outputs of 70b_down at M = 2048 fell outside the factor-8 bound used for K <= 8192;
the worst-case bound of any fp32 summation is (K - 1) * 2^-24).
The file sorts first so that it is collected before anything slow.
"""
import zlib

import numpy as np
import pytest
import torch

import q100
import q100_oracle as O
from q100 import _lib
from conftest import TDT, torch_to_bits

pytestmark = pytest.mark.gpu
EPS_OUT = {"f16": 2.0 ** -11, "bf16": 2.0 ** -8}
EPS_W = {"f16": 2.0 ** -12, "bf16": 2.0 ** -9}

SHAPES = {
    # name: (N, K)
    "8b_q_o": (4096, 4096), "8b_k_v": (1024, 4096), "8b_gate_up": (14336, 4096), "8b_down": (4096, 14336),
    "8b_qkv_fused": (6144, 4096), "8b_gateup_fused": (28672, 4096),
    "70b_q_o": (8192, 8192), "70b_k_v": (1024, 8192), "70b_gate_up": (28672, 8192), "70b_down": (8192, 28672),
    "70b_qkv_fused": (10240, 8192), "70b_gateup_fused": (57344, 8192),
    # the per-GPU shards of config 5 (70B, TP = 8): column shards of q / gate|up, row shards of o / down
    "70b_tp8_qkv": (1280, 8192), "70b_tp8_o": (8192, 1024), "70b_tp8_gateup": (7168, 8192), "70b_tp8_down": (8192, 3584),
}
DECODE_M = (1, 4, 8, 9, 16)
_CACHE = {}


def _weights(name, n_bits, dtype):
    """(PackedWeight, dequantised weight as float64 on the GPU) for a shape; cached for the M sweep."""
    key = (name, n_bits, dtype)
    if key not in _CACHE:
        _CACHE.clear()                                    # one big weight at a time
        torch.cuda.empty_cache()
        N, K = SHAPES[name]
        gen = torch.Generator(device="cuda").manual_seed(zlib.crc32(repr(key).encode()) & 0x3fffffff)
        w = (torch.randn(N, K, generator=gen, device="cuda") * 0.02).to(TDT[dtype])
        pw, r = q100.pack_weight(w, n_bits, 128, want_dequant=True)
        if name.startswith("8b") and n_bits == 4 and dtype == "f16":
            # the fake-quantised weight, scales and zero-points are the oracle's, bit for bit, at the real size
            o = O.quantize_groupwise(O.from_bits(torch_to_bits(w), dtype).reshape(N, K), dtype, 128, n_bits)
            assert np.array_equal(torch_to_bits(r.dequant), O.to_bits(o["dequant"], dtype)), "dequant != oracle"
            assert np.array_equal(torch_to_bits(r.scales), O.to_bits(o["scales"], dtype)), "scales != oracle"
            assert np.array_equal(torch_to_bits(r.zeros), O.to_bits(o["zeros"], dtype)), "zeros != oracle"
            assert np.array_equal(pw.packed.cpu().numpy().view(np.uint32), O.pack_codes(o["codes"], n_bits)), "codes"
        _CACHE[key] = (pw, r.dequant.double())
        del w, r
    return _CACHE[key]


def _check(y, x, w64, dtype, exact_weights):
    x64 = x.double()
    ref = x64 @ w64.T
    mag = x64.abs() @ w64.abs().T
    tol = 2 * EPS_OUT[dtype] * ref.abs() + max(8.0, x.shape[-1] / 1024.0) * 2.0 ** -24 * mag + 1e-7
    if not exact_weights:
        tol = tol + 8 * EPS_W[dtype] * torch.sqrt((x64 ** 2) @ (w64 ** 2).T)
    err = (y.double() - ref).abs()
    bad = err > tol
    assert not bool(bad.any()), (f"{int(bad.sum())}/{bad.numel()} outside tolerance, max abs err {float(err.max()):.3e} "
                                 f"(|ref| max {float(ref.abs().max()):.3e})")


@pytest.mark.parametrize("M", DECODE_M)
@pytest.mark.parametrize("n_bits", [4, 2, 3])
@pytest.mark.parametrize("name", sorted(SHAPES))
def test_decode_shapes(name, n_bits, M):
    """Decode sizes: the default kernel (integer mma.sync at M <= 2, fp16 mma.sync up to 8, tcgen05 streaming from 9)
    and every other kernel that can take the case."""
    dtype = "f16"
    pw, w64 = _weights(name, n_bits, dtype)
    gen = torch.Generator(device="cuda").manual_seed(M * 131 + n_bits)
    x = torch.randn(M, pw.K, generator=gen, device="cuda").to(TDT[dtype])
    paths = {"default": 0, "mma_sync": _lib.GEMM_FORCE_MMA_SYNC}
    if n_bits in (2, 4):
        paths["stream_tc"] = _lib.GEMM_FORCE_STREAM_TC
        if M <= 2:
            paths["mma_sync_f16"] = _lib.GEMM_FORCE_MMA_SYNC | _lib.GEMM_NO_IMMA
            if n_bits == 4:                       # default at these shapes: the row-owner kernel K3r; keep K3a covered
                paths["mma_sync_streamk"] = _lib.GEMM_FORCE_MMA_SYNC | _lib.GEMM_NO_ROWS
    for pname, fl in paths.items():
        y = q100.gemm_wna16(x, pw, None, flags=fl)
        torch.cuda.synchronize()
        try:
            _check(y, x, w64, dtype, exact_weights=False)
        except AssertionError as e:
            raise AssertionError(f"{name} W{n_bits} M={M} path={pname}: {e}") from None


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("n_bits", [4, 2, 3])
@pytest.mark.parametrize("name", sorted(SHAPES))
def test_prefill_shapes(name, n_bits, dtype):
    """M = 2048 (BASELINE configs[1]): the tcgen05 prefill kernel rebuilds the 16-bit fake-quantised weight exactly, so
    only the accumulation order differs from F.linear."""
    if dtype == "f16" and n_bits != 4 and not name.startswith("8b"):
        pytest.skip("70B shapes at W2 / W3: bf16 only (same kernel instantiation, keeps the sweep short)")
    pw, w64 = _weights(name, n_bits, dtype)
    gen = torch.Generator(device="cuda").manual_seed(2048 + n_bits)
    x = torch.randn(2048, pw.K, generator=gen, device="cuda").to(TDT[dtype])
    y = q100.gemm_wna16(x, pw, None)
    torch.cuda.synchronize()
    _check(y, x, w64, dtype, exact_weights=True)
