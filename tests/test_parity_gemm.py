"""GPU: fused dequant-GEMM (K3) through the C ABI against the oracle.

Tolerance (stated here, used everywhere).  ``ref`` = float64 matmul of x with the ORACLE's
fake-quantised weight (the reference's F.linear input), eps_out = 2^-11 (fp16) / 2^-8 (bf16)
the output rounding, eps_acc = 2^-24 the fp32 accumulator rounding:

* exact-weight paths (tcgen05 kernel; decode kernel with Q100_GEMM_EXACT_WEIGHTS or group < 128)
  rebuild r(r(q-zp)*scale) bit-identically, so only the accumulation order differs:
      |y - ref| <= 2*eps_out*|ref| + 8*eps_acc*sum_k|x_k*w_k|
* the default decode kernel keeps (q-zp) exact and applies the scale in fp32 per group, i.e. it
  does NOT round the weight to 16 bits (it is closer to the real-valued product than the
  reference).  Its distance to ``ref`` is the reference's own weight-rounding noise, a sum of K
  independent terms each bounded by eps_w*|x_k*w_k| with eps_w = 2^-12 (fp16) / 2^-9 (bf16):
      |y - ref| <= 2*eps_out*|ref| + 8*eps_w*sqrt(sum_k (x_k*w_k)^2) + 8*eps_acc*sum_k|x_k*w_k|
"""
import numpy as np
import pytest
import torch

import q100
import q100_oracle as O
from q100 import _lib
from conftest import TDT, bits_to_torch, load_golden, torch_to_bits

pytestmark = pytest.mark.gpu
EPS_OUT = {"f16": 2.0 ** -11, "bf16": 2.0 ** -8}
EPS_W = {"f16": 2.0 ** -12, "bf16": 2.0 ** -9}
# decode shapes have two kernels: the tcgen05 streaming kernel (M <= 16, W2/W4, group % 128 == 0; default
# for 8 < M <= 16) and the mma.sync streaming kernel (everything else); both are forced here in turn
# the mma.sync kernel itself has two operand forms for W4 at M <= 8: unsigned-byte codes x signed-byte limbs of x
# on the integer tensor-core path (default) and the fp16 magic form (GEMM_NO_IMMA)
PATHS = {"stream_tc": _lib.GEMM_FORCE_STREAM_TC, "mma_sync": _lib.GEMM_FORCE_MMA_SYNC,
         "mma_sync_f16": _lib.GEMM_FORCE_MMA_SYNC | _lib.GEMM_NO_IMMA,
         # W4 g128, M <= 2, K % 2048 == 0 run the row-owner kernel K3r by default (tests/test_rows.py); this path keeps
         # the stream-K kernel K3a covered on those shapes
         "mma_sync_streamk": _lib.GEMM_FORCE_MMA_SYNC | _lib.GEMM_NO_ROWS}


def _check(y, x, wdeq, bias, dtype, exact_weights=True):
    x64 = x.double().cpu().numpy().reshape(-1, x.shape[-1])
    w64 = wdeq.double().cpu().numpy()
    ref = x64 @ w64.T
    mag = np.abs(x64) @ np.abs(w64).T
    if bias is not None:
        ref = ref + bias.double().cpu().numpy()
        mag = mag + np.abs(bias.double().cpu().numpy())
    got = y.double().cpu().numpy().reshape(ref.shape)
    tol = 2 * EPS_OUT[dtype] * np.abs(ref) + 8 * 2.0 ** -24 * mag + 1e-7
    if not exact_weights:
        tol = tol + 8 * EPS_W[dtype] * np.sqrt((x64 ** 2) @ (w64 ** 2).T)
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{int(bad.sum())}/{bad.size} outside tolerance, max abs err {np.abs(got - ref).max():.3e}"
    return float(np.abs(got - ref).max()), float(np.abs(ref).max())


def _make(N, K, M, dtype, n_bits, group, seed, nozp=False, sym=False):
    gen = torch.Generator().manual_seed(seed)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(TDT[dtype]).cuda()
    x = torch.randn(M, K, generator=gen).to(TDT[dtype]).cuda()
    pw, r = q100.pack_weight(w, n_bits, group, sym, nozp, want_dequant=True)
    return x, w, pw, r


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("n_bits", [4, 2, 3, 8])
@pytest.mark.parametrize("M", [1, 2, 5, 8, 9, 16, 17, 32])
def test_skinny_vs_oracle(dtype, n_bits, M, path):
    if path == "stream_tc" and (n_bits in (3, 8) or M > 16):
        pytest.skip("same kernel as mma_sync for this case")
    if path == "mma_sync_f16" and (n_bits not in (2, 4) or M > 2):
        pytest.skip("same kernel as mma_sync for this case")
    if path == "mma_sync_streamk":
        pytest.skip("same kernel as mma_sync for this case (K = 1024)")
    N, K = 384, 1024
    x, w, pw, r = _make(N, K, M, dtype, n_bits, 128, 1000 + M + n_bits)
    # the packed twin is consistent with the oracle (bit-exact)
    o = O.quantize_groupwise(O.from_bits(torch_to_bits(w), dtype).reshape(N, K), dtype, 128, n_bits)
    assert np.array_equal(pw.packed.cpu().numpy().view(np.uint32), O.pack_codes(o["codes"], n_bits))
    wdeq = torch.from_numpy(o["dequant"])
    y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
    torch.cuda.synchronize()
    _check(y, x, wdeq, None, dtype, exact_weights=False)
    y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | _lib.GEMM_EXACT_WEIGHTS)
    torch.cuda.synchronize()
    _check(y, x, wdeq, None, dtype, exact_weights=True)


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("N,K,group", [(128, 256, 128), (136, 512, 128), (1024, 4096, 128), (4096, 4096, 128),
                                       (512, 14336, 128), (256, 512, 64), (256, 512, 32), (256, 1024, 256),
                                       (256, 1024, 0), (6144, 4096, 128)])
def test_skinny_shapes(dtype, N, K, group, path):
    if path == "stream_tc" and group in (32, 64):
        pytest.skip("same kernel as mma_sync for this case")
    if path == "mma_sync_streamk" and (group != 128 or K % 2048):
        pytest.skip("same kernel as mma_sync for this case")
    for M, bias in ((1, False), (3, True), (16, True)):
        x, w, pw, r = _make(N, K, M, dtype, 4, group or None, N + K + M)
        b = (torch.randn(N) * 0.1).to(TDT[dtype]).cuda() if bias else None
        y = q100.gemm_wna16(x, pw, b, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
        torch.cuda.synchronize()
        _check(y, x, r.dequant, b, dtype, exact_weights=(group in (32, 64)))


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_skinny_symmetric_and_nozp(dtype, path):
    for n_bits in (2, 3, 4, 8):
        for sym, nozp in ((True, False), (True, True)):
            x, w, pw, r = _make(256, 512, 4, dtype, n_bits, 128, 77 + n_bits, nozp=nozp, sym=sym)
            y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
            _check(y, x, r.dequant, None, dtype, exact_weights=False)
            y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | _lib.GEMM_EXACT_WEIGHTS)
            _check(y, x, r.dequant, None, dtype, exact_weights=True)


@pytest.mark.parametrize("path", sorted(PATHS))
def test_skinny_deterministic_and_workspace_clean(path):
    x, w, pw, r = _make(4096, 4096, 16, "bf16", 4, 128, 5)
    y1 = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
    y2 = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
    torch.cuda.synchronize()
    assert torch.equal(y1.view(torch.int16), y2.view(torch.int16))
    from q100 import ops
    for ws in ops._WORKSPACES.values():
        n_tiles_bytes = 4096 // 128 * 4
        assert int(ws[:n_tiles_bytes].abs().sum()) == 0, "tile counters must return to zero"


def test_weird_zero_points():
    """Zero-points outside [0, 2^n-1] (all-positive / all-negative / constant groups)."""
    torch.manual_seed(3)
    w = torch.randn(128, 256) * 0.02
    w[0:8] = w[0:8].abs() + 0.5
    w[8:16] = -(w[8:16].abs() + 0.5)
    w[16:24] = 0.3
    w[24:32] = 0.0
    w = w.half().cuda()
    x = torch.randn(3, 256).half().cuda()
    pw, r = q100.pack_weight(w, 4, 128, want_dequant=True)
    for path in sorted(PATHS):
        y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | PATHS[path])
        _check(y, x, r.dequant, None, "f16", exact_weights=False)
    y = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_FORCE_SKINNY | _lib.GEMM_EXACT_WEIGHTS)
    _check(y, x, r.dequant, None, "f16", exact_weights=True)


L = load_golden("linear")


@pytest.mark.parametrize("case", sorted(c for c in L if "f32" not in c and "a16" in c))
def test_golden_linear_through_module(case):
    """QuantLinear drop-in (weight re-quantised each call, int_linear.py:52-53) vs the
    reference's own output stored in the fixture."""
    import torch.nn as nn
    from quant.int_linear import QuantLinear
    rec = L[case]
    dtype = str(rec["dtype"])
    w = bits_to_torch(rec["w"], dtype)
    lin = nn.Linear(w.shape[1], w.shape[0], bias=True).to(TDT[dtype])
    with torch.no_grad():
        lin.weight.copy_(w)
        lin.bias.copy_(bits_to_torch(rec["bias"], dtype))
    lin = lin.cuda()
    ql = QuantLinear(lin, {"n_bits": int(rec["w_bits"]), "per_channel_axes": [0], "symmetric": False,
                           "dynamic_method": "per_channel", "group_size": int(rec["group"]) or None,
                           "lwc": False, "disable_zero_point": False},
                     {"n_bits": int(rec["a_bits"]), "per_channel_axes": [], "symmetric": False,
                      "dynamic_method": "per_token"})
    ql.set_quant_state(True, False)
    x = bits_to_torch(rec["x"], dtype, "cuda")
    with torch.no_grad():
        y = ql(x)
    yr = bits_to_torch(rec["y"], dtype).float()
    # N = 48 is a multiple of 8 and K = 512 of 128 => the fused kernel ran
    assert ql._requant_cache is not None
    err = float((y.float().cpu() - yr).abs().max())
    tol = 4 * EPS_OUT[dtype] * max(float(yr.abs().max()), 1.0)
    assert err <= tol, (err, tol)
    # reference side effects: scale / round_zero_point live on the quantizer
    assert ql.weight_quantizer.scale.shape == (w.shape[0] * (w.shape[1] // (int(rec["group"]) or w.shape[1])), 1)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("n_bits", [4, 2, 3])
@pytest.mark.parametrize("M", [1, 2, 5, 8])
def test_gate_up_staging_equals_separate_silu_mul(dtype, n_bits, M):
    """Q100_GEMM_X_GATE_UP: down_proj with act_fn(gate) * up applied while x is staged == K6 followed by the same
    GEMM, bit for bit (same kernel, same x bits)."""
    N, K = 256, 1024
    gen = torch.Generator().manual_seed(31 + M + n_bits)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(TDT[dtype]).cuda()
    gu = (torch.randn(M, 2 * K, generator=gen) * 2.0).to(TDT[dtype]).cuda()
    pw, _ = q100.pack_weight(w, n_bits, 128)
    y_ref = q100.gemm_wna16(q100.ops.silu_mul(gu), pw, None, flags=_lib.GEMM_FORCE_SKINNY)
    y = q100.gemm_wna16(gu, pw, None, flags=_lib.GEMM_FORCE_SKINNY | _lib.GEMM_X_GATE_UP)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16))
    with pytest.raises(q100.Q100Error):
        q100.gemm_wna16(torch.zeros(64, 2 * K, dtype=TDT[dtype], device="cuda"), pw, None, flags=_lib.GEMM_X_GATE_UP)
