"""GPU: K3c, q100_gemm_wna8 -- W8A8 / W4A4 (SmoothQuant-style configurations, main.py:326-349) at decode shapes: the
per-token activation quantisation fused into the GEMM as its prologue, weight codes x activation codes on the integer
tensor cores.

Two references, both built from the oracle-pinned quantize kernels (K1 weights, K2 activations; bit-exact vs the
oracle in tests/test_parity_quant.py):
  exact     y = sum_g s_w s_x sum_k (q_w - z_w)(q_x - z_x) in float64 from the integer codes: what the kernel computes,
            so it must agree to fp32 accumulation round-off -- this pins the activation CODES the prologue produces
            (they are the reference's: one wrong code moves an output by a whole quantisation step);
  reference F.linear of the two fake-quantised 16-bit tensors (quant/int_linear.py:59-62) in float64: the reference
            rounds (q - z) s to 16 bits on both sides before multiplying, the kernel does not; tolerance = the GEMM
            tolerance of tests/test_parity_gemm.py with the rounding term taken for both operands.
"""
import numpy as np
import pytest
import torch

import q100
from q100 import interop, ops
from conftest import TDT

pytestmark = pytest.mark.gpu
EPS_OUT = {"f16": 2.0 ** -11, "bf16": 2.0 ** -8}
EPS_OP = {"f16": 2.0 ** -12, "bf16": 2.0 ** -9}


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("wbits,abits,group", [(8, 8, None), (4, 4, None), (4, 4, 128), (8, 8, 128), (4, 8, 128), (8, 4, None)])
@pytest.mark.parametrize("M", [1, 2, 5, 7, 8, 12, 15])
@pytest.mark.parametrize("N,K", [(512, 1024), (4096, 4096), (1024, 14336)])
def test_gemm_wna8(dtype, wbits, abits, group, M, N, K):
    if (N, K) != (512, 1024) and (dtype == "bf16" or M not in (1, 7, 15)):
        pytest.skip("large shapes: fp16, M in {1, 7, 15}")
    dt = TDT[dtype]
    g = torch.Generator(device="cuda").manual_seed(N + K + M + wbits * 7 + abits)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(dt)
    x = (torch.randn(M, K, generator=g, device="cuda") * (1 + torch.rand(M, 1, generator=g, device="cuda") * 3)).to(dt)
    x[:, ::97] *= 6.0                                     # outlier channels, like the activations SmoothQuant targets
    pw, rw = q100.pack_weight(w, wbits, group, want_dequant=True)
    xq = ops.fakequant_rows(x, abits)
    y = ops.gemm_wna8(x, pw, abits)
    torch.cuda.synchronize()
    # ---- exact integer reference from the codes
    gK = K if group is None else group
    qw = interop.unpack_codes(pw.packed, N, K, wbits).double()
    sw = pw.scales.double().view(N, K // gK).repeat_interleave(gK, dim=1)
    zw = pw.zeros.double().view(N, K // gK).repeat_interleave(gK, dim=1)
    sx, zx = xq.scales.double().view(M, 1), xq.zeros.double().view(M, 1)
    # the activation codes K2 chose, taken from its packed output (never recovered as round(deq / s) + z: that is
    # ambiguous in bf16 at 8 bits)
    rq = ops.quantize_groupwise(x, 0, abits, want_dequant=True, want_packed=True)
    assert torch.equal(rq.dequant.view(torch.int16), xq.dequant.view(torch.int16))
    qx = interop.unpack_codes(rq.packed, M, K, abits).double()
    assert float(qx.min()) >= 0 and float(qx.max()) <= 2 ** abits - 1
    wd, xd = (qw - zw) * sw, (qx - zx) * sx
    exact = xd @ wd.T
    mag = xd.abs() @ wd.abs().T
    err = (y.double() - exact).abs()
    tol = 2 * EPS_OUT[dtype] * exact.abs() + 16 * 2.0 ** -24 * mag + 1e-6
    assert bool((err <= tol).all()), f"vs exact integer product: max err {float(err.max()):.3e}, {int((err > tol).sum())} outside"
    # ---- the reference's product of the two fake-quantised 16-bit tensors
    ref = xq.dequant.double() @ rw.dequant.double().T
    tol2 = 2 * EPS_OUT[dtype] * ref.abs() + 16 * EPS_OP[dtype] * torch.sqrt((xd ** 2) @ (wd ** 2).T) + 16 * 2.0 ** -24 * mag + 1e-6
    err2 = (y.double() - ref).abs()
    assert bool((err2 <= tol2).all()), f"vs fake-quant F.linear: max err {float(err2.max()):.3e}, {int((err2 > tol2).sum())} outside"


def test_quantlinear_takes_the_fused_path_at_decode_shapes():
    """QuantLinear.forward with act quant on: rows <= 15 run K3c, larger inputs the reference's two steps; both agree."""
    from quant.int_linear import QuantLinear
    lin = torch.nn.Linear(1024, 512, bias=False, device="cuda", dtype=torch.float16)
    ql = QuantLinear(lin, {"n_bits": 8, "per_channel_axes": [0], "symmetric": False, "dynamic_method": "per_channel",
                           "group_size": None, "lwc": False, "disable_zero_point": False},
                     {"n_bits": 8, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"})
    with torch.no_grad():
        ql.weight = ql.weight_quantizer(ql.weight)
        ql.set_quant_state(False, True)
        x = torch.randn(2, 4, 1024, device="cuda").half()
        y_fused = ql(x)                                   # 8 rows: K3c
        ql.use_fused_gemm = False
        y_ref = ql(x)                                     # act fake-quant + F.linear on the dense fake-quantised weight
    assert y_fused.shape == y_ref.shape
    assert float((y_fused.float() - y_ref.float()).abs().max()) <= 2e-2 * float(y_ref.float().abs().max())


# ------------------------------------------------------------------------------------------------ K3d (prefill sizes)
@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K", [(16, 512, 1024), (100, 520, 1024), (257, 6144, 4096), (2048, 4096, 4096),
                                   (2048, 14336, 4096), (2048, 4096, 14336), (4100, 1024, 8192)])
def test_gemm_w8a8_prefill(dtype, M, N, K):
    """K3d, q100_gemm_w8a8: 8-bit codes x 8-bit codes on tcgen05 kind::i8 (per-channel weights, per-token activations).
    Same two references as K3c above; the token rows include one-signed rows (zero-points outside [0, 255])."""
    if dtype == "bf16" and M * N * K > 2048 * 4096 * 4096:
        pytest.skip("largest shapes: fp16 only")
    dt = TDT[dtype]
    g = torch.Generator(device="cuda").manual_seed(N + K + M)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(dt)
    x = (torch.randn(M, K, generator=g, device="cuda") * (1 + torch.rand(M, 1, generator=g, device="cuda") * 3)).to(dt)
    x[:, ::97] *= 6.0
    x[3] = x[3].abs() + 2.0                               # zero-point < 0
    x[5] = -x[5].abs() - 8.0                              # zero-point > 255
    w[7] = w[7].abs() + 0.02
    bias = (torch.randn(N, generator=g, device="cuda") * 0.1).to(dt) if N == 520 else None
    pw, rw = q100.pack_weight(w, 8, None, want_dequant=True)
    xq = ops.fakequant_rows(x, 8)
    y = ops.gemm_w8a8(x, pw, bias=bias)
    torch.cuda.synchronize()
    qw = interop.unpack_codes(pw.packed, N, K, 8).double()
    assert torch.equal(ops.code_rowsums(pw.packed, N, K).long(), qw.sum(1).long())
    sw, zw = pw.scales.double().view(N, 1), pw.zeros.double().view(N, 1)
    sx, zx = xq.scales.double().view(M, 1), xq.zeros.double().view(M, 1)
    rq = ops.quantize_groupwise(x, 0, 8, want_dequant=True, want_packed=True)
    assert torch.equal(rq.dequant.view(torch.int16), xq.dequant.view(torch.int16))
    qx = interop.unpack_codes(rq.packed, M, K, 8).double()
    wd, xd = (qw - zw) * sw, (qx - zx) * sx
    exact = xd @ wd.T
    mag = xd.abs() @ wd.abs().T
    if bias is not None:
        exact = exact + bias.double()
        mag = mag + bias.double().abs()
    err = (y.double() - exact).abs()
    # the integer part is exact; what is left is one fp32 product of the scales, one fma and the rounding to 16 bits
    tol = 1.01 * EPS_OUT[dtype] * exact.abs() + 4 * 2.0 ** -24 * mag + 1e-7
    assert bool((err <= tol).all()), f"vs exact integer product: max err {float(err.max()):.3e}, {int((err > tol).sum())} outside"
    ref = xq.dequant.double() @ rw.dequant.double().T
    if bias is not None:
        ref = ref + bias.double()
    tol2 = 2 * EPS_OUT[dtype] * ref.abs() + 16 * EPS_OP[dtype] * torch.sqrt((xd ** 2) @ (wd ** 2).T) + 16 * 2.0 ** -24 * mag + 1e-6
    err2 = (y.double() - ref).abs()
    assert bool((err2 <= tol2).all()), f"vs fake-quant F.linear: max err {float(err2.max()):.3e}, {int((err2 > tol2).sum())} outside"


def test_gemm_w8a8_external_stats_match_single_pass():
    """Row statistics supplied from outside (the tensor-parallel route: MIN / MAX all-reduced) give the same bits."""
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(1024, 2048, generator=g, device="cuda") * 0.02).half()
    x = torch.randn(300, 2048, generator=g, device="cuda").half()
    pw, _ = q100.pack_weight(w, 8, None)
    y0 = ops.gemm_w8a8(x, pw)
    y1 = ops.gemm_w8a8(x, pw, minmax=ops.rows_minmax(x))
    assert torch.equal(y0, y1)


def test_quantlinear_takes_k3d_at_prefill_shapes():
    """QuantLinear.forward with W8A8 per-channel and more than 15 rows runs K3d; agrees with the reference's two steps."""
    from quant.int_linear import QuantLinear
    lin = torch.nn.Linear(1024, 512, bias=True, device="cuda", dtype=torch.float16)
    ql = QuantLinear(lin, {"n_bits": 8, "per_channel_axes": [0], "symmetric": False, "dynamic_method": "per_channel",
                           "group_size": None, "lwc": False, "disable_zero_point": False},
                     {"n_bits": 8, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"})
    calls = []
    orig = ops.gemm_w8a8
    with torch.no_grad():
        ql.weight = ql.weight_quantizer(ql.weight)
        ql.set_quant_state(False, True)
        x = torch.randn(3, 50, 1024, device="cuda").half()
        try:
            ops.gemm_w8a8 = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
            y_fused = ql(x)
        finally:
            ops.gemm_w8a8 = orig
        ql.use_fused_gemm = False
        y_ref = ql(x)
    assert calls, "K3d was not taken"
    assert y_fused.shape == y_ref.shape
    assert float((y_fused.float() - y_ref.float()).abs().max()) <= 2e-2 * float(y_ref.float().abs().max())
