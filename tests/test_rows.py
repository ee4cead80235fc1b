"""GPU: K3r, the row-owner decode kernel (csrc/q100_gemm_rows.cu; W4 g128, M <= 2, K % 2048 == 0), through the C ABI.

K3r is what ``q100_gemm_wna16`` runs by default where it applies; ``Q100_GEMM_NO_ROWS`` selects the stream-K kernel
(K3a) for the same call.  Tolerance: the one stated in test_parity_gemm.py for the default decode path (the kernel
keeps (q - zp) exact and applies the scale in fp32 per group; x is rounded once to 2^-22 -- 2^-18 for the positions that
meet the high nibbles -- of the token's largest magnitude).
"""
import numpy as np
import pytest
import torch

import q100
from q100 import _lib
from conftest import TDT
from test_parity_gemm import _check, _make

pytestmark = pytest.mark.gpu


# N: one / several slabs per CTA, ranges that straddle a 128-row tile, N % 16 != 0, fewer row quads than CTAs;
# K: the 8B decode shapes (4096, 14336) and the smallest supported one
SHAPES = [(8, 2048), (136, 2048), (2368, 2048), (4096, 4096), (6144, 4096), (9472, 2048), (18952, 2048),
          (28672, 4096), (4096, 14336), (1024, 8192)]


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("N,K", SHAPES)
def test_rows_vs_oracle_and_streamk(dtype, N, K):
    for M, bias in ((1, False), (2, True)):
        x, w, pw, r = _make(N, K, M, dtype, 4, 128, N + K + M)
        b = (torch.randn(N) * 0.1).to(TDT[dtype]).cuda() if bias else None
        y = q100.gemm_wna16(x, pw, b)
        y_sk = q100.gemm_wna16(x, pw, b, flags=_lib.GEMM_NO_ROWS)
        torch.cuda.synchronize()
        _check(y, x, r.dequant, b, dtype, exact_weights=False)
        _check(y_sk, x, r.dequant, b, dtype, exact_weights=False)
        # the two kernels differ by fp32 summation order and the block / token exponent of x only
        d = (y.float() - y_sk.float()).abs().max().item()
        assert d <= 4 * (2.0 ** -8 if dtype == "bf16" else 2.0 ** -11) * max(y_sk.float().abs().max().item(), 1e-3)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_rows_symmetric_and_nozp(dtype):
    for sym, nozp in ((True, False), (True, True)):
        for M in (1, 2):
            x, w, pw, r = _make(512, 2048, M, dtype, 4, 128, 77 + M, nozp=nozp, sym=sym)
            y = q100.gemm_wna16(x, pw, None)
            _check(y, x, r.dequant, None, dtype, exact_weights=False)


def test_rows_deterministic_graph_and_pdl():
    x, w, pw, r = _make(4096, 4096, 2, "bf16", 4, 128, 5)
    y1 = q100.gemm_wna16(x, pw, None)
    y2 = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_PDL)
    torch.cuda.synchronize()
    assert torch.equal(y1.view(torch.int16), y2.view(torch.int16))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_PDL)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        y3 = q100.gemm_wna16(x, pw, None, flags=_lib.GEMM_PDL)
        y4 = q100.gemm_wna16(y3[:, :4096].contiguous(), pw, None, flags=_lib.GEMM_PDL)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y1.view(torch.int16), y3.view(torch.int16))
    y4_ref = q100.gemm_wna16(y1[:, :4096].contiguous(), pw, None)
    torch.cuda.synchronize()
    assert torch.equal(y4.view(torch.int16), y4_ref.view(torch.int16))


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("M", [1, 2])
def test_rows_gate_up_staging_equals_separate_silu_mul(dtype, M):
    """Q100_GEMM_X_GATE_UP on the row-owner kernel == K6 followed by the same GEMM, bit for bit."""
    N, K = 512, 4096
    gen = torch.Generator().manual_seed(131 + M)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(TDT[dtype]).cuda()
    gu = (torch.randn(M, 2 * K, generator=gen) * 2.0).to(TDT[dtype]).cuda()
    pw, _ = q100.pack_weight(w, 4, 128)
    y_ref = q100.gemm_wna16(q100.ops.silu_mul(gu), pw, None)
    y = q100.gemm_wna16(gu, pw, None, flags=_lib.GEMM_X_GATE_UP)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16))


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_rows_outliers_and_specials(dtype):
    """A token with one huge element (the per-token exponent is set by it), tiny tokens, inf / NaN poisoning."""
    N, K = 256, 2048
    x, w, pw, r = _make(N, K, 2, dtype, 4, 128, 909)
    x[0, 7] = 3000.0
    x[1] *= 1e-3
    y = q100.gemm_wna16(x, pw, None)
    _check(y, x, r.dequant, None, dtype, exact_weights=False)
    x2 = x.clone()
    x2[1, 100] = float("inf")
    y2 = q100.gemm_wna16(x2, pw, None)
    torch.cuda.synchronize()
    assert torch.equal(y2[0].view(torch.int16), y[0].view(torch.int16))
    assert not torch.isfinite(y2[1].float()).any()
    x3 = torch.zeros_like(x)
    y3 = q100.gemm_wna16(x3, pw, None)
    assert float(y3.float().abs().max()) == 0.0


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("M", [1, 2])
def test_block_glue_folded_into_the_gemm(dtype, M):
    """q100_gemm_wna16_block: RMSNorm in the x staging == K4 + GEMM up to the order of the fp32 sum of squares (at most
    one 16-bit ulp on a few staged elements); residual add in the epilogue == the add on the two 16-bit tensors, bit
    for bit; both together; in place."""
    N, K = 4096, 4096
    gen = torch.Generator().manual_seed(17 + M)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(TDT[dtype]).cuda()
    h = (torch.randn(M, K, generator=gen) * 1.5).to(TDT[dtype]).cuda()
    nw = (1.0 + 0.1 * torch.randn(K, generator=gen)).to(TDT[dtype]).cuda()
    res = torch.randn(M, N, generator=gen).to(TDT[dtype]).cuda()
    pw, _ = q100.pack_weight(w, 4, 128)
    assert q100.ops.gemm_block_fusable(M, pw)
    eps_out = 2.0 ** -8 if dtype == "bf16" else 2.0 ** -11
    n_ref = q100.ops.add_rmsnorm(h, nw, 1e-5)[1]
    y_ref = q100.gemm_wna16(n_ref, pw, None)
    y = q100.ops.gemm_wna16_block(h, pw, norm_weight=nw, eps=1e-5)
    torch.cuda.synchronize()
    assert (y.float() - y_ref.float()).abs().max().item() <= 4 * eps_out * y_ref.float().abs().max().item()
    # residual in the epilogue: same kernel, same accumulator bits, then the 16-bit add
    y2_ref = res + q100.gemm_wna16(n_ref, pw, None)
    y2 = q100.ops.gemm_wna16_block(n_ref, pw, residual=res)
    torch.cuda.synchronize()
    assert torch.equal(y2.view(torch.int16), y2_ref.view(torch.int16))
    # both, in place on the residual buffer, with a bias
    b = (torch.randn(N, generator=gen) * 0.1).to(TDT[dtype]).cuda()
    y3_ref = res + q100.gemm_wna16(n_ref, pw, b)
    buf = res.clone()
    y3 = q100.ops.gemm_wna16_block(h, pw, b, norm_weight=nw, eps=1e-5, residual=buf, out=buf)
    torch.cuda.synchronize()
    assert y3.data_ptr() == buf.data_ptr()
    assert (y3.float() - y3_ref.float()).abs().max().item() <= 4 * eps_out * max(y3_ref.float().abs().max().item(), 1.0)
    # not a K3r case: refused, not silently different
    pw3, _ = q100.pack_weight(w, 3, 128)
    assert not q100.ops.gemm_block_fusable(M, pw3)
    with pytest.raises(q100.Q100Error):
        q100.ops.gemm_wna16_block(h, pw3, norm_weight=nw, eps=1e-5)


def test_decoder_folded_route_matches_the_k4_route(monkeypatch):
    """GraphDecoder at batch 1 / 2: the folded route (4 GEMM launches with the norms and residual adds inside) against the
    K4 route on the same model: same greedy tokens, logits within 16-bit round-off."""
    from q100 import llama as QL
    dev = torch.device("cuda", 0)
    cfg = QL.LLAMA3_8B
    model = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, seed=0, n_layers=2,
                                           fast_attention=True)
    for B in (1, 2):
        outs = {}
        for fold in (True, False):
            if fold:
                monkeypatch.delenv("Q100_NO_FOLD", raising=False)
            else:
                monkeypatch.setenv("Q100_NO_FOLD", "1")
            dec = QL.GraphDecoder(model, B, 64)
            dec.prime(8)
            tok = torch.arange(B, device=dev).view(B, 1) + 11
            toks, logits = [], []
            for i in range(4):
                toks.append(dec.step(tok if i == 0 else None).clone())
                logits.append(dec.last_logits.float().clone())
            dec.check()
            outs[fold] = (toks, logits)
        for la, lb in zip(outs[True][1], outs[False][1]):
            assert (la - lb).abs().max().item() <= 0.02 * max(lb.abs().max().item(), 1.0)
