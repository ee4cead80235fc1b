"""GPU: the tcgen05 path (K3b).  First the TMA / UMMA / TMEM plumbing alone on a dense
16-bit weight, then the fused dequant-GEMM against the oracle, same tolerance as the
skinny kernel (tests/test_parity_gemm.py)."""
import numpy as np
import pytest
import torch

import q100
import q100_oracle as O
from q100 import _lib, ops
from conftest import TDT, torch_to_bits
from test_parity_gemm import _check, _make

pytestmark = pytest.mark.gpu
# two prefill kernels: K3b' (weights as the TMEM A operand, default) and K3b (weights staged in shared memory)
KERN = {"tmem_a": _lib.GEMM_FORCE_TC, "smem_b": _lib.GEMM_FORCE_TC | _lib.GEMM_FORCE_TC_SMEM}


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 64), (128, 256, 256), (256, 512, 1024),
                                   (200, 384, 512), (2048, 1024, 4096)])
def test_dense_plumbing(dtype, M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=gen).to(TDT[dtype]).cuda()
    w = (torch.randn(N, K, generator=gen) * 0.05).to(TDT[dtype]).cuda()
    y = ops.debug_gemm_dense_tc(x, w)
    torch.cuda.synchronize()
    _check(y, x, w, None, dtype)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,N,K,group", [(128, 128, 128, 128), (128, 256, 128, 128), (64, 256, 512, 128),
                                         (200, 384, 1024, 128), (2048, 1024, 4096, 128), (256, 512, 512, 64),
                                         (256, 512, 1024, 256), (256, 512, 1024, 0), (2048, 4096, 4096, 128)])
@pytest.mark.parametrize("kern", sorted(KERN))
def test_tc_vs_oracle(dtype, M, N, K, group, kern):
    x, w, pw, r = _make(N, K, M, dtype, 4, group or None, M + N + K)
    b = (torch.randn(N) * 0.1).to(TDT[dtype]).cuda()
    y = q100.gemm_wna16(x, pw, b, flags=KERN[kern])
    torch.cuda.synchronize()
    _check(y, x, r.dequant, b, dtype)
    if N * K <= 1 << 20:       # oracle-side check that the fake-quantised weight itself is right
        o = O.quantize_groupwise(O.from_bits(torch_to_bits(w), dtype).reshape(N, K), dtype, group, 4)
        assert np.array_equal(torch_to_bits(r.dequant), O.to_bits(o["dequant"], dtype))


def test_tc_matches_skinny_bits_of_weight():
    """Both GEMM paths rebuild the same weight: on an identity-like x they must agree exactly."""
    N, K = 256, 256
    gen = torch.Generator().manual_seed(9)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(torch.bfloat16).cuda()
    pw, r = q100.pack_weight(w, 4, 128, want_dequant=True)
    x = torch.eye(K, dtype=torch.bfloat16, device="cuda")          # y = dequant(W)^T exactly
    for kern in sorted(KERN):
        y_tc = q100.gemm_wna16(x, pw, None, flags=KERN[kern])
        assert torch.equal(y_tc.t().contiguous().view(torch.int16), r.dequant.view(torch.int16)), kern
    y_sk = q100.gemm_wna16(x[:32].contiguous(), pw, None, flags=_lib.GEMM_FORCE_SKINNY)
    assert torch.equal(y_sk.view(torch.int16), y_tc[:32].contiguous().view(torch.int16))


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("n_bits", [2, 3, 8])
@pytest.mark.parametrize("kern", sorted(KERN))
def test_tc_other_bits(dtype, n_bits, kern):
    if n_bits == 3 and kern != "tmem_a":
        pytest.skip("W3 (2-bit + 1-bit planes) runs on the TMEM-A kernel only")
    for (M, N, K, group) in ((128, 128, 128, 128), (300, 384, 1024, 128), (512, 256, 512, 0)):
        x, w, pw, r = _make(N, K, M, dtype, n_bits, group or None, M + N + K + n_bits)
        y = q100.gemm_wna16(x, pw, None, flags=KERN[kern])
        torch.cuda.synchronize()
        _check(y, x, r.dequant, None, dtype)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,N,K", [(33, 136, 256), (129, 392, 512), (700, 1000, 1024), (2048, 6144, 4096)])
def test_tmem_a_ragged_shapes(dtype, M, N, K):
    """N not a multiple of the 128-row tile, M not a multiple of the token tile (TMEM-A kernel only)."""
    x, w, pw, r = _make(N, K, M, dtype, 4, 128, M + N + K)
    b = (torch.randn(N) * 0.1).to(TDT[dtype]).cuda()
    y = q100.gemm_wna16(x, pw, b, flags=_lib.GEMM_FORCE_TC)
    torch.cuda.synchronize()
    _check(y, x, r.dequant, b, dtype)


def test_tc_weird_zero_points_and_nozp():
    """Slow dequant path (zero-points far outside the code range) and the no-zero-point mode."""
    torch.manual_seed(3)
    w = torch.randn(128, 256) * 0.02
    w[0:8] = w[0:8].abs() + 0.5
    w[8:16] = -(w[8:16].abs() + 0.5)
    w[16:24] = 0.3
    w[24:32] = 0.0
    for dt in (torch.float16, torch.bfloat16):
        wd = w.to(dt).cuda()
        x = torch.randn(64, 256).to(dt).cuda()
        for kern in sorted(KERN):
            pw, r = q100.pack_weight(wd, 4, 128, want_dequant=True)
            y = q100.gemm_wna16(x, pw, None, flags=KERN[kern])
            _check(y, x, r.dequant, None, "f16" if dt == torch.float16 else "bf16")
            pw, r = q100.pack_weight(wd, 4, 128, True, True, want_dequant=True)
            y = q100.gemm_wna16(x, pw, None, flags=KERN[kern])
            _check(y, x, r.dequant, None, "f16" if dt == torch.float16 else "bf16")


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_tca_w3_bits_of_weight(dtype):
    """W3 on the prefill kernel: with x = identity the output is the fake-quantised weight itself, bit for bit."""
    N, K = 256, 256
    gen = torch.Generator().manual_seed(11)
    w = (torch.randn(N, K, generator=gen) * 0.02).to(TDT[dtype]).cuda()
    pw, r = q100.pack_weight(w, 3, 128, want_dequant=True)
    x = torch.eye(K, dtype=TDT[dtype], device="cuda")
    y = q100.gemm_wna16(x, pw, None)
    assert torch.equal(y.t().contiguous().view(torch.int16), r.dequant.view(torch.int16))
