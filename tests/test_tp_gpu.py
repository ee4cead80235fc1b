"""GPU: tensor-parallel arithmetic of the path on real hardware.

One GPU is enough for the first group: the ranks' shards are computed one after the other in this process through the
same C-ABI calls a rank makes, and the collectives are replaced by their definition (element-wise min / max of the
[rows, 2] statistics, an fp32 sum of the partial outputs in rank order -- what K7 and the chain kernel's peer epilogue
compute).  That pins what sharding does to the bits: activation scales / zero-points / fake-quantised values and
per-channel weight codes equal the single-GPU ones exactly, outputs agree within the GEMM tolerance.

The second group launches real ranks with torchrun (NCCL + peer memory) and needs >= 2 GPUs (skipped otherwise):
tools/tp_parity_check.py compares a tensor-parallel block stack with the single-GPU one on the same weights.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import q100
import q100_oracle as O
from q100 import ops, tp
from conftest import TDT, torch_to_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_minmax(parts):
    mm = [ops.rows_minmax(p) for p in parts]
    return torch.stack((torch.stack([m[:, 0] for m in mm]).amin(0), torch.stack([m[:, 1] for m in mm]).amax(0)), dim=1)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("n_bits", [8, 4])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("rows,K", [(1, 4096), (16, 14336), (2048, 4096), (300, 28672)])
def test_sharded_activation_quant_bit_exact(dtype, n_bits, world, rows, K):
    """abits < 16 in front of a row-parallel linear: every rank quantises its K / world slice with the all-reduced
    row statistics; the slices concatenated are the single-GPU fake-quantised activation, bit for bit, and so are the
    per-token scales and zero-points (quantizer.py:140-159)."""
    gen = torch.Generator(device="cuda").manual_seed(rows + K + world)
    x = torch.randn(rows, K, generator=gen, device="cuda").to(TDT[dtype])
    whole = ops.fakequant_rows(x, n_bits)
    parts = [tp.shard_input(x, tp.ROW, r, world) for r in range(world)]
    mm = _global_minmax(parts)
    assert torch.equal(mm[:, 0], x.amin(-1)) and torch.equal(mm[:, 1], x.amax(-1))
    outs = [ops.quantize_rows_stats(p, mm, n_bits) for p in parts]
    got = torch.cat([o.dequant for o in outs], dim=1)
    assert torch.equal(got.view(torch.int16), whole.dequant.view(torch.int16))
    for o in outs:
        assert torch.equal(o.scales.view(torch.int16), whole.scales.view(torch.int16))
        assert torch.equal(o.zeros.view(torch.int16), whole.zeros.view(torch.int16))
    if rows <= 16:        # and the single-GPU result is the oracle's
        oo = O.fakequant_rows(O.from_bits(torch_to_bits(x), dtype).reshape(rows, K), dtype, n_bits)
        assert np.array_equal(torch_to_bits(whole.dequant), O.to_bits(oo["dequant"], dtype))


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("n_bits,abits", [(8, 8), (4, 4)])
@pytest.mark.parametrize("world", [2, 8])
def test_per_channel_row_parallel_linear(dtype, n_bits, abits, world):
    """SURVEY 8d config 4 under tensor parallelism: per-channel weights (group_size=None) AND per-token activations of a
    row-parallel layer (o_proj shape).  Codes of the weight shards equal the slices of the unsharded codes; the fp32
    sum of the ranks' partial GEMMs matches the single-GPU GEMM within the GEMM tolerance."""
    N, K, M = 4096, 4096, 16
    gen = torch.Generator(device="cuda").manual_seed(world * 10 + n_bits)
    w = (torch.randn(N, K, generator=gen, device="cuda") * 0.02).to(TDT[dtype])
    x = torch.randn(M, K, generator=gen, device="cuda").to(TDT[dtype])
    pw_full, r_full = q100.pack_weight(w, n_bits, None, want_dequant=True)
    xq_full = ops.fakequant_rows(x, abits).dequant
    y_full = q100.gemm_wna16(xq_full, pw_full)
    wparts = [tp.shard_weight(w, tp.ROW, r, world, 0) for r in range(world)]
    xparts = [tp.shard_input(x, tp.ROW, r, world) for r in range(world)]
    wmm, xmm = _global_minmax(wparts), _global_minmax(xparts)
    acc = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    part_mag = torch.zeros(M, N, dtype=torch.float64, device="cuda")
    for r in range(world):
        rw = ops.quantize_rows_stats(wparts[r], wmm, n_bits, want_dequant=True, want_packed=True, want_meta=True)
        k0 = r * (K // world)
        assert torch.equal(rw.dequant.view(torch.int16), r_full.dequant[:, k0:k0 + K // world].contiguous().view(torch.int16))
        assert torch.equal(rw.scales.view(torch.int16), r_full.scales.view(torch.int16))
        pw = ops.PackedWeight(rw.packed, rw.meta, rw.scales, rw.zeros, n_bits, K // world, N, K // world, False)
        xq = ops.quantize_rows_stats(xparts[r], xmm, abits).dequant
        assert torch.equal(xq.view(torch.int16), xq_full[:, k0:k0 + K // world].contiguous().view(torch.int16))
        acc += q100.gemm_wna16(xq, pw).float()                # the all-reduce: fp32 sum of 16-bit partials, rank order
        part_mag += (xq.double() @ rw.dequant.double().T).abs()
    x64, w64 = xq_full.double(), r_full.dequant.double()
    ref = x64 @ w64.T
    eps_out = {"f16": 2.0 ** -11, "bf16": 2.0 ** -8}[dtype]
    eps_w = {"f16": 2.0 ** -12, "bf16": 2.0 ** -9}[dtype]
    # tests/test_parity_gemm.py tolerance of the default decode kernel, plus the 16-bit rounding of every rank's partial
    base = 8 * 2.0 ** -24 * (x64.abs() @ w64.abs().T) + 8 * eps_w * torch.sqrt((x64 ** 2) @ (w64 ** 2).T) + 1e-7
    assert bool(((y_full.double() - ref).abs() <= 2 * eps_out * ref.abs() + base).all())
    assert bool(((acc.double() - ref).abs() <= 2 * eps_out * part_mag + base).all())


def test_quantizer_module_ignores_stat_group_without_dist():
    """A quantizer with stat_group set but no process group initialised behaves like the single-GPU one."""
    from quant.quantizer import UniformAffineQuantizer
    q = UniformAffineQuantizer(n_bits=8, per_channel_axes=[], symmetric=False, dynamic_method="per_token")
    x = torch.randn(4, 1, 4096, device="cuda").half()
    a = q(x)
    q.stat_group = object()
    b = q(x)
    assert torch.equal(a, b)


def _torchrun(script, nproc, port, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("cfg", ["w4a16", "w8a8", "w4a4"])
def test_tp2_block_stack_matches_single_gpu(cfg):
    r = _torchrun("tp_parity_check.py", 2, 29547, "--config", cfg)
    assert r.returncode == 0 and "TP PARITY OK" in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]


def test_gemm_on_second_device_after_first():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device property: a process that used cuda:0 must
    still be able to launch every GEMM kernel on cuda:1 (the reference's parallel_utils layer placement is such a
    process)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            w = (torch.randn(512, 1024, device=dev) * 0.02).half()
            pw, r = q100.pack_weight(w, 4, 128, want_dequant=True)
            for M in (1, 12, 160):
                x = torch.randn(M, 1024, device=dev).half()
                y = q100.gemm_wna16(x, pw)
                ref = x.double() @ r.dequant.double().T
                assert float((y.double() - ref).abs().max()) <= 2 ** -9 * max(float(ref.abs().max()), 1.0)
