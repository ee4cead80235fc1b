"""GPU: K1/K2 through the C ABI against the golden fixtures (reference outputs) and the
numpy oracle.  Everything here is bit-exact: dequantised values, scales, zero-points,
integer codes and the packed words."""
import numpy as np
import pytest
import torch

import q100
import q100_oracle as O
from q100 import _lib
from conftest import TDT, assert_same_bits, bits_to_torch, load_golden, torch_to_bits

pytestmark = pytest.mark.gpu

W = load_golden("weights")
A = load_golden("activations")
MODES = {"packed": 0, "generic": _lib.FLAG_K1_GENERIC, "scalar": _lib.FLAG_SCALAR_MATH, "ieee": _lib.FLAG_IEEE_DIV}


def _codes_from_packed(packed, rows, cols, n_bits, nozp):
    return O.unpack_codes(packed.cpu().numpy().view(np.uint32), rows, cols, n_bits, nozp)


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("case", sorted(W))
def test_golden_weights(case, mode):
    rec = W[case]
    dtype = str(rec["dtype"])
    if dtype == "f32" and mode != "packed":
        pytest.skip("fp32 has a single (IEEE) math path")
    n_bits, group = int(rec["n_bits"]), int(rec["group"])
    sym, nozp, lwc, d = bool(rec["symmetric"]), bool(rec["disable_zp"]), bool(rec["lwc"]), int(rec["deficiency"])
    w = bits_to_torch(rec["w"], dtype, "cuda")
    if d:
        w = torch.cat([w, torch.zeros(w.shape[0], d, dtype=w.dtype, device="cuda")], dim=1)
    up = bits_to_torch(rec["lwc_up"], dtype, "cuda") if lwc else None
    low = bits_to_torch(rec["lwc_low"], dtype, "cuda") if lwc else None
    cols = w.shape[1]
    g = group or cols
    can_pack = cols % 64 == 0 and (g in (32, 64, 128, 256) or g == cols)
    can_meta = can_pack and dtype != "f32"
    r = q100.quantize_groupwise(w, group, n_bits, sym, nozp, up, low, want_dequant=True,
                                want_packed=can_pack, want_meta=can_meta, flags=MODES[mode])
    torch.cuda.synchronize()
    deq = r.dequant[:, : rec["w"].shape[1]]
    assert_same_bits(torch_to_bits(deq), rec["dequant"], dtype, "dequant")
    assert_same_bits(torch_to_bits(r.scales), rec["scales"], dtype, "scales")
    if nozp:
        assert r.zeros is None
    else:
        assert_same_bits(torch_to_bits(r.zeros), rec["zeros"], dtype, "zeros")
    if can_pack:
        codes = _codes_from_packed(r.packed, w.shape[0], cols, n_bits, nozp)
        want = rec["codes"].reshape(codes.shape).astype(np.float32)
        qmin, _ = O.qrange(n_bits, nozp)
        want = np.where(want == -32768, qmin, want)            # NaN code is stored as qmin
        assert np.array_equal(codes, want), "packed codes"
        # byte-identical to the oracle's packer
        assert np.array_equal(r.packed.cpu().numpy().view(np.uint32), O.pack_codes(want, n_bits, nozp))
    if can_meta:
        meta = r.meta.cpu()
        rows = w.shape[0]
        s = r.scales.cpu().view(rows, cols // g)
        assert torch.equal(meta[:, :, 0].view(torch.int16), s.t().contiguous().view(torch.int16))
        if not nozp:
            z = r.zeros.cpu().view(rows, cols // g)
            assert torch.equal(meta[:, :, 1].view(torch.int16), z.t().contiguous().view(torch.int16))


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("case", sorted(A))
def test_golden_activations(case, mode):
    rec = A[case]
    dtype = str(rec["dtype"])
    if dtype == "f32" and mode != "packed":
        pytest.skip("fp32 has a single (IEEE) math path")
    x = bits_to_torch(rec["x"], dtype, "cuda")
    r = q100.fakequant_rows(x, int(rec["n_bits"]), flags=MODES[mode])
    torch.cuda.synchronize()
    assert_same_bits(torch_to_bits(r.dequant), rec["dequant"], dtype, "dequant")
    assert_same_bits(torch_to_bits(r.scales), rec["scales"], dtype, "scales")
    assert_same_bits(torch_to_bits(r.zeros), rec["zeros"], dtype, "zeros")


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_fast_division_exhaustive(dtype):
    """rcp + 2 FMA == div.rn.f32 for every (x, scale) pair the quantizer can see."""
    assert q100.selftest_div(TDT[dtype]) == 0


@pytest.mark.parametrize("dtype", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("n_bits", [2, 3, 4, 8])
@pytest.mark.parametrize("shape,group", [((256, 1024), 128), ((96, 4096), 128), ((64, 2048), 64), ((33, 512), 32),
                                         ((40, 1024), 256), ((48, 4096), 0), ((5, 14336), 0), ((130, 128), 0),
                                         ((17, 768), 0), ((3, 28672), 0), ((9, 1000), 0)])
def test_random_vs_oracle(dtype, n_bits, shape, group):
    """Fresh seeded inputs at sizes the oracle handles in well under a second."""
    import zlib
    gen = torch.Generator().manual_seed(zlib.crc32(repr((dtype, n_bits, shape, group)).encode()))
    x = (torch.randn(*shape, generator=gen) * 0.05).to(TDT[dtype])
    # sprinkle exact ties and a few large values
    x[0, :8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 3.0, 0.0, -0.0]).to(TDT[dtype]) * 0.01
    xn = O.from_bits(torch_to_bits(x), dtype).reshape(shape)
    o = O.quantize_groupwise(xn, dtype, group, n_bits)
    cols = shape[1]
    g = group or cols
    esz = 4 if dtype == "f32" else 2
    if (g * esz) % 16 != 0:
        with pytest.raises(q100.Q100Error):
            q100.quantize_groupwise(x.cuda(), group, n_bits)
        return
    can_pack = cols % 64 == 0 and (g in (32, 64, 128, 256) or g == cols)
    r = q100.quantize_groupwise(x.cuda(), group, n_bits, want_packed=can_pack)
    torch.cuda.synchronize()
    assert_same_bits(torch_to_bits(r.dequant), O.to_bits(o["dequant"], dtype), dtype, "dequant")
    assert_same_bits(torch_to_bits(r.scales), O.to_bits(o["scales"], dtype), dtype, "scales")
    assert_same_bits(torch_to_bits(r.zeros), O.to_bits(o["zeros"], dtype), dtype, "zeros")
    if can_pack:
        assert np.array_equal(r.packed.cpu().numpy().view(np.uint32), O.pack_codes(o["codes"], n_bits))
        # dequantize(packed) reproduces the fake-quantised tensor bit for bit
        back = q100.dequantize(r.packed, r.scales, r.zeros, shape[0], cols, group, n_bits, TDT[dtype])
        assert_same_bits(torch_to_bits(back), torch_to_bits(r.dequant), dtype, "dequantize(pack(x))")
        raw = q100.dequantize(r.packed, None, None, shape[0], cols, group, n_bits, torch.float32)
        assert np.array_equal(raw.cpu().numpy(), o["codes"])


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_symmetric_and_nozp_vs_oracle(dtype):
    gen = torch.Generator().manual_seed(11)
    x = (torch.randn(64, 1024, generator=gen) * 0.03).to(TDT[dtype])
    xn = O.from_bits(torch_to_bits(x), dtype).reshape(64, 1024)
    for n_bits in (2, 3, 4, 8):
        for sym, nozp in ((True, False), (True, True), (False, True)):
            o = O.quantize_groupwise(xn, dtype, 128, n_bits, sym, nozp)
            r = q100.quantize_groupwise(x.cuda(), 128, n_bits, sym, nozp, want_packed=True)
            assert_same_bits(torch_to_bits(r.dequant), O.to_bits(o["dequant"], dtype), dtype, f"deq w{n_bits} {sym} {nozp}")
            assert_same_bits(torch_to_bits(r.scales), O.to_bits(o["scales"], dtype), dtype, "scales")
            assert np.array_equal(r.packed.cpu().numpy().view(np.uint32), O.pack_codes(o["codes"], n_bits, nozp))
            back = q100.dequantize(r.packed, r.scales, r.zeros, 64, 1024, 128, n_bits, TDT[dtype], nozp)
            assert_same_bits(torch_to_bits(back), torch_to_bits(r.dequant), dtype, "dequantize")


def test_empty_and_ragged():
    x = torch.empty(0, 128, dtype=torch.float16, device="cuda")
    r = q100.quantize_groupwise(x, 128, 4)
    assert r.dequant.shape == (0, 128) and r.scales.shape == (0, 1)
    # 1 group, 1 row
    x = torch.randn(1, 128, dtype=torch.float16, device="cuda")
    o = O.quantize_groupwise(x.float().cpu().numpy(), "f16", 128, 4)
    r = q100.quantize_groupwise(x, 128, 4)
    assert_same_bits(torch_to_bits(r.dequant), O.to_bits(o["dequant"], "f16"), "f16", "1 group")
    # group count not a multiple of the groups-per-warp
    x = torch.randn(7, 384, dtype=torch.bfloat16, device="cuda")
    o = O.quantize_groupwise(x.float().cpu().numpy(), "bf16", 128, 3)
    r = q100.quantize_groupwise(x, 128, 3, want_packed=True)
    assert_same_bits(torch_to_bits(r.dequant), O.to_bits(o["dequant"], "bf16"), "bf16", "ragged")
    assert np.array_equal(r.packed.cpu().numpy().view(np.uint32), O.pack_codes(o["codes"], 3))
    with pytest.raises(ValueError):
        q100.quantize_groupwise(torch.randn(4, 100, device="cuda"), 128, 4)


@pytest.mark.parametrize("dtype", ["f16", "bf16", "f32"])
def test_full_size_properties(dtype):
    """BASELINE config shape (LLaMA3-8B up_proj 14336 x 4096, W4 g128): size-independent
    properties instead of a 58 M-element oracle run, plus an oracle check on a row sample."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    w = (torch.randn(14336, 4096, generator=gen, device="cuda") * 0.02).to(TDT[dtype])
    r = q100.quantize_groupwise(w, 128, 4, want_packed=True, want_meta=(dtype != "f32"))
    torch.cuda.synchronize()
    # (1) decode(pack) == fake-quant output, bit for bit
    back = q100.dequantize(r.packed, r.scales, r.zeros, 14336, 4096, 128, 4, TDT[dtype])
    assert torch.equal(back.view(torch.int32 if dtype == "f32" else torch.int16),
                       r.dequant.view(torch.int32 if dtype == "f32" else torch.int16))
    # (2) all three math paths agree on every bit
    for flag in (_lib.FLAG_SCALAR_MATH, _lib.FLAG_IEEE_DIV):
        r2 = q100.quantize_groupwise(w, 128, 4, want_packed=True, flags=flag)
        assert torch.equal(r2.packed, r.packed) and torch.equal(r2.scales, r.scales) and torch.equal(r2.zeros, r.zeros)
        assert torch.equal(r2.dequant.view(torch.int16 if dtype != "f32" else torch.int32),
                           r.dequant.view(torch.int16 if dtype != "f32" else torch.int32))
    # (3) codes span the range and the error is bounded by scale/2 (+ storage rounding)
    codes = q100.dequantize(r.packed, None, None, 14336, 4096, 128, 4, torch.float32)
    assert float(codes.min()) == 0.0 and float(codes.max()) == 15.0
    err = (r.dequant.float() - w.float()).abs().view(-1, 128)
    assert bool((err <= r.scales.float() * 1.01 + 1e-6).all())   # <= scale/2 rounding + <= scale/2 from the rounded zero-point
    # (4) row-sharding invariance: quantising a slice == slicing the result
    sl = slice(1024, 1024 + 512)
    r3 = q100.quantize_groupwise(w[sl].contiguous(), 128, 4, want_packed=True)
    assert torch.equal(r3.packed, r.packed[1024 // 128:(1024 + 512) // 128]) and \
        torch.equal(r3.scales.view(512, -1), r.scales.view(14336, -1)[sl])
    # (5) K-sharding invariance on multiples of the group
    r4 = q100.quantize_groupwise(w[:, 2048:].contiguous(), 128, 4)
    assert torch.equal(r4.dequant.view(torch.int16 if dtype != "f32" else torch.int32),
                       r.dequant[:, 2048:].contiguous().view(torch.int16 if dtype != "f32" else torch.int32))
    # (6) oracle on a sample of rows
    idx = torch.tensor([0, 1, 777, 4095, 8191, 14335], device="cuda")
    ws = w[idx].contiguous()
    o = O.quantize_groupwise(O.from_bits(torch_to_bits(ws), dtype).reshape(6, 4096), dtype, 128, 4)
    assert_same_bits(torch_to_bits(r.dequant[idx]), O.to_bits(o["dequant"], dtype), dtype, "sample rows")
    codes_s = q100.dequantize(r.packed, None, None, 14336, 4096, 128, 4, torch.float32)[idx]
    assert np.array_equal(codes_s.cpu().numpy(), o["codes"])


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("n_bits", [2, 3, 4, 8])
@pytest.mark.parametrize("group", [64, 128, 256])
def test_k1_fast_vs_generic_stress(dtype, n_bits, group):
    """k1_fast (default) against k1_groups (FLAG_K1_GENERIC) and the scalar IEEE path on rows built to flip
    between the exact-integer fast step and the generic step: magnitudes from subnormal to overflow, one-sided
    groups (large zero-points), constant groups, NaN / inf, and symmetric / no-zero-point / LWC variants."""
    g = torch.Generator(device="cpu").manual_seed(1234 + n_bits + group)
    rows, cols = 384, 1024
    x = torch.randn(rows, cols, generator=g)
    mag = torch.exp2(torch.randint(-26, 16, (rows, cols // group, 1), generator=g).float())
    off = torch.randn(rows, cols // group, 1, generator=g) * torch.exp2(torch.randint(-4, 6, (rows, cols // group, 1), generator=g).float())
    x = (x.view(rows, cols // group, group) * mag + off * mag * (torch.rand(rows, cols // group, 1, generator=g) < 0.4)).view(rows, cols)
    x[5, :group] = 3.0                                   # constant group
    x[6, :group] = 0.0
    x[7, 3] = float("nan")
    x[8, 70] = float("inf")
    x[9, 130] = float("-inf")
    x[10, :group] = -0.0
    x = x.to(TDT[dtype]).cuda()
    up = (torch.rand(rows * cols // group, 1, generator=g) * 0.5 + 0.5).to(TDT[dtype]).cuda()
    low = (torch.rand(rows * cols // group, 1, generator=g) * 0.5 + 0.5).to(TDT[dtype]).cuda()
    for sym, nozp, lwc in ((False, False, False), (True, False, False), (False, True, False), (True, True, False),
                           (False, False, True)):
        outs = []
        for flags in (0, _lib.FLAG_K1_GENERIC, _lib.FLAG_IEEE_DIV):
            r = q100.quantize_groupwise(x, group, n_bits, sym, nozp, up if lwc else None, low if lwc else None,
                                        want_dequant=True, want_packed=True, want_meta=True, flags=flags)
            torch.cuda.synchronize()
            outs.append(r)
        a = outs[0]
        for b in outs[1:]:
            tag = f"sym={sym} nozp={nozp} lwc={lwc}"
            assert torch.equal(a.dequant.view(torch.int16), b.dequant.view(torch.int16)), "dequant " + tag
            assert torch.equal(a.scales.view(torch.int16), b.scales.view(torch.int16)), "scales " + tag
            if not nozp:
                assert torch.equal(a.zeros.view(torch.int16), b.zeros.view(torch.int16)), "zeros " + tag
            assert torch.equal(a.packed, b.packed), "packed " + tag
            assert torch.equal(a.meta.view(torch.int16), b.meta.view(torch.int16)), "meta " + tag


@pytest.mark.parametrize("dtype", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("n_bits", [2, 3, 4, 8])
@pytest.mark.parametrize("group", [32, 64, 128, 256])
def test_backward_kernel_vs_op_by_op(dtype, n_bits, group):
    """K1b (q100_quantize_backward) against quant.quantizer.fake_quant_backward -- the op-by-op form that
    tests/test_host_logic.py pins to torch autograd through the reference op sequence.  The STE mask comes from
    the same per-op arithmetic as the forward (exact); the gradient arithmetic is fp32 in the kernel and
    per-op-rounded in torch, hence the dtype-dependent tolerance."""
    from quant.quantizer import fake_quant_backward
    gen = torch.Generator(device="cpu").manual_seed(77 + n_bits + group)
    rows, cols = 96, 1024
    x = (torch.randn(rows, cols, generator=gen) * 0.05).to(TDT[dtype]).cuda()
    x[3, :group] = 0.25                                     # constant group: every element is both min and max
    x[4, 5] = x[4, 6] = x[4, :group].max()                  # tied maxima
    x[6, group + 3] = float("nan")                          # NaN group: every gradient of the group is NaN
    go = torch.randn(rows, cols, generator=gen).to(TDT[dtype]).cuda()
    ng = rows * cols // group
    su = (torch.rand(ng, 1, generator=gen) * 0.4 + 0.6).to(TDT[dtype]).cuda()
    sl = (torch.rand(ng, 1, generator=gen) * 0.4 + 0.6).to(TDT[dtype]).cuda()
    tol = {"f32": 2e-5, "f16": 4e-3, "bf16": 3e-2}[dtype]
    for sym, nozp, lwc in ((False, False, False), (False, False, True), (True, False, True), (False, True, False),
                           (True, True, True)):
        up, low = (su, sl) if lwc else (None, None)
        r = q100.quantize_groupwise(x, group, n_bits, sym, nozp, up, low)
        qmin, qmax = O.qrange(n_bits, nozp)
        want = fake_quant_backward(x.reshape(-1, group), go.reshape(-1, group), r.scales, r.zeros, up, low,
                                   n_bits, sym, qmin, qmax, True, lwc)
        f32 = lambda t: None if t is None else t.float()                                    # noqa: E731
        want32 = fake_quant_backward(x.float().reshape(-1, group), go.float().reshape(-1, group), f32(r.scales),
                                     f32(r.zeros), f32(up), f32(low), n_bits, sym, qmin, qmax, True, lwc)
        got = q100.ops.quantize_backward(x, go, group, n_bits, sym, nozp, up, low, True, lwc)
        torch.cuda.synchronize()
        tag = f"sym={sym} nozp={nozp} lwc={lwc}"
        for name, g_, w_ in (("grad_x", got[0], want[0].reshape(rows, cols)), ("grad_up", got[1], want[1]),
                             ("grad_low", got[2], want[2])):
            if w_ is None:
                assert g_ is None, name + " " + tag
                continue
            g64, w64 = g_.double(), w_.double()
            # NaN: a NaN in a group makes every gradient of that group NaN (autograd through the reference op sequence
            # does the same).  The 16-bit op-by-op form can ALSO lose a finite group to overflow (a clamped constant
            # group sums gout * 1e4 past 65504: inf - inf between the max and the min share); the kernel accumulates
            # in fp32 and keeps those -- the pattern it must reproduce is the one of the form evaluated in fp32.
            n32 = {"grad_x": 0, "grad_up": 1, "grad_low": 2}[name]
            nan_ref = want32[n32].reshape(w64.shape).isnan()
            assert torch.equal(g64.isnan(), nan_ref), f"{name} {tag}: NaN pattern differs"
            assert bool((w64.isnan() | ~nan_ref).all()), f"{name} {tag}: the 16-bit form is finite where fp32 is NaN"
            skip = w64.isnan()
            g64, w64 = torch.where(skip, 0.0, g64), torch.where(skip, 0.0, w64)
            err = (g64 - w64).abs()
            # 16-bit: torch rounds every product of the group sum to the storage type before adding, the kernel keeps
            # fp32 -- the reference's own noise scales with the largest term, not with the (cancelling) sum
            ref_mag = w64.abs().mean() if dtype == "f32" else w64.abs().max()
            bound = tol * (w64.abs() + ref_mag + 1e-6)
            bad = err > bound
            assert bad.sum().item() <= (0 if dtype == "f32" else bad.numel() // 1000), \
                f"{name} {tag}: {int(bad.sum())}/{bad.numel()} outside tolerance, max err {err.max().item():.3e}"


def test_backward_through_module_matches_autograd_of_port():
    """End to end: UniformAffineQuantizer (LWC on) under autograd uses K1b; compare with torch autograd through
    the torch restatement of the reference op sequence (oracle/torch_port.py)."""
    import torch_port as TPORT
    from quant.quantizer import UniformAffineQuantizer
    torch.manual_seed(5)
    w = (torch.randn(64, 512) * 0.03).cuda().requires_grad_(True)
    q = UniformAffineQuantizer(n_bits=4, per_channel_axes=[0], symmetric=False, dynamic_method="per_channel",
                               group_size=128, lwc=True, shape=w.shape).cuda()
    go = torch.randn(64, 512, device="cuda")
    out = q(w)
    out.backward(go)
    gw, gu, gl = w.grad.clone(), q.upbound_factor.grad.clone(), q.lowbound_factor.grad.clone()
    w2 = w.detach().clone().requires_grad_(True)
    up = q.upbound_factor.detach().clone().requires_grad_(True)
    low = q.lowbound_factor.detach().clone().requires_grad_(True)
    ref, _, _ = TPORT.fake_quant(w2, 4, 128, False, False, up, low)
    ref.backward(go)
    for a, b, nm in ((gw, w2.grad, "weight"), (gu, up.grad, "upbound"), (gl, low.grad, "lowbound")):
        assert torch.allclose(a, b, rtol=2e-4, atol=1e-6 + 1e-4 * b.abs().mean().item()), \
            f"{nm}: max err {(a - b).abs().max().item():.3e}"
