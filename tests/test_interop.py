"""Packed checkpoint format and AutoGPTQ v1 interop.

CPU: the torch restatement of the q100 v2 layout equals the oracle's, AutoGPTQ v1 packing round-trips for 2/3/4/8 bits
(including the 3-bit 32-values-in-3-words pattern) and its dequantisation formula reproduces the fake-quantised weight.
GPU: the kernel's packed words equal interop.pack_codes, a weight exported to AutoGPTQ tensors and imported back runs
through the fused dequant-GEMM with identical results, and a QuantLinear tree survives save_packed / load_packed.
AutoGPTQ itself is not installed (and not vendored by the reference): parity with ITS kernels is unpinned, the layout
is restated from its published pack() (see q100/interop.py).
"""
import numpy as np
import pytest
import torch

import q100_oracle as O
from q100 import interop


def _quantized(N, K, bits, group, seed, positive_zero=True):
    rng = np.random.RandomState(seed)
    w = (rng.randn(N, K) * 0.02).astype(np.float16).astype(np.float32)
    if positive_zero:
        # AutoGPTQ v1 holds zero in [1, 2^b]: make every group straddle zero with min < -scale
        w[:, ::2] = -np.abs(w[:, ::2]) - 0.01
        w[:, 1::2] = np.abs(w[:, 1::2]) + 0.01
    return w, O.quantize_groupwise(w, "f16", group, bits)


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("N,K", [(128, 128), (200, 256), (384, 1024)])
def test_pack_codes_matches_oracle_layout(bits, N, K):
    _, o = _quantized(N, K, bits, 128, bits + N)
    want = O.pack_codes(o["codes"], bits)
    got = interop.pack_codes(torch.from_numpy(o["codes"].astype(np.int64)), bits)
    assert np.array_equal(got.numpy().view(np.uint32), want)
    back = interop.unpack_codes(got, N, K, bits)
    assert np.array_equal(back.numpy(), o["codes"].astype(np.int64))


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_gptq_row_packing_round_trip(bits):
    g = torch.Generator().manual_seed(bits)
    iw = torch.randint(0, 1 << bits, (256, 96), generator=g)
    qw = interop._gptq_pack_rows(iw, bits)
    assert qw.shape == (256 // 32 * bits, 96) and qw.dtype == torch.int32
    assert torch.equal(interop._gptq_unpack_rows(qw, bits), iw)
    if bits == 4:       # known answer: eight nibbles little-end first
        one = interop._gptq_pack_rows(torch.arange(8).reshape(8, 1), 4)
        assert int(one[0, 0]) == 0x76543210
    if bits == 3:       # known answer of the 10 | 1 | 10 | 1 | 10 split: value 7 everywhere sets every bit
        full = interop._gptq_pack_rows(torch.full((32, 1), 7), 3)
        assert [int(v) & 0xFFFFFFFF for v in full[:, 0]] == [0xFFFFFFFF] * 3


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_autogptq_export_import_cpu(bits):
    N, K, g = 256, 512, 128
    w, o = _quantized(N, K, bits, g, 40 + bits)
    s = torch.from_numpy(o["scales"]).half()
    z = torch.from_numpy(o["zeros"]).half()
    pw = interop.PackedWeight(interop.pack_codes(torch.from_numpy(o["codes"].astype(np.int64)), bits),
                              interop.make_meta(s, z, N, K, g), s, z, bits, g, N, K, False)
    t = interop.to_autogptq(pw)
    assert t["qweight"].shape == (K // 32 * bits, N) and t["qzeros"].shape == (K // g, N // 32 * bits)
    assert t["scales"].shape == (K // g, N) and t["g_idx"].shape == (K,)
    # AutoGPTQ's dequantisation of those tensors is the reference's fake-quantised weight (fp32 product of the same
    # 16-bit scale and integer (q - zero); the reference rounds that product to fp16)
    deq = interop.autogptq_dequant(t)
    assert np.array_equal(deq.half().float().numpy(), o["dequant"])
    back = interop.from_autogptq(t["qweight"], t["qzeros"], t["scales"], t["g_idx"], bits, g)
    assert torch.equal(back.packed, pw.packed) and torch.equal(back.scales, pw.scales)
    assert torch.equal(back.zeros, pw.zeros) and torch.equal(back.meta, pw.meta)


def test_autogptq_rejects_zero_points_it_cannot_hold():
    N, K, g, bits = 128, 256, 128, 4
    w, o = _quantized(N, K, bits, g, 3, positive_zero=False)
    w = np.abs(w) + 0.05                                   # all-positive groups: negative zero-points
    o = O.quantize_groupwise(w, "f16", g, bits)
    assert o["zeros"].min() < 0
    s, z = torch.from_numpy(o["scales"]).half(), torch.from_numpy(o["zeros"]).half()
    pw = interop.PackedWeight(interop.pack_codes(torch.from_numpy(o["codes"].astype(np.int64)), bits),
                              interop.make_meta(s, z, N, K, g), s, z, bits, g, N, K, False)
    with pytest.raises(ValueError, match="do not fit AutoGPTQ v1"):
        interop.to_autogptq(pw)
    t_ok = {"g_idx": torch.arange(K) // 64}
    with pytest.raises(ValueError, match="activation-reordered"):
        interop.from_autogptq(torch.zeros(K // 8, N, dtype=torch.int32), torch.zeros(K // g, N // 8, dtype=torch.int32),
                              torch.ones(K // g, N).half(), t_ok["g_idx"], bits, g)


def test_packed_state_round_trip_cpu():
    N, K, g, bits = 128, 256, 64, 4
    w, o = _quantized(N, K, bits, g, 5)
    s, z = torch.from_numpy(o["scales"]).half(), torch.from_numpy(o["zeros"]).half()
    pw = interop.PackedWeight(interop.pack_codes(torch.from_numpy(o["codes"].astype(np.int64)), bits),
                              interop.make_meta(s, z, N, K, g), s, z, bits, g, N, K, False)
    st = interop.packed_state(pw, "layer.")
    assert all(isinstance(v, torch.Tensor) for v in st.values())
    back = interop.packed_from_state(st, "layer.")
    for f in ("packed", "meta", "scales", "zeros"):
        assert torch.equal(getattr(back, f), getattr(pw, f))
    assert (back.n_bits, back.group, back.N, back.K, back.disable_zero_point) == (bits, g, N, K, False)
    st["layer.q100_header"][0] = 1
    with pytest.raises(ValueError, match="format version"):
        interop.packed_from_state(st, "layer.")


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_kernel_packing_equals_interop_and_imported_weights_run(bits, dtype):
    import q100
    N, K, g = 512, 1024, 128
    gen = torch.Generator().manual_seed(bits)
    w = torch.randn(N, K, generator=gen) * 0.02
    w[:, ::2] = -w[:, ::2].abs() - 0.01
    w[:, 1::2] = w[:, 1::2].abs() + 0.01
    w = w.to(dtype).cuda()
    pw, r = q100.pack_weight(w, bits, g, want_dequant=True)
    codes = interop.unpack_codes(pw.packed, N, K, bits)
    assert torch.equal(interop.pack_codes(codes, bits), pw.packed)
    raw = q100.dequantize(pw.packed, None, None, N, K, g, bits, torch.float32)
    assert torch.equal(raw.to(torch.int64), codes)
    t = interop.to_autogptq(pw)
    back = interop.from_autogptq(t["qweight"], t["qzeros"], t["scales"].to(dtype), t["g_idx"], bits, g, dtype)
    if dtype == torch.float16:          # AutoGPTQ stores fp16 scales: lossless for an fp16 model
        assert torch.equal(back.packed, pw.packed) and torch.equal(back.meta, pw.meta)
    x = torch.randn(8, K, generator=gen).to(dtype).cuda()
    if dtype == torch.float16:
        assert torch.equal(q100.gemm_wna16(x, back), q100.gemm_wna16(x, pw))
        x2 = torch.randn(256, K, generator=gen).to(dtype).cuda()
        assert torch.equal(q100.gemm_wna16(x2, back), q100.gemm_wna16(x2, pw))
        assert torch.equal(interop.autogptq_dequant(t).to(dtype), r.dequant)


@pytest.mark.gpu
def test_save_and_load_packed_quantlinear_tree(tmp_path):
    import dataclasses
    from q100 import llama as QL
    cfg = dataclasses.replace(QL.LLAMA3_8B, hidden_size=256, intermediate_size=512, num_attention_heads=4,
                              num_key_value_heads=2, num_hidden_layers=2, vocab_size=500)
    dev = torch.device("cuda:0")
    m = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, fast_attention=True)
    toks = torch.randint(0, 500, (1, 48), device=dev)
    want = m.logits(toks)
    path = str(tmp_path / "packed.pt")
    st = interop.save_packed(m, path)
    assert sum(k.endswith("q100_header") for k in st) == 14
    # a fresh, UNQUANTISED model of the same shape (different weights) takes the packed twins
    m2 = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, seed=1, fast_attention=True,
                                        quantize=False)
    m2.embed_tokens, m2.norm_weight, m2.lm_head = m.embed_tokens, m.norm_weight, m.lm_head
    for a, b in zip(m.layers, m2.layers):
        b.input_layernorm.weight, b.post_attention_layernorm.weight = a.input_layernorm.weight, a.post_attention_layernorm.weight
    assert interop.load_packed(m2, path) == 14
    assert torch.equal(m2.logits(toks), want)
    m3 = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, seed=1, fast_attention=True,
                                        quantize=False)
    interop.load_packed(m3, st, release_dense=False)          # dense weight rebuilt by the dequantize kernel
    for a, b in zip(m.layers, m3.layers):
        assert b.self_attn.q_proj.weight.shape == (256, 256)
