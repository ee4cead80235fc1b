"""K4-K6 (residual add + RMSNorm, rotary + KV append, SiLU*up) against the reference's eager op sequence
(quant/omni_norm.py:52-63, models/int_llama_layer.py:45, 114-133, 246-251) on the GPU, and the fused decode step
of a block / of a model against the module-path forward."""
import math

import pytest
import torch
import torch.nn.functional as F

import llama_port as LP

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]


def _ulp(dtype):
    return 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols", [(1, 4096), (16, 4096), (3, 8192), (5, 256), (2, 1000)])
@pytest.mark.parametrize("with_res", [False, True])
def test_add_rmsnorm(dtype, rows, cols, with_res):
    from q100 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows * 131 + cols)
    x = torch.randn(rows, 1, cols, generator=g, device=dev).to(dtype)
    r = (torch.randn(rows, 1, cols, generator=g, device=dev) * 3).to(dtype) if with_res else None
    w = (1 + 0.1 * torch.randn(cols, generator=g, device=dev)).to(dtype)
    h, n = ops.add_rmsnorm(x, w, 1e-5, r)
    h_ref = r + x if with_res else x
    assert torch.equal(h, h_ref)                                   # the add is one rounded op: bit-exact
    n_ref = LP.rms_norm(h_ref, w, 1e-5)
    # only the order of the fp32 sum inside the variance differs: at most one ulp of the 16-bit output
    assert torch.allclose(n.float(), n_ref.float(), rtol=_ulp(dtype), atol=1e-7)
    assert (n != n_ref).float().mean().item() < 0.02


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bsz,heads,kvh,hd", [(1, 32, 8, 128), (16, 32, 8, 128), (3, 4, 2, 64), (2, 8, 8, 128)])
def test_rope_kv_decode_bit_exact(dtype, bsz, heads, kvh, hd):
    from q100 import ops
    dev = torch.device("cuda:0")
    max_len, pos = 40, 17
    g = torch.Generator(device=dev).manual_seed(bsz + heads)
    qkv = torch.randn(bsz, 1, (heads + 2 * kvh) * hd, generator=g, device=dev).to(dtype)
    cos, sin = LP.rotary_cache(hd, max_len, 500000.0, dev, dtype)
    kc = torch.zeros(bsz, kvh, max_len, hd, dtype=dtype, device=dev)
    vc = torch.zeros_like(kc)
    p = torch.tensor([[pos]], device=dev)
    q = ops.rope_kv_decode(qkv, cos.contiguous(), sin.contiguous(), p, kc, vc, heads, kvh)
    # reference formulation (int_llama_layer.py:114-125 + the cast our layer applies outside autocast)
    qs, ks, vs = qkv.split([heads * hd, kvh * hd, kvh * hd], dim=-1)
    q_ref = qs.view(bsz, 1, heads, hd).transpose(1, 2).type(torch.float32)
    k_ref = ks.view(bsz, 1, kvh, hd).transpose(1, 2)
    v_ref = vs.view(bsz, 1, kvh, hd).transpose(1, 2)
    q_ref, k_ref = LP.apply_rotary_pos_emb(q_ref, k_ref, cos, sin, p)
    assert torch.equal(q, q_ref.to(dtype))
    assert torch.equal(kc[:, :, pos:pos + 1], k_ref) and torch.equal(vc[:, :, pos:pos + 1], v_ref)
    assert kc[:, :, :pos].abs().max().item() == 0 and kc[:, :, pos + 1:].abs().max().item() == 0


@pytest.mark.parametrize("pos", [-1, 40, 41, 10 ** 6, 35])
def test_rope_kv_decode_position_outside_cache(pos):
    """A static cache can overflow (the reference's torch.cat cache cannot): K5 must write nothing and raise the flag
    when the device-side position is outside the cache or the rotary table (here: table 32 rows, cache 40 slots)."""
    from q100 import ops
    dev = torch.device("cuda:0")
    bsz, heads, kvh, hd, max_len, rope_len = 2, 4, 2, 64, 40, 32
    qkv = torch.randn(bsz, 1, (heads + 2 * kvh) * hd, device=dev).half()
    cos, sin = LP.rotary_cache(hd, rope_len, 500000.0, dev, torch.float16)
    guard = torch.full((bsz * kvh * max_len * hd + 4096,), 7.0, dtype=torch.float16, device=dev)
    kc = guard[2048:2048 + bsz * kvh * max_len * hd].view(bsz, kvh, max_len, hd)
    vc = torch.full_like(kc, 7.0)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.rope_kv_decode(qkv, cos.contiguous(), sin.contiguous(), torch.tensor([[pos]], device=dev), kc, vc, heads, kvh, err)
    torch.cuda.synchronize()
    assert int(err.item()) == 1
    assert bool((guard == 7.0).all()) and bool((vc == 7.0).all())


def test_graph_decoder_refuses_to_run_past_the_cache():
    from q100 import llama as QL
    import dataclasses
    cfg = dataclasses.replace(QL.LLAMA3_8B, hidden_size=256, intermediate_size=512, num_attention_heads=4,
                              num_key_value_heads=2, num_hidden_layers=1, vocab_size=1000)
    m = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), torch.device("cuda:0"), torch.float16,
                                       fast_attention=True)
    dec = QL.GraphDecoder(m, 1, 8, use_graph=False)
    dec.prime(4)
    for _ in range(4):
        dec.step()
    dec.check()
    with pytest.raises(RuntimeError):
        dec.step()
    with pytest.raises(ValueError):
        dec.prime(7)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,inter", [(1, 14336), (16, 14336), (3, 512), (2, 1000)])
def test_silu_mul(dtype, rows, inter):
    from q100 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows + inter)
    gu = (torch.randn(rows, 1, 2 * inter, generator=g, device=dev) * 3).to(dtype)
    out = ops.silu_mul(gu)
    ref = F.silu(gu[..., :inter]) * gu[..., inter:]
    assert torch.allclose(out.float(), ref.float(), rtol=_ulp(dtype), atol=1e-7)
    assert (out != ref).float().mean().item() < 0.01


@pytest.mark.parametrize("dtype", DTYPES)
def test_block_fused_decode_step_matches_forward(dtype):
    """decode_step_fused against forward() with the same static cache contents (W4 g128)."""
    from q100 import llama as QL
    from models.int_llama_layer import StaticKV
    dev = torch.device("cuda:0")
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 1, 64)
    args = QL.quant_args(4, 128)
    w = QL.random_block_weights(cfg, 0, 5, dev, dtype)
    layer = QL.build_block(cfg, w, args)
    layer.self_attn.fast_attention = True
    bsz, max_len, ctx = 3, 32, 9
    g = torch.Generator(device=dev).manual_seed(0)
    kv_a = StaticKV(bsz, 2, max_len, 64, dtype, dev)
    kv_a.k[:, :, :ctx] = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    kv_a.v[:, :, :ctx] = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    kv_b = StaticKV(bsz, 2, max_len, 64, dtype, dev)
    kv_b.k.copy_(kv_a.k)
    kv_b.v.copy_(kv_a.v)
    x = torch.randn(bsz, 1, 512, generator=g, device=dev).to(dtype)
    res = torch.randn(bsz, 1, 512, generator=g, device=dev).to(dtype)
    pos = torch.tensor([[ctx]], device=dev)
    slots = torch.arange(max_len, device=dev).view(1, 1, 1, max_len)
    mask = torch.where(slots <= ctx, 0.0, torch.finfo(dtype).min).to(dtype).expand(bsz, 1, 1, max_len)
    with torch.no_grad():
        ref = layer(res + x, attention_mask=mask, position_ids=pos, past_key_value=kv_a, use_cache=False)[0]
        out = layer.decode_step_fused(x, res, kv_b, pos, mask)
    assert out is not None
    d, h1 = out
    got = h1 + d
    assert torch.equal(kv_a.v, kv_b.v)
    assert torch.equal(kv_a.k, kv_b.k)                         # rotary bit-exact, same GEMM underneath
    err = (got.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    assert err <= 4 * _ulp(dtype), err


def test_graph_decoder_fused_matches_module_path():
    from q100 import llama as QL
    dev = torch.device("cuda:0")
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 3, 1024)
    model = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128), dev, torch.float16, seed=2,
                                           fast_attention=True)
    toks = torch.randint(0, cfg.vocab_size, (2, 1), generator=torch.Generator().manual_seed(0)).to(dev)
    outs = []
    for fused, graph in ((False, False), (True, False), (True, True)):
        dec = QL.GraphDecoder(model, 2, 24, use_graph=graph, fused=fused)
        dec.prime(0)
        t, seq = toks, []
        for _ in range(6):
            t = dec.step(t).clone()
            seq.append(t)
        outs.append(torch.cat(seq, dim=1).cpu())
    # random-init logits are nearly flat, so the argmax may legitimately flip on last-bit differences;
    # eager-fused and graph-fused run the very same kernels and must agree exactly
    assert torch.equal(outs[1], outs[2])
    assert (outs[0] == outs[1]).float().mean().item() >= 0.5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bsz,q_len,heads,kvh,hd,per_batch_pos", [
    (1, 2048, 32, 8, 128, False), (2, 37, 32, 8, 128, True), (3, 5, 4, 2, 64, False), (1, 1, 8, 8, 128, False),
    (2, 300, 8, 1, 128, False), (1, 64, 6, 2, 48, False)])
def test_rope_qkv_prefill_bit_exact(dtype, bsz, q_len, heads, kvh, hd, per_batch_pos):
    """K5p against the reference formulation (int_llama_layer.py:114-127: view / transpose, q in fp32, rotary on
    q and k) + the cast back to the model dtype our layer applies outside autocast: bit-exact, all three outputs."""
    from q100 import ops
    dev = torch.device("cuda:0")
    past = 11
    g = torch.Generator(device=dev).manual_seed(bsz * 7 + q_len)
    qkv = torch.randn(bsz, q_len, (heads + 2 * kvh) * hd, generator=g, device=dev).to(dtype)
    cos, sin = LP.rotary_cache(hd, past + q_len + 3, 500000.0, dev, dtype)
    pos = torch.arange(past, past + q_len, device=dev).unsqueeze(0)
    if per_batch_pos:
        pos = torch.stack([pos[0] - b for b in range(bsz)])
    q, k, v = ops.rope_qkv_prefill(qkv, cos.contiguous(), sin.contiguous(), pos, heads, kvh)
    qs, ks, vs = qkv.split([heads * hd, kvh * hd, kvh * hd], dim=-1)
    q_ref = qs.view(bsz, q_len, heads, hd).transpose(1, 2).type(torch.float32)
    k_ref = ks.view(bsz, q_len, kvh, hd).transpose(1, 2)
    v_ref = vs.view(bsz, q_len, kvh, hd).transpose(1, 2)
    q_ref, k_ref = LP.apply_rotary_pos_emb(q_ref, k_ref, cos, sin, pos)
    assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    assert torch.equal(q, q_ref.to(dtype))
    assert torch.equal(k, k_ref) and torch.equal(v, v_ref)


def test_rope_qkv_prefill_position_outside_table():
    from q100 import ops
    dev = torch.device("cuda:0")
    qkv = torch.randn(1, 4, (4 + 4) * 64, device=dev).half()
    cos, sin = LP.rotary_cache(64, 8, 500000.0, dev, torch.float16)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.rope_qkv_prefill(qkv, cos.contiguous(), sin.contiguous(), torch.tensor([[5, 6, 7, 8]], device=dev), 4, 2, err)
    torch.cuda.synchronize()
    assert int(err.item()) == 1
    with pytest.raises(ValueError):
        ops.rope_qkv_prefill(qkv, cos.contiguous(), sin.contiguous(), torch.zeros(3, 4, dtype=torch.int64, device=dev), 4, 2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("use_cache", [False, True])
def test_block_prefill_kernels_match_eager_formulation(dtype, use_cache, monkeypatch):
    """QuantLlamaDecoderLayer.forward at prefill sizes: the K4 / K5p / K6 route (inference, no autocast) against the
    same layer with those kernels switched off (the reference's eager op sequence on the same packed GEMMs)."""
    from q100 import llama as QL
    from quant import omni_norm
    import models.int_llama_layer as ILL
    dev = torch.device("cuda:0")
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 1, 64)
    layer = QL.build_block(cfg, QL.random_block_weights(cfg, 0, 3, dev, dtype), QL.quant_args(4, 128))
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(2, 96, 512, generator=g, device=dev).to(dtype)
    mask = QL.QuantLlamaModel.causal_mask(2, 96, 96, dtype, dev)
    pos = torch.arange(96, device=dev).unsqueeze(0)
    with torch.no_grad():
        out = layer(x, attention_mask=mask, position_ids=pos, use_cache=use_cache)
        with torch.enable_grad():            # grad mode switches every fused glue kernel off (nothing requires grad)
            monkeypatch.setattr(omni_norm.OmniLlamaRMSNorm, "_k4_ok", lambda self, t: False)
            ref = layer(x, attention_mask=mask, position_ids=pos, use_cache=use_cache)
    assert torch.allclose(out[0].float(), ref[0].float(), rtol=4 * _ulp(dtype), atol=4 * _ulp(dtype) * ref[0].float().abs().max().item())
    if use_cache:
        (k, v), (k_ref, v_ref) = out[-1], ref[-1]
        # the norm in front differs by the order of its fp32 sum (<= 1 ulp of n1), so k / v agree to GEMM round-off
        assert torch.allclose(v.float(), v_ref.float(), rtol=0, atol=8 * _ulp(dtype) * v_ref.float().abs().max().item())
        assert torch.allclose(k.float(), k_ref.float(), rtol=0, atol=8 * _ulp(dtype) * k_ref.float().abs().max().item())


# ------------------------------------------------------------------------------------------------ K8 (decode attention)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bsz,heads,kvh", [(1, 32, 8), (2, 32, 8), (16, 32, 8), (1, 64, 8), (3, 4, 1), (2, 8, 8), (1, 6, 2)])
@pytest.mark.parametrize("max_len,pos", [(512, 0), (512, 5), (512, 15), (512, 16), (512, 63), (512, 64), (512, 256),
                                         (512, 511), (4096, 1000), (4096, 4095), (8192, 200), (8192, 8191), (8192, 4243)])
def test_attn_decode_matches_k5_plus_reference_attention(dtype, bsz, heads, kvh, max_len, pos):
    """K8 = K5 (bit-exact cache append) + softmax(q k^T / sqrt(d)) v over slots 0..pos in fp32 from the 16-bit q, k, v
    (int_llama_layer.py:136-163 on the static cache); the output agrees within the rounding of the 16-bit result."""
    from q100 import ops
    if bsz == 16 and max_len >= 4096 and (dtype == torch.bfloat16 or max_len == 8192):
        pytest.skip("covered by fp16 / the smaller cache")
    dev = torch.device("cuda:0")
    hd = 128
    g = torch.Generator(device=dev).manual_seed(bsz * 131 + heads + pos)
    qkv = torch.randn(bsz, (heads + 2 * kvh) * hd, generator=g, device=dev).to(dtype)
    cos, sin = LP.rotary_cache(hd, max_len, 500000.0, dev, dtype)
    cos, sin = cos.contiguous(), sin.contiguous()
    kc = torch.randn(bsz, kvh, max_len, hd, generator=g, device=dev).to(dtype)
    vc = torch.randn(bsz, kvh, max_len, hd, generator=g, device=dev).to(dtype)
    kc[:, :, pos + 1:] = 1000.0                          # slots behind the new token must not be read
    vc[:, :, pos + 1:] = 1000.0
    kc2, vc2 = kc.clone(), vc.clone()
    p = torch.tensor([[pos]], device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    out = ops.attn_decode(qkv, cos, sin, p, kc, vc, heads, kvh, err)
    q = ops.rope_kv_decode(qkv, cos, sin, p, kc2, vc2, heads, kvh)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    assert torch.equal(kc, kc2) and torch.equal(vc, vc2)
    rep = heads // kvh
    k = kc2[:, :, :pos + 1].float().repeat_interleave(rep, dim=1)
    v = vc2[:, :, :pos + 1].float().repeat_interleave(rep, dim=1)
    s = (q.float() @ k.transpose(-1, -2)) / math.sqrt(hd)
    pr = torch.softmax(s, dim=-1)
    ref = (pr @ v).reshape(bsz, heads * hd)
    assert out.shape == ref.shape
    # the kernel rounds the (un-normalised) probabilities to the model dtype for its second MMA, like the reference
    # (int_llama_layer.py:152-153): half a unit per term of sum p |v|, plus the rounding of the result
    tol = _ulp(dtype) * (pr @ v.abs()).reshape(bsz, heads * hd) + 1e-6
    bad = (out.float() - ref).abs() > tol
    assert not bool(bad.any()), f"{int(bad.sum())} outside, max err {(out.float() - ref).abs().max().item():.3e}"
    # a second call on the same workspace (the counters must have been left zeroed)
    out2 = ops.attn_decode(qkv, cos, sin, p, kc, vc, heads, kvh, err)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("pos", [-1, 512, 10 ** 6, 300])
def test_attn_decode_position_outside_cache(pos):
    from q100 import ops
    dev = torch.device("cuda:0")
    bsz, heads, kvh, hd, max_len, rope_len = 2, 8, 2, 128, 512, 256
    qkv = torch.randn(bsz, (heads + 2 * kvh) * hd, device=dev).half()
    cos, sin = LP.rotary_cache(hd, rope_len, 500000.0, dev, torch.float16)
    kc = torch.full((bsz, kvh, max_len, hd), 7.0, dtype=torch.float16, device=dev)
    vc = torch.full_like(kc, 7.0)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.attn_decode(qkv, cos.contiguous(), sin.contiguous(), torch.tensor([[pos]], device=dev), kc, vc, heads, kvh, err)
    torch.cuda.synchronize()
    assert int(err.item()) == 1
    assert bool((kc == 7.0).all()) and bool((vc == 7.0).all())
    # the next good call works (no counter was left behind)
    err.zero_()
    out = ops.attn_decode(qkv, cos.contiguous(), sin.contiguous(), torch.tensor([[3]], device=dev), kc, vc, heads, kvh, err)
    assert int(err.item()) == 0 and bool(torch.isfinite(out).all())


# ------------------------------------------------------------------------------------------------ K10 (argmax)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols", [(1, 128256), (16, 128256), (3, 1000), (1, 1003), (2, 8), (5, 32000)])
def test_argmax_rows(dtype, rows, cols):
    from q100 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows * 7 + cols)
    x = torch.randn(rows, 1, cols, generator=g, device=dev).to(dtype)
    got = ops.argmax_rows(x)
    assert got.shape == (rows, 1) and got.dtype == torch.int64
    # 16-bit values collide: the maximum may occur more than once -- the kernel returns the lowest index
    mx = x.amax(-1, keepdim=True)
    first = torch.where(x == mx, torch.arange(cols, device=dev).expand_as(x), cols).amin(-1)
    assert torch.equal(got, first)
    # ties planted at known places, NaN as the maximum, preallocated output, second call on the same workspace
    x[..., :] = 0.0
    if cols > 600:
        x[0, 0, 517] = 3.0
        x[0, 0, 133] = 3.0
    out = torch.empty(rows, 1, dtype=torch.int64, device=dev)
    ops.argmax_rows(x, out=out)
    assert int(out[0, 0]) == (133 if cols > 600 else 0)
    assert all(int(v) == 0 for v in out[1:, 0])
    x[0, 0, cols - 1] = float("nan")
    assert int(ops.argmax_rows(x)[0, 0]) == cols - 1 == int(torch.argmax(x[0, 0]))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("abits", [8, 4])
def test_kv_quantised_before_repeat_kv_is_exact(dtype, abits):
    """QuantMatMul.quant_x2 works per (head, token) row over head_dim and repeat_kv copies whole heads, so the block
    quantises the kv heads before the GQA expansion: the same bits as the reference's order
    (int_llama_layer.py:134-141: repeat_kv first), a quarter of the rows."""
    from models.int_llama_layer import repeat_kv
    from quant.int_matmul import QuantMatMul
    dev = torch.device("cuda:0")
    act = {"n_bits": abits, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"}
    mm = QuantMatMul(act, act, matmul_func=torch.matmul).to(dev)
    mm.set_quant_state(False, True)
    g = torch.Generator(device=dev).manual_seed(abits)
    k = (torch.randn(2, 2, 37, 128, generator=g, device=dev) * 1.7).to(dtype)
    k[0, 1, 5] = 0                                       # a constant row (scale at its lower clamp)
    a = repeat_kv(mm.quant_x2(k), 4)
    b = mm.quant_x2(repeat_kv(k, 4))
    assert torch.equal(a, b)


@pytest.mark.parametrize("wbits,abits,group", [(8, 8, None), (4, 4, None), (4, 8, 128)])
def test_fast_attention_with_act_quant_matches_reference_formulation(wbits, abits, group):
    """``fast_attention`` with activation quantisation on: SDPA on the fake-quantised Q / K / V (kv heads, no
    expansion) against the reference formulation of the same block -- prefill with a causal mask, then one decode
    step on a static cache.  The same fake-quantised tensors enter softmax(QK^T)V on both sides; what differs is the
    library kernel's online softmax against the fp32 softmax rounded to 16 bits.
    (i) attention module with o_proj's input quantiser off (nothing discrete behind the core): <= 16 16-bit ulps of
    the largest output; (ii) the whole block with every quantiser on: a 1-ulp difference may flip codes of o_proj's
    and the MLP's per-token quantisers (one A4 step is 1/15 of a row's range), so the bound is on the relative
    Frobenius distance -- 1e-2 (A8) / 6e-2 (A4); a wrong mask or head mapping gives O(1)."""
    from q100 import llama as QL
    from models.int_llama_layer import StaticKV
    dev, dtype = torch.device("cuda:0"), torch.float16
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 1, 64)
    args = QL.quant_args(wbits, group, abits)
    layer = QL.build_block(cfg, QL.random_block_weights(cfg, 0, 11, dev, dtype), args)
    attn = layer.self_attn
    assert layer.act_quant_active() and attn.qkt_matmul.use_act_quant and attn.pv_matmul.use_act_quant
    g = torch.Generator(device=dev).manual_seed(3)
    T = 48
    x = torch.randn(2, T, 512, generator=g, device=dev).to(dtype)
    pos = torch.arange(T, device=dev).unsqueeze(0)
    mask = QL.QuantLlamaModel.causal_mask(2, T, T, dtype, dev)
    fro_tol = 1e-2 if abits == 8 else 6e-2

    def rel_max(a, b):
        return (a.float() - b.float()).abs().max().item() / b.float().abs().max().item()

    def rel_fro(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    def both(fn):
        out = []
        for fast in (False, True):
            attn.fast_attention = fast
            assert attn.sdpa_route() == fast
            with torch.no_grad():
                out.append(fn())
        return out

    # (i) the core alone
    attn.o_proj.use_act_quant = False
    with torch.no_grad():
        h = layer.input_layernorm(x)
    ref, got = both(lambda: attn(h, attention_mask=mask, position_ids=pos)[0])
    attn.fast_attention = True
    with torch.no_grad():
        got_causal = attn(h, attention_mask=None, position_ids=pos)[0]        # what QuantLlamaModel.forward passes
    assert rel_max(got, ref) <= 16 * _ulp(dtype), rel_max(got, ref)
    assert rel_max(got_causal, ref) <= 16 * _ulp(dtype), rel_max(got_causal, ref)
    attn.o_proj.use_act_quant = True

    # (ii) the whole block, every quantiser on
    ref, got = both(lambda: layer(x, attention_mask=mask, position_ids=pos)[0])
    assert rel_fro(got, ref) <= fro_tol, rel_fro(got, ref)

    # decode step on a static cache: the cache rows are quantised on the kv heads, unwritten slots are masked
    bsz, max_len, ctx = 2, 32, 9
    kvs = []
    k0 = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    v0 = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    for _ in range(4):
        kv = StaticKV(bsz, 2, max_len, 64, dtype, dev)
        kv.k[:, :, :ctx] = k0
        kv.v[:, :, :ctx] = v0
        kvs.append(kv)
    x1 = torch.randn(bsz, 1, 512, generator=g, device=dev).to(dtype)
    p1 = torch.tensor([[ctx]], device=dev)
    slots = torch.arange(max_len, device=dev).view(1, 1, 1, max_len)
    m1 = torch.where(slots <= ctx, 0.0, torch.finfo(dtype).min).to(dtype).expand(bsz, 1, 1, max_len)
    attn.o_proj.use_act_quant = False
    with torch.no_grad():
        h1 = layer.input_layernorm(x1)
    it = iter(kvs)
    ref1, got1 = both(lambda: attn(h1, attention_mask=m1, position_ids=p1, past_key_value=next(it))[0])
    assert rel_max(got1, ref1) <= 16 * _ulp(dtype), rel_max(got1, ref1)
    attn.o_proj.use_act_quant = True
    ref2, got2 = both(lambda: layer(x1, attention_mask=m1, position_ids=p1, past_key_value=next(it), use_cache=False)[0])
    assert rel_fro(got2, ref2) <= fro_tol, rel_fro(got2, ref2)
    for kv in kvs[1:]:
        assert torch.equal(kv.k, kvs[0].k) and torch.equal(kv.v, kvs[0].v)

    # a softmax-output quantiser of its own keeps the reference formulation
    attn.fast_attention = True
    attn.pv_matmul.x1_quantizer.change_n_bits(8)
    assert not attn.sdpa_route()
    attn.pv_matmul.x1_quantizer.change_n_bits(16)
    assert attn.sdpa_route()


@pytest.mark.parametrize("wbits,abits,group,bsz", [(8, 8, None, 1), (8, 8, None, 3), (4, 4, None, 2), (4, 8, 128, 16)])
def test_block_fused_decode_step_with_act_quant(wbits, abits, group, bsz):
    """decode_step_fused with activation quantisation on (W8A8 / W4A4: K3c GEMMs, SDPA on the fake-quantised Q / K / V)
    against forward() in the reference formulation on the same static cache.  The cache rows are bit-identical (same
    GEMM, rotary bit-exact); the block output differs by the attention core's rounding pushed through per-token
    quantisers (tolerance of test_fast_attention_with_act_quant_matches_reference_formulation)."""
    from q100 import llama as QL
    from models.int_llama_layer import StaticKV
    dev, dtype = torch.device("cuda:0"), torch.float16
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 1, 64)
    layer = QL.build_block(cfg, QL.random_block_weights(cfg, 0, 13, dev, dtype), QL.quant_args(wbits, group, abits))
    assert layer.act_quant_active() and layer.self_attn.pv_matmul.use_act_quant
    max_len, ctx = 32, 9
    g = torch.Generator(device=dev).manual_seed(bsz)
    k0 = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    v0 = torch.randn(bsz, 2, ctx, 64, generator=g, device=dev).to(dtype)
    kvs = []
    for _ in range(3):
        kv = StaticKV(bsz, 2, max_len, 64, dtype, dev)
        kv.k[:, :, :ctx] = k0
        kv.v[:, :, :ctx] = v0
        kvs.append(kv)
    x = torch.randn(bsz, 1, 512, generator=g, device=dev).to(dtype)
    res = torch.randn(bsz, 1, 512, generator=g, device=dev).to(dtype)
    pos = torch.tensor([[ctx]], device=dev)
    slots = torch.arange(max_len, device=dev).view(1, 1, 1, max_len)
    mask = torch.where(slots <= ctx, 0.0, torch.finfo(dtype).min).to(dtype).expand(bsz, 1, 1, max_len)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.no_grad():
        layer.self_attn.fast_attention = False
        ref = layer(res + x, attention_mask=mask, position_ids=pos, past_key_value=kvs[0], use_cache=False)[0]
        layer.self_attn.fast_attention = True
        mod = layer(res + x, attention_mask=mask, position_ids=pos, past_key_value=kvs[1], use_cache=False)[0]
        out = layer.decode_step_fused(x, res, kvs[2], pos, mask, err, causal_cache=True)
    assert out is not None, "the fused decode step must take W8A8 / W4A4 blocks"
    d, h1 = out
    got = (h1 + d).view_as(ref)
    assert int(err.item()) == 0
    for kv in kvs[1:]:
        assert torch.equal(kv.k, kvs[0].k) and torch.equal(kv.v, kvs[0].v)
    tol = 1e-2 if abits == 8 else 6e-2
    for name, t in (("forward + SDPA", mod), ("decode_step_fused", got)):
        e = ((t.float() - ref.float()).norm() / ref.float().norm()).item()
        assert e <= tol, (name, e)

    # a quantiser of its own on the softmax output: the fused route declines, forward() keeps the reference formulation
    layer.self_attn.pv_matmul.x1_quantizer.change_n_bits(8)
    with torch.no_grad():
        assert layer.decode_step_fused(x, res, kvs[2], pos, mask, err, causal_cache=True) is None


@pytest.mark.parametrize("wbits,abits", [(8, 8), (4, 4)])
def test_graph_decoder_with_act_quant(wbits, abits):
    """GraphDecoder on a W8A8 / W4A4 model: the fused route now takes these blocks.  Eager-fused and graph-fused run the
    same kernels on the same state and agree exactly; the logits of the first step agree with the module path
    (forward() per block) within the act-quant block tolerance."""
    from q100 import llama as QL
    dev = torch.device("cuda:0")
    cfg = QL.LlamaCfg(512, 1024, 8, 2, 3, 1024)
    model = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(wbits, None, abits), dev, torch.float16, seed=4,
                                           fast_attention=True)
    toks = torch.randint(0, cfg.vocab_size, (2, 1), generator=torch.Generator().manual_seed(0)).to(dev)
    outs, first = [], []
    for fused, graph in ((False, False), (True, False), (True, True)):
        dec = QL.GraphDecoder(model, 2, 24, use_graph=graph, fused=fused)
        dec.prime(0)
        t, seq = toks, []
        for i in range(5):
            t = dec.step(t).clone()
            if i == 0:
                first.append(dec.last_logits.float().clone())
            seq.append(t)
        assert int(dec.err.item()) == 0
        outs.append(torch.cat(seq, dim=1).cpu())
    assert torch.equal(outs[1], outs[2])
    assert torch.equal(first[1], first[2])
    e = ((first[1] - first[0]).norm() / first[0].norm()).item()
    assert e <= (2e-2 if abits == 8 else 1e-1), e
