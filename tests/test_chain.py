"""GPU: the decode chain (K9, q100_chain_build / q100_chain_run): one persistent launch for the GEMMs of a decoder
block (or a whole model) at decode shapes with RMSNorm / residual / SiLU*up folded in and the single-token attention
as phases of the same launch.

* every GEMM phase flavour against a float64 reference built from the quantize kernel's fake-quantised weight (the
  tolerance of tests/test_parity_gemm.py; RMSNorm and SiLU steps to 16-bit round-off);
* the attention phases against K5 (rotary + cache append: bit-exact cache) + an fp32 softmax reference;
* a chain of all blocks of a random-init model against the per-kernel decode route (K4 / K5 / K3 / SDPA) over several
  tokens, eagerly and as CUDA-graph replays (the grid-barrier counters must come back to zero after every launch).
"""
import dataclasses
import math

import pytest
import torch
import torch.nn.functional as F

import q100
from q100 import _lib, chain as QC, llama as QL, ops
from conftest import TDT

pytestmark = pytest.mark.gpu
EPS = {"f16": 2.0 ** -11, "bf16": 2.0 ** -8}


def _pw(N, K, bits, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(TDT[dtype])
    pw, r = q100.pack_weight(w, bits, 128, want_dequant=True)
    return pw, r.dequant.double()


def _close(got, ref, dtype, mag, extra=0.0):
    tol = 2 * EPS[dtype] * ref.abs() + 8 * EPS[dtype] / 2 * mag + extra + 1e-6
    bad = (got.double() - ref).abs() > tol
    assert not bool(bad.any()), f"{int(bad.sum())}/{bad.numel()} outside tolerance, max err {float((got.double() - ref).abs().max()):.3e}"


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("bits", [4, 2, 3, 8])
@pytest.mark.parametrize("M", [1, 2, 5, 8, 12, 16])
def test_gemm_phases(dtype, bits, M):
    """o_proj-like (residual + partial sums) -> gate|up-like (RMSNorm staged from the partial sums) -> down-like
    (SiLU*up staged, residual + partial sums), i.e. the MLP half of a block, on ragged-free small shapes."""
    H, I = 512, 1024
    dt = TDT[dtype]
    wo, wo64 = _pw(H, H, bits, dtype, 1)
    wgu, wgu64 = _pw(2 * I, H, bits, dtype, 2)
    wd, wd64 = _pw(H, I, bits, dtype, 3)
    g = torch.Generator(device="cuda").manual_seed(M)
    att = torch.randn(M, H, generator=g, device="cuda").to(dt)
    h0 = torch.randn(M, H, generator=g, device="cuda").to(dt)
    nw = (1 + 0.1 * torch.randn(H, generator=g, device="cuda")).to(dt)
    h = h0.clone()
    gu = torch.zeros(M, 2 * I, dtype=dt, device="cuda")
    ssq_a = torch.zeros(QC.ssq_parts(H), M, dtype=torch.float32, device="cuda")
    ssq_b = torch.zeros_like(ssq_a)
    ch = QC.Chain(M, 128, bits, False, dt, torch.device("cuda:0"))
    ch.gemm(att, wo, h, epi_mode=_lib.EPI_ADD_SSQ, residual=h, ssq_out=ssq_b)
    ch.gemm(h, wgu, gu, x_mode=_lib.X_NORM, norm_weight=nw, ssq_in=ssq_b, eps=1e-5)
    ch.gemm(gu, wd, h, x_mode=_lib.X_GATE_UP, epi_mode=_lib.EPI_ADD_SSQ, residual=h, ssq_out=ssq_a)
    ch.build()
    for rep in range(3):                       # the barrier counters and the fix-up workspace reset themselves
        h.copy_(h0)
        ch.run()
        torch.cuda.synchronize()
        ch.check()
        # reference, float64 with the 16-bit roundings of the reference's op sequence
        o = (att.double() @ wo64.T).to(dt)
        h1 = (h0.double() + o.double()).to(dt)
        mag1 = att.double().abs() @ wo64.abs().T
        n2 = (nw.double() * (h1.double() * torch.rsqrt(h1.double().pow(2).mean(-1, keepdim=True) + 1e-5))).to(dt)
        gu_ref = n2.double() @ wgu64.T
        _close(gu, gu_ref, dtype, n2.double().abs() @ wgu64.abs().T * 0 + 0, extra=4 * EPS[dtype] * (n2.double().abs() @ wgu64.abs().T))
        act = (F.silu(gu[:, :I].float()).to(dt).double() * gu[:, I:].double()).to(dt)
        d_ref = act.double() @ wd64.T
        h2_ref = (h1.double() + d_ref.to(dt).double())
        magd = act.double().abs() @ wd64.abs().T
        _close(h, h2_ref, dtype, magd * 0, extra=4 * EPS[dtype] * (magd + h1.double().abs()))
        # the partial sums of squares the next RMSNorm would use
        assert torch.allclose(ssq_a.sum(0).double(), h.double().pow(2).sum(-1), rtol=1e-5)
        assert torch.allclose(ssq_b.sum(0).double(), h1.double().pow(2).sum(-1), rtol=2e-2)   # h1 itself is only close


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("bsz,heads,kvh", [(1, 32, 8), (4, 8, 2), (16, 32, 8), (2, 8, 1), (3, 2, 2), (1, 4, 2)])
@pytest.mark.parametrize("pos,max_len", [(0, 64), (17, 64), (63, 64), (64, 300), (256, 320), (2100, 4096)])
def test_attention_phases(dtype, bsz, heads, kvh, pos, max_len):
    dt, hd = TDT[dtype], 128
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(bsz * 100 + heads + pos)
    qkv = torch.randn(bsz, (heads + 2 * kvh) * hd, generator=g, device=dev).to(dt)
    import llama_port as LP
    cos, sin = LP.rotary_cache(hd, max_len, 500000.0, dev, dt)
    cos, sin = cos.contiguous(), sin.contiguous()
    kc = (torch.randn(bsz, kvh, max_len, hd, generator=g, device=dev) * 0.5).to(dt)
    vc = (torch.randn(bsz, kvh, max_len, hd, generator=g, device=dev) * 0.5).to(dt)
    kc[:, :, pos:] = 0
    vc[:, :, pos:] = 0
    kc2, vc2 = kc.clone(), vc.clone()
    p = torch.tensor([[pos]], device=dev)
    out = torch.zeros(bsz, heads * hd, dtype=dt, device=dev)
    ch = QC.Chain(bsz, 128, 4, False, dt, dev)
    ch.attention(qkv, cos, sin, p, kc, vc, out, heads, kvh)
    ch.build()
    ch.run()
    torch.cuda.synchronize()
    ch.check()
    # K5 on the same input: the cache must be bit-identical; attention in fp32 on K5's q and the cache
    q = ops.rope_kv_decode(qkv.view(bsz, 1, -1), cos, sin, p, kc2, vc2, heads, kvh)
    assert torch.equal(kc, kc2) and torch.equal(vc, vc2)
    rep = heads // kvh
    k = kc2[:, :, :pos + 1].float().repeat_interleave(rep, dim=1)
    v = vc2[:, :, :pos + 1].float().repeat_interleave(rep, dim=1)
    s = (q.float() @ k.transpose(2, 3)) / math.sqrt(hd)
    ref = (torch.softmax(s, dim=-1) @ v).reshape(bsz, heads * hd)
    assert torch.allclose(out.float(), ref, rtol=4 * EPS[dtype], atol=4 * EPS[dtype] * float(ref.abs().max()))


def _tiny(cfg_kw, layers, bits=4, dtype=torch.float16, seed=0):
    cfg = dataclasses.replace(QL.LLAMA3_8B, num_hidden_layers=layers, vocab_size=2000, **cfg_kw)
    return QL.QuantLlamaModel.random_init(cfg, QL.quant_args(bits, 128, 16), torch.device("cuda:0"), dtype, seed=seed,
                                          fast_attention=True)


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("bsz", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("shape", ["gqa2", "gqa4", "gqa8"])
def test_model_chain_matches_per_kernel_route(shape, bsz, use_graph):
    kw = {"gqa2": dict(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=2),
          "gqa4": dict(hidden_size=1024, intermediate_size=1536, num_attention_heads=8, num_key_value_heads=2),
          "gqa8": dict(hidden_size=1024, intermediate_size=2048, num_attention_heads=8, num_key_value_heads=1)}[shape]
    m = _tiny(kw, 3)
    dev = torch.device("cuda:0")
    start, steps, max_len = 5, 6, 80
    decs = [QL.GraphDecoder(m, bsz, max_len, use_graph=use_graph, chain=c) for c in (True, False)]
    g = torch.Generator(device=dev).manual_seed(3)
    for kv_a, kv_b in zip(decs[0].caches, decs[1].caches):       # the same "prefill" in both caches
        kv_a.k[:, :, :start] = (torch.randn(kv_a.k[:, :, :start].shape, generator=g, device=dev) * 0.3).half()
        kv_a.v[:, :, :start] = (torch.randn(kv_a.v[:, :, :start].shape, generator=g, device=dev) * 0.3).half()
        kv_b.k.copy_(kv_a.k)
        kv_b.v.copy_(kv_a.v)
    # prime() runs warm-up steps that write positions start, start + 1: identical garbage in both, overwritten below
    for d in decs:
        d.prime(start)
    assert decs[0].chain is not None and decs[1].chain is None
    tok = torch.randint(0, 2000, (bsz, 1), generator=g, device=dev)
    for s in range(steps):
        outs = []
        for d in decs:
            d.step(tok)
            outs.append(d.last_logits.float().clone())
        torch.cuda.synchronize()
        ref = outs[1]
        err = float((outs[0] - ref).abs().max() / ref.abs().max())
        assert err <= 2e-2, f"step {s}: chain logits differ from the per-kernel route by {err:.3e}"
        tok = decs[1].next_tok.clone()
    for d in decs:
        d.check()
    for kv_a, kv_b in zip(decs[0].caches, decs[1].caches):
        ka, kb = kv_a.k[:, :, :start + steps].float(), kv_b.k[:, :, :start + steps].float()
        assert float((ka - kb).abs().max()) <= 2e-2 * float(kb.abs().max())


def test_chain_8b_block_shapes_against_per_kernel_route():
    """One block at the real LLaMA3-8B shapes (the shapes BASELINE.json names), batch 1 and 16."""
    m = _tiny({}, 1)
    for bsz in (1, 16):
        decs = [QL.GraphDecoder(m, bsz, 300, use_graph=False, chain=c) for c in (True, False)]
        for d in decs:
            d.prime(256)
        tok = torch.randint(0, 2000, (bsz, 1), device="cuda")
        for s in range(3):
            outs = []
            for d in decs:
                d.step(tok)
                outs.append(d.last_logits.float().clone())
            err = float((outs[0] - outs[1]).abs().max() / outs[1].abs().max())
            assert err <= 2e-2, f"bsz {bsz} step {s}: {err:.3e}"
        decs[0].check()


def test_chain_position_outside_cache_sets_error():
    m = _tiny(dict(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_key_value_heads=2), 1)
    d = QL.GraphDecoder(m, 1, 16, use_graph=False, chain=True)
    d.prime(4)
    d.pos.fill_(16)                       # behind the host-side counter's back
    d.step()
    with pytest.raises(RuntimeError):
        d.check()
