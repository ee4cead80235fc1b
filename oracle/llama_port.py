"""Torch restatement of the reference's quantised LLaMA decoder block and perplexity loop.
TEST / BASELINE INFRASTRUCTURE ONLY (same rule as oracle/torch_port.py: imported by tests/,
__graft_entry__.smoke() and bench.py's checker legs, never by the product path).

Follows, op for op,
  * ``QuantLlamaMLP.forward``            models/int_llama_layer.py:44-45
  * ``QuantLlamaAttention.forward``      models/int_llama_layer.py:103-179
  * ``QuantLlamaDecoderLayer.forward``   models/int_llama_layer.py:213-262
  * ``OmniLlamaRMSNorm.forward``         quant/omni_norm.py:52-63
  * the perplexity loop of ``evaluate``  main.py:122-151
with the rotary embedding of transformers 4.37 (the version the reference pins; cos/sin cache
``[seq, head_dim]``, ``apply_rotary_pos_emb(q, k, cos, sin, position_ids)`` indexing the cache by
position) written out, because the installed transformers 5.x no longer has that signature.

It is functional: weights come in as a dict of dense tensors that are fake-quantised with
``torch_port.fake_quant`` (the reference op sequence).  ``oracle/make_golden_block.py`` pins it
against the unmodified reference module on CPU (fp32): tests/golden/block.npz.

One dtype rule is ours: the reference casts Q to fp32 (int_llama_layer.py:117) and then calls
``torch.matmul(q_fp32, k_fp16)``, which only runs under autocast; outside autocast this port (and the
product layer) bring Q back to the K/V dtype before the first matmul, which is what autocast does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

import torch_port as TP

LINEAR_NAMES = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def rotary_cache(head_dim: int, seq_len: int, base: float, device, dtype):
    """transformers 4.37 LlamaRotaryEmbedding: cos/sin of shape [seq_len, head_dim]."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float().to(device) / head_dim))
    t = torch.arange(seq_len, device=device, dtype=torch.int64).type_as(inv_freq)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids):
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    b, h, s, d = x.shape
    if n_rep == 1:
        return x
    return x[:, :, None, :, :].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


def rms_norm(x, weight, eps):
    dt = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    return (weight * x).to(dt)


def causal_mask(bsz, q_len, kv_len, dtype, device):
    """transformers 4.37 _prepare_4d_causal_attention_mask: [bsz, 1, q_len, kv_len], 0 / finfo.min."""
    m = torch.full((q_len, kv_len), torch.finfo(dtype).min, dtype=dtype, device=device)
    past = kv_len - q_len
    i = torch.arange(q_len, device=device).unsqueeze(1)
    j = torch.arange(kv_len, device=device).unsqueeze(0)
    m = m.masked_fill(j <= i + past, 0.0)
    return m[None, None].expand(bsz, 1, q_len, kv_len)


class BlockCfg:
    def __init__(self, hidden, inter, heads, kv_heads, eps=1e-5, rope_theta=500000.0, w_bits=4, group=128,
                 a_bits=16, symmetric=False, disable_zero_point=False):
        self.hidden, self.inter, self.heads, self.kv_heads = hidden, inter, heads, kv_heads
        self.head_dim = hidden // heads
        self.eps, self.rope_theta = eps, rope_theta
        self.w_bits, self.group, self.a_bits = w_bits, group, a_bits
        self.symmetric, self.disable_zero_point = symmetric, disable_zero_point


def fake_quant_weights(w: Dict[str, torch.Tensor], cfg: BlockCfg) -> Dict[str, torch.Tensor]:
    """smooth_and_quant_inplace with let=False (int_llama_layer.py:307-324): weight <- quantizer(weight)."""
    out = dict(w)
    if cfg.w_bits < 16:
        for n in LINEAR_NAMES:
            out[n] = TP.fake_quant(w[n], cfg.w_bits, cfg.group, cfg.symmetric, cfg.disable_zero_point)[0]
    return out


def _aq(x, cfg: BlockCfg, act_quant: bool):
    if act_quant and cfg.a_bits < 16:
        return TP.fake_quant(x, cfg.a_bits, None)[0]
    return x


def block_forward(x, w: Dict[str, torch.Tensor], cfg: BlockCfg, attention_mask=None, position_ids=None,
                  past_key_value: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, use_cache=False,
                  act_quant=False, linear=None):
    """QuantLlamaDecoderLayer.forward on (already fake-quantised) dense weights ``w``.  ``linear`` replaces
    ``F.linear`` (the reference's ``fwd_func``, int_linear.py:26): ``linear_fp32_once`` below is the same product with
    another, equally valid, accumulation -- used to measure how far the REFERENCE's own perplexity moves with the GEMM's
    rounding (the yardstick for abits < 16, where one flipped activation code is a discrete step)."""
    lin = linear if linear is not None else F.linear
    bsz, q_len, _ = x.shape
    residual = x
    h = rms_norm(x, w["input_layernorm"], cfg.eps)
    hq = _aq(h, cfg, act_quant)
    q = lin(hq, w["q_proj"]).view(bsz, q_len, cfg.heads, cfg.head_dim).transpose(1, 2).type(torch.float32)
    k = lin(hq, w["k_proj"]).view(bsz, q_len, cfg.kv_heads, cfg.head_dim).transpose(1, 2)
    v = lin(hq, w["v_proj"]).view(bsz, q_len, cfg.kv_heads, cfg.head_dim).transpose(1, 2)
    kv_len = k.shape[-2] + (past_key_value[0].shape[-2] if past_key_value is not None else 0)
    cos, sin = rotary_cache(cfg.head_dim, kv_len, cfg.rope_theta, x.device, v.dtype)
    if position_ids is None:
        position_ids = torch.arange(kv_len - q_len, kv_len, device=x.device).unsqueeze(0)
    q, k = apply_rotary_pos_emb(q, k, cos, sin, position_ids)
    if past_key_value is not None:
        k = torch.cat([past_key_value[0], k], dim=2)
        v = torch.cat([past_key_value[1], v], dim=2)
    present = (k, v) if use_cache else None
    n_rep = cfg.heads // cfg.kv_heads
    k2, v2 = repeat_kv(k, n_rep), repeat_kv(v, n_rep)
    q = _aq(q, cfg, act_quant)
    k2 = _aq(k2, cfg, act_quant)
    if not torch.is_autocast_enabled():
        q = q.to(k2.dtype)                       # see module docstring
    att = torch.matmul(q, k2.transpose(2, 3)) / math.sqrt(cfg.head_dim)
    if attention_mask is not None:
        att = att + attention_mask
        att = torch.max(att, torch.tensor(torch.finfo(att.dtype).min))
    att = F.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
    v2 = _aq(v2, cfg, act_quant)
    o = torch.matmul(att, v2).transpose(1, 2).reshape(bsz, q_len, cfg.hidden)
    o = lin(_aq(o, cfg, act_quant), w["o_proj"])
    x = residual + o
    residual = x
    h = rms_norm(x, w["post_attention_layernorm"], cfg.eps)
    hq = _aq(h, cfg, act_quant)
    m = F.silu(lin(hq, w["gate_proj"])) * lin(hq, w["up_proj"])
    x = residual + lin(_aq(m, cfg, act_quant), w["down_proj"])
    return x, present


def random_block_weights(cfg: BlockCfg, seed: int, dtype=torch.float32, device="cpu", std=0.02):
    """Seeded weights of one block: N(0, std) linears, norm weights 1 + N(0, 0.1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    kv = cfg.kv_heads * cfg.head_dim
    shapes = {"q_proj": (cfg.hidden, cfg.hidden), "k_proj": (kv, cfg.hidden), "v_proj": (kv, cfg.hidden),
              "o_proj": (cfg.hidden, cfg.hidden), "gate_proj": (cfg.inter, cfg.hidden),
              "up_proj": (cfg.inter, cfg.hidden), "down_proj": (cfg.hidden, cfg.inter)}
    w = {n: (torch.randn(s, generator=g) * std).to(dtype).to(device) for n, s in shapes.items()}
    for n in ("input_layernorm", "post_attention_layernorm"):
        w[n] = (1.0 + 0.1 * torch.randn(cfg.hidden, generator=g)).to(dtype).to(device)
    return w


# ------------------------------------------------------------------------------------ whole model
def linear_fp32_once(x, w, b=None):
    """x @ w.T accumulated in fp32 by an fp32 GEMM (no tensor-core split of the K loop) and rounded once."""
    y = torch.matmul(x.to(torch.float32), w.to(torch.float32).t())
    if b is not None:
        y = y + b.to(torch.float32)
    return y.to(x.dtype)


def model_forward(tokens, embed, blocks, final_norm, cfg: BlockCfg, act_quant=False, linear=None):
    """LlamaModel.forward without cache: embedding -> blocks (causal mask) -> final RMSNorm."""
    bsz, q_len = tokens.shape
    x = F.embedding(tokens, embed)
    mask = causal_mask(bsz, q_len, q_len, x.dtype, x.device)
    pos = torch.arange(q_len, device=x.device).unsqueeze(0)
    for w in blocks:
        x, _ = block_forward(x, w, cfg, attention_mask=mask, position_ids=pos, act_quant=act_quant, linear=linear)
    return rms_norm(x, final_norm, cfg.eps)


def perplexity(testenc, seqlen, logits_fn):
    """main.py:122-151: ``logits_fn(batch) -> [1, seqlen, vocab]``; returns (ppl, mean nll per token)."""
    nsamples = testenc.numel() // seqlen
    nlls = []
    for i in range(nsamples):
        batch = testenc[:, (i * seqlen): ((i + 1) * seqlen)]
        logits = logits_fn(batch)
        shift_logits = logits[:, :-1, :]
        shift_labels = batch[:, 1:].to(logits.device)
        loss = F.cross_entropy(shift_logits.reshape(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
        nlls.append(loss.float() * seqlen)
    mean_nll = torch.stack(nlls).sum() / (nsamples * seqlen)
    return torch.exp(mean_nll).item(), mean_nll.item()
