"""A LLaMA stack of q100 decoder blocks with seeded random weights, for the whole-model rows of the
benchmark contract: synthetic-token perplexity exactly as
``main.py:122-151`` computes it, prefill tokens/s and per-token decode latency with a KV cache
(methodology of the reference's ``llama.py:409-465``: median per-token latency).

There is no checkpoint and no dataset on the box, so weights are N(0, 0.02) from a per-tensor seed
(``seed * 1000 + layer * 7 + j``; identical on every tensor-parallel rank, each rank keeps its shard) and
tokens are ``randint``.  Blocks are built one at a time, quantised in place (``smooth_and_quant_inplace``,
LET off -- what ``main.py`` leaves behind after calibration) and only the packed codes + side band are
kept: an 8B model is 4.3 GB of W4 codes, a 70B shard at TP=8 is 4.5 GB per GPU.

Tensor parallel layout per block (q100/tp.py): q/k/v/gate/up column shards (heads split across ranks),
o/down row shards, one all-reduce after o_proj and after down_proj.  Embedding, norms and lm_head are
replicated (lm_head is unquantised fp16/bf16 in the reference as well; ``QuantLlamaModel.quantize_head`` is the
opt-in that puts it on packed codes too).
"""
from __future__ import annotations

import math
import types
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tp as _tp
from models.int_llama_layer import QuantLlamaDecoderLayer


@dataclass
class LlamaCfg:
    hidden_size: int
    intermediate_size: int
    num_attention_heads: int
    num_key_value_heads: int
    num_hidden_layers: int
    vocab_size: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_position_embeddings: int = 8192
    hidden_act: str = "silu"


LLAMA3_8B = LlamaCfg(4096, 14336, 32, 8, 32, 128256)
LLAMA3_70B = LlamaCfg(8192, 28672, 64, 8, 80, 128256)

_LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def quant_args(wbits=4, group_size=128, abits=16, symmetric=False, lwc=False, disable_zero_point=False):
    """The parameter dicts main.py:317-353 builds from the command line."""
    act = {"n_bits": abits, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"}
    return types.SimpleNamespace(
        weight_quant_params={"n_bits": wbits, "per_channel_axes": [0], "symmetric": symmetric,
                             "dynamic_method": "per_channel", "group_size": group_size, "lwc": lwc,
                             "disable_zero_point": disable_zero_point},
        act_quant_params=dict(act), q_quant_params=dict(act), k_quant_params=dict(act), v_quant_params=dict(act),
        p_quant_params={"n_bits": 16, "metric": "fix0to1"}, wbits=wbits, abits=abits, group_size=group_size)


class _Norm(nn.Module):
    def __init__(self, w, eps):
        super().__init__()
        self.weight = nn.Parameter(w, requires_grad=False)
        self.variance_epsilon = eps


def _linear(w: torch.Tensor) -> nn.Linear:
    m = nn.Linear(w.shape[1], w.shape[0], bias=False, device="meta")
    m.weight = nn.Parameter(w, requires_grad=False)
    return m


def linear_shapes(cfg: LlamaCfg):
    kv = cfg.num_key_value_heads * (cfg.hidden_size // cfg.num_attention_heads)
    return {"q_proj": (cfg.hidden_size, cfg.hidden_size), "k_proj": (kv, cfg.hidden_size),
            "v_proj": (kv, cfg.hidden_size), "o_proj": (cfg.hidden_size, cfg.hidden_size),
            "gate_proj": (cfg.intermediate_size, cfg.hidden_size),
            "up_proj": (cfg.intermediate_size, cfg.hidden_size),
            "down_proj": (cfg.hidden_size, cfg.intermediate_size)}


def random_block_weights(cfg: LlamaCfg, layer: int, seed: int, device, dtype, rank=0, world=1, group=128,
                         std=0.02):
    """This rank's shard of every linear of block ``layer`` plus the two norm weights."""
    gen = torch.Generator(device=device)
    out = {}
    for j, (name, (n, k)) in enumerate(linear_shapes(cfg).items()):
        gen.manual_seed(seed * 1000 + layer * 7 + j)
        w = (torch.randn(n, k, generator=gen, device=device) * std).to(dtype)
        if world > 1:
            w = _tp.shard_weight(w, _tp.LLAMA_PLAN[name], rank, world, group or k)
        out[name] = w
    gen.manual_seed(seed * 1000 + layer * 7 + 900_000)
    for name in ("input_layernorm", "post_attention_layernorm"):
        out[name] = (1.0 + 0.1 * torch.randn(cfg.hidden_size, generator=gen, device=device)).to(dtype)
    return out


def build_block(cfg: LlamaCfg, weights: dict, args, rank=0, world=1, tp_group=None, quantize=True,
                drop_dense=True) -> QuantLlamaDecoderLayer:
    """QuantLlamaDecoderLayer from dense weights; quantised in place like main.py leaves it."""
    attn = types.SimpleNamespace(**{n: _linear(weights[n]) for n in ("q_proj", "k_proj", "v_proj", "o_proj")})
    mlp = types.SimpleNamespace(**{n: _linear(weights[n]) for n in ("gate_proj", "up_proj", "down_proj")})
    ori = types.SimpleNamespace(self_attn=attn, mlp=mlp,
                                input_layernorm=_Norm(weights["input_layernorm"], cfg.rms_norm_eps),
                                post_attention_layernorm=_Norm(weights["post_attention_layernorm"],
                                                               cfg.rms_norm_eps))
    layer = QuantLlamaDecoderLayer(cfg, ori, args)
    layer.self_attn.tp_world, layer.self_attn.tp_group = world, tp_group
    layer.mlp.tp_world, layer.mlp.tp_group = world, tp_group
    if world > 1:
        # row-parallel layers see a K-slice: per-token activation statistics (abits < 16) and per-channel weight
        # statistics (group_size=None) span all ranks -> one [rows, 2] MIN/MAX all-reduce inside the quantizer
        import torch.distributed as dist
        sg = tp_group if tp_group is not None else dist.group.WORLD
        from . import peer as _peer
        for m in (layer.self_attn.o_proj, layer.mlp.down_proj):
            m.tp_scatter = _peer.get_scatter(tp_group, weights["o_proj"].dtype)     # None unless enable_scatter() ran
            if m.act_quantizer is not None:
                m.act_quantizer.stat_group = sg
            if not m.weight_quantizer.group_size:
                m.weight_quantizer.stat_group = sg
    if quantize and args.weight_quant_params["n_bits"] < 16:
        layer.smooth_and_quant_inplace()
        if drop_dense:
            for m in layer._quant_linears():
                m.release_dense_weight()          # inference only needs the packed twin
    layer.set_quant_state(weight_quant=False, act_quant=args.act_quant_params["n_bits"] < 16)
    return layer


class QuantLlamaModel(nn.Module):
    def __init__(self, cfg: LlamaCfg, layers: List[QuantLlamaDecoderLayer], embed: torch.Tensor,
                 norm_w: torch.Tensor, lm_head: Optional[torch.Tensor]):
        super().__init__()
        self.cfg = cfg
        self.layers = nn.ModuleList(layers)
        self.register_buffer("embed_tokens", embed)
        self.register_buffer("norm_weight", norm_w)
        self.register_buffer("lm_head", lm_head)
        self.head_linear = None          # set by quantize_head(): lm_head as a QuantLinear on packed codes

    # ------------------------------------------------------------------ lm_head
    @torch.no_grad()
    def quantize_head(self, wbits: int = 4, group_size: Optional[int] = 128, keep_dense: bool = False):
        """Opt-in extension: the reference leaves ``lm_head`` in 16 bits (1.05 GB of an 8B model's 5.4 GB decode
        stream at W4).  This wraps it in the drop-in ``QuantLinear`` with the reference's weight-quantizer
        parameters (main.py:317-325), quantises it in place the way ``smooth_and_quant_inplace`` treats a block's
        linears (int_llama_layer.py:307-324: ``weight <- weight_quantizer(weight)``) and keeps the packed twin, so
        the logits come from the same fused dequant-GEMM as every other linear.  ``lm_head`` becomes the
        fake-quantised dense weight (``keep_dense``) or is dropped."""
        from quant.int_linear import QuantLinear
        if self.lm_head is None or self.lm_head.numel() == 0:
            raise RuntimeError("quantize_head: the model has no dense lm_head to quantise")
        wq = quant_args(wbits, group_size).weight_quant_params
        m = QuantLinear(_linear(self.lm_head), wq, {}, disable_input_quant=True)
        m.weight = m.weight_quantizer(m.weight)
        m.set_quant_state(weight_quant=False, act_quant=False)
        if m.weight_quantizer.packed_for(m.weight) is None:
            raise RuntimeError("quantize_head: no packed form for this lm_head (needs CUDA, 16-bit, K % 128 == 0)")
        self.lm_head = m.weight if keep_dense else torch.empty(0, dtype=m.weight.dtype, device=m.weight.device)
        if not keep_dense:
            m.release_dense_weight()
        self.head_linear = m
        return m

    def head(self, x: torch.Tensor) -> torch.Tensor:
        if self.head_linear is not None:
            return self.head_linear(x)
        return F.linear(x, self.lm_head)

    # ------------------------------------------------------------------ construction
    @classmethod
    def random_init(cls, cfg: LlamaCfg, args, device, dtype=torch.float16, seed=0, rank=0, world=1,
                    tp_group=None, n_layers: Optional[int] = None, with_head=True, fast_attention=False,
                    quantize=True, drop_dense=True):
        gen = torch.Generator(device=device).manual_seed(seed * 1000 + 999_001)
        embed = (torch.randn(cfg.vocab_size, cfg.hidden_size, generator=gen, device=device) * 0.02).to(dtype)
        head = (torch.randn(cfg.vocab_size, cfg.hidden_size, generator=gen, device=device) * 0.02).to(dtype) \
            if with_head else None
        norm_w = (1.0 + 0.1 * torch.randn(cfg.hidden_size, generator=gen, device=device)).to(dtype)
        layers = []
        grp = args.weight_quant_params["group_size"]
        for li in range(n_layers if n_layers is not None else cfg.num_hidden_layers):
            w = random_block_weights(cfg, li, seed, device, dtype, rank, world, grp)
            blk = build_block(cfg, w, args, rank, world, tp_group, quantize, drop_dense)
            blk.self_attn.fast_attention = fast_attention
            layers.append(blk)
            del w
        return cls(cfg, layers, embed, norm_w, head)

    # ------------------------------------------------------------------ forward
    def _norm(self, x):
        if x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and self.norm_weight.dtype == x.dtype \
                and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192:
            from . import ops as _ops
            return _ops.add_rmsnorm(x, self.norm_weight, self.cfg.rms_norm_eps)[1]      # K4
        dt = x.dtype
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        return (self.norm_weight * (x * torch.rsqrt(var + self.cfg.rms_norm_eps))).to(dt)

    @staticmethod
    def causal_mask(bsz, q_len, kv_len, dtype, device):
        m = torch.full((q_len, kv_len), torch.finfo(dtype).min, dtype=dtype, device=device)
        i = torch.arange(q_len, device=device).unsqueeze(1)
        j = torch.arange(kv_len, device=device).unsqueeze(0)
        m = m.masked_fill(j <= i + (kv_len - q_len), 0.0)
        return m[None, None].expand(bsz, 1, q_len, kv_len)

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, past=None, use_cache=False):
        """LlamaModel.forward: hidden states after the final norm (+ the new cache)."""
        bsz, q_len = tokens.shape
        past_len = 0 if past is None else past[0][0].shape[-2]
        x = F.embedding(tokens, self.embed_tokens)
        fast = all(l.self_attn.sdpa_route() for l in self.layers)
        mask = None
        if not (fast and (q_len == 1 or past_len == 0)):
            mask = self.causal_mask(bsz, q_len, past_len + q_len, x.dtype, x.device)
        pos = torch.arange(past_len, past_len + q_len, device=x.device).unsqueeze(0)
        new_past = []
        for i, layer in enumerate(self.layers):
            out = layer(x, attention_mask=mask, position_ids=pos,
                        past_key_value=None if past is None else past[i], use_cache=use_cache)
            x = out[0]
            if use_cache:
                new_past.append(out[-1])
        x = self._norm(x)
        return (x, new_past) if use_cache else x

    @torch.no_grad()
    def logits(self, tokens):
        return self.head(self.forward(tokens))

    @torch.no_grad()
    def perplexity(self, testenc: torch.Tensor, seqlen: int):
        """main.py:122-151 on a [1, n] token tensor; returns (ppl, mean nll per token)."""
        nsamples = testenc.numel() // seqlen
        nlls = []
        for i in range(nsamples):
            batch = testenc[:, (i * seqlen): ((i + 1) * seqlen)].to(self.embed_tokens.device)
            lg = self.logits(batch)
            shift_logits = lg[:, :-1, :]
            shift_labels = batch[:, 1:]
            loss = F.cross_entropy(shift_logits.reshape(-1, shift_logits.size(-1)).float(), shift_labels.reshape(-1))
            nlls.append(loss.float() * seqlen)
        mean_nll = torch.stack(nlls).sum() / (nsamples * seqlen)
        return math.exp(mean_nll.item()), mean_nll.item()


class GraphDecoder:
    """One-token decode step of a QuantLlamaModel with fixed shapes (static KV cache, device-side position
    counter), captured in a CUDA graph: embedding -> blocks -> norm -> lm_head -> argmax.  The reference
    measures per-token latency with a growing ``torch.cat`` cache and one Python dispatch per op
    (llama.py:409-465); the arithmetic per token is the same, the launch overhead is not."""

    def __init__(self, model: QuantLlamaModel, bsz: int, max_len: int, use_graph: bool = True,
                 fused: bool = True, chain: Optional[bool] = None):
        """``chain``: run all blocks of a step as ONE persistent launch (q100/chain.py, K9; the model must qualify:
        single GPU, bsz <= 16, weight-only quantisation with group % 128 == 0, head_dim 128).  None = only when
        Q100_CHAIN=1 is set: measured on B200 the chain is correct but not yet faster than the per-kernel route
        (2.9 vs 2.1 ms per token for LLaMA3-8B at batch 1), so it stays opt-in."""
        from models.int_llama_layer import StaticKV
        self.fused = fused
        self.want_chain = chain
        self.chain = None
        # programmatic dependent launch for the decode GEMMs: since K4 / K5 release their successor at once, the GEMM
        # behind them fetches its first weight blocks while they run (Q100_DECODE_PDL=0: off)
        import os as _os
        self.pdl = _os.environ.get("Q100_DECODE_PDL", "1") != "0"
        self.model, self.bsz, self.max_len = model, bsz, max_len
        cfg = model.cfg
        dev, dt = model.embed_tokens.device, model.embed_tokens.dtype
        hd = cfg.hidden_size // cfg.num_attention_heads
        self.caches = []
        for layer in model.layers:
            nkv = max(cfg.num_key_value_heads // layer.self_attn.tp_world, 1)
            self.caches.append(StaticKV(bsz, nkv, max_len, hd, dt, dev))
        self.tok = torch.zeros(bsz, 1, dtype=torch.long, device=dev)
        self.pos = torch.zeros(1, 1, dtype=torch.long, device=dev)
        self.next_tok = torch.zeros(bsz, 1, dtype=torch.long, device=dev)
        self._slots = torch.arange(max_len, device=dev).view(1, 1, 1, max_len)
        self._neg = torch.finfo(dt).min
        self.graph = None
        self.use_graph = use_graph
        # the static cache cannot grow: count the steps on the host and refuse to run past max_len (K5 also checks
        # the device-side position and raises this flag instead of writing outside the cache)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._host_pos = 0
        # tensor parallel: K7n (all-reduce inside the norm kernel) when the peer buffers of K7 are up -- collective
        # set-up, every rank builds its GraphDecoder at the same point
        self._reduce_norm = None
        import os as _os2
        tpw = max((getattr(l.self_attn, "tp_world", 1) for l in model.layers), default=1)
        if tpw > 1 and not _os2.environ.get("Q100_NO_REDUCE_NORM"):
            from . import peer as _peer
            grp = model.layers[0].self_attn.tp_group
            if _peer.get(grp) is not None:
                self._reduce_norm = _peer.enable_reduce_norm(grp, dev, cfg.hidden_size)

    # ------------------------------------------------------------------ the chain route (K9)
    def _build_chain(self):
        """Lay the decode step of every block out as phases of one persistent launch.  Returns None (and the
        per-kernel route stays) when a block does not qualify."""
        import os
        from . import _lib, chain as _chain, ops as _ops
        m, cfg = self.model, self.model.cfg
        dev, dt = m.embed_tokens.device, m.embed_tokens.dtype
        B, H, I = self.bsz, cfg.hidden_size, cfg.intermediate_size
        hd = H // cfg.num_attention_heads
        if os.environ.get("Q100_NO_CHAIN") or B > 16 or hd != 128 or dt not in (torch.float16, torch.bfloat16):
            return None
        packs = []
        xh = torch.zeros(B, H, dtype=dt, device=dev)
        xi = torch.zeros(B, I, dtype=dt, device=dev)
        for layer in m.layers:
            attn, mlp = layer.self_attn, layer.mlp
            if attn.tp_world > 1 or layer.act_quant_active() or mlp.act_fn is not F.silu \
                    or layer.input_layernorm.use_temporary_parameter or layer.post_attention_layernorm.use_temporary_parameter:
                return None
            qkv = attn._qkv.packed((attn.q_proj, attn.k_proj, attn.v_proj), xh)
            gu = mlp._gate_up.packed((mlp.gate_proj, mlp.up_proj), xh)
            o = attn.o_proj.packed_operand(xh) if attn.o_proj.bias is None else None
            d = mlp.down_proj.packed_operand(xi) if mlp.down_proj.bias is None else None
            if qkv is None or gu is None or o is None or d is None:
                return None
            ws = (qkv[0], o, gu[0], d)
            w0 = ws[0]
            if any((w.n_bits, w.group, w.disable_zero_point, w.dtype) != (w0.n_bits, w0.group, w0.disable_zero_point, dt)
                   for w in ws) or w0.group % 128 != 0 or H % 128 != 0:
                return None
            packs.append(ws)
        w0 = packs[0][0]
        ch = _chain.Chain(B, w0.group, w0.n_bits, w0.disable_zero_point, dt, dev)
        nh, nkv = cfg.num_attention_heads, cfg.num_key_value_heads
        parts = _chain.ssq_parts(H)
        b = self._cb = {"h": torch.zeros(B, H, dtype=dt, device=dev), "n1": torch.zeros(B, H, dtype=dt, device=dev),
                        "qkv": torch.zeros(B, (nh + 2 * nkv) * hd, dtype=dt, device=dev),
                        "att": torch.zeros(B, nh * hd, dtype=dt, device=dev),
                        "gu": torch.zeros(B, 2 * I, dtype=dt, device=dev),
                        "ssq_a": torch.zeros(parts, B, dtype=torch.float32, device=dev),
                        "ssq_b": torch.zeros(parts, B, dtype=torch.float32, device=dev)}
        n_layers = len(m.layers)
        for li, (layer, kv, (wqkv, wo, wgu, wd)) in enumerate(zip(m.layers, self.caches, packs)):
            attn = layer.self_attn
            cos, sin = attn.rotary_emb(kv.v, seq_len=self.max_len)
            cos, sin = cos.to(dt).contiguous(), sin.to(dt).contiguous()
            if li == 0:       # block 0 reads K4's output (no partial sums of squares exist for the embedding)
                ch.gemm(b["n1"], wqkv, b["qkv"])
            else:
                ch.gemm(b["h"], wqkv, b["qkv"], x_mode=_lib.X_NORM, norm_weight=layer.input_layernorm.weight,
                        ssq_in=b["ssq_a"], eps=layer.input_layernorm.variance_epsilon)
            at = ch.attention(b["qkv"], cos, sin, self.pos, kv.k, kv.v, b["att"], nh, nkv, combine=False)
            ch.gemm(b["att"], wo, b["h"], x_mode=_lib.X_ATTN, attn=at, epi_mode=_lib.EPI_ADD_SSQ, residual=b["h"],
                    ssq_out=b["ssq_b"])
            ch.gemm(b["h"], wgu, b["gu"], x_mode=_lib.X_NORM, norm_weight=layer.post_attention_layernorm.weight,
                    ssq_in=b["ssq_b"], eps=layer.post_attention_layernorm.variance_epsilon)
            ch.gemm(b["gu"], wd, b["h"], x_mode=_lib.X_GATE_UP, epi_mode=_lib.EPI_ADD_SSQ, residual=b["h"],
                    ssq_out=b["ssq_a"])
        return ch.build()

    def _step_chain(self):
        from . import ops as _ops
        m, b = self.model, self._cb
        torch.index_select(m.embed_tokens, 0, self.tok.view(-1), out=b["h"])
        l0 = m.layers[0]
        _ops.add_rmsnorm(b["h"], l0.input_layernorm.weight, l0.input_layernorm.variance_epsilon, out_norm=b["n1"])
        self.chain.run()
        x = _ops.add_rmsnorm(b["h"], m.norm_weight, m.cfg.rms_norm_eps)[1]
        logits = m.head(x.view(self.bsz, 1, -1))
        self.last_logits = logits
        if logits.dtype in (torch.float16, torch.bfloat16) and logits.shape[-1] % 8 == 0:
            from . import ops as _ops
            _ops.argmax_rows(logits, out=self.next_tok)              # K10: 64 CTAs per row instead of torch's one
        else:
            self.next_tok.copy_(logits.argmax(-1))
        self.pos.add_(1)

    def _step(self):
        if self.chain is not None:
            return self._step_chain()
        m = self.model
        x = F.embedding(self.tok, m.embed_tokens)
        mask = torch.where(self._slots <= self.pos.view(1, 1, 1, 1), 0.0, self._neg).to(x.dtype)
        mask = mask.expand(self.bsz, 1, 1, self.max_len)
        res = None                  # residual stream, kept apart from the last MLP output on the fused route
        partial = False             # tensor parallel: x is this rank's un-reduced down_proj output (K7n reduces it)
        for layer, kv in zip(m.layers, self.caches):
            # `mask` IS the plain static-cache mask: the block may run its attention core as K8 (one launch)
            out = layer.decode_step_fused(x, res, kv, self.pos, mask, self.err, causal_cache=True,
                                          x_partial=partial) if self.fused else None
            if out is None and partial:
                x, partial = self._reduce_partial(x, layer), False
                out = layer.decode_step_fused(x, res, kv, self.pos, mask, self.err, causal_cache=True)
            if out is not None:
                x, res = out
                partial = bool(getattr(layer, "last_output_partial", False))
            else:
                if res is not None:
                    x, res = x + res, None
                x = layer(x, attention_mask=mask, position_ids=self.pos, past_key_value=kv, use_cache=False)[0]
        if res is not None:
            from . import ops as _ops
            rn = self._reduce_norm if partial else None
            if rn is not None and rn.fits(x):
                x = rn(x, m.norm_weight, m.cfg.rms_norm_eps, res)[1]
            else:
                if partial:
                    x = self._reduce_partial(x, m.layers[-1])
                x = _ops.add_rmsnorm(x, m.norm_weight, m.cfg.rms_norm_eps, res)[1]
        else:
            x = m._norm(x)
        logits = m.head(x)
        self.last_logits = logits
        if logits.dtype in (torch.float16, torch.bfloat16) and logits.shape[-1] % 8 == 0:
            from . import ops as _ops
            _ops.argmax_rows(logits, out=self.next_tok)              # K10: 64 CTAs per row instead of torch's one
        else:
            self.next_tok.copy_(logits.argmax(-1))
        self.pos.add_(1)

    @staticmethod
    def _reduce_partial(x, layer):
        from . import tp as _tp
        return _tp.reduce_output(x, _tp.ROW, layer.mlp.tp_group)

    @torch.no_grad()
    def prime(self, start_pos: int = 0):
        """Warm up (builds rotary caches, fused packs, workspaces) and capture."""
        if not 0 <= start_pos or start_pos + 2 >= self.max_len:
            raise ValueError(f"GraphDecoder.prime: start_pos={start_pos} leaves no room in a cache of {self.max_len}")
        self.pos.fill_(start_pos)
        self._host_pos = start_pos
        import os as _os
        want = self.want_chain if self.want_chain is not None else _os.environ.get("Q100_CHAIN") == "1"
        if self.chain is None and self.fused and want:
            self.chain = self._build_chain()
            if self.chain is None and self.want_chain:
                raise RuntimeError("GraphDecoder: this model / batch does not qualify for the chain route")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step_pdl()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.use_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_pdl()
                self.graph = g
            except Exception as ex:                # e.g. a collective that cannot be captured on this setup
                import warnings
                warnings.warn(f"q100 GraphDecoder: CUDA-graph capture failed, decoding eagerly ({ex!r})")
                self.graph = None
                torch.cuda.synchronize()
        self.pos.fill_(start_pos)

    @torch.no_grad()
    def step(self, tokens: Optional[torch.Tensor] = None):
        if self._host_pos >= self.max_len:
            raise RuntimeError(f"GraphDecoder: KV cache of {self.max_len} positions is full")
        self._host_pos += 1
        if tokens is not None:
            self.tok.copy_(tokens)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_pdl()
        return self.next_tok

    def check(self):
        """Raise if a kernel saw a position outside the cache (synchronises)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("GraphDecoder: a decode step ran past the KV cache / rotary table")
        if self.chain is not None:
            self.chain.check()

    def _step_pdl(self):
        # the weights are static during decode, so every GEMM may start streaming them while its
        # predecessor drains (programmatic dependent launch, include/q100.h Q100_GEMM_PDL)
        from . import ops as _ops, _lib
        old = _ops.set_default_gemm_flags(_lib.GEMM_PDL if self.pdl else 0)
        try:
            self._step()
        finally:
            _ops.set_default_gemm_flags(old)
