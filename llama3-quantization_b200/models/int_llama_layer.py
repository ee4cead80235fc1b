"""Drop-in for the reference's ``models/int_llama_layer.py``: QuantLlamaMLP (lines 20-45),
QuantLlamaAttention (48-188) and QuantLlamaDecoderLayer (190-344), the callers of the quantised-linear
hot path.

Same classes, constructor signatures, attribute names and methods (``set_quant_state``,
``smooth_and_quant_inplace`` / ``smooth_and_quant_temporary`` / ``clear_temp_variable``,
``register_scales_and_zeros``, ``lwc_parameters`` / ``let_parameters`` / ``omni_parameters`` /
``omni_state_dict``), so ``quant/omniquant.py`` and ``main.py`` can build blocks from here.  What is
B200-specific sits underneath:

* the three projections that read the same normalised activation (q, k, v) and the two of the MLP
  (gate, up) run as ONE fused dequant-GEMM launch each on the concatenation of their packed weights
  (concatenating along N is exact: quantisation groups run along K), so a decoder block is four GEMM
  launches instead of seven -- the k/v shapes are sub-2-microsecond at the HBM roofline and cannot be
  efficient as launches of their own ("fuse q|k|v and gate|up");
* the per-token activation fake-quant of a shared input runs once (K2), not once per consumer;
* under tensor parallelism (``tp_world > 1``) q/k/v/gate/up hold column shards, o/down row shards, and
  the one collective is the all-reduce after o_proj and down_proj (q100/tp.py).

The attention core between the two ``QuantMatMul``s is the reference's eager formulation (fp32 softmax), with one
exact re-ordering: K and V pass their per-(head, token) quantisers BEFORE ``repeat_kv`` (a quarter of the rows, the
same bits).  With ``fast_attention = True`` the core is ``scaled_dot_product_attention`` on the (fake-quantised, when
activation quantisation is on) Q / K / V with ``enable_gqa`` -- no ``[B, heads, T, T]`` tensor, no expanded K / V (a
library kernel; the prefill attention core is SURVEY section 8 row f3, the single-token form is K8).

The rotary embedding follows transformers 4.37 (the version the reference pins): ``rotary_emb(x, seq_len)``
returns the ``[seq_len, head_dim]`` cos/sin cache which ``apply_rotary_pos_emb`` indexes by position.  If
``org_module.rotary_emb`` exists it is deep-copied like the reference does (int_llama_layer.py:70); otherwise
one is built from ``config.rope_theta``.

LET (``self.let = True``: learnable equivalent transformation) needs ``models/transformation.py``, a
calibration-side caller that is out of scope here; those branches import it from the reference checkout
(``Q100_REF_PATH`` / ``sys.path``) and raise if it is absent.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from q100 import ops as _ops
from q100 import tp as _tp
from quant.int_linear import QuantLinear
from quant.int_matmul import QuantMatMul
from quant.omni_norm import OmniLlamaRMSNorm


def _act_fn(name: str):
    if name in ("silu", "swish"):
        return F.silu
    if name == "gelu":
        return F.gelu
    if name == "relu":
        return F.relu
    from transformers.activations import ACT2FN
    return ACT2FN[name]


# ------------------------------------------------------------------------------------ rotary (4.37 form)
class LlamaRotaryEmbedding(nn.Module):
    """cos/sin cache ``[seq_len, dim]``; ``forward(x, seq_len)`` like transformers 4.37."""

    def __init__(self, dim, max_position_embeddings=8192, base=500000.0, device=None):
        super().__init__()
        self.dim, self.base = dim, base
        self.max_seq_len_cached = 0
        self._device = device
        self.register_buffer("cos_cached", torch.empty(0), persistent=False)
        self.register_buffer("sin_cached", torch.empty(0), persistent=False)

    def _build(self, seq_len, device):
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.int64).float().to(device) / self.dim))
        t = torch.arange(seq_len, device=device, dtype=torch.int64).type_as(inv_freq)
        freqs = torch.outer(t, inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        self.cos_cached, self.sin_cached = emb.cos(), emb.sin()
        self.max_seq_len_cached = seq_len

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached or self.cos_cached.device != x.device:
            self._build(max(seq_len, 2048), x.device)
        return self.cos_cached[:seq_len].to(dtype=x.dtype), self.sin_cached[:seq_len].to(dtype=x.dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids):
    cos = cos[position_ids].unsqueeze(1)
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    batch, num_key_value_heads, slen, head_dim = hidden_states.shape
    if n_rep == 1:
        return hidden_states
    hidden_states = hidden_states[:, :, None, :, :].expand(batch, num_key_value_heads, n_rep, slen, head_dim)
    return hidden_states.reshape(batch, num_key_value_heads * n_rep, slen, head_dim)


class StaticKV:
    """Extension for CUDA-graph decode: a preallocated ``[bsz, kv_heads, max_len, head_dim]`` K/V pair written
    in place at ``position_ids`` (a device tensor), so a decode step has fixed shapes and can be captured.
    Passing one as ``past_key_value`` replaces the reference's ``torch.cat`` cache (int_llama_layer.py:130-133);
    the caller supplies an attention mask over ``max_len`` that hides the slots not written yet."""

    def __init__(self, bsz, kv_heads, max_len, head_dim, dtype, device):
        self.k = torch.zeros(bsz, kv_heads, max_len, head_dim, dtype=dtype, device=device)
        self.v = torch.zeros_like(self.k)

    def update(self, k_new, v_new, position_ids):
        idx = position_ids.reshape(-1)[-k_new.shape[2]:]
        self.k.index_copy_(2, idx, k_new.to(self.k.dtype))
        self.v.index_copy_(2, idx, v_new.to(self.v.dtype))
        return self.k, self.v


# ------------------------------------------------------------------------------------ fused launches
from q100._lib import GEMM_X_GATE_UP as _GEMM_X_GATE_UP


class _FusedLinears:
    """Several QuantLinears that consume the same input, run as one dequant-GEMM on the concatenation
    (along N) of their packed weights.  The concatenation is cached on the identity of the parts."""

    def __init__(self):
        self._key = None
        self._pw = None

    def _concat(self, parts):
        p0 = parts[0]
        zeros = None if p0.zeros is None else torch.cat([p.zeros for p in parts], dim=0)
        return _ops.PackedWeight(torch.cat([p.packed for p in parts], dim=0),
                                 torch.cat([p.meta for p in parts], dim=1).contiguous(),
                                 torch.cat([p.scales for p in parts], dim=0), zeros,
                                 p0.n_bits, p0.group, sum(p.N for p in parts), p0.K, p0.disable_zero_point)

    def packed(self, mods, x: torch.Tensor):
        """(concatenated PackedWeight, per-module N list) for ``mods`` on an input like ``x``, or None when the
        modules cannot be fused (then the caller runs them one by one exactly like the reference)."""
        if not x.is_cuda or x.dtype not in (torch.float16, torch.bfloat16):
            return None
        parts = []
        m0 = mods[0]
        for m in mods:
            if m.bias is not None or m.use_act_quant != m0.use_act_quant \
                    or m.disable_input_quant != m0.disable_input_quant:
                return None
            p = m.packed_operand(x)
            if p is None or p.dtype != x.dtype or p.N % _ops.TILE_ROWS != 0:
                return None
            parts.append(p)
        p0 = parts[0]
        for p in parts[1:]:
            if (p.K, p.n_bits, p.group, p.disable_zero_point) != (p0.K, p0.n_bits, p0.group, p0.disable_zero_point):
                return None
        key = tuple((p.packed.data_ptr(), p.packed._version, p.meta.data_ptr()) for p in parts)
        if key != self._key:
            self._pw = self._concat(parts)
            self._key = key
        return self._pw, [p.N for p in parts]

    def __call__(self, mods, x: torch.Tensor, split: bool = True):
        """Returns the outputs of ``mods`` on ``x`` (a tuple; the unsplit ``[..., sum N]`` tensor with
        ``split=False``), or None when the modules cannot be fused."""
        got = self.packed(mods, x)
        if got is None:
            return None
        pw, ns = got
        m0 = mods[0]
        if m0.use_act_quant and not m0.disable_input_quant:
            aq0 = m0.act_quantizer
            same = all(m.act_quantizer.n_bits == aq0.n_bits and m.act_quantizer.enable == aq0.enable for m in mods)
            if same and (x.numel() // x.shape[-1] <= 15 or aq0.n_bits == 8) and aq0.stat_group is None:
                # decode shapes: activation quantisation as the prologue of the integer GEMM (K3c), one launch;
                # W8A8 per-channel at prefill sizes: codes x codes on the integer tensor cores (K3d)
                saved = m0.bias
                y = m0._fused_act_gemm(x, pw, None) if saved is None else None
                if y is not None:
                    return y.split(ns, dim=-1) if split else y
            xq = m0.act_quantizer(x)                  # same input, same per-token statistics for every consumer
            for m in mods[1:]:
                aq, a0 = m.act_quantizer, m0.act_quantizer
                if aq.n_bits != a0.n_bits or aq.enable != a0.enable:
                    return None
                if aq.n_bits < 16 and aq.enable:
                    aq.scale, aq.round_zero_point = a0.scale, a0.round_zero_point
            x = xq
        y = _ops.gemm_wna16(x, pw)
        return y.split(ns, dim=-1) if split else y


class QuantLlamaMLP(nn.Module):
    def __init__(self, org_module: nn.Module, hidden_size: int, intermediate_size: int, hidden_act: str,
                 args=None):
        super().__init__()
        self.gate_proj = QuantLinear(org_module.gate_proj, args.weight_quant_params, args.act_quant_params)
        self.down_proj = QuantLinear(org_module.down_proj, args.weight_quant_params, args.act_quant_params)
        self.up_proj = QuantLinear(org_module.up_proj, args.weight_quant_params, args.act_quant_params)
        self.act_fn = _act_fn(hidden_act)
        self._gate_up = _FusedLinears()
        self.tp_group = None

    def forward(self, x):
        out = None
        if self.act_fn is F.silu and not torch.is_grad_enabled():
            # inference: gate | up stays one tensor and act_fn(gate) * up is K6 (or the down_proj GEMM's x staging at
            # decode shapes) instead of two strided element-wise passes
            gu = self._gate_up((self.gate_proj, self.up_proj), x, split=False)
            if gu is not None:
                out = self.down_proj.forward_gate_up(gu)
        if out is None:
            fused = self._gate_up((self.gate_proj, self.up_proj), x)
            if fused is not None:
                gate, up = fused
            else:
                gate, up = self.gate_proj(x), self.up_proj(x)
            out = self.down_proj(self.act_fn(gate) * up)
        if getattr(self, "tp_world", 1) > 1 and not self.down_proj.reduced:
            out = _tp.reduce_output(out, _tp.ROW, self.tp_group)
        return out


class QuantLlamaAttention(nn.Module):
    """Multi-headed attention with quantised projections (reference int_llama_layer.py:48-188)."""

    def __init__(self, org_module: nn.Module, config, args=None):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.max_position_embeddings = config.max_position_embeddings
        if (self.head_dim * self.num_heads) != self.hidden_size:
            raise ValueError(
                f"hidden_size must be divisible by num_heads (got `hidden_size`: {self.hidden_size}"
                f" and `num_heads`: {self.num_heads})."
            )
        rot = getattr(org_module, "rotary_emb", None)
        if rot is not None:
            self.rotary_emb = copy.deepcopy(rot)
        else:
            self.rotary_emb = LlamaRotaryEmbedding(self.head_dim, self.max_position_embeddings,
                                                   float(getattr(config, "rope_theta", 500000.0)))
        self.k_proj = QuantLinear(org_module.k_proj, args.weight_quant_params, args.act_quant_params)
        self.v_proj = QuantLinear(org_module.v_proj, args.weight_quant_params, args.act_quant_params)
        self.q_proj = QuantLinear(org_module.q_proj, args.weight_quant_params, args.act_quant_params)
        self.o_proj = QuantLinear(org_module.o_proj, args.weight_quant_params, args.act_quant_params)
        self.qkt_matmul = QuantMatMul(args.q_quant_params, args.k_quant_params, matmul_func=torch.matmul)
        self.pv_matmul = QuantMatMul(args.p_quant_params, args.v_quant_params, matmul_func=torch.matmul)
        self.use_weight_quant = False
        self.use_act_quant = False
        self.fast_attention = False
        self._qkv = _FusedLinears()
        # tensor parallel: this rank's share of the heads (set by q100.llama when tp_world > 1)
        self.tp_world = 1
        self.tp_group = None

    def _shape(self, tensor: torch.Tensor, seq_len: int, bsz: int):
        return tensor.view(bsz, seq_len, self.num_heads, self.head_dim).transpose(1, 2).contiguous()

    def forward(
        self,
        hidden_states: torch.Tensor,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.LongTensor] = None,
        past_key_value: Optional[Tuple[torch.Tensor]] = None,
        output_attentions: bool = False,
        use_cache: bool = False,
    ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[Tuple[torch.Tensor]]]:
        bsz, q_len, _ = hidden_states.size()
        nh = self.num_heads // self.tp_world
        nkv = max(self.num_key_value_heads // self.tp_world, 1)

        static_kv = past_key_value if isinstance(past_key_value, StaticKV) else None
        kv_seq_len = q_len
        if static_kv is not None:
            kv_seq_len = static_kv.k.shape[-2]
        elif past_key_value is not None:
            kv_seq_len += past_key_value[0].shape[-2]
        if position_ids is None:
            position_ids = torch.arange(kv_seq_len - q_len, kv_seq_len, device=hidden_states.device).unsqueeze(0)

        hd = self.head_dim
        qkv = None
        if (hidden_states.is_cuda and hidden_states.dtype in (torch.float16, torch.bfloat16) and hd % 16 == 0
                and not torch.is_grad_enabled() and not torch.is_autocast_enabled() and position_ids.dim() == 2):
            qkv = self._qkv((self.q_proj, self.k_proj, self.v_proj), hidden_states, split=False)
        if qkv is not None and qkv.shape[-1] == (nh + 2 * nkv) * hd:
            # inference outside autocast: rotary + the three head-major transposes as one launch (K5p); q is rotated
            # in fp32 and rounded once -- what the fp32 cast of int_llama_layer.py:117 followed by the cast back below
            # computes -- k with every op rounded to the 16-bit type
            rc = getattr(self, "_rope16", None)
            if rc is None or rc[0].dtype != qkv.dtype or rc[0].shape[0] < kv_seq_len or rc[0].device != qkv.device:
                cos, sin = self.rotary_emb(qkv, seq_len=kv_seq_len)
                rc = self._rope16 = (cos.to(qkv.dtype).contiguous(), sin.to(qkv.dtype).contiguous())
            query_states, key_states, value_states = _ops.rope_qkv_prefill(qkv, rc[0], rc[1], position_ids, nh, nkv)
        else:
            if qkv is not None:
                q, k, v = qkv.split([nh * hd, nkv * hd, nkv * hd], dim=-1)
            else:
                fused = self._qkv((self.q_proj, self.k_proj, self.v_proj), hidden_states)
                if fused is not None:
                    q, k, v = fused
                else:
                    q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
            query_states = q.view(bsz, q_len, nh, self.head_dim).transpose(1, 2).type(torch.float32)
            key_states = k.view(bsz, q_len, nkv, self.head_dim).transpose(1, 2)
            value_states = v.view(bsz, q_len, nkv, self.head_dim).transpose(1, 2)
            cos, sin = self.rotary_emb(value_states, seq_len=kv_seq_len)
            query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin, position_ids)

        if static_kv is not None:
            key_states, value_states = static_kv.update(key_states, value_states, position_ids)
        elif past_key_value is not None:
            key_states = torch.cat([past_key_value[0], key_states], dim=2)
            value_states = torch.cat([past_key_value[1], value_states], dim=2)
        if static_kv is None:
            past_key_value = (key_states, value_states) if use_cache else None

        if not torch.is_autocast_enabled():
            # the reference leaves Q in fp32 (int_llama_layer.py:117) and relies on autocast to bring both
            # matmul operands to the model dtype; outside autocast do that cast here
            query_states = query_states.to(key_states.dtype)
        n_rep = nh // nkv

        # K / V are quantised per (head, token) row over head_dim, and repeat_kv only copies whole heads: quantising the
        # nkv heads BEFORE the GQA expansion gives bit for bit the rows the reference gets from quantising the expanded
        # tensor (int_llama_layer.py:136-141), on a quarter of the rows
        if self.sdpa_route() and not output_attentions:
            query_states = self.qkt_matmul.quant_x1(query_states)
            key_states = self.qkt_matmul.quant_x2(key_states)
            value_states = self.pv_matmul.quant_x2(value_states)
            attn_output = F.scaled_dot_product_attention(
                query_states, key_states, value_states, attn_mask=attention_mask,
                is_causal=attention_mask is None and q_len > 1, enable_gqa=n_rep > 1)
            attn_weights = None
        else:
            key_states = repeat_kv(self.qkt_matmul.quant_x2(key_states), n_rep)
            query_states = self.qkt_matmul.quant_x1(query_states)
            attn_weights = self.qkt_matmul(query_states, key_states.transpose(2, 3)) / math.sqrt(self.head_dim)
            if attn_weights.size() != (bsz, nh, q_len, kv_seq_len):
                raise ValueError(
                    f"Attention weights should be of size {(bsz, nh, q_len, kv_seq_len)}, but is"
                    f" {attn_weights.size()}"
                )
            if attention_mask is not None:
                if attention_mask.size() != (bsz, 1, q_len, kv_seq_len):
                    raise ValueError(
                        f"Attention mask should be of size {(bsz, 1, q_len, kv_seq_len)}, but is {attention_mask.size()}"
                    )
                attn_weights = attn_weights + attention_mask
                attn_weights = torch.max(attn_weights, torch.tensor(torch.finfo(attn_weights.dtype).min))
            # upcast attention to fp32
            attn_weights = nn.functional.softmax(attn_weights, dim=-1, dtype=torch.float32).to(query_states.dtype)
            attn_weights = self.pv_matmul.quant_x1(attn_weights)
            value_states = repeat_kv(self.pv_matmul.quant_x2(value_states), n_rep)
            attn_output = self.pv_matmul(attn_weights, value_states)

        if attn_output.size() != (bsz, nh, q_len, self.head_dim):
            raise ValueError(
                f"`attn_output` should be of size {(bsz, nh, q_len, self.head_dim)}, but is"
                f" {attn_output.size()}"
            )
        attn_output = attn_output.transpose(1, 2)
        attn_output = attn_output.reshape(bsz, q_len, nh * self.head_dim)
        attn_output = self.o_proj(attn_output)
        if self.tp_world > 1 and not self.o_proj.reduced:
            attn_output = _tp.reduce_output(attn_output, _tp.ROW, self.tp_group)
        if not output_attentions:
            attn_weights = None
        return attn_output, attn_weights, past_key_value

    def sdpa_route(self) -> bool:
        """``fast_attention`` (extension) sends the attention core through ``scaled_dot_product_attention`` -- on the
        fake-quantised Q / K / V when activation quantisation is on -- unless the softmax output has a quantiser of its
        own to pass (the reference's ``p_quant_params`` are 16 bits: identity)."""
        pv = self.pv_matmul
        pq = pv.x1_quantizer
        return bool(self.fast_attention) and (not pv.use_act_quant or pq.n_bits >= 16 or not pq.enable)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        # setting weight quantization here does not affect actual forward pass
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, (QuantLinear, QuantMatMul)):
                m.set_quant_state(weight_quant, act_quant)


def _transformation():
    """models/transformation.py of the reference (LET smoothing; a calibration-side caller, out of scope)."""
    try:
        import importlib
        return importlib.import_module("models.transformation")
    except Exception as ex:  # pragma: no cover - depends on the reference checkout being importable
        raise NotImplementedError(
            "LET (let=True) needs the reference's models/transformation.py on sys.path / Q100_REF_PATH") from ex


class QuantLlamaDecoderLayer(nn.Module):
    def __init__(self, config, ori_layer, args):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = QuantLlamaAttention(org_module=ori_layer.self_attn, config=config, args=args)
        self.mlp = QuantLlamaMLP(
            org_module=ori_layer.mlp,
            hidden_size=self.hidden_size,
            intermediate_size=config.intermediate_size,
            hidden_act=config.hidden_act,
            args=args,
        )
        self.input_layernorm = OmniLlamaRMSNorm(ori_layer.input_layernorm,
                                                eps=ori_layer.input_layernorm.variance_epsilon)
        self.post_attention_layernorm = OmniLlamaRMSNorm(ori_layer.post_attention_layernorm,
                                                         eps=ori_layer.post_attention_layernorm.variance_epsilon)
        self.let = False

    def forward(
        self,
        hidden_states: torch.Tensor,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.LongTensor] = None,
        past_key_value: Optional[Tuple[torch.Tensor]] = None,
        output_attentions: Optional[bool] = False,
        use_cache: Optional[bool] = False,
    ):
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, self_attn_weights, present_key_value = self.self_attn(
            hidden_states=hidden_states,
            attention_mask=attention_mask,
            position_ids=position_ids,
            past_key_value=past_key_value,
            output_attentions=output_attentions,
            use_cache=use_cache,
        )
        # residual add + post-attention norm (one launch, K4, for 16-bit CUDA inference; the two eager steps otherwise)
        residual, hidden_states = self.post_attention_layernorm.add_norm(hidden_states, residual)
        hidden_states = self.mlp(hidden_states)
        hidden_states = residual + hidden_states

        outputs = (hidden_states,)
        if output_attentions:
            outputs += (self_attn_weights,)
        if use_cache:
            outputs += (present_key_value,)
        return outputs

    # ------------------------------------------------------------------ fused decode step (extension)
    @torch.no_grad()
    def _decode_attention(self, qkv, rc, position, kv, nh, nkv, attention_mask, err_flag, causal_cache):
        """rotary + cache append + single-token attention on the fused q|k|v row: ONE launch (K8, q100_attn_decode) when
        the caller states that ``attention_mask`` is the plain static-cache mask (``causal_cache``: every slot up to the
        position visible), else K5 + scaled_dot_product_attention with the mask.  Returns [bsz, nh * head_dim]."""
        import os
        attn = self.self_attn
        hd = attn.head_dim
        act_q = attn.qkt_matmul.use_act_quant or attn.pv_matmul.use_act_quant
        k8_max = int(os.environ.get("Q100_K8_MAX_BATCH", "64"))
        if not act_q and causal_cache and qkv.shape[0] <= k8_max and _ops.attn_decode_supported(nh, nkv, hd) \
                and not os.environ.get("Q100_NO_K8"):
            return _ops.attn_decode(qkv, rc[0], rc[1], position, kv.k, kv.v, nh, nkv, err_flag)
        q = _ops.rope_kv_decode(qkv, rc[0], rc[1], position, kv.k, kv.v, nh, nkv, err_flag)
        k, v = kv.k, kv.v
        if act_q:
            # W8A8 / W4A4: the cache holds what the reference's cache holds (unquantised K / V); Q and the kv heads pass
            # their per-(head, token) quantisers on the way into the product, as in forward() (identity when off)
            q = attn.qkt_matmul.quant_x1(q)
            k = attn.qkt_matmul.quant_x2(k)
            v = attn.pv_matmul.quant_x2(v)
        att = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, enable_gqa=nh != nkv)
        return att.reshape(q.shape[0], nh * hd)

    def decode_step_fused(self, x, residual, kv: "StaticKV", position, attention_mask, err_flag=None,
                          causal_cache: bool = False, x_partial: bool = False):
        """One new token per sequence through this block in 4 GEMM launches + 4 small kernels + the
        attention core, on a static KV cache (CUDA-graph friendly).  Same arithmetic as ``forward`` with
        ``past_key_value=kv`` (tests/test_block.py compares the two); what moves is where the residual add
        happens: this block receives the previous block's MLP output ``x`` and residual stream ``residual``
        separately, adds them inside its first norm kernel (K4), and returns its own (mlp_out, residual)
        pair.  With activation quantisation on (W8A8 / W4A4) the GEMMs are K3c launches (the per-token quantisation
        as their prologue) and the attention core is SDPA on the fake-quantised Q / K / V.  Returns None when the
        block cannot take this route (a quantiser of its own on the softmax output, dense weights, non-SiLU
        activation, fp32): the caller then uses ``forward``."""
        attn, mlp = self.self_attn, self.mlp
        if x.dtype not in (torch.float16, torch.bfloat16) or mlp.act_fn is not F.silu \
                or self.input_layernorm.use_temporary_parameter or self.post_attention_layernorm.use_temporary_parameter:
            return None
        # (the quantisation state lives on the QuantLinear / QuantMatMul modules: like the reference's, this layer's
        # set_quant_state does not touch the attention module's own ``use_act_quant`` attribute)
        pq = attn.pv_matmul.x1_quantizer
        if attn.pv_matmul.use_act_quant and pq.n_bits < 16 and pq.enable:
            return None
        bsz = x.shape[0]
        nh = attn.num_heads // attn.tp_world
        nkv = max(attn.num_key_value_heads // attn.tp_world, 1)
        folded = self._decode_step_folded(x, residual, kv, position, attention_mask, err_flag, nh, nkv, causal_cache)
        if folded is not None:
            return folded
        # tensor parallel: the all-reduce after a row-parallel linear runs inside the norm kernel that consumes it (K7n,
        # q100.peer.ReduceNorm) -- ``x_partial``: the previous block left its down_proj output un-reduced for this one
        self.last_output_partial = False
        rn = None
        if attn.tp_world > 1 and getattr(mlp, "tp_world", 1) == attn.tp_world and not torch.is_grad_enabled():
            from q100 import peer as _peer
            rn = _peer.get_reduce_norm(attn.tp_group, x.shape[-1])
            if rn is not None and not rn.fits(x):
                rn = None
        if x_partial and (rn is None or residual is None):
            return None
        ln1 = self.input_layernorm
        if x_partial:
            h0, n1 = rn(x, ln1.weight, ln1.variance_epsilon, residual)
        else:
            h0, n1 = _ops.add_rmsnorm(x, ln1.weight, ln1.variance_epsilon, residual)
        qkv = attn._qkv((attn.q_proj, attn.k_proj, attn.v_proj), n1, split=False)
        if qkv is None:
            return None
        rc = getattr(attn, "_rope16", None)
        if rc is None or rc[0].dtype != x.dtype or rc[0].shape[0] < kv.k.shape[-2] or rc[0].device != x.device:
            cos, sin = attn.rotary_emb(kv.v, seq_len=kv.k.shape[-2])
            rc = attn._rope16 = (cos.to(x.dtype).contiguous(), sin.to(x.dtype).contiguous())
        att = self._decode_attention(qkv, rc, position, kv, nh, nkv, attention_mask, err_flag, causal_cache)
        o = attn.o_proj(att.reshape(bsz, 1, nh * attn.head_dim))
        ln2 = self.post_attention_layernorm
        if rn is not None and not attn.o_proj.reduced:
            h1, n2 = rn(o.contiguous(), ln2.weight, ln2.variance_epsilon, h0)
        else:
            if attn.tp_world > 1 and not attn.o_proj.reduced:
                o = _tp.reduce_output(o, _tp.ROW, attn.tp_group)
            h1, n2 = _ops.add_rmsnorm(o, ln2.weight, ln2.variance_epsilon, h0)
        gu = mlp._gate_up((mlp.gate_proj, mlp.up_proj), n2, split=False)
        if gu is None:
            return None
        d = mlp.down_proj.forward_gate_up(gu)              # SiLU * up inside the down_proj GEMM's x staging
        if getattr(mlp, "tp_world", 1) > 1 and not mlp.down_proj.reduced:
            if rn is not None:
                self.last_output_partial = True            # the consumer's norm kernel reduces it (x_partial=True)
                return d.contiguous(), h1
            d = _tp.reduce_output(d, _tp.ROW, mlp.tp_group)
        return d, h1

    def _decode_step_folded(self, x, residual, kv, position, attention_mask, err_flag, nh, nkv, causal_cache=False):
        """The decode step in 4 GEMM launches + K5 + the attention core: both RMSNorms run inside the x staging of the
        GEMM that consumes them and both residual adds inside the epilogue of the GEMM that produces the summand
        (q100_gemm_wna16_block: the row-owner kernel K3r, batch <= 2, W4 g128, one GPU).  Takes and returns the
        residual stream: ``x`` is this block's input hidden state (``residual`` None), the result is (h_out, None).
        None when the block does not qualify -- the caller then runs the K4 route."""
        import os
        attn, mlp = self.self_attn, self.mlp
        bsz = x.shape[0]
        if residual is not None or bsz > 2 or x.numel() > 16384 or attn.tp_world > 1 or getattr(mlp, "tp_world", 1) > 1 \
                or os.environ.get("Q100_NO_FOLD") or torch.is_grad_enabled():
            return None
        h = x.reshape(bsz, -1)
        got_qkv = attn._qkv.packed((attn.q_proj, attn.k_proj, attn.v_proj), h)
        got_gu = mlp._gate_up.packed((mlp.gate_proj, mlp.up_proj), h)
        if got_qkv is None or got_gu is None or attn.q_proj.use_act_quant or mlp.gate_proj.use_act_quant:
            return None
        wqkv, wgu = got_qkv[0], got_gu[0]
        wo = attn.o_proj.packed_operand(h)
        wd = mlp.down_proj.packed_operand(h)
        if wo is None or wd is None or attn.o_proj.use_act_quant or mlp.down_proj.use_act_quant \
                or not all(_ops.gemm_block_fusable(bsz, w) and w.dtype == h.dtype for w in (wqkv, wo, wgu, wd)):
            return None
        rc = getattr(attn, "_rope16", None)
        if rc is None or rc[0].dtype != x.dtype or rc[0].shape[0] < kv.k.shape[-2] or rc[0].device != x.device:
            cos, sin = attn.rotary_emb(kv.v, seq_len=kv.k.shape[-2])
            rc = attn._rope16 = (cos.to(x.dtype).contiguous(), sin.to(x.dtype).contiguous())
        ln1, ln2 = self.input_layernorm, self.post_attention_layernorm
        qkv = _ops.gemm_wna16_block(h, wqkv, norm_weight=ln1.weight, eps=ln1.variance_epsilon)
        att = self._decode_attention(qkv, rc, position, kv, nh, nkv, attention_mask, err_flag, causal_cache)
        h1 = _ops.gemm_wna16_block(att, wo, attn.o_proj.bias, residual=h)
        gu = _ops.gemm_wna16_block(h1, wgu, norm_weight=ln2.weight, eps=ln2.variance_epsilon)
        h2 = _ops.gemm_wna16_block(gu, wd, mlp.down_proj.bias, flags=_GEMM_X_GATE_UP, residual=h1)
        attn.o_proj.reduced = mlp.down_proj.reduced = False
        return h2.view(x.shape), None

    # ------------------------------------------------------------------ quantisation state
    def act_quant_active(self) -> bool:
        """True when any linear or attention matmul of the block fake-quantises its input (the state set_quant_state
        leaves on the QuantLinear / QuantMatMul modules -- the attention module's own flag is not part of it)."""
        for m in self.modules():
            if isinstance(m, QuantLinear) and m.use_act_quant and not m.disable_input_quant:
                return True
            if isinstance(m, QuantMatMul) and m.use_act_quant:
                return True
        return False

    def _quant_linears(self):
        return [m for _, m in self.named_modules() if isinstance(m, QuantLinear)]

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        # setting weight quantization here does not affect actual forward pass
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for name, m in self.named_modules():
            if isinstance(m, (QuantLinear, QuantMatMul)):
                m.set_quant_state(weight_quant, act_quant)

    def _let_smooth(self, suffix: str):
        t = _transformation()
        with torch.no_grad():
            for name, module in self.named_parameters():
                if "smooth_scale" in name:
                    module.data = t.truncate_number(module)
        ln_fcs = getattr(t, "smooth_ln_fcs_" + suffix)
        fc_fc = getattr(t, "smooth_fc_fc_" + suffix)
        q_k = getattr(t, "smooth_q_k_" + suffix)
        ln_fcs(self.input_layernorm, [self.self_attn.q_proj, self.self_attn.k_proj, self.self_attn.v_proj],
               self.qkv_smooth_scale, self.qkv_smooth_shift)
        ln_fcs(self.post_attention_layernorm, [self.mlp.up_proj, self.mlp.gate_proj],
               self.fc1_smooth_scale, self.fc1_smooth_shift)
        fc_fc(self.self_attn.v_proj, self.self_attn.o_proj, self.out_smooth_scale, self.out_smooth_shift)
        q_k(self.self_attn.q_proj, self.self_attn.k_proj, self.qkt_smooth_scale)

    def smooth_and_quant_temporary(self):
        if self.let:
            self._let_smooth("temporary")
            self.mlp.down_proj.temp_weight = self.mlp.down_proj.weight
        else:
            for module in self._quant_linears():
                module.temp_weight = module.weight
        for module in self._quant_linears():
            if hasattr(module, "temp_weight"):
                module.temp_weight = module.weight_quantizer(module.temp_weight)
            else:
                module.temp_weight = module.weight_quantizer(module.weight)
            if not hasattr(module, "temp_bias"):
                module.temp_bias = module.bias
            module.use_temporary_parameter = True

    def clear_temp_variable(self):
        for module in self._quant_linears():
            del module.temp_weight
            del module.temp_bias

    @torch.no_grad()
    def smooth_and_quant_inplace(self):
        if self.let:
            self._let_smooth("inplace")
        for module in self._quant_linears():
            module.weight = module.weight_quantizer(module.weight)
            module.use_temporary_parameter = False

    def let_parameters(self, use_shift=True):
        template = "smooth" if use_shift else "smooth_scale"
        return iter([m for n, m in self.named_parameters() if n.find(template) > -1])

    def lwc_parameters(self):
        return iter([m for n, m in self.named_parameters() if n.find('bound_factor') > -1])

    def omni_parameters(self, use_shift=True):
        template = "smooth" if use_shift else "smooth_scale"
        return iter([m for n, m in self.named_parameters()
                     if n.find('bound_factor') > -1 or n.find(template) > -1])

    def omni_state_dict(self, destination=None, prefix='', keep_vars=False):
        if destination is None:
            destination = OrderedDict()
        for name, param in self.named_parameters():
            if name.find('smooth') > -1 or name.find('bound_factor') > -1:
                destination[prefix + name] = param if keep_vars else param.detach()
        return destination

    def register_scales_and_zeros(self):
        for module in self._quant_linears():
            module.weight_quantizer.register_scales_and_zeros()
