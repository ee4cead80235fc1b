#!/usr/bin/env python
"""Generate tests/golden/block.npz from the UNMODIFIED reference decoder block (build container only).

    python oracle/make_golden_block.py            # REF_PATH defaults to /root/reference

Builds the reference's ``QuantLlamaDecoderLayer`` (models/int_llama_layer.py:190-344) on CPU in fp32 from
seeded weights, exactly the way main.py / omniquant.py use it after calibration with LET off
(``smooth_and_quant_inplace`` -> ``set_quant_state(False, act_quant)`` -> forward), runs a causal prefill with
``use_cache`` and one decode step on the returned cache, and stores inputs and outputs.

The reference layer was written against transformers 4.37; with the installed 5.x two things are supplied
from outside, as SURVEY.md section 8c describes: ``ori_layer`` is a SimpleNamespace tree of ``nn.Linear`` /
norm stubs with a 4.37-style ``rotary_emb`` (``forward(x, seq_len) -> cos, sin``), and the module-level name
``apply_rotary_pos_emb`` of the reference file is bound to the 4.37 form.  Nothing of the reference is edited.

The script FAILS unless ``oracle/llama_port.py`` reproduces every reference output to fp32 round-off
(1e-5 relative), so the fixture pins the port as well as the product layer.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_PATH = os.environ.get("REF_PATH", "/root/reference")
sys.path.insert(0, REF_PATH)
sys.path.insert(0, HERE)

import llama_port as LP  # noqa: E402

import models.int_llama_layer as ref_layer  # noqa: E402  (reference)

assert os.path.realpath(ref_layer.__file__).startswith(os.path.realpath(REF_PATH)), ref_layer.__file__
ref_layer.apply_rotary_pos_emb = LP.apply_rotary_pos_emb


class Rotary437(nn.Module):
    def __init__(self, head_dim, base):
        super().__init__()
        self.head_dim, self.base = head_dim, base

    def forward(self, x, seq_len=None):
        return LP.rotary_cache(self.head_dim, seq_len, self.base, x.device, x.dtype)


class _Norm(nn.Module):
    def __init__(self, w, eps):
        super().__init__()
        self.weight = nn.Parameter(w.clone(), requires_grad=False)
        self.variance_epsilon = eps


def _lin(w):
    m = nn.Linear(w.shape[1], w.shape[0], bias=False)
    m.weight = nn.Parameter(w.clone(), requires_grad=False)
    return m


def build_reference_layer(cfg: LP.BlockCfg, w):
    attn = types.SimpleNamespace(q_proj=_lin(w["q_proj"]), k_proj=_lin(w["k_proj"]), v_proj=_lin(w["v_proj"]),
                                 o_proj=_lin(w["o_proj"]), rotary_emb=Rotary437(cfg.head_dim, cfg.rope_theta))
    mlp = types.SimpleNamespace(gate_proj=_lin(w["gate_proj"]), up_proj=_lin(w["up_proj"]),
                                down_proj=_lin(w["down_proj"]))
    ori = types.SimpleNamespace(self_attn=attn, mlp=mlp,
                                input_layernorm=_Norm(w["input_layernorm"], cfg.eps),
                                post_attention_layernorm=_Norm(w["post_attention_layernorm"], cfg.eps))
    hf_cfg = types.SimpleNamespace(hidden_size=cfg.hidden, num_attention_heads=cfg.heads,
                                   num_key_value_heads=cfg.kv_heads, max_position_embeddings=8192,
                                   intermediate_size=cfg.inter, hidden_act="silu")
    act = {"n_bits": cfg.a_bits, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"}
    args = types.SimpleNamespace(
        weight_quant_params={"n_bits": cfg.w_bits, "per_channel_axes": [0], "symmetric": cfg.symmetric,
                             "dynamic_method": "per_channel", "group_size": cfg.group, "lwc": False,
                             "disable_zero_point": cfg.disable_zero_point},
        act_quant_params=dict(act), q_quant_params=dict(act), k_quant_params=dict(act), v_quant_params=dict(act),
        p_quant_params={"n_bits": 16, "metric": "fix0to1"})
    layer = ref_layer.QuantLlamaDecoderLayer(hf_cfg, ori, args)
    layer.let = False
    return layer


CASES = {
    # name: (w_bits, group, a_bits, symmetric, disable_zp)
    "fp": (16, None, 16, False, False),
    "w4g128": (4, 128, 16, False, False),
    "w3g128": (3, 128, 16, False, False),
    "w2g64": (2, 64, 16, False, False),
    "w8pc_a8": (8, None, 8, False, False),
    "w4g128_a4": (4, 128, 4, False, False),
    "w4g128_sym": (4, 128, 16, True, False),
}
HIDDEN, INTER, HEADS, KVH = 256, 512, 4, 2
BSZ, T_PREFILL = 2, 24


def main():
    torch.manual_seed(0)
    out = {}
    for ci, (name, (wb, grp, ab, sym, nozp)) in enumerate(CASES.items()):
        cfg = LP.BlockCfg(HIDDEN, INTER, HEADS, KVH, eps=1e-5, rope_theta=500000.0, w_bits=wb, group=grp,
                          a_bits=ab, symmetric=sym, disable_zero_point=nozp)
        w = LP.random_block_weights(cfg, seed=100 + ci)
        g = torch.Generator().manual_seed(200 + ci)
        x0 = torch.randn(BSZ, T_PREFILL, HIDDEN, generator=g)
        x1 = torch.randn(BSZ, 1, HIDDEN, generator=g)
        layer = build_reference_layer(cfg, w)
        act_q = ab < 16
        with torch.no_grad():
            if wb < 16:
                layer.smooth_and_quant_inplace()
            layer.set_quant_state(weight_quant=False, act_quant=act_q)
            mask0 = LP.causal_mask(BSZ, T_PREFILL, T_PREFILL, x0.dtype, x0.device)
            pos0 = torch.arange(T_PREFILL).unsqueeze(0)
            y0, present = layer(x0, attention_mask=mask0, position_ids=pos0, use_cache=True)
            mask1 = LP.causal_mask(BSZ, 1, T_PREFILL + 1, x0.dtype, x0.device)
            pos1 = torch.tensor([[T_PREFILL]])
            y1, present1 = layer(x1, attention_mask=mask1, position_ids=pos1, past_key_value=present,
                                 use_cache=True)
            # the port on the same inputs
            wq = LP.fake_quant_weights(w, cfg)
            if wb < 16:
                for n in LP.LINEAR_NAMES:       # weights: bit-identical to the reference's in-place result
                    ref_w = getattr(layer.self_attn if n[0] in "qkvo" else layer.mlp, n).weight
                    assert torch.equal(ref_w, wq[n]), f"{name}: fake-quantised {n} differs from the reference"
            p0, pres = LP.block_forward(x0, wq, cfg, mask0, pos0, None, True, act_q)
            p1, _ = LP.block_forward(x1, wq, cfg, mask1, pos1, pres, True, act_q)
        for a, b, what in ((p0, y0, "prefill"), (p1, y1, "decode"), (pres[0], present[0], "k cache"),
                           (pres[1], present[1], "v cache")):
            err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-9)
            assert err <= 1e-5, f"{name}: llama_port {what} deviates from the reference by {err:.3e}"
        wsum = float(sum(w[n].double().sum().item() for n in sorted(w)))
        out[f"{name}/cfg"] = np.array([HIDDEN, INTER, HEADS, KVH, wb, grp or 0, ab, int(sym), int(nozp),
                                       100 + ci], dtype=np.int64)
        out[f"{name}/weights_checksum"] = np.array([wsum])
        out[f"{name}/x0"] = x0.numpy()
        out[f"{name}/x1"] = x1.numpy()
        out[f"{name}/y0"] = y0.numpy()
        out[f"{name}/y1"] = y1.numpy()
        out[f"{name}/k_cache"] = present[0].numpy()
        out[f"{name}/v_cache"] = present[1].numpy()
        print(f"{name:12s} ok  |y0| max {y0.abs().max().item():.4f}  port-vs-reference rel err <= 1e-5")
    path = os.path.join(ROOT, "tests", "golden", "block.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
