#!/usr/bin/env python
"""Run the reference's UNMODIFIED calibration driver -- quant/omniquant.py::omniquant (the LWC loop,
omniquant.py:261-311; quant/utils.py::smooth_and_quant_temporary / smooth_and_quant_inplace; utils.py's AMP scaler)
-- on one tiny LLaMA-shaped block on the GPU, with the modules underneath chosen by --impl:

    ref             everything from oracle/_ref (the vendored reference): quant/quantizer.py, int_linear.py,
                    int_matmul.py and models/int_llama_layer.py are the reference's own eager-torch code
    dropin          quant/{quantizer,int_linear,int_matmul,omni_norm}.py and models/int_llama_layer.py from
                    llama3-quantization_b200/ (the q100 kernels); omniquant.py, quant/utils.py,
                    models/transformation.py, utils.py still the reference's files
    dropin_refblock as dropin, but models/int_llama_layer.py is ALSO the reference's file: its unmodified
                    decoder block running on the drop-in quant/ package

Test infrastructure (tests/test_reference_callers.py launches it once per impl in a subprocess -- the module names
collide -- and compares what comes back).  Writes a torch file with the per-epoch loss the driver logs, the learned
upbound / lowbound factors, the in-place fake-quantised weights, the registered scales / zeros and the block output.
"""
import argparse
import logging
import os
import re
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "llama3-quantization_b200")
REF = os.path.join(ROOT, "oracle", "_ref")


def _setup_path(impl: str):
    entries = [REF, os.path.join(REF, "stubs"), os.path.join(ROOT, "oracle")]
    if impl == "dropin":
        entries.insert(0, PKG)
    elif impl == "dropin_refblock":
        # only quant/ and q100/ of the drop-in are visible; models/* resolves to the reference's files
        d = tempfile.mkdtemp(prefix="q100_dropin_")
        for name in ("quant", "q100"):
            os.symlink(os.path.join(PKG, name), os.path.join(d, name))
        entries.insert(0, d)
    elif impl != "ref":
        raise SystemExit(f"unknown --impl {impl}")
    sys.path[:0] = entries


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--wbits", type=int, default=4)
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--nsamples", type=int, default=4)
    ap.add_argument("--seqlen", type=int, default=64)
    a = ap.parse_args()
    _setup_path(a.impl)

    import torch
    import torch.nn as nn
    import llama_port as LP                      # rotary helpers of the 4.37 API (oracle/; test infrastructure)
    import models.int_llama_layer as layer_mod
    import quant.quantizer as quantizer_mod
    import quant.omniquant as omniquant_mod

    where = {"int_llama_layer": os.path.realpath(layer_mod.__file__), "quantizer": os.path.realpath(quantizer_mod.__file__),
             "omniquant": os.path.realpath(omniquant_mod.__file__)}
    assert os.path.realpath(omniquant_mod.__file__).startswith(os.path.realpath(REF)), where
    if a.impl == "ref":
        assert os.path.realpath(quantizer_mod.__file__).startswith(os.path.realpath(REF)), where
    else:
        assert os.path.realpath(quantizer_mod.__file__).startswith(os.path.realpath(PKG)), where
    if os.path.realpath(layer_mod.__file__).startswith(os.path.realpath(REF)):
        # the reference block was written against transformers 4.37: bind the 4.37 form of the rotary helper
        # (nothing of the file is edited)
        layer_mod.apply_rotary_pos_emb = LP.apply_rotary_pos_emb

    dev = torch.device("cuda:0")
    H, I, HEADS, KVH, V = 512, 1024, 4, 2, 1000
    hd = H // HEADS

    class Rotary437(nn.Module):
        def forward(self, x, seq_len=None):
            return LP.rotary_cache(hd, seq_len, 500000.0, x.device, x.dtype)

    class Norm(nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = nn.Parameter(w, requires_grad=False)
            self.variance_epsilon = 1e-5

    def lin(n, k, g):
        m = nn.Linear(k, n, bias=False)
        m.weight = nn.Parameter((torch.randn(n, k, generator=g) * 0.02).half(), requires_grad=False)
        return m

    class Attn(nn.Module):
        def __init__(self, g):
            super().__init__()
            self.q_proj, self.k_proj = lin(H, H, g), lin(KVH * hd, H, g)
            self.v_proj, self.o_proj = lin(KVH * hd, H, g), lin(H, H, g)
            self.rotary_emb = Rotary437()

    class MLP(nn.Module):
        def __init__(self, g):
            super().__init__()
            self.gate_proj, self.up_proj, self.down_proj = lin(I, H, g), lin(I, H, g), lin(H, I, g)

    class OriLayer(nn.Module):
        def __init__(self, g):
            super().__init__()
            self.self_attn, self.mlp = Attn(g), MLP(g)
            self.input_layernorm = Norm((1 + 0.1 * torch.randn(H, generator=g)).half())
            self.post_attention_layernorm = Norm((1 + 0.1 * torch.randn(H, generator=g)).half())

        def forward(self, x, attention_mask=None, position_ids=None):      # never reached: omniquant replaces the layer
            raise RuntimeError("fp layer forward is not part of the check")

    class Inner(nn.Module):
        def __init__(self, g):
            super().__init__()
            self.embed_tokens = nn.Embedding(V, H)
            self.embed_tokens.weight = nn.Parameter(torch.randn(V, H, generator=g).half(), requires_grad=False)
            self.layers = nn.ModuleList([OriLayer(g)])
            self.norm = Norm(torch.ones(H).half())

        def forward(self, tokens):
            x = self.embed_tokens(tokens)
            t = tokens.shape[1]
            mask = LP.causal_mask(tokens.shape[0], t, t, x.dtype, x.device)
            pos = torch.arange(t, device=x.device).unsqueeze(0)
            return self.layers[0](x, attention_mask=mask, position_ids=pos)

    class Model(nn.Module):
        def __init__(self, g):
            super().__init__()
            self.model = Inner(g)
            self.config = types.SimpleNamespace(use_cache=False, hidden_size=H, num_attention_heads=HEADS,
                                                num_key_value_heads=KVH, max_position_embeddings=8192,
                                                intermediate_size=I, hidden_act="silu", rope_theta=500000.0)

        def forward(self, tokens):
            return self.model(tokens)

    g = torch.Generator().manual_seed(0)
    model = Model(g)
    lm = types.SimpleNamespace(model=model, device=dev, _device=dev, seqlen=a.seqlen)
    act = {"n_bits": 16, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"}
    outdir = tempfile.mkdtemp(prefix="q100_omni_")
    args = types.SimpleNamespace(
        net="llama-3-tiny", resume=None, nsamples=a.nsamples, epochs=a.epochs, batch_size=1, deactive_amp=False,
        aug_loss=False, let=False, lwc=True, let_lr=5e-3, lwc_lr=1e-2, wd=0.0, alpha=0.5, wbits=a.wbits, abits=16,
        real_quant=False, output_dir=outdir, group_size=a.group, symmetric=False, disable_zero_point=False,
        weight_quant_params={"n_bits": a.wbits, "per_channel_axes": [0], "symmetric": False,
                             "dynamic_method": "per_channel", "group_size": a.group, "lwc": True,
                             "disable_zero_point": False},
        act_quant_params=dict(act), q_quant_params=dict(act), k_quant_params=dict(act), v_quant_params=dict(act),
        p_quant_params={"n_bits": 16, "metric": "fix0to1"})
    dataloader = [(torch.randint(0, V, (1, a.seqlen), generator=g),) for _ in range(a.nsamples)]

    losses = []

    class Grab(logging.Handler):
        def emit(self, rec):
            m = re.search(r"layer (\d+) iter (\d+) loss:([0-9.eE+-]+|nan|inf)", rec.getMessage())
            if m:
                losses.append(float(m.group(3)))

    logger = logging.getLogger("omni_check")
    logger.setLevel(logging.INFO)
    logger.addHandler(Grab())
    torch.manual_seed(0)
    omniquant_mod.omniquant(lm, args, dataloader, None, None, logger)

    qlayer = model.model.layers[0]              # the calibrated QuantLlamaDecoderLayer (moved to the CPU by the driver)
    out = {"impl": a.impl, "where": where, "losses": losses, "weights": {}, "scales": {}, "zeros": {}, "bounds": {}}
    for name, m in qlayer.named_modules():
        if type(m).__name__ == "QuantLinear":
            out["weights"][name] = m.weight.detach().cpu()
            out["scales"][name] = m.weight_quantizer.scales.detach().cpu()
            out["zeros"][name] = m.weight_quantizer.zeros.detach().cpu()
    for name, p in qlayer.named_parameters():
        if "bound_factor" in name:
            out["bounds"][name] = p.detach().float().cpu()
    # the calibrated block on a fixed input (fp16, under autocast like the driver's own propagation step)
    qlayer = qlayer.to(dev)
    gx = torch.Generator().manual_seed(5)
    x = torch.randn(1, a.seqlen, H, generator=gx).half().to(dev)
    mask = LP.causal_mask(1, a.seqlen, a.seqlen, x.dtype, dev)
    pos = torch.arange(a.seqlen, device=dev).unsqueeze(0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        out["y"] = qlayer(x, attention_mask=mask, position_ids=pos)[0].float().cpu()
    torch.save(out, a.out)
    print("OMNIQUANT RUN OK", a.impl, "losses", losses)


if __name__ == "__main__":
    main()
