#!/usr/bin/env python
"""Whole-model rows of the benchmark contract: random-init LLaMA3-8B,
W{2,3,4}A16 g128, synthetic tokens -- perplexity against the torch port on the same weights, prefill tokens/s at
seq 2048, per-token decode latency (median) at batch 1-16 with a static KV cache in a CUDA graph.

    python tools/bench_model.py [--layers 32] [--wbits 4] [--ppl-samples 2] [--batches 1,4,16]

Prints one JSON line.  Under torchrun the blocks are tensor parallel (q100/tp.py)."""
import argparse, json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "llama3-quantization_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def default_args(**kw):
    a = argparse.Namespace(layers=None, wbits=4, abits=16, group=128, dtype="f16", seqlen=2048, ppl_samples=1,
                           ppl_ref=True, batches="1,4,16", decode_steps=48, ctx=256, no_graph=False, model="8b",
                           ppl_layers=None, quant_head=False)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=None, help="blocks (default: all of the model)")
    ap.add_argument("--model", default="8b", choices=["8b", "70b"])
    ap.add_argument("--ppl-layers", type=int, default=None,
                    help="blocks of the stack the perplexity comparison runs on (default: all; the fp16 reference "
                         "of the full 70B does not fit next to the model)")
    ap.add_argument("--wbits", type=int, default=4)
    ap.add_argument("--abits", type=int, default=16)
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--seqlen", type=int, default=2048)
    ap.add_argument("--ppl-samples", type=int, default=2)
    ap.add_argument("--ppl-ref", action="store_true", help="also run the torch port (F.linear on fake-quantised weights)")
    ap.add_argument("--batches", default="1,4,16")
    ap.add_argument("--decode-steps", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=256, help="KV positions already filled when decode is timed")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--quant-head", action="store_true",
                    help="extra rows with lm_head quantised like the block linears (opt-in extension)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if not os.environ.get("Q100_NO_PEER"):
            from q100 import peer
            peer.enable(None, dev)          # K7: decode-size all-reduces over peer memory (falls back to NCCL by itself)
    out = run(a, dev, rank, world)
    if world > 1:
        from q100 import peer as _peer
        if _peer.get(None) is not None:
            _peer.get(None).check()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import gc, threading
        threading.Timer(45.0, lambda: os._exit(0)).start()           # a stuck teardown must not keep the job alive
        from q100 import peer
        peer.shutdown()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def _port_ppl(a, cfg, model, testenc, dev, dt, ppl_layers, head, order_noise=False):
    """Perplexity of the torch port (oracle/llama_port.py, checker leg) on the same seeded weights; with
    ``order_noise`` also the port's mean NLL when its linears run as fp32 GEMMs rounded once."""
    from q100 import llama as QL
    import llama_port as LP
    bc = LP.BlockCfg(cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads,
                     cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta, a.wbits, a.group or None, a.abits)
    blocks = []
    for li in range(ppl_layers):
        w = QL.random_block_weights(cfg, li, 0, dev, dt)
        blocks.append(LP.fake_quant_weights(w, bc))
        del w

    def run_port(linear):
        def logits_fn(batch):
            with torch.no_grad():
                h = LP.model_forward(batch.to(dev), model.embed_tokens, blocks, model.norm_weight, bc,
                                     act_quant=a.abits < 16, linear=linear)
                return torch.nn.functional.linear(h, head).float()
        return LP.perplexity(testenc, a.seqlen, logits_fn)
    ppl_r, nll_r = run_port(None)
    nll_o = run_port(LP.linear_fp32_once)[1] if order_noise else None
    del blocks
    torch.cuda.empty_cache()
    return ppl_r, nll_r, nll_o


def run(a, dev, rank=0, world=1):
    from q100 import llama as QL
    dt = {"f16": torch.float16, "bf16": torch.bfloat16}[a.dtype]
    cfg = {"8b": QL.LLAMA3_8B, "70b": QL.LLAMA3_70B}[getattr(a, "model", "8b")]
    n_layers = a.layers if a.layers else cfg.num_hidden_layers
    ppl_layers = getattr(a, "ppl_layers", None) or n_layers
    args = QL.quant_args(a.wbits, a.group or None, a.abits)
    t0 = time.time()
    model = QL.QuantLlamaModel.random_init(cfg, args, dev, dt, seed=0, rank=rank, world=world, n_layers=n_layers,
                                           fast_attention=True)
    torch.cuda.synchronize()
    out = {"model": f"llama3-{getattr(a, 'model', '8b')} random-init, {n_layers} blocks, W{a.wbits}A{a.abits} "
                    f"g{a.group}, {a.dtype}",
           "tp": world, "build_s": round(time.time() - t0, 2),
           "hbm_GB_after_build": round(torch.cuda.memory_allocated() / 1e9, 2)}
    g = torch.Generator().manual_seed(1)
    testenc = torch.randint(0, cfg.vocab_size, (1, max(a.ppl_samples, 1) * a.seqlen), generator=g)
    # ---- perplexity (main.py:122-151) ----
    if a.ppl_samples > 0:
        for l in model.layers:
            l.self_attn.fast_attention = False        # reference attention formulation for the parity number
        all_layers = model.layers
        if ppl_layers < n_layers:                     # the comparison stack: the first ppl_layers blocks
            model.layers = torch.nn.ModuleList(list(all_layers)[:ppl_layers])
        ppl, nll = model.perplexity(testenc, a.seqlen)
        model.layers = all_layers
        out["ppl"] = {"ppl": ppl, "mean_nll": nll, "tokens": a.ppl_samples * a.seqlen, "blocks": ppl_layers}
        if a.ppl_ref and rank == 0:      # the single-GPU torch port on the full weights (rank 0; the others wait below)
            try:                         # rank 0 only: a failure here must still reach the barrier the others wait at
                ppl_r, nll_r, nll_o = _port_ppl(a, cfg, model, testenc, dev, dt, ppl_layers, model.lm_head,
                                                order_noise=True)
                out["ppl"].update({"ref_ppl": ppl_r, "ref_mean_nll": nll_r, "delta_mean_nll": nll - nll_r,
                                   "delta_ppl_rel": (ppl - ppl_r) / ppl_r,
                                   "delta_ppl_at_wikitext2_scale(6.1)": 6.1 * (ppl - ppl_r) / ppl_r,
                                   # the yardstick for abits < 16: the reference against ITSELF with every F.linear
                                   # accumulated by an fp32 GEMM and rounded once (oracle/llama_port.linear_fp32_once)
                                   "ref_vs_ref_fp32_gemm_delta_mean_nll": nll_r - nll_o})
            except Exception as ex:                                   # noqa: BLE001
                import traceback
                traceback.print_exc(file=sys.stderr)
                out["ppl"]["ref_error"] = f"{type(ex).__name__}: {ex}"[:400]
        if world > 1:
            dist.barrier()
        for l in model.layers:
            l.self_attn.fast_attention = True
    # ---- prefill tokens/s ----
    batch = testenc[:, :a.seqlen].to(dev)
    for _ in range(2):
        model.forward(batch)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    iters = 5
    for _ in range(iters):
        model.forward(batch)
    e.record(); e.synchronize()
    ms = s.elapsed_time(e) / iters
    out["prefill"] = {"seq": a.seqlen, "ms": ms, "tokens_per_s": a.seqlen / ms * 1e3}
    # ---- decode ----
    out["decode"] = {}
    for B in [int(b) for b in a.batches.split(",")]:
        dec = QL.GraphDecoder(model, B, a.ctx + a.decode_steps + 8, use_graph=not a.no_graph)
        dec.prime(a.ctx)
        tok = torch.randint(0, cfg.vocab_size, (B, 1), generator=g).to(dev)
        dec.step(tok)
        torch.cuda.synchronize()
        lat = []
        for i in range(a.decode_steps):
            s.record()
            dec.step()
            e.record(); e.synchronize()
            lat.append(s.elapsed_time(e))
        med = statistics.median(lat)
        w4_bytes = sum(m.weight_quantizer._packed.nbytes() for l in model.layers for m in l._quant_linears()
                       if m.weight_quantizer._packed is not None)
        head_bytes = model.lm_head.numel() * model.lm_head.element_size()
        out["decode"][f"B{B}"] = {"median_ms_per_token": med, "tokens_per_s": B / med * 1e3,
                                  "weight_stream_GBps": (w4_bytes + head_bytes) / med / 1e6,
                                  "cuda_graph": dec.graph is not None}
        del dec
    # ---- opt-in: lm_head on packed codes as well (not the reference's configuration) ----
    if getattr(a, "quant_head", False) and a.wbits < 16:
        try:
            out["quant_head"] = _quant_head_rows(a, cfg, model, testenc, dev, dt, ppl_layers, n_layers, rank, world, g)
        except Exception as ex:                                       # noqa: BLE001 -- an extra: the row above stands
            import traceback
            traceback.print_exc(file=sys.stderr)
            out["quant_head"] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
    del model
    torch.cuda.empty_cache()
    return out


def _quant_head_rows(a, cfg, model, testenc, dev, dt, ppl_layers, n_layers, rank, world, g):
    from q100 import llama as QL
    import llama_port as LP
    dense_head = model.lm_head
    hg = a.group or 128
    model.quantize_head(a.wbits, hg)
    res = {"what": f"lm_head W{a.wbits} g{hg} through QuantLinear (the reference keeps it in 16 bits)",
           "head_MB": round(model.head_linear.weight_quantizer._packed.nbytes() / 1e6, 1),
           "dense_head_MB": round(dense_head.numel() * dense_head.element_size() / 1e6, 1)}
    if a.ppl_samples > 0:
        for l in model.layers:
            l.self_attn.fast_attention = False
        all_layers = model.layers
        if ppl_layers < n_layers:
            model.layers = torch.nn.ModuleList(list(all_layers)[:ppl_layers])
        ppl, nll = model.perplexity(testenc, a.seqlen)
        model.layers = all_layers
        for l in model.layers:
            l.self_attn.fast_attention = True
        res["ppl"] = {"ppl": ppl, "mean_nll": nll}
        if a.ppl_ref and rank == 0:       # the port with the head fake-quantised by the reference's op sequence
            try:                          # rank 0 only: see above
                import torch_port as TPo
                head_fq = TPo.fake_quant(dense_head, a.wbits, hg)[0]
                ppl_r, nll_r, _ = _port_ppl(a, cfg, model, testenc, dev, dt, ppl_layers, head_fq)
                res["ppl"].update({"ref_ppl": ppl_r, "ref_mean_nll": nll_r, "delta_mean_nll": nll - nll_r})
                del head_fq
            except Exception as ex:                                   # noqa: BLE001
                import traceback
                traceback.print_exc(file=sys.stderr)
                res["ppl"]["ref_error"] = f"{type(ex).__name__}: {ex}"[:400]
        if world > 1:
            dist.barrier()
    del dense_head
    torch.cuda.empty_cache()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res["decode"] = {}
    for B in [int(b) for b in a.batches.split(",")]:
        dec = QL.GraphDecoder(model, B, a.ctx + a.decode_steps + 8, use_graph=not a.no_graph)
        dec.prime(a.ctx)
        tok = torch.randint(0, cfg.vocab_size, (B, 1), generator=g).to(dev)
        dec.step(tok)
        torch.cuda.synchronize()
        lat = []
        for i in range(a.decode_steps):
            s.record()
            dec.step()
            e.record(); e.synchronize()
            lat.append(s.elapsed_time(e))
        med = statistics.median(lat)
        res["decode"][f"B{B}"] = {"median_ms_per_token": med, "tokens_per_s": B / med * 1e3,
                                  "cuda_graph": dec.graph is not None}
        del dec
    return res


if __name__ == "__main__":
    main()
