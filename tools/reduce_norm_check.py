"""torchrun --nproc-per-node N tools/reduce_norm_check.py -- K7n (q100_allreduce_add_rmsnorm, q100/peer.py ReduceNorm)
against the two launches it replaces: K7 (one-shot all-reduce) + K4 (residual add + RMSNorm).  Both h and the normalised
row must be bit-identical, eagerly, over many calls (slot parity) and inside a CUDA graph; then a tensor-parallel
GraphDecoder step with and without K7n must give the same logits.  Prints "REDUCE NORM OK"."""
import dataclasses, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import torch, torch.distributed as dist
from q100 import peer, ops, llama as QL

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
one = peer.enable(None, dev)
assert one is not None, "symmetric memory unavailable"
res = {"world": world}
for cols in (4096, 8192):
    rn = peer.enable_reduce_norm(None, dev, cols)
    assert rn is not None
    for dt in (torch.float16, torch.bfloat16):
        for rows in (1, 2, 16, 64):
            g = torch.Generator(device=dev).manual_seed(100 * rank + rows)
            w = (1 + 0.1 * torch.randn(cols, generator=torch.Generator(device=dev).manual_seed(7), device=dev)).to(dt)
            for it in range(5):
                x = torch.randn(rows, 1, cols, generator=g, device=dev).to(dt)
                r = torch.randn(rows, 1, cols, generator=torch.Generator(device=dev).manual_seed(it), device=dev).to(dt)
                for resid in (r, None):
                    # K7's definition (tools/peer_allreduce_check.py): fp32 sum of the gathered partials in rank order
                    parts = [torch.empty_like(x) for _ in range(world)]
                    dist.all_gather(parts, x)
                    y = sum((p.float() for p in parts[1:]), parts[0].float()).to(dt)
                    if x.numel() <= one.slot_elems:
                        assert torch.equal(one(x.clone()).view(torch.int16), y.view(torch.int16))
                    h_ref, n_ref = ops.add_rmsnorm(y, w, 1e-5, resid)
                    h, n = rn(x, w, 1e-5, resid)
                    torch.cuda.synchronize()
                    assert torch.equal(h.view(torch.int16), h_ref.view(torch.int16)), (cols, dt, rows, it, "h")
                    assert torch.equal(n.view(torch.int16), n_ref.view(torch.int16)), (cols, dt, rows, it, "norm")
    # graph capture: 6 dependent calls
    x = torch.randn(2, 1, cols, device=dev).half()
    w = torch.ones(cols, device=dev).half()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rn(x, w, 1e-5, x)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        h = x
        for _ in range(6):
            h, n = rn(n if _ else x, w, 1e-5, h)
    gr.replay(); torch.cuda.synchronize()
    h_e = x; n_e = x
    for i in range(6):
        y = one((n_e if i else x).clone())
        h_e, n_e = ops.add_rmsnorm(y, w, 1e-5, h_e)
    # (2 x cols elements fit a K7 slot)
    torch.cuda.synchronize()
    assert torch.equal(n.view(torch.int16), n_e.view(torch.int16)), "graph replay differs"
    rn.check()

    def t(fn, n_=50):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / n_ * 1e3
    res[f"cols{cols}_graph6_us_per_call"] = round(t(gr.replay, 20) / 6, 2)

# ---- a tensor-parallel decode step with and without K7n: same logits
cfg = dataclasses.replace(QL.LLAMA3_8B, num_hidden_layers=4, vocab_size=4096)
m = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, seed=3, rank=rank, world=world,
                                   tp_group=None, fast_attention=True)
tok = torch.randint(0, cfg.vocab_size, (2, 1), generator=torch.Generator().manual_seed(1)).to(dev)
outs = []
for no_rn in ("", "1"):
    if no_rn:
        os.environ["Q100_NO_REDUCE_NORM"] = "1"
    dec = QL.GraphDecoder(m, 2, 64, use_graph=False)
    dec.pos.fill_(0)
    for _ in range(3):
        dec.step(tok)
    outs.append(dec.last_logits.clone())
    res["k7n_active" if not no_rn else "k7n_off_active"] = dec._reduce_norm is not None
os.environ.pop("Q100_NO_REDUCE_NORM", None)
res["tp_world_of_model"] = getattr(m.layers[0].self_attn, "tp_world", 1)
assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16)), "decode logits differ with K7n"
ok = torch.ones(1, device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps(res))
    print("REDUCE NORM OK")
import threading
threading.Timer(45.0, lambda: os._exit(0)).start()
peer.shutdown()
import gc; gc.collect()
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
os._exit(0)
