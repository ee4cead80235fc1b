"""torchrun --nproc-per-node N tools/peer_allreduce_check.py -- K7 (one-shot all-reduce over peer memory) against an
fp32 reference built from an all-gather, eager and inside a CUDA graph, plus latency next to NCCL."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch
import torch.distributed as dist
from q100 import peer, tp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
one = peer.enable(None, dev)
assert one is not None, "symmetric memory setup failed"
ok = True
for dt in (torch.float16, torch.bfloat16):
    for n in (8, 4096, 4 * 4096, 16 * 4096, 16 * 8192, 1000 * 8):
        for it in range(20):
            g = torch.Generator(device=dev).manual_seed(1000 * rank + it + n)
            x = torch.randn(n, generator=g, device=dev).to(dt)
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x)
            ref = torch.stack([p.float() for p in parts]).sum(0).to(dt)          # rank order, fp32, one rounding
            y = tp.reduce_output(x.clone(), tp.ROW)
            if not torch.equal(y, ref):
                ok = False
                print(f"rank {rank}: MISMATCH dt={dt} n={n} it={it} max err {(y.float() - ref.float()).abs().max().item()}")
                break
# CUDA graph: 8 calls per replay, data changes between replays
x = torch.zeros(16 * 4096, device=dev, dtype=torch.float16)
outs = [torch.empty_like(x) for _ in range(8)]
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        for o in outs:
            o.copy_(x); one(o)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i, o in enumerate(outs):
        o.copy_(x * (i + 1)); one(o)
for rep in range(5):
    x.copy_(torch.full_like(x, float(rank + 1 + rep)))
    g.replay()
    torch.cuda.synchronize()
    want = sum(float(r + 1 + rep) for r in range(world))
    for i, o in enumerate(outs):
        if not torch.all(o == want * (i + 1)):
            ok = False
            print(f"rank {rank}: graph MISMATCH rep {rep} call {i}: got {o[0].item()} want {want * (i + 1)}")
            break
assert int(one.err.item()) == 0, "a peer timed out"
# latency: 64 calls in a graph, [M, 4096] fp16
for M in (1, 4, 16):
    y = torch.randn(M * 4096, device=dev).half()
    res = {}
    for name, fn in (("k7", lambda: one(y)), ("nccl", lambda: dist.all_reduce(y))):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            for _ in range(64): fn()
        gg.replay(); torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): gg.replay()
        b.record(); b.synchronize()
        res[name] = a.elapsed_time(b) / (5 * 64) * 1e3
    if rank == 0:
        print(f"all-reduce [{M}, 4096] fp16, world {world}: K7 {res['k7']:.2f} us   NCCL {res['nccl']:.2f} us")
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER ALLREDUCE", "OK" if t.item() == 1 else "FAILED", flush=True)
# teardown order matters: graphs and symmetric-memory handles first, then the process group
import gc, time
del g, gg, one
peer.shutdown()
gc.collect()
torch.cuda.synchronize()
dist.barrier()
print(f"rank {rank}: released peer buffers", flush=True)
t0 = time.time()
dist.destroy_process_group()
print(f"rank {rank}: destroy_process_group took {time.time() - t0:.1f} s", flush=True)
