"""torchrun --nproc-per-node N tools/peer_scatter_check.py -- the fused row-parallel reduce (q100_gemm_wna16_scatter +
q100_reduce_bcast, q100/peer.py ScatterReduce) against its definition: every rank's partial GEMM output gathered and
summed in fp32 in rank order (bit-exact, identical on every rank), eagerly and inside a CUDA graph, and timed against
GEMM + NCCL all-reduce.  Prints "PEER SCATTER OK"."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import torch, torch.distributed as dist
import q100
from q100 import peer, tp

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dt = torch.bfloat16
M_MAX = 2048 * world
sc = peer.enable_scatter(None, dev, M_MAX, 4096, dt)
assert sc is not None, "symmetric memory unavailable"
res = {}
for (name, N, K, M) in (("o_proj", 4096, 4096, 2048 * world), ("down_proj", 4096, 14336, 2048 * world),
                        ("o_proj_strong", 4096, 4096, 2048 if 2048 % (world * 256) == 0 else 256 * world)):
    g = torch.Generator(device=dev).manual_seed(5)
    w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(dt)
    x = torch.randn(M, K, generator=g, device=dev).to(dt)
    wl = tp.shard_weight(w, tp.ROW, rank, world, 128)
    xl = tp.shard_input(x, tp.ROW, rank, world)
    pw, _ = q100.pack_weight(wl, 4, 128)
    assert sc.fits(M, pw)
    part = q100.gemm_wna16(xl, pw)                                   # this rank's partial, 16-bit
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    ref = torch.zeros(M, N, dtype=torch.float32, device=dev)
    for p in parts:                                                   # fp32 sum in rank order, one rounding
        ref += p.float()
    ref32 = ref.clone()
    ref = ref.to(dt)
    multicast = isinstance(sc, peer.MulticastReduce)

    def same(y, what):
        if not multicast:                                             # K7b: fp32 sum in rank order, bit for bit
            assert torch.equal(y.view(torch.int16), ref.view(torch.int16)), f"{name}: {what} differs from the fp32 sum of the partials"
            return
        # K7c: the switch adds in fp32 in ITS order -- within one rounding of the 16-bit result, and the same bits on
        # every rank (one reduction, then a broadcast)
        err = (y.float() - ref32).abs()
        ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10       # one unit in the last place of the 16-bit result
        tol = ulp * ref32.abs() + 1e-6
        res.setdefault("max_err_ulps", 0.0)
        res["max_err_ulps"] = max(res["max_err_ulps"], float((err / (ulp * ref32.abs() + 1e-6)).max()))
        res["frac_not_rn"] = float((y.view(torch.int16) != ref.view(torch.int16)).float().mean())
        assert bool((err <= tol).all()), f"{name}: {what}: max err {float(err.max()):.3e}"
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y.contiguous())
        assert all(torch.equal(ys[0].view(torch.int16), t.view(torch.int16)) for t in ys), f"{name}: {what}: ranks disagree"
    y = sc(xl, pw).clone()
    torch.cuda.synchronize()
    same(y, "fused result")
    y2 = sc(xl, pw).clone()                                           # the other buffer
    torch.cuda.synchronize()
    same(y2, "second buffer")
    # graph capture
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sc(xl, pw); sc(xl, pw)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(4):
            yg = sc(xl, pw)
    gr.replay(); torch.cuda.synchronize()
    same(yg.clone(), "graph replay")
    # timing: fused vs GEMM + NCCL all-reduce
    def t(fn, n=10):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / n * 1e3
    def nccl():
        yy = q100.gemm_wna16(xl, pw)
        dist.all_reduce(yy)
    for _ in range(3):
        nccl()
    us_f, us_n, us_g = t(lambda: sc(xl, pw)), t(nccl), t(lambda: q100.gemm_wna16(xl, pw))
    res["route"] = type(sc).__name__
    res[name] = {"M": M, "fused_us": round(us_f, 1), "gemm_plus_nccl_us": round(us_n, 1), "gemm_alone_us": round(us_g, 1)}
sc.check()
ok = torch.ones(1, device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, **res}))
    print("PEER SCATTER OK")
import threading
threading.Timer(45.0, lambda: os._exit(0)).start()
peer.shutdown()
import gc; gc.collect()
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
os._exit(0)
