"""K3r (row-owner decode kernel) against K3a (stream-K): agreement and timings on the decode shapes of a LLaMA3-8B block.

    python tools/exp_rows.py [M ...]
Per shape and M: max |y_rows - y_streamk| relative to max |y|, per-launch time of both kernels (three weight sets
rotated, 6 launches per CUDA graph with PDL), and the four launches of a block in one graph.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
SH = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gateup", 28672, 4096), ("down", 4096, 14336)]
torch.manual_seed(0)
sets = [[q100.pack_weight((torch.randn(n, k, device=dev) * 0.02).to(dt), 4, 128)[0] for _, n, k in SH] for _ in range(3)]
Ms = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]


def timed(fn, reps=6, outer=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(outer): g.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / (reps * outer) * 1e3


for M in Ms:
    xs = [torch.randn(M, k, device=dev).to(dt) for _, _, k in SH]
    line = [f"M={M}"]
    for j, (name, n, k) in enumerate(SH):
        y_r = q100.gemm_wna16(xs[j], sets[0][j], None, _lib.GEMM_PDL)
        y_s = q100.gemm_wna16(xs[j], sets[0][j], None, _lib.GEMM_PDL | _lib.GEMM_NO_ROWS)
        torch.cuda.synchronize()
        err = (y_r.float() - y_s.float()).abs().max().item() / y_s.float().abs().max().item()
        t_r = timed(lambda i: q100.gemm_wna16(xs[j], sets[i % 3][j], None, _lib.GEMM_PDL))
        t_s = timed(lambda i: q100.gemm_wna16(xs[j], sets[i % 3][j], None, _lib.GEMM_PDL | _lib.GEMM_NO_ROWS))
        mb = (n * k // 2 + n * k // 128 * 4) / 1e6
        line.append(f"{name}: rel {err:.1e} rows {t_r:.2f} us ({mb / t_r:.0f} GB/s) streamk {t_s:.2f} us")
    def blk(i, fl):
        for j in range(4): q100.gemm_wna16(xs[j], sets[i % 3][j], None, fl)
    b_r = timed(lambda i: blk(i, _lib.GEMM_PDL)) 
    b_s = timed(lambda i: blk(i, _lib.GEMM_PDL | _lib.GEMM_NO_ROWS))
    line.append(f"block: rows {b_r:.2f} us streamk {b_s:.2f} us (HBM floor 17.7)")
    print(" | ".join(line), flush=True)
