"""Prefill kernel on the block's GEMM shapes (M = 2048), graph-timed."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
M = int(os.environ.get("M", "2048"))
for name, N, K in (("q/o", 4096, 4096), ("k/v", 1024, 4096), ("gate/up", 14336, 4096), ("down", 4096, 14336), ("qkv", 6144, 4096)):
    pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(dt), 4, 128)[0] for _ in range(3)]
    x = torch.randn(M, K, device=dev).to(dt)
    fn = lambda i: q100.gemm_wna16(x, pws[i % 3], None, _lib.GEMM_FORCE_TC)
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(12): fn(i)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): g.replay()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 36
    print(f"{name:8s} N={N:6d} K={K:6d} M={M}: {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
