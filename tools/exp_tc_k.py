"""Prefill kernel: time vs K at fixed M, N -> per-k-block slope and per-tile overhead."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
M = 2048
for N in (14336, 18944):
  for K in (1024, 2048, 4096, 8192):
    pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(dt), 4, 128)[0] for _ in range(2)]
    x = torch.randn(M, K, device=dev).to(dt)
    fn = lambda i: q100.gemm_wna16(x, pws[i % 2], None, _lib.GEMM_FORCE_TC)
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(8): fn(i)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): g.replay()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 24
    tiles = (N // 128) * (M // 256); waves = -(-tiles // 148)
    print(f"N={N} K={K:5d}: {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s  tiles={tiles} waves={waves}  us/wave={ms*1e3/waves:6.2f}  kblocks={K//64}", flush=True)
