import sys, torch
sys.path.insert(0, "/root/repo/llama3-quantization_b200")
import q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
def run(N, K, M=1, tag=""):
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(dt) for _ in range(3)]
    pws = [q100.pack_weight(w, 4, 128)[0] for w in ws]
    x = torch.randn(M, K, device=dev).to(dt)
    fn = lambda i: q100.gemm_wna16(x, pws[i % 3], None, _lib.GEMM_FORCE_SKINNY)
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(12): fn(i)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): g.replay()
    e.record(); e.synchronize()
    ms = s.elapsed_time(e) / 36
    print(f"{tag:28s} N={N:6d} K={K:5d} M={M:2d}: {ms*1e3:6.1f} us  {pws[0].nbytes()/ms/1e6:6.0f} GB/s", flush=True)
run(37888, 2048, 1, "whole tiles (1 tile/CTA)")
run(18944, 4096, 1, "half tiles (2 CTAs/tile)")
run(14336, 4096, 1, "up_proj")
run(28672, 4096, 1, "gate|up fused")
run(4096, 14336, 1, "down_proj")
run(4096, 4096, 1, "o_proj")
run(6144, 4096, 1, "qkv fused")
run(128*296*4, 2048, 1, "whole tiles x4 (155MB)")
