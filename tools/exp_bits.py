import sys, torch
sys.path.insert(0, "/root/repo/llama3-quantization_b200")
import q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.float16
N, K = 14336, 4096
for bits in (2, 4, 8):
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(dt) for _ in range(3)]
    pws = [q100.pack_weight(w, bits, 128)[0] for w in ws]
    x = torch.randn(1, K, device=dev).to(dt)
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
    print(f"W{bits}: {ms*1e3:.1f} us  {pws[0].nbytes()/ms/1e6:.0f} GB/s", flush=True)
