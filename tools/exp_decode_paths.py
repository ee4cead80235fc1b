"""gate|up decode launch (28672 x 4096, W4 g128) at M = 1..16 through each decode kernel, timed in a CUDA graph
with three rotating weight sets (HBM, not L2)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0")
N, K = 28672, 4096
PATHS = {"default": 0, "mma_sync": _lib.GEMM_FORCE_MMA_SYNC, "mma_f16": _lib.GEMM_FORCE_MMA_SYNC | _lib.GEMM_NO_IMMA,
         "stream_tc": _lib.GEMM_FORCE_STREAM_TC}
for dt in (torch.bfloat16, torch.float16):
    pws = []
    for i in range(3):
        w = (torch.randn(N, K, device=dev) * 0.02).to(dt); pws.append(q100.pack_weight(w, 4, 128)[0]); del w
    nbytes = pws[0].nbytes()
    for M in (1, 2, 4, 8, 12, 16):
        x = torch.randn(M, K, device=dev).to(dt)
        line = f"{str(dt)[6:]:9s} M={M:2d}"
        for name, fl in PATHS.items():
            flags = _lib.GEMM_FORCE_SKINNY | fl | _lib.GEMM_PDL
            try:
                for i in range(3): q100.gemm_wna16(x, pws[i % 3], None, flags)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(12): q100.gemm_wna16(x, pws[i % 3], None, flags)
                g.replay(); torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(5): g.replay()
                e.record(); e.synchronize()
                us = s.elapsed_time(e) / 60 * 1e3
                line += f" | {name} {us:6.2f} us {nbytes / us / 1e3:5.0f} GB/s"
            except Exception as ex:
                line += f" | {name} n/a"
        print(line, flush=True)
