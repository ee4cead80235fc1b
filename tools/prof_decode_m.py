"""gate|up decode launch at a given M through the mma.sync kernel (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
pws = [q100.pack_weight((torch.randn(28672, 4096, device=dev) * 0.02).half(), 4, 128)[0] for _ in range(3)]
x = torch.randn(M, 4096, device=dev).half()
for i in range(6):
    q100.gemm_wna16(x, pws[i % 3], None, _lib.GEMM_FORCE_SKINNY | _lib.GEMM_FORCE_MMA_SYNC)
torch.cuda.synchronize()
