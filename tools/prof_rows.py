"""One decode GEMM shape through K3r, a few launches (for ncu): python tools/prof_rows.py N K M"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
N, K, M = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda:0"); dt = torch.bfloat16
pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(dt), 4, 128)[0] for _ in range(3)]
x = torch.randn(M, K, device=dev).to(dt)
for i in range(6):
    q100.gemm_wna16(x, pws[i % 3], None, 0)
torch.cuda.synchronize()
