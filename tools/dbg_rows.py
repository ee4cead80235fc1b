import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
shapes = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1:]]
for (N, K) in shapes:
    for M in (1,):
        torch.manual_seed(N + K + M)
        pw = q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(dt), 4, 128)[0]
        x = torch.randn(M, K, device=dev).to(dt)
        ys = q100.gemm_wna16(x, pw, None, _lib.GEMM_NO_ROWS); torch.cuda.synchronize()
        yr = q100.gemm_wna16(x, pw, None, 0); torch.cuda.synchronize()
        err = (yr.float() - ys.float()).abs().max().item() / ys.float().abs().max().item()
        print(N, K, M, "rows ok, rel", err, flush=True)
