"""Prefill kernel with and without stream-K (Q100_TCA_NO_STREAMK=1): TFLOP/s per 8B / 70B shape at M = 2048, and equality
of the two results within fp32 summation order."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
dev = torch.device("cuda:0"); dt = torch.bfloat16
M = 2048
shapes = {"8b_q_o": (4096, 4096), "8b_k_v": (1024, 4096), "8b_up": (14336, 4096), "8b_down": (4096, 14336),
          "70b_q_o": (8192, 8192), "70b_up": (28672, 8192), "70b_down": (8192, 28672)}
out = {}
for name, (N, K) in shapes.items():
    g = torch.Generator(device=dev).manual_seed(1)
    pws = [q100.pack_weight((torch.randn(N, K, generator=g, device=dev) * 0.02).to(dt), 4, 128)[0] for _ in range(3)]
    x = torch.randn(M, K, generator=g, device=dev).to(dt)
    for i in range(3): y = q100.gemm_wna16(x, pws[i % 3])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        s.record()
        for i in range(12): q100.gemm_wna16(x, pws[i % 3])
        e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e) / 12)
    out[name] = {"us": round(best * 1e3, 1), "tflops": round(2.0 * M * N * K / (best * 1e-3) / 1e12, 1),
                 "checksum": float(q100.gemm_wna16(x, pws[0]).float().abs().sum())}
print(json.dumps(out))
