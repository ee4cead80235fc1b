"""K1b timing: fused quantizer backward vs the op-by-op torch form on one up_proj weight (14336 x 4096, W4 g128)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import ops
from quant.quantizer import fake_quant_backward
dev = torch.device("cuda:0")
N, K, G = 14336, 4096, 128
for dt in (torch.float32, torch.float16):
    x = (torch.randn(N, K, device=dev) * 0.02).to(dt)
    go = torch.randn(N, K, device=dev).to(dt)
    ng = N * K // G
    su = (torch.rand(ng, 1, device=dev) * 0.4 + 0.6).to(dt); sl = (torch.rand(ng, 1, device=dev) * 0.4 + 0.6).to(dt)
    r = q100.quantize_groupwise(x, G, 4, False, False, su, sl)
    def t(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): fn()
        e.record(); e.synchronize()
        return s.elapsed_time(e) / n
    ms_f = t(lambda: q100.quantize_groupwise(x, G, 4, False, False, su, sl))
    fb = N * K * 2 * x.element_size() + ng * 4 * x.element_size()
    print(f"{str(dt)[6:]:8s} K1 forward (LWC) {ms_f*1e3:8.1f} us  {fb/ms_f/1e6:7.1f} GB/s ({fb/ms_f/1e6/6550.7:.2%} of HBM)")
    ms_k = t(lambda: ops.quantize_backward(x, go, G, 4, False, False, su, sl, True, True))
    ms_t = t(lambda: fake_quant_backward(x.view(-1, G), go.view(-1, G), r.scales, r.zeros, su, sl, 4, False, 0, 15, True, True), 3)
    b = x.element_size()
    algo = N * K * 3 * b + ng * 4 * b
    print(f"{str(dt)[6:]:8s} K1b {ms_k*1e3:8.1f} us  {algo/ms_k/1e6:7.1f} GB/s ({algo/ms_k/1e6/6550.7:.2%} of HBM) | op-by-op torch {ms_t*1e3:9.1f} us  -> {ms_t/ms_k:5.1f}x")
