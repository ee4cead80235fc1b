#!/usr/bin/env python
"""Launch the kernels added in round 2 a few times at the BASELINE shapes (for ncu captures):
    python tools/profile_r02.py chain|w8a8"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import dataclasses
import torch
import q100
from q100 import llama as QL, ops

which = sys.argv[1] if len(sys.argv) > 1 else "chain"
dev = torch.device("cuda:0")
if which == "chain":
    cfg = dataclasses.replace(QL.LLAMA3_8B, num_hidden_layers=4, vocab_size=2000)
    m = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, torch.float16, fast_attention=True)
    d = QL.GraphDecoder(m, 1, 320, use_graph=False, chain=True)
    d.prime(256)
    for _ in range(3):
        d.step()
    torch.cuda.synchronize()
    d.check()
else:
    N, K = 14336, 4096
    pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).half(), 8, None)[0] for _ in range(3)]
    x = torch.randn(8, K, device=dev).half()
    for i in range(6):
        ops.gemm_wna8(x, pws[i % 3], 8)
    torch.cuda.synchronize()
