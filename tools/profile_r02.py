#!/usr/bin/env python
"""Launch the kernels added in round 2 a few times at the BASELINE shapes (for ncu captures):
    python tools/profile_r02.py chain|w8a8|i8|attn"""
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
elif which == "i8":
    # K3d: W8A8 per-channel at prefill sizes, gate|up of the 8B block, M = 2048
    N, K, M = 28672, 4096, 2048
    pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16), 8, None)[0] for _ in range(2)]
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    for i in range(4):
        ops.gemm_w8a8(x, pws[i % 2])
    torch.cuda.synchronize()
elif which == "attn":
    # K8: decode attention, 8B head shapes: batch 16 on a 4096-slot cache at position 2048, batch 1 at position 256
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import llama_port as LP
    H, KVH, D = 32, 8, 128
    for B, max_len, ctx in ((16, 4096, 2048), (1, 512, 256)):
        cos, sin = LP.rotary_cache(D, max_len, 500000.0, dev, torch.float16)
        qkv = torch.randn(B, (H + 2 * KVH) * D, device=dev).half()
        kc = [torch.randn(B, KVH, max_len, D, device=dev).half() for _ in range(3)]
        vc = [torch.randn(B, KVH, max_len, D, device=dev).half() for _ in range(3)]
        pos = torch.tensor([[ctx]], device=dev)
        for i in range(3):
            ops.attn_decode(qkv, cos.contiguous(), sin.contiguous(), pos, kc[i], vc[i], H, KVH)
        torch.cuda.synchronize()
else:
    N, K = 14336, 4096
    pws = [q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).half(), 8, None)[0] for _ in range(3)]
    x = torch.randn(8, K, device=dev).half()
    for i in range(6):
        ops.gemm_wna8(x, pws[i % 3], 8)
    torch.cuda.synchronize()
