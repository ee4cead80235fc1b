#!/usr/bin/env python
"""Per-launch time of the decode attention core: K8 (q100_attn_decode, one launch) against K5 + scaled_dot_product_attention
(two launches), 32 launches on 32 different caches inside one CUDA graph (a decode step's worth), LLaMA3-8B head shapes."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "llama3-quantization_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
from q100 import ops  # noqa: E402
import llama_port as LP  # noqa: E402


def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    dt = torch.float16
    H, KVH, D, L = 32, 8, 128, 32
    cfgs = ((512, 256), (4096, 2048), (8192, 8000))
    if os.environ.get("SHORT"):
        cfgs = ((512, 256), (2048, 1024))
    for max_len, ctx in cfgs:
        for B in (1, 2, 4, 16):
            if max_len > 512 and B == 4:
                continue
            cos, sin = LP.rotary_cache(D, max_len, 500000.0, dev, dt)
            cos, sin = cos.contiguous(), sin.contiguous()
            qkv = [torch.randn(B, (H + 2 * KVH) * D, device=dev).to(dt) for _ in range(L)]
            kc = [torch.randn(B, KVH, max_len, D, device=dev).to(dt) for _ in range(L)]
            vc = [torch.randn(B, KVH, max_len, D, device=dev).to(dt) for _ in range(L)]
            pos = torch.tensor([[ctx]], device=dev)
            slots = torch.arange(max_len, device=dev).view(1, 1, 1, max_len)
            mask = torch.where(slots <= pos.view(1, 1, 1, 1), 0.0, torch.finfo(dt).min).to(dt).expand(B, 1, 1, max_len)

            def k8():
                for i in range(L):
                    ops.attn_decode(qkv[i], cos, sin, pos, kc[i], vc[i], H, KVH)

            def k5_sdpa():
                for i in range(L):
                    q = ops.rope_kv_decode(qkv[i], cos, sin, pos, kc[i], vc[i], H, KVH)
                    F.scaled_dot_product_attention(q, kc[i], vc[i], attn_mask=mask, enable_gqa=True)

            t8, t5 = graph_time(k8) / L, graph_time(k5_sdpa) / L
            mb = 2 * B * KVH * (ctx + 1) * D * 2 / 1e6
            print(f"max_len {max_len:5d} ctx {ctx:5d} B {B:2d}: K8 {t8:6.2f} us   K5 + SDPA {t5:6.2f} us   "
                  f"(cache read {mb:6.1f} MB = {mb / 1e6 / (t8 * 1e-6):5.2f} TB/s in K8)")
            del qkv, kc, vc


if __name__ == "__main__":
    main()
