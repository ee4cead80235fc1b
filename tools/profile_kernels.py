#!/usr/bin/env python
"""Launch each q100 kernel a few times at the BASELINE shapes (for ncu captures / quick timing).

    python tools/profile_kernels.py [--which k1,k1pack,k2,skinny1,skinny16,tc] [--iters 3] [--skip-tc]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import torch  # noqa: E402
import q100  # noqa: E402
from q100 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="k1,k1pack,k2,skinny1,skinny16,tc")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--graph", action="store_true", help="time each kernel inside a CUDA graph (no launch gaps)")
    a = ap.parse_args()
    which = a.which.split(",")
    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    N, K = 14336, 4096
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(dt) for _ in range(3)]
    pws = [q100.pack_weight(w, 4, 128)[0] for w in ws]
    x1 = torch.randn(1, K, device=dev).to(dt)
    x16 = torch.randn(16, K, device=dev).to(dt)
    x2k = torch.randn(2048, K, device=dev).to(dt)
    act = torch.randn(2048, 4096, device=dev).to(dt)
    torch.cuda.synchronize()

    def run(name, fn, algo_bytes=None, flops=None):
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(a.iters):
            fn(i)
        e.record()
        e.synchronize()
        ms = s.elapsed_time(e) / a.iters
        if a.graph:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    fn(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for i in range(a.iters):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            s.record()
            for _ in range(3):
                g.replay()
            e.record()
            e.synchronize()
            ms = s.elapsed_time(e) / (3 * a.iters)
        msg = f"{name:10s} {ms * 1e3:9.1f} us"
        if algo_bytes:
            msg += f"  {algo_bytes / ms / 1e6:8.1f} GB/s"
        if flops:
            msg += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
        print(msg, flush=True)

    if "k1" in which:
        run("k1_fake", lambda i: q100.quantize_groupwise(ws[i % 3], 128, 4), algo_bytes=N * K * 4.03125)
    if "k1pack" in which:
        run("k1_pack", lambda i: q100.quantize_groupwise(ws[i % 3], 128, 4, want_dequant=False, want_packed=True,
                                                         want_meta=True), algo_bytes=N * K * 2.53125)
    if "k2" in which:
        run("k2_act8", lambda i: q100.fakequant_rows(act, 8), algo_bytes=act.numel() * 4)
    wbytes = pws[0].nbytes()
    if "skinny1" in which:
        run("skinny_m1", lambda i: q100.gemm_wna16(x1, pws[i % 3], None, _lib.GEMM_FORCE_SKINNY), algo_bytes=wbytes)
    if "skinny16" in which:
        run("skinny_m16", lambda i: q100.gemm_wna16(x16, pws[i % 3], None, _lib.GEMM_FORCE_SKINNY), algo_bytes=wbytes)
    if "dense" in which:
        from q100 import ops
        wd = ws[0]
        run("dense_tc", lambda i: ops.debug_gemm_dense_tc(x2k, ws[i % 3]), flops=2.0 * 2048 * N * K)
    if "contig" in which:
        # same bytes as up_proj but every [128 rows x 256 k] unit is one contiguous 16 KB run in HBM
        Nc, Kc = N * 16, K // 16
        wc = [(torch.randn(Nc, Kc, device=dev) * 0.02).to(dt) for _ in range(3)]
        pwc = [q100.pack_weight(w, 4, 128)[0] for w in wc]
        xc = torch.randn(1, Kc, device=dev).to(dt)
        run("skinny_ctg", lambda i: q100.gemm_wna16(xc, pwc[i % 3], None, _lib.GEMM_FORCE_SKINNY), algo_bytes=pwc[0].nbytes())
        del wc, pwc
    if "tc" in which:
        run("tc_m2048", lambda i: q100.gemm_wna16(x2k, pws[i % 3], None, _lib.GEMM_FORCE_TC), flops=2.0 * 2048 * N * K)


if __name__ == "__main__":
    main()
