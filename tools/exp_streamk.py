"""Decode-GEMM shape sweep (CUDA-graph timed, three weight sets rotated so the stream comes from HBM).

    python tools/exp_streamk.py [--dtype bf16|f16] [--m 1,16] [--pdl] [--shapes up,gateup,...]
"""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch
import q100
from q100 import _lib

SHAPES = {
    "up": (14336, 4096), "gateup": (28672, 4096), "down": (4096, 14336), "o": (4096, 4096), "qkv": (6144, 4096),
    "up70": (28672, 8192), "gateup70": (57344, 8192), "down70": (8192, 28672), "qkv70": (10240, 8192),
    "down_tp8": (4096, 1792), "up_tp8": (1792, 4096),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--m", default="1")
    ap.add_argument("--pdl", action="store_true")
    ap.add_argument("--mma-sync", action="store_true", help="force the mma.sync streaming kernel")
    ap.add_argument("--stream-tc", action="store_true", help="force the tcgen05 streaming kernel")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--shapes", default="up,gateup,down,o,qkv,gateup70")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    flags = _lib.GEMM_FORCE_SKINNY | (_lib.GEMM_PDL if a.pdl else 0) | (_lib.GEMM_FORCE_MMA_SYNC if a.mma_sync else 0) | (_lib.GEMM_FORCE_STREAM_TC if a.stream_tc else 0)
    for name in a.shapes.split(","):
        N, K = SHAPES[name]
        nset = max(2, -(-340_000_000 // (N * K * a.bits // 8)))   # rotate through > 2.5x the 126 MB L2
        pws = []
        for i in range(nset):
            w = (torch.randn(N, K, device=dev) * 0.02).to(dt)
            pws.append(q100.pack_weight(w, a.bits, 128)[0])
            del w
        for M in [int(m) for m in a.m.split(",")]:
            x = torch.randn(M, K, device=dev).to(dt)
            fn = lambda i: q100.gemm_wna16(x, pws[i % nset], None, flags)
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(12):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                g.replay()
            e.record()
            e.synchronize()
            ms = s.elapsed_time(e) / 60
            nb = pws[0].nbytes()
            print(f"{name:9s} N={N:6d} K={K:5d} M={M:2d} {a.dtype} W{a.bits}{' pdl' if a.pdl else ''}: {ms*1e3:7.2f} us "
                  f"{nb/ms/1e6:6.0f} GB/s ({nb/ms/1e6/6550.7*100:4.1f}% of 6550.7)", flush=True)
        del pws


if __name__ == "__main__":
    main()
