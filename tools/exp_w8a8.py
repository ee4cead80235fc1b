#!/usr/bin/env python
"""Timing of K3d (q100_gemm_w8a8, tcgen05 kind::i8) against the two-step route (K2 fake-quant + the W8A16 dequant-GEMM)
at the LLaMA3-8B block shapes, M = 2048.  Prints TOP/s of the integer GEMM alone and of the whole linear."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "llama3-quantization_b200"))
import q100  # noqa: E402
from q100 import ops  # noqa: E402
from q100._lib import lib, check  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    M = int(os.environ.get("M", "2048"))
    dt = torch.bfloat16
    shapes = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gateup", 28672, 4096), ("down", 4096, 14336)]
    tot = {"k3d": 0.0, "k3d_full": 0.0, "two_step": 0.0}
    for name, N, K in shapes:
        ws = [q100.pack_weight((torch.randn(N, K, device="cuda") * 0.02).to(dt), 8, None)[0] for _ in range(3)]
        x = torch.randn(M, K, device="cuda").to(dt)
        r = ops.quantize_groupwise(x, 0, 8, want_dequant=False, want_packed=True, want_meta=True)
        xs = ops.code_rowsums(r.packed, M, K)
        for w in ws:
            w.code_sums = ops.code_rowsums(w.packed, N, K)
        y = torch.empty(M, N, device="cuda", dtype=dt)
        st = torch.cuda.current_stream().cuda_stream
        i = [0]

        def gemm_only():
            w = ws[i[0] % 3]
            i[0] += 1
            check(lib.q100_gemm_w8a8(r.packed.data_ptr(), r.meta.data_ptr(), xs.data_ptr(), w.packed.data_ptr(),
                                     w.meta.data_ptr(), w.code_sums.data_ptr(), None, y.data_ptr(), M, N, K, 1, st))

        def full():
            w = ws[i[0] % 3]
            i[0] += 1
            ops.gemm_w8a8(x, w)

        def two_step():
            w = ws[i[0] % 3]
            i[0] += 1
            ops.gemm_wna16(ops.fakequant_rows(x, 8).dequant, w)

        t0, t1, t2 = timeit(gemm_only), timeit(full), timeit(two_step)
        fl = 2.0 * M * N * K
        print(f"{name:7s} N={N:6d} K={K:6d}  K3d {t0:8.1f} us ({fl / t0 / 1e6:7.0f} TOP/s)   K2-pack + sums + K3d {t1:8.1f} us"
              f"   K2 + W8A16 k3_tca {t2:8.1f} us ({fl / t2 / 1e6:7.0f} TFLOP/s)")
        tot["k3d"] += t0
        tot["k3d_full"] += t1
        tot["two_step"] += t2
    print("block:", {k: round(v, 1) for k, v in tot.items()})


if __name__ == "__main__":
    main()
