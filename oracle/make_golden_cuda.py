#!/usr/bin/env python
"""Generate tests/golden/cuda_eager.npz: the UNMODIFIED reference quantizer (oracle/_ref/quant/quantizer.py, vendored by
oracle/build_ref.py) run on CUDA tensors -- the reference as it is actually deployed.  Must run on a GPU box:

    gpurun -- 'python oracle/make_golden_cuda.py gpurun_out/cuda_eager.npz'   (then copy into tests/golden/)

Why a second fixture: torch's CUDA kernels compute `tensor / python_scalar` as a multiplication by the fp32 reciprocal,
its CPU kernels divide, so `scale = (xmax - xmin) / (2^n - 1)` (quantizer.py:147,152) can differ by one ulp between the
reference's own CPU and GPU runs.  tests/golden/weights.npz pins the CPU semantics (the default of the kernels); this
file pins the CUDA semantics (Q100_FLAG_CUDA_SCALE).  Same case grid as make_golden.py, smaller tensors.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_ref"), os.path.join(HERE, "_ref", "stubs"), HERE]
import quant.quantizer as RQ  # noqa: E402  (the reference's)

assert os.path.realpath(RQ.__file__).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), RQ.__file__
TDT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def bits(t):
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32).copy()
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def main(path):
    dev = torch.device("cuda:0")
    out = {}
    n_diff = 0
    for dt in ("f16", "bf16", "f32"):
        for n_bits in (2, 3, 4, 8):
            for mode in ("asym", "sym", "nozp", "lwc"):
                for group in (128, None):
                    name = f"{dt}_w{n_bits}_{mode}_g{group or 0}"
                    g = torch.Generator().manual_seed(sum(map(ord, name)))
                    rows, cols = 16, 256
                    x = torch.randn(rows, cols, generator=g) * 0.02
                    x[0] = x[0].abs() + 0.01            # all-positive groups: negative zero-point
                    x[1] = 0.3                          # constant groups: scale clamps to CLIPMIN
                    x[2, :128] = 0.0
                    x = x.to(TDT[dt]).to(dev)
                    q = RQ.UniformAffineQuantizer(n_bits=n_bits, symmetric=mode == "sym", per_channel_axes=[0],
                                                  dynamic_method="per_channel", group_size=group, shape=x.shape,
                                                  lwc=mode == "lwc", disable_zero_point=mode == "nozp").to(dev)
                    if mode == "lwc":
                        with torch.no_grad():
                            q.upbound_factor.copy_(torch.randn(q.upbound_factor.shape, generator=g) + 3.0)
                            q.lowbound_factor.copy_(torch.randn(q.lowbound_factor.shape, generator=g) + 3.0)
                    with torch.no_grad():
                        y = q(x)
                        # the same module on the CPU: how often do the two runs of the reference disagree?
                        qc = RQ.UniformAffineQuantizer(n_bits=n_bits, symmetric=mode == "sym", per_channel_axes=[0],
                                                       dynamic_method="per_channel", group_size=group, shape=x.shape,
                                                       lwc=mode == "lwc", disable_zero_point=mode == "nozp")
                        if mode == "lwc":
                            qc.upbound_factor.copy_(q.upbound_factor.cpu())
                            qc.lowbound_factor.copy_(q.lowbound_factor.cpu())
                        yc = qc(x.cpu())
                    n_diff += int((bits(y) != bits(yc)).sum())
                    out[f"{name}/x"] = bits(x)
                    out[f"{name}/y"] = bits(y)
                    out[f"{name}/scale"] = bits(q.scale)
                    if q.round_zero_point is not None:
                        out[f"{name}/zp"] = bits(q.round_zero_point)
                    if mode == "lwc":
                        # post-sigmoid multipliers in the dtype of sigmoid(f) * xmax (what the kernel receives)
                        cdt = torch.promote_types(x.dtype, q.upbound_factor.dtype)
                        out[f"{name}/up"] = bits(torch.sigmoid(q.upbound_factor).to(cdt))
                        out[f"{name}/low"] = bits(torch.sigmoid(q.lowbound_factor).to(cdt))
                        out[f"{name}/factors"] = np.stack([q.upbound_factor.detach().cpu().numpy(),
                                                           q.lowbound_factor.detach().cpu().numpy()])
                    out[f"{name}/cfg"] = np.array([n_bits, group or 0, int(mode == "sym"), int(mode == "nozp"),
                                                   int(mode == "lwc")], dtype=np.int64)
    out["_meta/elements_where_cuda_and_cpu_reference_differ"] = np.array([n_diff], dtype=np.int64)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes; "
          f"CUDA and CPU runs of the reference differ in {n_diff} output elements over all cases")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "tests", "golden", "cuda_eager.npz"))
