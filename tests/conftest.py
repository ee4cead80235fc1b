import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "llama3-quantization_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TDT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """tests/golden/<name>.npz -> {case: {field: ndarray}}"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        case, field = k.split("/", 1)
        out.setdefault(case, {})[field] = z[k]
    return out


def bits_to_torch(bits: np.ndarray, dtype: str, device="cpu") -> torch.Tensor:
    if dtype == "f32":
        t = torch.from_numpy(bits.astype(np.uint32).view(np.int32).copy()).view(torch.float32)
    else:
        t = torch.from_numpy(bits.astype(np.uint16).view(np.int16).copy()).view(TDT[dtype])
    return t.to(device)


def torch_to_bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32).copy()
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def assert_same_bits(got: np.ndarray, want: np.ndarray, dtype: str, what=""):
    """Bit equality, NaNs compared as NaNs (payload / sign of a NaN is not part of the contract)."""
    import q100_oracle as O
    got = got.reshape(want.shape)
    if np.array_equal(got, want):
        return
    a, b = O.from_bits(got, dtype), O.from_bits(want, dtype)
    both_nan = np.isnan(a) & np.isnan(b)
    bad = (got != want) & ~both_nan
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {int(bad.sum())} / {bad.size} elements differ, first at {tuple(i)}: "
                             f"got {a[tuple(i)]!r} ({got[tuple(i)]:#x}) want {b[tuple(i)]!r} ({want[tuple(i)]:#x})")
