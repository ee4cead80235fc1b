"""GPU: parity with the reference AS DEPLOYED (its modules on CUDA tensors), Q100_FLAG_CUDA_SCALE.

tests/golden/cuda_eager.npz was written by oracle/make_golden_cuda.py on a B200 from the unmodified reference quantizer
running on the GPU.  torch's CUDA kernels turn `tensor / python_scalar` into a multiplication by the fp32 reciprocal
(its CPU kernels divide), so the reference's scale differs by up to one ulp between its own two deployments; the
kernels follow the CPU one by default (tests/golden/weights.npz) and the CUDA one with the flag.  Bit-exact in both.
When oracle/_ref travelled to the box the same comparison also runs live on fresh inputs at a real weight size.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, TDT, bits_to_torch, torch_to_bits
from q100 import _lib, ops

pytestmark = pytest.mark.gpu
PATH = os.path.join(GOLDEN, "cuda_eager.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    if not os.path.exists(PATH):
        return {}
    z = np.load(PATH)
    out = {}
    for k in z.files:
        case, field = k.split("/", 1)
        out.setdefault(case, {})[field] = z[k]
    out.pop("_meta", None)
    return out


CASES = _cases()


@pytest.mark.skipif(not CASES, reason="tests/golden/cuda_eager.npz not generated yet (oracle/make_golden_cuda.py)")
@pytest.mark.parametrize("case", sorted(CASES))
def test_fixture_from_reference_on_cuda(case):
    c = CASES[case]
    dt = case.split("_")[0]
    n_bits, group, sym, nozp, lwc = [int(v) for v in c["cfg"]]
    x = bits_to_torch(c["x"], dt, "cuda")
    up = low = None
    if lwc:
        cdt = "f32" if c["up"].dtype == np.uint32 else dt
        up, low = bits_to_torch(c["up"], cdt, "cuda"), bits_to_torch(c["low"], cdt, "cuda")
        x = x.to(up.dtype)
    for flags in (_lib.FLAG_CUDA_SCALE, _lib.FLAG_CUDA_SCALE | _lib.FLAG_K1_GENERIC,
                  _lib.FLAG_CUDA_SCALE | _lib.FLAG_SCALAR_MATH):
        r = ops.quantize_groupwise(x, group, n_bits, bool(sym), bool(nozp), up, low, flags=flags)
        want_dt = "f32" if x.dtype == torch.float32 else dt
        got = torch_to_bits(r.dequant)
        want = c["y"].reshape(got.shape)
        if want_dt != dt:                       # lwc promoted the 16-bit input to fp32: the fixture's y is fp32 too
            assert c["y"].dtype == np.uint32
        nan_ok = np.isnan(r.dequant.float().cpu().numpy())
        assert np.array_equal(got[~nan_ok], want[~nan_ok]), f"{case} flags={flags:#x}: dequant differs"
        assert np.array_equal(torch_to_bits(r.scales).ravel(), c["scale"].ravel()), f"{case}: scale differs"
        if not nozp:
            assert np.array_equal(torch_to_bits(r.zeros).ravel(), c["zp"].ravel()), f"{case}: zero-point differs"


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "quant", "quantizer.py")),
                    reason="oracle/_ref (vendored reference, oracle/build_ref.py) is not on this box")
@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("n_bits", [4, 3, 2, 8])
def test_live_against_reference_module_on_cuda(dt, n_bits):
    """up_proj-sized weight through the reference's UniformAffineQuantizer on the GPU vs the drop-in with
    cuda_eager_scale: identical bits (subprocess: the two `quant` packages cannot share an interpreter)."""
    import subprocess
    code = f"""
import sys, torch
sys.path[:0] = [r"{ROOT}/oracle/_ref", r"{ROOT}/oracle/_ref/stubs"]
import quant.quantizer as RQ
torch.manual_seed(0)
x = (torch.randn(14336, 4096, device="cuda") * 0.02).to(torch.{str(TDT[dt]).split('.')[-1]})
q = RQ.UniformAffineQuantizer(n_bits={n_bits}, symmetric=False, per_channel_axes=[0], dynamic_method="per_channel",
                              group_size=128, shape=x.shape, lwc=False, disable_zero_point=False)
with torch.no_grad():
    y = q(x)
torch.save({{"x": x.cpu(), "y": y.cpu(), "scale": q.scale.cpu(), "zp": q.round_zero_point.cpu()}}, sys.argv[1])
"""
    import tempfile
    f = tempfile.mktemp(suffix=".pt")
    r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(f)
    os.remove(f)
    x = ref["x"].cuda()
    got = ops.quantize_groupwise(x, 128, n_bits, flags=_lib.FLAG_CUDA_SCALE)
    assert torch.equal(got.scales.cpu().view(-1), ref["scale"].view(-1)), "scale differs from the reference on CUDA"
    assert torch.equal(got.zeros.cpu().view(-1), ref["zp"].view(-1)), "zero-point differs from the reference on CUDA"
    assert np.array_equal(torch_to_bits(got.dequant), torch_to_bits(ref["y"])), "dequant differs from the reference on CUDA"
    # and how far the default (CPU semantics: true division) is from it: never more than one ulp of the scale.  In fp32
    # most groups ARE one ulp apart (measured on B200: 72 % at W4 -- x / 15 against x * RN(1 / 15)); after the rounding
    # to 16 bits almost none are.
    d = ops.quantize_groupwise(x, 128, n_bits)
    a, b = d.scales.cpu().view(-1).float(), ref["scale"].view(-1).float()
    ulp = torch.finfo(TDT[dt]).eps * b.abs()
    assert bool(((a - b).abs() <= ulp).all())
