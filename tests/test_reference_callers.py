"""GPU: the reference's UNMODIFIED callers on top of the drop-in (SURVEY.md section 7 step 1 "done when", section 8b).

tests/run_omniquant.py runs quant/omniquant.py::omniquant -- the LWC calibration loop with its AdamW / AMP scaler, then
smooth_and_quant_inplace, the propagation step and register_scales_and_zeros -- from the vendored reference
(oracle/_ref, oracle/build_ref.py) on one tiny LLaMA block, three times in subprocesses:
   ref              reference quantizer / QuantLinear / QuantMatMul / decoder block (eager torch on the GPU)
   dropin           q100 quant/* and models/int_llama_layer.py under the same unmodified driver
   dropin_refblock  q100 quant/* under the reference's own models/int_llama_layer.py
The drop-in runs with Q100_CUDA_SCALE=1 (the reference's CUDA scale semantics), so the forward of every fake-quant is
bit-identical; what differs is the backward (K1b sums the group's gradient in one pass, autograd in a dozen ops) and
therefore, slightly, the AdamW trajectory.  Tolerances are stated at the assertions.
"""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "quant", "omniquant.py"))
_RUNS = {}


def _run(impl, wbits=4):
    key = (impl, wbits)
    if key not in _RUNS:
        out = tempfile.mktemp(suffix=f"_{impl}.pt")
        env = dict(os.environ, Q100_CUDA_SCALE="1", PYTHONHASHSEED="0")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_omniquant.py"), "--impl", impl,
                            "--out", out, "--wbits", str(wbits)], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0 and "OMNIQUANT RUN OK" in r.stdout, (r.stdout[-2000:] + r.stderr[-4000:])
        _RUNS[key] = torch.load(out)
        os.remove(out)
    return _RUNS[key]


def _compare(a, b, what):
    # per-epoch mean loss the driver logs (omniquant.py:292): same data, same init; the trajectories may only
    # separate through the backward's summation order
    assert len(a["losses"]) == len(b["losses"]) == 2
    for la, lb in zip(a["losses"], b["losses"]):
        assert abs(la - lb) <= 0.02 * abs(lb) + 1e-9, f"{what}: loss {la} vs {lb}"
    # learned clipping factors: AdamW moves each by ~lr = 1e-2 per step (8 steps); a gradient whose sign is in the
    # noise may take a different step, so single entries may differ by a few lr, the bulk must agree closely
    for n in a["bounds"]:
        d = (a["bounds"][n] - b["bounds"][n]).abs()
        assert float(d.mean()) <= 2e-3, f"{what}: {n} mean |delta| {float(d.mean()):.3e}"
        assert float(d.max()) <= 8e-2, f"{what}: {n} max |delta| {float(d.max()):.3e}"
    # in-place fake-quantised weights (quant/utils.py:133-136) on the learned bounds: a clipping factor that differs
    # by 1e-3 moves a few codes of its group by one step
    for n in a["weights"]:
        wa, wb = a["weights"][n].float(), b["weights"][n].float()
        frac = float((wa != wb).float().mean())
        assert frac <= 0.05, f"{what}: {n}: {frac:.4f} of the weights differ"
        assert float((wa - wb).abs().max()) <= 1.01 * float(b["scales"][n].float().max()), f"{what}: {n} differs by more than one step"
        assert a["scales"][n].shape == b["scales"][n].shape and a["zeros"][n].shape == b["zeros"][n].shape
    rel = float((a["y"] - b["y"]).abs().max() / b["y"].abs().max())
    assert rel <= 5e-2, f"{what}: calibrated block output differs by {rel:.3e}"


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (vendored reference, oracle/build_ref.py) is not on this box")
def test_unmodified_omniquant_loop_on_dropin_matches_reference():
    ref, drop = _run("ref"), _run("dropin")
    assert "oracle/_ref" in ref["where"]["quantizer"] and "llama3-quantization_b200" in drop["where"]["quantizer"]
    assert "llama3-quantization_b200" in drop["where"]["int_llama_layer"]
    _compare(drop, ref, "drop-in vs reference")


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (vendored reference, oracle/build_ref.py) is not on this box")
def test_reference_decoder_block_runs_unmodified_on_dropin_quant():
    ref, drop = _run("ref"), _run("dropin_refblock")
    assert "oracle/_ref" in drop["where"]["int_llama_layer"] and "llama3-quantization_b200" in drop["where"]["quantizer"]
    _compare(drop, ref, "reference block on drop-in quant vs reference")
