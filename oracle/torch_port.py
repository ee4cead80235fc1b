"""Torch restatement of the reference hot path, op for op.  TEST / BASELINE INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path never does.

It issues the same eager torch op sequence as the reference
(``quant/quantizer.py:94-159``, ``quant/int_linear.py:48-65``) so that

* timed on the host cores it costs what the reference's CPU path costs (same ATen
  kernels, same number of passes over the tensor) -- this is ``bench.py``'s
  ``cpu_baseline`` (kind "port") and the ``--impl reference`` arm, because the
  reference checkout itself does not exist on the GPU box;
* run under autograd it yields the gradients torch derives for the reference, which
  is what the drop-in's custom backward is compared against.

It is functional (no nn.Module state) and is pinned against the unmodified
reference by ``oracle/make_golden.py`` via the numpy oracle: the golden fixtures
hold the reference's outputs, and tests check this port reproduces them exactly.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CLIPMIN = 1e-5


def calibrate(xg: torch.Tensor, n_bits: int, symmetric: bool = False, disable_zero_point: bool = False,
              up=None, low=None):
    """quantizer.py:140-159 on [groups, g] (or any shape, reducing the last dim)."""
    xmin = xg.amin([-1], keepdim=True)
    xmax = xg.amax([-1], keepdim=True)
    if up is not None:
        xmax = torch.sigmoid(up) * xmax
        xmin = torch.sigmoid(low) * xmin
    if symmetric:
        abs_max = torch.max(xmax.abs(), xmin.abs())
        scale = (abs_max / (2 ** (n_bits - 1) - 1)).clamp(min=CLIPMIN, max=1e4)
        zero_point = (2 ** (n_bits - 1) - 1) * torch.ones_like(scale)
    else:
        scale = ((xmax - xmin) / (2 ** n_bits - 1)).clamp(min=CLIPMIN, max=1e4)
        zero_point = -(xmin) / scale
    if disable_zero_point:
        return scale, None
    return scale, zero_point.clamp(min=-1e4, max=1e4).round()


def fake_quant(x: torch.Tensor, n_bits: int, group=None, symmetric: bool = False,
               disable_zero_point: bool = False, up=None, low=None):
    """UniformAffineQuantizer.forward (quantizer.py:118-130).  Returns (dequant, scale, zp)."""
    if disable_zero_point:
        qmin, qmax = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
    else:
        qmin, qmax = 0, 2 ** n_bits - 1
    shape = x.shape
    xg = x.reshape(-1, group) if group else x
    scale, zp = calibrate(xg, n_bits, symmetric, disable_zero_point, up, low)
    v = xg / scale
    x_int = (v.round() - v).detach() + v
    if zp is not None:
        x_int = x_int.add(zp)
    x_int = x_int.clamp(qmin, qmax)
    deq = x_int
    if zp is not None:
        deq = deq.sub(zp)
    deq = deq.mul(scale)
    return deq.reshape(shape), scale, zp


def quant_linear(x, w, bias, w_bits, group, a_bits=16, symmetric=False, disable_zero_point=False,
                 requantize=True):
    """QuantLinear.forward (int_linear.py:48-65): optional weight re-quantisation, optional
    per-token activation fake-quant, then F.linear on the dense fake-quantised weight."""
    if requantize and w_bits < 16:
        w = fake_quant(w, w_bits, group, symmetric, disable_zero_point)[0]
    if a_bits < 16:
        x = fake_quant(x, a_bits, None)[0]
    return F.linear(x, w, bias)
