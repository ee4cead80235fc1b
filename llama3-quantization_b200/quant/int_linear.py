"""Drop-in for the reference's ``quant/int_linear.py`` (QuantLinear, lines 11-69).

Identical module surface (buffers ``weight`` / ``bias``, ``use_weight_quant``,
``use_act_quant``, ``use_temporary_parameter``, ``temp_weight`` / ``temp_bias``,
``weight_quantizer`` / ``act_quantizer``, ``fwd_func`` / ``fwd_kwargs``,
``set_quant_state``).  ``forward`` keeps the reference's three-way weight choice
(int_linear.py:48-65); what changes is what runs underneath:

* the per-token activation fake-quant is one K2 kernel;
* when the weight is quantised and no gradient is needed the matmul reads the PACKED
  codes through the fused dequant-GEMM (K3) instead of ``F.linear`` on a dequantised
  fp16 copy: 0.5 B/weight instead of 2 B/weight at W4.  The dequantised weight the
  kernel reconstructs in registers is bit-identical to the reference's fake-quantised
  weight, so only the fp32 accumulation order differs from ``F.linear``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from q100 import ops as _ops
from q100._lib import GEMM_X_GATE_UP as _GEMM_X_GATE_UP
from quant.quantizer import UniformAffineQuantizer


class QuantLinear(nn.Module):
    """
    Quantized Module that can perform quantized convolution or normal convolution.
    To activate quantization, please use set_quant_state function.
    """
    def __init__(
        self,
        org_module: nn.Linear,
        weight_quant_params: dict = {},
        act_quant_params: dict = {},
        disable_input_quant=False,
    ):
        super().__init__()
        self.fwd_kwargs = dict()
        self.fwd_func = F.linear
        self.register_buffer('weight', org_module.weight)
        if org_module.bias is not None:
            self.register_buffer('bias', org_module.bias)
        else:
            self.bias = None
        self.in_features = org_module.in_features
        self.out_features = org_module.out_features
        # de-activate the quantized forward default
        self.use_weight_quant = False
        self.use_act_quant = False
        # initialize quantizer
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params, shape=org_module.weight.shape)
        self.weight_quantizer.keep_packed = True
        if not disable_input_quant:
            self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        else:
            self.act_quantizer = None

        self.disable_input_quant = disable_input_quant
        self.use_temporary_parameter = False
        # fused dequant-GEMM switch (extension; True keeps reference results within GEMM tolerance)
        self.use_fused_gemm = True
        self._requant_cache = None
        # tensor parallelism (extension): a q100.peer.ScatterReduce makes forward() of a row-parallel layer return the
        # ALL-REDUCED output at prefill sizes (reduce-scatter as the GEMM's epilogue); `reduced` tells the caller
        # whether the last forward did, so that it skips its own all-reduce
        self.tp_scatter = None
        self.reduced = False

    # ------------------------------------------------------------------ fused path plumbing
    def _fused_ok(self, x: torch.Tensor, packed) -> bool:
        return (packed is not None and self.use_fused_gemm and self.fwd_func is F.linear
                and not self.fwd_kwargs and x.is_cuda and x.dtype == packed.dtype
                and packed.K % 128 == 0 and packed.N % 8 == 0
                and (packed.group in (32, 64) or packed.group % 128 == 0))

    def _packed_requant(self):
        """use_weight_quant=True re-quantises ``self.weight`` on every call in the reference
        (int_linear.py:52-53); the result only depends on the weight and the quantizer state, so
        it is cached on (storage, version, quantizer config)."""
        wq = self.weight_quantizer
        w = self.weight
        if wq.n_bits not in (2, 3, 4, 8) or w.dtype not in (torch.float16, torch.bfloat16) \
                or not w.is_cuda or w.dim() != 2 or wq.deficiency != 0 or w.shape[1] % 128 != 0 \
                or not wq.enable or wq.metric == "fix0to1":
            return None
        lwc_ver = (wq.upbound_factor._version, wq.lowbound_factor._version) if wq.lwc else None
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype, wq.n_bits, wq.group_size,
               wq.symmetric, wq.disable_zero_point, lwc_ver, wq.cuda_eager_scale)
        if self._requant_cache is not None and self._requant_cache[0] == key:
            pw, scale, zp = self._requant_cache[1]
            wq.scale, wq.round_zero_point = scale, zp      # same side effects as a real call
            return pw
        with torch.no_grad():
            out = wq(w)                                    # K1: dequant + packed + side band, one pass
        pw = wq.packed_for(out)
        if pw is None:
            return None
        self._requant_cache = (key, (pw, wq.scale, wq.round_zero_point))
        return pw

    def packed_operand(self, input: torch.Tensor):
        """The PackedWeight ``forward(input)`` would hand to the fused dequant-GEMM, or None when
        forward would take the dense ``F.linear`` route (temporary parameters, autograd, custom
        ``fwd_func``, unsupported shape).  Lets a caller that owns several linears on the same
        input (q|k|v, gate|up) run them as one launch (models/int_llama_layer.py)."""
        if self.use_temporary_parameter or not self.use_fused_gemm or self.fwd_func is not F.linear \
                or self.fwd_kwargs:
            return None
        if self.use_weight_quant:
            if torch.is_grad_enabled() and (
                    input.requires_grad or self.weight.requires_grad
                    or any(p.requires_grad for p in self.weight_quantizer.parameters())):
                return None
            packed = self._packed_requant()
        else:
            if torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad):
                return None
            packed = self.weight_quantizer.packed_for(self.weight)
        if packed is None or not input.is_cuda:
            return None
        if not (packed.K % 128 == 0 and packed.N % 8 == 0
                and (packed.group in (32, 64) or packed.group % 128 == 0)):
            return None
        return packed

    def _fused_act_gemm(self, x: torch.Tensor, packed, bias):
        """W8A8 / W4A4 at decode shapes: the per-token activation quantisation runs as the prologue of the integer
        dequant-GEMM (K3c, q100_gemm_wna8) instead of as a K2 launch in front of a 16-bit-activation GEMM.  W8A8 with
        per-channel weights at more than 15 rows: K2 in pack mode + the integer GEMM K3d (q100_gemm_w8a8).  None when
        the case does not qualify (then forward() takes the reference's two steps)."""
        aq = self.act_quantizer
        rows = x.numel() // x.shape[-1] if x.dim() else 0
        if (packed is None or aq is None or not aq.enable or aq.n_bits > 8 or aq.symmetric or aq.group_size
                or aq.disable_zero_point or aq.lwc or aq.metric == "fix0to1" or packed.n_bits not in (4, 8)
                or packed.group % 128 != 0 or not self._fused_ok(x, packed)
                or (torch.is_grad_enabled() and x.requires_grad)):
            return None
        # more than 15 rows: K3d (codes x codes on tcgen05 kind::i8) for the W8A8 per-channel configuration; everything
        # else keeps the reference's two steps at these sizes
        prefill = rows > 15
        if prefill and not (aq.n_bits == 8 and _ops.gemm_w8a8_supported(rows, packed)):
            return None
        from q100._lib import FLAG_CUDA_SCALE
        fl = FLAG_CUDA_SCALE if aq.cuda_eager_scale else 0
        mm = None
        tp_stats = False
        if aq.stat_group is not None:
            import torch.distributed as dist
            tp_stats = dist.is_initialized() and dist.get_world_size(aq.stat_group) > 1
        if tp_stats or not prefill:
            mm = _ops.rows_minmax(x)
            if tp_stats:
                from q100 import tp as _tp
                mm = _tp.allreduce_minmax(mm, aq.stat_group)
        if prefill:
            return _ops.gemm_w8a8(x, packed, mm, bias, fl)
        return _ops.gemm_wna8(x, packed, aq.n_bits, mm, bias, fl)

    def release_dense_weight(self) -> bool:
        """Inference-only extension: free the dense fake-quantised ``weight`` buffer and keep only its
        packed twin (0.5 B/weight at W4 instead of 2.5).  ``weight`` becomes an empty tensor of the same
        dtype/device; forward() keeps working through the fused dequant-GEMM.  Returns False (and changes
        nothing) when there is no packed twin."""
        pw = self.weight_quantizer.packed_for(self.weight)
        if pw is None:
            return False
        self.weight = torch.empty(0, dtype=self.weight.dtype, device=self.weight.device)
        self.weight_quantizer._packed_of = (self.weight.data_ptr(), self.weight._version,
                                            tuple(self.weight.shape), self.weight.dtype)
        return True

    def forward_gate_up(self, gate_up: torch.Tensor):
        """``self(act_fn(gate) * up)`` for ``gate_up = gate | up`` (SiLU): at decode shapes the packed GEMM applies the
        element-wise step while it stages x (Q100_GEMM_X_GATE_UP), otherwise K6 + the normal forward."""
        rows = gate_up.numel() // gate_up.shape[-1]
        self.reduced = False
        if (rows <= 8 and not self.use_temporary_parameter and not self.use_weight_quant
                and not (self.use_act_quant and not self.disable_input_quant) and not torch.is_grad_enabled()):
            packed = self.weight_quantizer.packed_for(self.weight)
            if packed is not None and gate_up.dtype == packed.dtype and gate_up.shape[-1] == 2 * packed.K \
                    and packed.group % 128 == 0 and self._fused_ok(gate_up, packed):
                return _ops.gemm_wna16(gate_up, packed, self.bias, flags=_GEMM_X_GATE_UP)
        return self.forward(_ops.silu_mul(gate_up))

    def forward(self, input: torch.Tensor):
        packed = None
        weight = None
        self.reduced = False
        if self.use_temporary_parameter:
            weight = self.temp_weight
            bias = self.temp_bias
        elif self.use_weight_quant:
            bias = self.bias
            needs_grad = torch.is_grad_enabled() and (
                input.requires_grad or self.weight.requires_grad
                or any(p.requires_grad for p in self.weight_quantizer.parameters()))
            if not needs_grad and self.use_fused_gemm:
                packed = self._packed_requant()
            if packed is None:
                weight = self.weight_quantizer(self.weight)
        else:
            weight = self.weight
            bias = self.bias
            if not (torch.is_grad_enabled() and (input.requires_grad or weight.requires_grad)):
                packed = self.weight_quantizer.packed_for(weight)

        if self.use_act_quant and not self.disable_input_quant:
            y = self._fused_act_gemm(input, packed, bias)
            if y is not None:
                return y
            input = self.act_quantizer(input)

        if packed is not None:
            x = input
            if x.dtype != packed.dtype and torch.is_autocast_enabled():
                x = x.to(packed.dtype)                     # what autocast does in front of F.linear
            if self._fused_ok(x, packed):
                sc = self.tp_scatter
                if sc is not None and bias is None and sc.fits(x.numel() // x.shape[-1], packed):
                    self.reduced = True
                    return sc(x, packed)
                return _ops.gemm_wna16(x, packed, bias)
            if weight is None:
                weight = self.weight_quantizer(self.weight)

        out = self.fwd_func(input, weight, bias, **self.fwd_kwargs)
        return out

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
