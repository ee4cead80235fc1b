"""Drop-in for the reference's ``quant/omni_norm.py`` (OmniLayerNorm lines 11-39, OmniLlamaRMSNorm 42-63).

Module surface kept (buffer ``weight``, ``bias``, ``variance_epsilon`` / ``eps``, ``use_temporary_parameter``
with ``temp_weight`` / ``temp_bias``, ``set_quant_state``) because ``models/transformation.py`` and
``quant/utils.py`` of the reference poke these attributes.  The normalisation is tiny and memory-bound and
is a handful of torch elementwise ops in the reference's order (fp32 variance, scale in the promoted dtype, affine,
cast back) wherever autograd or fp32 is involved; 16-bit CUDA inference rows run K4 (q100_add_rmsnorm: one launch,
optionally with the residual add in front) -- same per-op rounding.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _TemporaryAffine(nn.Module):
    """weight/bias selection shared by both norms: LET calibration swaps in temp_weight / temp_bias."""

    use_temporary_parameter = False

    def _affine(self):
        if self.use_temporary_parameter:
            return self.temp_weight, self.temp_bias
        return self.weight, getattr(self, "bias", None)


class OmniLlamaRMSNorm(_TemporaryAffine):
    def __init__(self, ori_norm, eps=1e-6):
        super().__init__()
        self.register_buffer("weight", ori_norm.weight)
        self.bias = None
        self.variance_epsilon = eps

    def _k4_ok(self, x):
        """K4 (q100_add_rmsnorm) computes this module bit for bit up to the order of the fp32 sum of squares: 16-bit
        CUDA rows of at most 8192 columns, no bias, nothing to differentiate."""
        weight, bias = self._affine()
        return (bias is None and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and weight.dtype == x.dtype
                and weight.device == x.device and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192 and x.numel() > 0
                and weight.numel() == x.shape[-1]
                and not (torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)))

    def add_norm(self, x, residual):
        """(residual + x, self(residual + x)): the residual add of int_llama_layer.py:246 / :251 and the norm that
        follows it as one launch (K4) where that applies."""
        if self._k4_ok(x) and residual.dtype == x.dtype and residual.shape == x.shape and residual.is_cuda \
                and not (torch.is_grad_enabled() and residual.requires_grad):
            from q100 import ops as _ops
            return _ops.add_rmsnorm(x, self._affine()[0], self.variance_epsilon, residual)
        h = residual + x
        return h, self.forward(h)

    def forward(self, hidden_states):
        if self._k4_ok(hidden_states):
            from q100 import ops as _ops
            return _ops.add_rmsnorm(hidden_states, self._affine()[0], self.variance_epsilon)[1]
        in_dtype = hidden_states.dtype
        mean_sq = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        normed = hidden_states * torch.rsqrt(mean_sq + self.variance_epsilon)
        weight, bias = self._affine()
        out = weight * normed
        return (out if bias is None else out + bias).to(in_dtype)


class OmniLayerNorm(_TemporaryAffine):
    def __init__(self, ori_layer_norm) -> None:
        super().__init__()
        self.register_buffer("weight", ori_layer_norm.weight)
        if ori_layer_norm.bias is None:
            self.bias = None
        else:
            self.register_buffer("bias", ori_layer_norm.bias)
        self.eps, self.normalized_shape = ori_layer_norm.eps, ori_layer_norm.normalized_shape
        self.norm_func = F.layer_norm
        self.use_act_quant = True

    def forward(self, x):
        weight, bias = self._affine()
        return self.norm_func(x, self.normalized_shape, weight, bias, eps=self.eps)

    def set_quant_state(self, use_weight_quant, use_act_quant):
        self.use_act_quant = use_act_quant
