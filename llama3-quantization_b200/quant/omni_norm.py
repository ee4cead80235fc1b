"""Drop-in for the reference's ``quant/omni_norm.py`` (OmniLayerNorm lines 11-39, OmniLlamaRMSNorm 42-63).

Module surface kept (buffer ``weight``, ``bias``, ``variance_epsilon`` / ``eps``, ``use_temporary_parameter``
with ``temp_weight`` / ``temp_bias``, ``set_quant_state``) because ``models/transformation.py`` and
``quant/utils.py`` of the reference poke these attributes.  The normalisation is tiny and memory-bound and
stays a handful of torch elementwise ops in the reference's order (fp32 variance, scale in the promoted
dtype, affine, cast back); fusing it with the activation fake-quant is SURVEY.md section 8 row f4.
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

    def forward(self, hidden_states):
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
