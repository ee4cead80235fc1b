"""Drop-in for the reference's ``quant/int_matmul.py`` (QuantMatMul, lines 7-43).

Surface kept: ctor ``(x1_quant_params, x2_quant_params, disable_act_quant, matmul_func)``, attributes
``use_act_quant`` / ``x1_quantizer`` / ``x2_quantizer`` / ``matmul_func`` / ``disable_act_quant`` /
``i_cluster_counts``, methods ``set_quant_state`` / ``quant_x1`` / ``quant_x2`` / ``forward``.
The two operand quantisers are q100 ``UniformAffineQuantizer``s, i.e. one K2 launch per operand
(per-token, per-head statistics over the last dim); the product itself is the caller's ``matmul_func``
(``torch.matmul`` for LLaMA, models/int_llama_layer.py:90-95).
"""
import torch
import torch.nn as nn

from quant.quantizer import UniformAffineQuantizer


class QuantMatMul(nn.Module):
    def __init__(self, x1_quant_params: dict = {}, x2_quant_params: dict = {}, disable_act_quant=False,
                 matmul_func=torch.bmm):
        super().__init__()
        self.matmul_func = matmul_func
        self.disable_act_quant = disable_act_quant
        self.i_cluster_counts = None
        self.use_act_quant = False            # quantised forward is opt-in (set_quant_state)
        for slot, params in (("x1_quantizer", x1_quant_params), ("x2_quantizer", x2_quant_params)):
            setattr(self, slot, UniformAffineQuantizer(**params))

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant

    def _maybe_quant(self, quantizer, t):
        return quantizer(t) if self.use_act_quant else t

    def quant_x1(self, x1):
        return self._maybe_quant(self.x1_quantizer, x1)

    def quant_x2(self, x2):
        return self._maybe_quant(self.x2_quantizer, x2)

    def forward(self, x1, x2):
        return self.matmul_func(x1, x2)
