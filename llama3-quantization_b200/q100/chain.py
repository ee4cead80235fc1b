"""Decode chains (K9, include/q100.h): one persistent launch for the dequant-GEMMs of one or many decoder blocks
at decode shapes, with RMSNorm / residual adds / SiLU*up folded into the GEMMs' x staging and epilogues and the
single-token attention as phases of the same launch.

:class:`Chain` is the thin host object: it owns the device-side descriptor, the scratch workspace and the small
fp32 side buffers (partial sums of squares for the RMSNorms, flash-decoding partials), keeps every tensor the phases
point to alive, and launches.  ``Chain.for_model`` lays out the phases of a whole ``QuantLlamaModel`` decode step.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import Phase, check, lib

_DT = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def ssq_parts(N: int) -> int:
    return int(lib.q100_chain_ssq_parts(N))


class Chain:
    def __init__(self, M: int, group: int, n_bits: int, disable_zp: bool, dtype: torch.dtype, device: torch.device):
        self.M, self.group, self.n_bits, self.disable_zp = int(M), int(group), int(n_bits), bool(disable_zp)
        self.dtype, self.device = dtype, device
        self.phases: List[Phase] = []
        self._keep: list = []                 # every tensor a phase points to
        self.desc = self.ws = None
        self.built = False

    # ------------------------------------------------------------------ phase constructors
    def _p(self, t: Optional[torch.Tensor]):
        if t is None:
            return None
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("q100.chain: tensors must be contiguous CUDA tensors")
        self._keep.append(t)
        return t.data_ptr()

    def gemm(self, x, w, y, bias=None, x_mode=_lib.X_PLAIN, norm_weight=None, ssq_in=None, eps=0.0,
             epi_mode=_lib.EPI_STORE, residual=None, ssq_out=None, attn=None):
        """y = x . dequant(w)^T; ``w`` is a q100.ops.PackedWeight.  ``attn``: the dict :meth:`attention` returned, for
        ``x_mode=X_ATTN`` (x merged from the attention partials while it is staged)."""
        if (w.n_bits, w.group, w.disable_zero_point, w.dtype) != (self.n_bits, self.group, self.disable_zp, self.dtype):
            raise ValueError("q100.chain: every weight of a chain must share bit width, group, zero-point mode and dtype")
        kx = 2 * w.K if x_mode == _lib.X_GATE_UP else w.K
        if tuple(x.shape) != (self.M, kx) or tuple(y.shape) != (self.M, w.N) or x.dtype != self.dtype or y.dtype != self.dtype:
            raise ValueError(f"q100.chain: x must be [{self.M}, {kx}], y [{self.M}, {w.N}] in {self.dtype}")
        ph = Phase()
        ph.type = _lib.PH_GEMM
        ph.x, ph.packed, ph.meta, ph.bias, ph.y = self._p(x), self._p(w.packed), self._p(w.meta), self._p(bias), self._p(y)
        ph.N, ph.K = w.N, w.K
        ph.x_mode, ph.epi_mode, ph.eps = x_mode, epi_mode, float(eps)
        if x_mode == _lib.X_NORM:
            if ssq_in.dtype != torch.float32 or ssq_in.shape[1] != self.M:
                raise ValueError("q100.chain: ssq_in must be fp32 [parts, M]")
            ph.norm_weight, ph.ssq_in, ph.ssq_in_parts = self._p(norm_weight.to(self.dtype)), self._p(ssq_in), ssq_in.shape[0]
        if x_mode == _lib.X_ATTN:
            if attn["heads"] * 128 != w.K:
                raise ValueError("q100.chain: X_ATTN needs K == heads * 128")
            ph.att_partials, ph.position = self._p(attn["partials"]), self._p(attn["position"])
            ph.max_len, ph.heads = attn["max_len"], attn["heads"]
        if epi_mode == _lib.EPI_ADD_SSQ:
            if tuple(residual.shape) != (self.M, w.N) or tuple(ssq_out.shape) != (ssq_parts(w.N), self.M) \
                    or ssq_out.dtype != torch.float32:
                raise ValueError("q100.chain: residual must be [M, N], ssq_out fp32 [ssq_parts(N), M]")
            ph.residual, ph.ssq_out = self._p(residual), self._p(ssq_out)
        self.phases.append(ph)
        return ph

    def attention(self, qkv, cos, sin, position, k_cache, v_cache, att_out, heads, kv_heads, combine=True):
        """Rotary + cache append + attention of the new token.  ``combine=True`` adds the phase that merges the
        position chunks into ``att_out``; with False the consumer is a GEMM phase with ``x_mode=X_ATTN`` that takes the
        returned dict as ``attn``."""
        bsz, kvh, max_len, hd = k_cache.shape
        n = int(lib.q100_chain_attn_partial_floats(bsz, heads, max_len))
        part = torch.zeros(n, dtype=torch.float32, device=self.device)
        for t in ((_lib.PH_ATTN, _lib.PH_COMBINE) if combine else (_lib.PH_ATTN,)):
            ph = Phase()
            ph.type = t
            ph.qkv, ph.cos_cache, ph.sin_cache, ph.position = self._p(qkv), self._p(cos), self._p(sin), self._p(position)
            ph.k_cache, ph.v_cache, ph.att_partials, ph.att_out = self._p(k_cache), self._p(v_cache), self._p(part), self._p(att_out)
            ph.batch, ph.heads, ph.kv_heads, ph.head_dim = bsz, heads, kv_heads, hd
            ph.max_len, ph.rope_len = max_len, cos.shape[0]
            self.phases.append(ph)
        return {"partials": part, "position": position, "max_len": max_len, "heads": heads}

    # ------------------------------------------------------------------ build / run
    def build(self):
        n = len(self.phases)
        arr = (Phase * n)(*self.phases)
        nbytes = int(lib.q100_chain_desc_bytes(n))
        self.desc = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        need = C.c_size_t(0)
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            # first pass: size the workspace
            rc = lib.q100_chain_build(arr, n, self.M, self.group, self.n_bits, int(self.disable_zp), _DT[self.dtype],
                                      self.desc.data_ptr(), nbytes, None, 0, C.byref(need), st)
            if rc not in (_lib.OK, _lib.EWORKSPACE):
                check(rc)
            self.ws = torch.zeros(max(int(need.value), 16), dtype=torch.uint8, device=self.device)
            check(lib.q100_chain_build(arr, n, self.M, self.group, self.n_bits, int(self.disable_zp), _DT[self.dtype],
                                       self.desc.data_ptr(), nbytes, self.ws.data_ptr(), self.ws.numel(),
                                       C.byref(need), st))
        self.built = True
        return self

    def run(self):
        if not self.built:
            raise RuntimeError("q100.chain: build() first")
        check(lib.q100_chain_run(self.desc.data_ptr(), len(self.phases), self.M, self.group, self.n_bits,
                                 int(self.disable_zp), _DT[self.dtype], self.ws.data_ptr(),
                                 torch.cuda.current_stream(self.device).cuda_stream))

    def check(self):
        """Raise if a launch reported an error (grid-barrier time-out, position outside the KV cache); synchronises."""
        check(lib.q100_chain_error(self.desc.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
