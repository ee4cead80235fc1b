"""One-shot all-reduce over peer memory (K7, csrc/q100_peer.cu) for the small row-parallel outputs of a
tensor-parallel decode step.  Staging buffers and flags live in torch symmetric memory (peer-mapped over NVLink);
if that cannot be set up (single rank, no peer access, older torch) `get()` returns None and the caller stays on
NCCL.  Setup is collective: every rank of the group must call :func:`enable` at the same point, outside CUDA-graph
capture."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib

_DT = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}
_STATE: dict = {}
FLAG_WORDS = 32                       # K7_MAX_CTAS


class OneShotAllReduce:
    def __init__(self, group, device: torch.device, slot_elems: int):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = device
        self.slot_elems = int(slot_elems)
        # 1) everything that can fail locally, 2) agree on it with ONE collective, 3) only then the collectives of
        # the setup itself (rendezvous, barrier) -- a rank that raised in (1) must not leave its peers in a barrier
        err = None
        try:
            self.buf = symm.empty(2 * self.world * self.slot_elems, dtype=torch.float16, device=device)
            self.flags = symm.empty(2 * self.world * FLAG_WORDS, dtype=torch.int32, device=device)
            self.flags.zero_()
        except Exception as e:                                     # noqa: BLE001
            err = e
        ok = torch.tensor([0.0 if err is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0:
            raise RuntimeError(f"symmetric-memory allocation failed on some rank ({err!r})")
        self.h_buf = symm.rendezvous(self.buf, self.group)
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.epoch = torch.zeros(FLAG_WORDS, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                                  # every rank's flags are zero before the first call

    def __call__(self, y: torch.Tensor) -> torch.Tensor:
        """In-place all-reduce(sum) of a contiguous fp16 / bf16 tensor of at most slot_elems elements."""
        n = y.numel()
        if y.data_ptr() % 16:
            raise ValueError("q100.peer: tensor must be 16-byte aligned")    # never a per-rank path choice (fits())
        with torch.cuda.device(self.device):
            check(lib.q100_allreduce_oneshot(
                y.data_ptr(), y.data_ptr(), n, _DT[y.dtype], self.h_buf.buffer_ptrs_dev, self.h_flags.buffer_ptrs_dev,
                self.epoch.data_ptr(), self.slot_elems, self.rank, self.world, self.err.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return y

    def check(self) -> None:
        """Raise if any call so far gave up waiting for a peer (the kernel sets err_flag instead of hanging)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("q100.peer: a one-shot all-reduce timed out waiting for a peer; its results are invalid")

    def fits(self, y: torch.Tensor) -> bool:
        # only properties that are identical on every rank may decide the path (NCCL vs peer): shape, dtype, layout
        return (y.is_cuda and y.dtype in _DT and y.is_contiguous() and y.numel() % 8 == 0
                and 0 < y.numel() <= self.slot_elems)


class ScatterReduce:
    """Row-parallel all-reduce at prefill sizes fused with the GEMM (K7b, include/q100.h q100_gemm_wna16_scatter +
    q100_reduce_bcast): the GEMM's epilogue stores every output tile over NVLink into the staging buffer of the rank
    that owns those token rows (the reduce-scatter travels while the tensor cores work), one more kernel sums the
    slots in fp32 and stores the result into every rank's y.  ``__call__(x, w)`` returns the COMPLETE ``[M, N]``
    output (a view of one of two alternating symmetric buffers: consume it before the second next call).
    Collective set-up: every rank constructs it at the same point with the same sizes."""

    FLAG_WORDS = 4 * 128                  # [barrier 2][parity 2][rank][128 CTAs] -> times world

    def __init__(self, group, device: torch.device, max_tokens: int, max_cols: int, dtype: torch.dtype):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device, self.dtype = device, dtype
        self.cap = int(max_tokens) * int(max_cols)
        err = None
        try:
            self.stage = symm.empty(self.cap, dtype=dtype, device=device)
            self.y = [symm.empty(self.cap, dtype=dtype, device=device) for _ in range(2)]
            self.flags = symm.empty(self.FLAG_WORDS * self.world, dtype=torch.int32, device=device)
            self.flags.zero_()
        except Exception as e:                                     # noqa: BLE001
            err = e
        ok = torch.tensor([0.0 if err is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0:
            raise RuntimeError(f"symmetric-memory allocation failed on some rank ({err!r})")
        self.h_stage = symm.rendezvous(self.stage, self.group)
        self.h_y = [symm.rendezvous(t, self.group) for t in self.y]
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.epoch = torch.zeros(128, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.turn = 0
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def fits(self, M: int, w) -> bool:
        """Only properties that are identical on every rank decide the path."""
        return (M > 32 and M % (self.world * 256) == 0 and M * w.N <= self.cap and w.dtype == self.dtype
                and w.K % 64 == 0 and w.N % 8 == 0 and w.group % 64 == 0)

    def __call__(self, x: torch.Tensor, w) -> torch.Tensor:
        shp = x.shape
        x2 = x.contiguous().view(-1, w.K)
        M = x2.shape[0]
        st = torch.cuda.current_stream(self.device).cuda_stream
        k = self.turn
        self.turn ^= 1
        with torch.cuda.device(self.device):
            check(lib.q100_gemm_wna16_scatter(x2.data_ptr(), w.packed.data_ptr(), w.meta.data_ptr(), M, w.N, w.K,
                                              w.group, w.n_bits, int(w.disable_zero_point), _DT[x.dtype],
                                              self.h_stage.buffer_ptrs_dev, self.rank, self.world, st))
            check(lib.q100_reduce_bcast(self.stage.data_ptr(), self.h_y[k].buffer_ptrs_dev, self.h_flags.buffer_ptrs_dev,
                                        self.epoch.data_ptr(), M // self.world, w.N, _DT[x.dtype], self.rank,
                                        self.world, self.err.data_ptr(), st))
        return self.y[k][:M * w.N].view(*shp[:-1], w.N)

    def check(self) -> None:
        if int(self.err.item()) != 0:
            raise RuntimeError("q100.peer: a fused reduce timed out waiting for a peer; its results are invalid")


class MulticastReduce:
    """Row-parallel all-reduce at prefill sizes through the switch (K7c, q100_allreduce_multimem): the GEMM writes its
    partial ``[M, N]`` into one of two alternating symmetric buffers that have a multicast mapping, then ONE kernel
    reduces this rank's token rows with ``multimem.ld_reduce`` (the switch adds the ``world`` copies) and broadcasts
    them with ``multimem.st``, in place.  Same interface as :class:`ScatterReduce`; chosen by :func:`enable_scatter`
    when every rank reports a multicast address (NVSwitch).  Collective set-up."""

    FLAG_WORDS = 4 * 128

    def __init__(self, group, device: torch.device, max_tokens: int, max_cols: int, dtype: torch.dtype):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device, self.dtype = device, dtype
        self.cap = int(max_tokens) * int(max_cols)
        err = None
        try:
            self.y = [symm.empty(self.cap, dtype=dtype, device=device) for _ in range(2)]
            self.flags = symm.empty(self.FLAG_WORDS * self.world, dtype=torch.int32, device=device)
            self.flags.zero_()
        except Exception as e:                                     # noqa: BLE001
            err = e
        ok = torch.tensor([0.0 if err is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0:
            raise RuntimeError(f"symmetric-memory allocation failed on some rank ({err!r})")
        self.h_y = [symm.rendezvous(t, self.group) for t in self.y]
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.h_y]
        have = torch.tensor([1.0 if all(self.mc) else 0.0], device=device)
        dist.all_reduce(have, op=dist.ReduceOp.MIN, group=self.group)
        if have.item() == 0:
            raise RuntimeError("no multicast mapping for the symmetric buffers (not an NVSwitch fabric?)")
        self.epoch = torch.zeros(128, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.turn = 0
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def fits(self, M: int, w) -> bool:
        """Only properties that are identical on every rank decide the path."""
        return (M > 32 and M % self.world == 0 and (M // self.world) * w.N % 8 == 0 and M * w.N <= self.cap
                and w.dtype == self.dtype and w.K % 128 == 0 and w.N % 8 == 0)

    def reduce_(self, k: int, M: int, N: int) -> None:
        """In-place all-reduce of the first M * N elements of buffer k (every rank has written its partial there)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            check(lib.q100_allreduce_multimem(self.mc[k], self.h_flags.buffer_ptrs_dev, self.epoch.data_ptr(),
                                              M // self.world, N, _DT[self.dtype], self.rank, self.world,
                                              self.err.data_ptr(), st))

    def __call__(self, x: torch.Tensor, w) -> torch.Tensor:
        from . import ops as _ops
        shp = x.shape
        x2 = x.contiguous().view(-1, w.K)
        M = x2.shape[0]
        k = self.turn
        self.turn ^= 1
        y = self.y[k][:M * w.N].view(M, w.N)
        _ops.gemm_wna16(x2, w, out=y)
        self.reduce_(k, M, w.N)
        return y.view(*shp[:-1], w.N)

    def check(self) -> None:
        if int(self.err.item()) != 0:
            raise RuntimeError("q100.peer: a multicast all-reduce timed out waiting for a peer; its results are invalid")


class ReduceNorm:
    """All-reduce + residual add + RMSNorm in one launch for the row-parallel outputs of a tensor-parallel decode step
    (K7n, q100_allreduce_add_rmsnorm): K4's CTA per token row does K7's exchange itself.  ``cols`` (the hidden size)
    is fixed per instance; at most 64 rows per call.  Collective set-up, outside CUDA-graph capture."""

    MAX_ROWS = 64

    def __init__(self, group, device: torch.device, cols: int):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device, self.cols = device, int(cols)
        err = None
        try:
            self.buf = symm.empty(2 * self.world * self.MAX_ROWS * self.cols, dtype=torch.float16, device=device)
            self.flags = symm.empty(2 * self.world * self.MAX_ROWS, dtype=torch.int32, device=device)
            self.flags.zero_()
        except Exception as e:                                     # noqa: BLE001
            err = e
        ok = torch.tensor([0.0 if err is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0:
            raise RuntimeError(f"symmetric-memory allocation failed on some rank ({err!r})")
        self.h_buf = symm.rendezvous(self.buf, self.group)
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.epoch = torch.zeros(self.MAX_ROWS, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def fits(self, x: torch.Tensor) -> bool:
        """Only properties that are identical on every rank decide the path."""
        return (x.is_cuda and x.dtype in _DT and x.is_contiguous() and x.shape[-1] == self.cols
                and 0 < x.numel() // self.cols <= self.MAX_ROWS)

    def __call__(self, x: torch.Tensor, weight: torch.Tensor, eps: float, residual: Optional[torch.Tensor] = None):
        """(h, norm): h = residual + all_reduce(x) (or the all-reduced x), norm = OmniLlamaRMSNorm(h)."""
        rows = x.numel() // self.cols
        h = torch.empty_like(x)
        norm = torch.empty_like(x)
        if residual is not None:
            residual = residual.contiguous()
        with torch.cuda.device(self.device):
            check(lib.q100_allreduce_add_rmsnorm(
                x.data_ptr(), None if residual is None else residual.data_ptr(), weight.data_ptr(), float(eps),
                h.data_ptr(), norm.data_ptr(), rows, self.cols, _DT[x.dtype], self.h_buf.buffer_ptrs_dev,
                self.h_flags.buffer_ptrs_dev, self.epoch.data_ptr(), self.rank, self.world, self.err.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return h, norm

    def check(self) -> None:
        if int(self.err.item()) != 0:
            raise RuntimeError("q100.peer: a fused all-reduce + norm timed out waiting for a peer; its results are invalid")


_REDUCE_NORM: dict = {}


def enable_reduce_norm(group=None, device: Optional[torch.device] = None, cols: int = 4096) -> Optional[ReduceNorm]:
    """Collective.  Sets up K7n for ``group`` and hidden size ``cols``; None when symmetric memory is not available (the
    caller keeps K7 / NCCL + K4)."""
    key = (group, int(cols))
    if key in _REDUCE_NORM:
        return _REDUCE_NORM[key]
    obj = None
    if dist.is_initialized() and dist.get_world_size(group) > 1 and torch.cuda.is_available():
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        try:
            obj = ReduceNorm(group, dev, cols)
        except Exception as e:                                     # noqa: BLE001
            import warnings
            warnings.warn(f"q100.peer: fused all-reduce + norm unavailable ({e!r})")
            obj = None
    _REDUCE_NORM[key] = obj
    return obj


def get_reduce_norm(group=None, cols: int = 4096) -> Optional[ReduceNorm]:
    return _REDUCE_NORM.get((group, int(cols)))


_SCATTER: dict = {}


def enable_scatter(group=None, device: Optional[torch.device] = None, max_tokens: int = 2048, max_cols: int = 8192,
                   dtype: torch.dtype = torch.bfloat16) -> Optional[ScatterReduce]:
    """Collective.  Sets up the fused row-parallel reduce for ``group``; None (NCCL stays) when symmetric memory is
    not available."""
    key = (group, dtype)
    if key in _SCATTER:
        return _SCATTER[key]
    obj = None
    if dist.is_initialized() and dist.get_world_size(group) > 1 and torch.cuda.is_available():
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        import os
        try:
            # K7c (through the switch, multimem) is opt-in: measured on 8 x B200 it ties K7b on o_proj and loses on
            # down_proj (the scatter of K7b travels under the GEMM; multimem moved 134 MB in 346 us = 0.44 TB/s per
            # direction, plain peer stores reach 0.68), and its sum is not the rank-order fp32 sum
            if os.environ.get("Q100_MULTIMEM") == "1" and not os.environ.get("Q100_NO_MULTIMEM"):
                try:
                    obj = MulticastReduce(group, dev, max_tokens, max_cols, dtype)
                except Exception as e:                             # noqa: BLE001 -- agreed across ranks inside
                    obj = None
                    if os.environ.get("Q100_PEER_VERBOSE"):
                        import warnings
                        warnings.warn(f"q100.peer: multicast all-reduce unavailable ({e!r}); using the scatter epilogue")
            if obj is None:
                obj = ScatterReduce(group, dev, max_tokens, max_cols, dtype)
        except Exception as e:                                     # noqa: BLE001
            import warnings
            warnings.warn(f"q100.peer: fused row-parallel reduce unavailable, NCCL all-reduce stays ({e!r})")
            obj = None
    _SCATTER[key] = obj
    return obj


def get_scatter(group=None, dtype: torch.dtype = torch.bfloat16) -> Optional[ScatterReduce]:
    return _SCATTER.get((group, dtype))


def enable(group=None, device: Optional[torch.device] = None, slot_elems: int = 16 * 8192) -> Optional[OneShotAllReduce]:
    """Collective.  Sets up the peer buffers for ``group`` (default: world); returns None (and remembers it) when
    symmetric memory is not available, in which case reduce_output keeps using NCCL."""
    key = group
    if key in _STATE:
        return _STATE[key]
    obj = None
    if dist.is_initialized() and dist.get_world_size(group) > 1 and torch.cuda.is_available():
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        try:
            obj = OneShotAllReduce(group, dev, slot_elems)         # agrees across ranks before its own collectives
        except Exception as e:                                     # noqa: BLE001 -- any setup failure => NCCL
            import warnings
            warnings.warn(f"q100.peer: symmetric memory unavailable, small all-reduces stay on NCCL ({e!r})")
            obj = None
    _STATE[key] = obj
    return obj


def get(group=None) -> Optional[OneShotAllReduce]:
    return _STATE.get(group)


def shutdown() -> None:
    """Drop the peer buffers of every group (call before dist.destroy_process_group).  Raises if a call timed out."""
    objs = [o for o in _STATE.values() if o is not None] + [o for o in _SCATTER.values() if o is not None] \
        + [o for o in _REDUCE_NORM.values() if o is not None]
    _STATE.clear()
    _SCATTER.clear()
    _REDUCE_NORM.clear()
    for o in objs:
        o.check()
