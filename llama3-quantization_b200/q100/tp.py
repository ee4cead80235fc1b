"""Tensor-parallel sharding of the quantised linears.

Megatron layout per decoder block: q/k/v/gate/up are column-parallel (output rows N split
across ranks), o_proj/down_proj are row-parallel (input features K split), and the only
collective on the data path is one all-reduce(sum) of the [M, hidden] output of each
row-parallel linear.  Quantisation groups are 128 consecutive K elements of one output
row, so any N split -- and any K split on a multiple of the group -- leaves every code,
scale and zero-point bit-identical to the unsharded tensor; each rank simply quantises
its own slice.

The reference has no counterpart (it only places whole layers on different GPUs,
parallel_utils.py:89-163); parity target is the single-GPU result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

COLUMN = "column"
ROW = "row"

# LLaMA decoder block: which linear is sharded how
LLAMA_PLAN = {"q_proj": COLUMN, "k_proj": COLUMN, "v_proj": COLUMN, "gate_proj": COLUMN, "up_proj": COLUMN,
              "o_proj": ROW, "down_proj": ROW}


@dataclass(frozen=True)
class Shard:
    """Half-open slices of the full [N, K] weight owned by one rank."""
    n0: int
    n1: int
    k0: int
    k1: int

    @property
    def shape(self) -> Tuple[int, int]:
        return self.n1 - self.n0, self.k1 - self.k0


def shard_bounds(N: int, K: int, mode: str, rank: int, world: int, group: int = 128,
                 n_align: int = 8) -> Shard:
    """Slice of an [N, K] linear for ``rank``.  Row-parallel cuts fall on multiples of the
    quantisation group (so groups never straddle ranks), column-parallel cuts on ``n_align``."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    if mode == COLUMN:
        if N % (world * n_align) != 0:
            raise ValueError(f"N={N} cannot be split over {world} ranks in multiples of {n_align}")
        per = N // world
        return Shard(rank * per, (rank + 1) * per, 0, K)
    if mode == ROW:
        g = K if group in (0, None) else group
        if g == K and world > 1:
            # per-channel (one group per output row): the row's min / max spans all ranks, so
            # the caller quantises its slice with all-reduced statistics (UniformAffineQuantizer.stat_group ->
            # allreduce_minmax below); the cut only has to respect the packed layout's 128-k blocks
            g = 128
        if K % (world * g) != 0:
            raise ValueError(f"K={K} cannot be split over {world} ranks in multiples of the group {g}")
        per = K // world
        return Shard(0, N, rank * per, (rank + 1) * per)
    raise ValueError(f"unknown shard mode {mode!r}")


def shard_weight(w: torch.Tensor, mode: str, rank: int, world: int, group: int = 128) -> torch.Tensor:
    s = shard_bounds(w.shape[0], w.shape[1], mode, rank, world, group)
    return w[s.n0:s.n1, s.k0:s.k1].contiguous()


def shard_input(x: torch.Tensor, mode: str, rank: int, world: int) -> torch.Tensor:
    """Row-parallel layers consume the K-slice of the activation produced by the preceding
    column-parallel layer; column-parallel layers see the full activation."""
    if mode == COLUMN:
        return x
    per = x.shape[-1] // world
    return x[..., rank * per:(rank + 1) * per].contiguous()


def reduce_output(y: torch.Tensor, mode: str, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """The one collective of the path: all-reduce(sum) after a row-parallel linear (NCCL over
    NVLink on GPU ranks; gloo in the CPU tests).  Column-parallel outputs stay sharded."""
    if mode == ROW and dist.is_initialized() and dist.get_world_size(group) > 1:
        if y.is_cuda:
            from . import peer                                   # K7: one-shot all-reduce over peer memory (decode sizes)
            one = peer.get(group)
            if one is not None and one.fits(y):
                return one(y)
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


def gather_columns(y: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all-gather of a column-parallel output along the feature dim (only needed when the
    consumer is not row-parallel, e.g. a vocab-sharded lm_head)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return y
    parts = [torch.empty_like(y) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, y.contiguous(), group=group)
    return torch.cat(parts, dim=-1)


def allreduce_minmax(mm: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """[rows, 2] (min, max) pairs of a row whose elements are spread over the ranks of ``group`` -> the pairs of the
    whole row, on every rank, with ONE collective: MAX over (-min, max) (negation is exact).  This is the extra
    all-reduce SURVEY 8e asks for when abits < 16 (inputs of o_proj / down_proj) and for per-channel weights of
    row-parallel layers; with it scales and zero-points are bit-identical to the single-GPU run
    (quantizer.py:140-159 sees the same amin / amax)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return mm
    t = torch.stack((-mm[:, 0], mm[:, 1]), dim=1).contiguous()
    work = t.float() if (not t.is_cuda and t.dtype != torch.float32) else t     # gloo has no 16-bit MAX on some builds
    dist.all_reduce(work, op=dist.ReduceOp.MAX, group=group)
    t = work.to(mm.dtype)
    return torch.stack((-t[:, 0], t[:, 1]), dim=1).contiguous()


def rowwise_minmax_allreduce(x_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Per-token min/max of an activation whose feature dim is sharded across ranks
    (inputs of row-parallel layers when abits < 16): a [rows, 1] MIN and MAX all-reduce
    keeps the activation scale / zero-point bit-identical to the single-GPU run."""
    mn = x_local.amin(dim=-1, keepdim=True)
    mx = x_local.amax(dim=-1, keepdim=True)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return mn, mx
