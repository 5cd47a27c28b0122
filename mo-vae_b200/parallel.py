"""P-sharded aggregation across GPUs (one process per GPU, torch.distributed).

The Gramian is additive over column blocks of J (G = sum_s J_s J_s^T), the solve is k x k and the
recombination is column-local, so the whole multi-GPU path needs exactly ONE collective per step: an
all_reduce(sum) of the k x k float64 Gramian between K1 and K2.  Every rank then runs
K2 on bit-identical input and obtains identical weights.  The reference has no multi-device code;
this module is the host-side logic, and it is backend-agnostic so that it is testable with gloo.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from .aggregation import Aggregator


def shard_columns(P: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """[lo, hi) column block of rank `rank`: contiguous, boundaries multiples of `align` columns
    (4 float32 = 16 bytes, so every shard keeps the float4 kernels), remainder on the last rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world size {world}")
    if P < 0 or align < 1:
        raise ValueError("P must be >= 0 and align >= 1")
    per = (P // world) // align * align
    lo = rank * per
    hi = P if rank == world - 1 else lo + per
    return lo, hi


def all_shards(P: int, world: int, align: int = 4) -> List[Tuple[int, int]]:
    return [shard_columns(P, r, world, align) for r in range(world)]


def gramian_allreduce(group: Optional[dist.ProcessGroup] = None) -> Callable[[torch.Tensor], None]:
    """In-place sum of the float64 [k, k] Gramian over the process group (<= 512 bytes: latency-bound)."""
    def reduce_(G: torch.Tensor) -> None:
        if G.dtype != torch.float64:
            raise TypeError("the Gramian exchanged between ranks must be float64")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return reduce_


def install_gramian_allreduce(aggregator: Aggregator, group: Optional[dist.ProcessGroup] = None) -> Aggregator:
    """Makes `aggregator(J_local)` aggregate the GLOBAL Jacobian whose column block this rank holds."""
    aggregator.weighting.gramian_reducer = gramian_allreduce(group)
    return aggregator


def check_replicated(t: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> bool:
    """True when `t` (e.g. the weights, or MGDA's losses) is bit-identical on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))
