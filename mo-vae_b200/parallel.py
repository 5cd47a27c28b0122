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

from . import _lib as L
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


class P2PGramianExchange:
    """k x k Gramian exchange over NVLink peer memory, fused into K1's tail and K2's head
    (include/movae_b200.h "P-sharded aggregation"): no collective launch on the per-step path.
    Construction is collective (all ranks of `group`, one GPU each, same node): every rank allocates an
    exchange buffer through the C ABI, the 64-byte CUDA IPC handles are all-gathered, peers are mapped."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("P2PGramianExchange needs an initialised torch.distributed process group")
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > L.MAX_WORLD:
            raise RuntimeError(f"movae_b200: world size {self.world} > {L.MAX_WORLD} is not supported by this CUDA build")
        lib = L.lib()
        import ctypes
        self._own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            L.check(lib.movae_p2p_alloc(lib.movae_p2p_exchange_bytes(), ctypes.byref(self._own), handle), "p2p_alloc")
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        self.ctx = L.P2PCtx(rank=self.rank, world=self.world)
        self._opened = []
        for r, t in enumerate(gathered):
            if r == self.rank:
                self.ctx.peers[r] = self._own.value
                continue
            peer = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                L.check(lib.movae_p2p_open(bytes(t.cpu().tolist()), ctypes.byref(peer)), "p2p_open")
            self.ctx.peers[r] = peer.value
            self._opened.append(peer)
        self._seq = 0
        dist.barrier(group=group)

    def next_seq(self) -> int:
        self._seq += 1
        return self._seq

    def close(self) -> None:
        lib = L.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.movae_p2p_close(p)
            self._opened = []
            if self._own:
                lib.movae_p2p_free(self._own)
                self._own = None


def install_p2p_gramian_exchange(aggregator: Aggregator, device: torch.device,
                                 group: Optional[dist.ProcessGroup] = None) -> P2PGramianExchange:
    """Like install_gramian_allreduce, but the exchange is fused into the kernels (CUDA only)."""
    ex = P2PGramianExchange(device, group)
    aggregator.weighting.p2p_exchange = ex
    return ex
