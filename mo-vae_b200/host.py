"""Aggregation of a Jacobian that lives in HOST memory (the `e2e` leg of bench.py): the C-ABI
host pipeline streams J to the GPU in column chunks overlapped with K1, solves on the device, and
streams K3's output back.  A P-sharded caller passes `gramian_reducer` (one k x k allreduce)."""
from __future__ import annotations

import ctypes
from typing import Callable, Optional

import torch

from . import _lib as L
from . import ops
from .aggregation import Aggregator


class HostAggregationPlan:
    """Owns the device staging buffers for one (k, P) so repeated steps do not allocate."""

    def __init__(self, k: int, P: int, device: torch.device, chunk_cols: int = 1 << 22):
        if not (1 <= k <= L.MAX_K):
            raise RuntimeError(f"movae_b200: k={k} objectives outside 1..{L.MAX_K} is not supported by this CUDA build")
        self.k, self.P, self.device = k, P, torch.device(device)
        self.ld = (P + 3) // 4 * 4
        self.chunk_cols = max(4, chunk_cols // 4 * 4)
        self.d_J = torch.empty((k, self.ld), dtype=torch.float32, device=device)
        self.d_grad = torch.empty(self.ld, dtype=torch.float32, device=device)
        self.d_G = torch.zeros((k, k), dtype=torch.float64, device=device)
        self.d_w = torch.empty(2 * k, dtype=torch.float32, device=device)      # COMFORT reports two weight vectors
        self.d_diag = torch.empty(L.DIAG_DOUBLES, dtype=torch.float64, device=device)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.kernel_launches = 0

    def run(self, h_J: torch.Tensor, aggregator: Aggregator, h_out: torch.Tensor,
            gramian_reducer: Optional[Callable[[torch.Tensor], None]] = None) -> torch.Tensor:
        k, P = self.k, self.P
        if h_J.is_cuda or h_out.is_cuda:
            raise ValueError("aggregate_host expects HOST tensors")
        if h_J.shape != (k, P) or h_J.dtype != torch.float32 or h_J.stride(1) != 1:
            raise ValueError(f"h_J must be float32 [{k},{P}] with contiguous rows")
        if h_out.shape != (P,) or h_out.dtype != torch.float32 or not h_out.is_contiguous():
            raise ValueError(f"h_out must be contiguous float32 [{P}]")
        lib = L.lib()
        with torch.cuda.device(self.device):
            cs = torch.cuda.current_stream(self.device)
            ws = ops._gram_workspace(self.device, k, cs.cuda_stream)
            L.check(lib.movae_host_gram_f32(h_J.data_ptr(), k, P, h_J.stride(0) if k > 1 else max(P, 1),
                                            self.d_J.data_ptr(), self.ld, self.d_G.data_ptr(), ws.data_ptr(), ws.numel(),
                                            self.chunk_cols, cs.cuda_stream, self.copy_stream.cuda_stream), "host_gram_f32")
            if gramian_reducer is not None:
                gramian_reducer(self.d_G)
            if hasattr(aggregator, "_ensure_coef"):
                aggregator._ensure_coef(self.device)          # COMFORT: the blend coefficients live on the device
            aggregator.weighting.prepare_step(self.device)
            spec, vec, aux = aggregator.weighting.solve_spec(k)
            vec = ops._dev_f32(vec, self.device, k, "pref_vector/losses")
            L.check(lib.movae_solve_aux(self.d_G.data_ptr(), k, ctypes.byref(spec), L.ptr(vec), L.ptr(aux), self.d_w.data_ptr(),
                                        self.d_diag.data_ptr(), cs.cuda_stream), "solve")
            L.check(lib.movae_host_recombine_f32(self.d_J.data_ptr(), k, P, self.ld, self.d_w.data_ptr(),
                                                 self.d_grad.data_ptr(), h_out.data_ptr(), self.chunk_cols,
                                                 cs.cuda_stream, self.copy_stream.cuda_stream), "host_recombine_f32")
        aggregator.weighting.last_gramian, aggregator.weighting.last_diag = self.d_G, self.d_diag
        n_chunks = (P + self.chunk_cols - 1) // self.chunk_cols
        self.kernel_launches = 2 * n_chunks + 1
        return h_out


def aggregate_host(h_J: torch.Tensor, aggregator: Aggregator, h_out: Optional[torch.Tensor] = None,
                   device: str | torch.device = "cuda", plan: Optional[HostAggregationPlan] = None) -> torch.Tensor:
    """One-shot convenience wrapper: host J[k,P] -> host g[P] (synchronous)."""
    k, P = h_J.shape
    plan = plan or HostAggregationPlan(k, P, torch.device(device))
    if h_out is None:
        h_out = torch.empty(P, dtype=torch.float32, pin_memory=True)
    return plan.run(h_J, aggregator, h_out)
