"""Inference-mode bulk code extraction.

The reference extracts the training set's codes batch by batch (`net.get_code_indices(images)` ->
`.cpu().numpy()`, /root/reference/utils/vq_codes_lmdb.py:58-96; on-the-fly variant main.py:1009-1018) and, in its
evaluation loops, concatenates every batch's int64 indices on the host to run `torch.unique` for the codebook
usage (main.py:261-330).  `CodeExtractor` does the device side of that:

  * K4 only per batch (no gather / losses), then one kernel narrows the indices to int16 / int32 (a quarter / half of
    the int64 D2H bytes) and ORs them into a K-bit usage bitmap that lives on the device across batches;
  * the narrowed codes are copied to pinned host memory on a side stream, `ring` batches deep, so the copy of batch i
    overlaps the search of batch i + 1 and nothing synchronises until `finish()`.

What is done with the codes afterwards (LMDB, pickles) is I/O and stays with the caller.  CUDA-only.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import _lib as L
from .quantizer import code_indices


class CodeExtractor:
    def __init__(self, quantizer, code_dtype: Optional[torch.dtype] = None, ring: int = 3):
        self.vq = quantizer
        K = int(quantizer.K)
        if code_dtype is None:
            code_dtype = torch.int16 if K <= 32768 else torch.int32
        if code_dtype not in (torch.int16, torch.int32, torch.int64):
            raise ValueError("code_dtype must be torch.int16, torch.int32 or torch.int64")
        if code_dtype == torch.int16 and K > 32768:
            raise ValueError(f"{K} codes do not fit int16")
        self.code_dtype = code_dtype
        self.K = K
        self.ring = max(1, int(ring))
        self._bitmap: Optional[Tensor] = None
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._dev_slots: List[Optional[Tensor]] = [None] * self.ring
        self._slot_free: List[Optional[torch.cuda.Event]] = [None] * self.ring
        self._arena: Optional[Tensor] = None         # current pinned block; batches are carved from it back to back
        self._arena_used = 0                         # (cudaHostAlloc per batch costs more than the search itself)
        self._chunks: List[Tensor] = []              # pinned host views, one per batch, in push order
        self._shapes: List[tuple] = []
        self._n = 0

    def push(self, latents: Tensor) -> None:
        """latents [B, D, H, W] (encoder output).  Enqueues search -> narrow + bitmap -> D2H; returns immediately."""
        L.require_cuda(latents, "latents")
        dev = latents.device
        w = self.vq.embedding.weight
        B, _, H, W = latents.shape
        n = B * H * W
        with torch.no_grad():
            idx = code_indices(latents, w, getattr(self.vq, "search_mode", L.VQ_AUTO))
        if self._bitmap is None:
            self._bitmap = torch.zeros((self.K + 31) // 32, dtype=torch.int32, device=dev)
            self._copy_stream = torch.cuda.Stream(device=dev)
        slot = self._n % self.ring
        cur = torch.cuda.current_stream(dev)
        if self._slot_free[slot] is not None:
            cur.wait_event(self._slot_free[slot])        # the copy that last used this device slot is done
        buf = self._dev_slots[slot]
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=self.code_dtype, device=dev)
            self._dev_slots[slot] = buf
        out = buf[:n]
        with torch.cuda.device(dev):
            L.check(L.lib().movae_vq_pack_codes(L.ptr(idx), n, self.K, L.ptr(out), out.element_size(), L.ptr(self._bitmap),
                                                cur.cuda_stream), "vq_pack_codes")
        host = self._host_chunk(n)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            host.copy_(out, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        out.record_stream(self._copy_stream)
        self._slot_free[slot] = done
        self._chunks.append(host)
        self._shapes.append((B, H, W))
        self._n += 1

    def reset(self, keep_usage: bool = False) -> None:
        """Start a new extraction run with the SAME pinned arena, device slots and copy stream (a fresh extractor pays a
        cudaHostAlloc of its arena -- milliseconds -- on its first push).  Call after `finish()`; tensors returned by
        `finish(per_batch=True)` are views of the arena and are overwritten by the next run."""
        if self._copy_stream is not None:
            self._copy_stream.synchronize()
        self._chunks, self._shapes, self._n, self._arena_used = [], [], 0, 0
        if self._bitmap is not None and not keep_usage:
            self._bitmap.zero_()

    def _host_chunk(self, n: int) -> Tensor:
        if self._arena is None or self._arena_used + n > self._arena.numel():
            self._arena = torch.empty(max(16 * n, 1 << 22), dtype=self.code_dtype, pin_memory=True)
            self._arena_used = 0
        view = self._arena[self._arena_used:self._arena_used + n]
        self._arena_used += n
        return view

    def usage_count(self) -> Tensor:
        """int32 device scalar: distinct codes seen over all batches pushed so far."""
        if self._bitmap is None:
            raise RuntimeError("no batch has been pushed yet")
        out = torch.empty(1, dtype=torch.int32, device=self._bitmap.device)
        with torch.cuda.device(self._bitmap.device):
            L.check(L.lib().movae_vq_bitmap_count(L.ptr(self._bitmap), self.K, L.ptr(out),
                                                  torch.cuda.current_stream(self._bitmap.device).cuda_stream), "vq_bitmap_count")
        return out

    def usage_percentage(self) -> float:
        return float(int(self.usage_count().item()) / self.K * 100.0)

    def finish(self, per_batch: bool = False):
        """Waits for the copies.  Returns all codes as one host tensor [total_rows] (or, with per_batch=True, a list of
        [B, H, W] host tensors in push order, the shape `get_code_indices` returns, vq_vae.py:419-421)."""
        if self._copy_stream is not None:
            self._copy_stream.synchronize()
        if per_batch:
            return [c.view(s) for c, s in zip(self._chunks, self._shapes)]
        return torch.cat(self._chunks) if self._chunks else torch.empty(0, dtype=self.code_dtype)
