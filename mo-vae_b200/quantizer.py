"""Drop-in for the reference's `VectorQuantizer` (/root/reference/models/vq_vae.py:11-124) over the
CUDA kernels K4 (tcgen05 nearest-codebook search), K5 (gather + losses + straight-through output +
usage bitmap) and K6 (backward).

Same constructor, attributes (`K`, `D`, `_summary_mode`, `embedding` = nn.Embedding(K, D) initialised
U(-1/K, 1/K), parameter name `embedding.weight`), same `forward` return tuple

    (quantized[B,D,H,W], commitment_loss, embedding_loss, encoding_inds[BHW] int64)

and the same helper methods, so the reference's model shells (vq_vae.py:329, vq_vae2.py:225,231) and
state dicts work unchanged.  CUDA-only: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L

_workspaces: dict = {}


def _workspace(device: torch.device, n_rows: int, K: int, D: int, stream: int) -> Tensor:
    """One zero-initialised workspace per (device, stream); grown when a bigger batch shows up."""
    need = L.lib().movae_vq_workspace_bytes(n_rows, K, D)
    if need == 0:
        raise RuntimeError(f"movae_b200: num_embeddings={K} is not supported by this CUDA build")
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_scratches: dict = {}


def _scratch(device: torch.device, nbytes: int, stream: int) -> Tensor:
    """Uninitialised per-(device, stream) scratch for K6's per-CTA codebook-gradient partials."""
    key = (device.index, stream)
    buf = _scratches.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _scratches[key] = buf
    return buf


def _check_inputs(latents: Tensor, weight: Tensor) -> Tuple[int, int, int, int]:
    L.require_cuda(latents, "latents")
    L.require_cuda(weight, "embedding.weight")
    if latents.dim() != 4:
        raise ValueError(f"latents must be [B, D, H, W], got shape {tuple(latents.shape)}")
    if latents.dtype != torch.float32 or weight.dtype != torch.float32:
        raise TypeError("movae_b200: the quantizer computes in float32 (got "
                        f"{latents.dtype} latents / {weight.dtype} codebook)")
    B, D, H, W = latents.shape
    K, De = weight.shape
    if De != D:
        raise RuntimeError(f"latents have {D} channels but the codebook has embedding_dim={De}")
    if latents.device != weight.device:
        raise RuntimeError("latents and the codebook must be on the same CUDA device")
    return B, D, H * W, K


def code_indices(latents: Tensor, weight: Tensor, mode: int = L.VQ_AUTO, debug_scores: Optional[Tensor] = None) -> Tensor:
    """K4 only: int64 [B*H*W] nearest-code indices (vq_vae.py:28-39; the inference-mode callers
    `get_code_indices` vq_vae.py:393-423 / vq_vae2.py:290-311 use exactly this)."""
    B, D, HW, K = _check_inputs(latents, weight)
    z = latents.detach().contiguous()
    E = weight.detach().contiguous()
    idx = torch.empty(B * HW, dtype=torch.int64, device=z.device)
    stream = L.stream_of(z)
    ws = _workspace(z.device, B * HW, K, D, stream)
    with torch.cuda.device(z.device):
        L.check(L.lib().movae_vq_argmin_f32(L.ptr(z), B, D, HW, L.ptr(E), K, L.ptr(idx), int(mode), L.ptr(debug_scores),
                                            L.ptr(ws), ws.numel(), stream), "vq_argmin_f32")
    return idx


def rechecked_rows(device: torch.device) -> int:
    """Rows the last tensor-path search on the current stream re-evaluated exactly (synchronises)."""
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    return 0 if ws is None else int(ws[:4].view(torch.int32).item())


def codebook_usage_count(encoding_inds: Tensor, K: int) -> Tensor:
    """int32 device scalar: |unique(encoding_inds)| via a K-bit bitmap (replaces torch.unique, vq_vae.py:121)."""
    L.require_cuda(encoding_inds, "encoding_inds")
    idx = encoding_inds.detach().reshape(-1).to(torch.int64).contiguous()
    out = torch.zeros(1, dtype=torch.int32, device=idx.device)
    stream = L.stream_of(idx)
    ws = _workspace(idx.device, 0, K, 1, stream)
    with torch.cuda.device(idx.device):
        L.check(L.lib().movae_vq_usage(L.ptr(idx), idx.numel(), K, L.ptr(out), L.ptr(ws), ws.numel(), stream), "vq_usage")
    return out


class _Quantize(torch.autograd.Function):
    """forward: K4 + K5; backward: K6.  Gradients (vq_vae.py:51-55): the straight-through output
    passes its gradient to the latents unchanged, commitment_loss adds 2 (z - q) / (N D) to them,
    embedding_loss sends 2 (q - z) / (N D) to the selected codebook rows; nothing flows through argmin."""

    @staticmethod
    def forward(ctx, latents: Tensor, weight: Tensor, mode: int):
        B, D, HW, K = _check_inputs(latents, weight)
        z = latents.detach().contiguous()
        E = weight.detach().contiguous()
        N = B * HW
        if N == 0:
            raise RuntimeError("movae_b200: empty latents")
        idx = torch.empty(N, dtype=torch.int64, device=z.device)
        quantized = torch.empty_like(z)
        losses = torch.empty(2, dtype=torch.float32, device=z.device)
        usage = torch.empty(1, dtype=torch.int32, device=z.device)
        stream = L.stream_of(z)
        ws = _workspace(z.device, N, K, D, stream)
        with torch.cuda.device(z.device):
            L.check(L.lib().movae_vq_forward_f32(L.ptr(z), B, D, HW, L.ptr(E), K, L.ptr(idx), L.ptr(quantized), L.ptr(losses),
                                                 L.ptr(usage), int(mode), L.ptr(ws), ws.numel(), stream), "vq_forward_f32")
        ctx.save_for_backward(z, E, idx)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(idx, usage)
        return quantized, losses[0], losses[1], idx, usage

    @staticmethod
    def backward(ctx, g_q, g_commit, g_embed, _g_idx, _g_usage):
        z, E, idx = ctx.saved_tensors
        B, D, H, W = z.shape
        K = E.shape[0]
        need_z, need_E = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dz = dE = None
        if need_z and not (g_q is None and g_commit is None):
            dz = torch.empty_like(z)      # else None: nothing reaches the latents (embedding_loss alone, vq_vae.py:52), and
                                          # autograd reports the latents as unused instead of back-propagating zeros
        run_E = need_E and g_embed is not None          # only embedding_loss reaches the codebook (vq_vae.py:51-55)
        if run_E:
            dE = torch.zeros_like(E)
        if dz is not None or run_E:
            f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()  # noqa: E731
            g_q, g_commit, g_embed = f32(g_q), f32(g_commit), f32(g_embed)
            scratch = None
            need = L.lib().movae_vq_backward_workspace_bytes(B * H * W, K, D) if run_E else 0
            if need:
                scratch = _scratch(z.device, need, L.stream_of(z))
            with torch.cuda.device(z.device):
                L.check(L.lib().movae_vq_backward_f32(L.ptr(g_q), L.ptr(g_commit), L.ptr(g_embed), L.ptr(z), B, D, H * W, L.ptr(E),
                                                      K, L.ptr(idx), L.ptr(dz), L.ptr(dE) if run_E else 0,
                                                      L.ptr(scratch), need, L.stream_of(z)), "vq_backward_f32")
        return dz, dE, None


class VectorQuantizer(nn.Module):
    """Vector Quantization module for VQ-VAE (same interface as the reference's, vq_vae.py:11-124)."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.K = num_embeddings
        self.D = embedding_dim
        self._summary_mode = False
        self.embedding = nn.Embedding(self.K, self.D)
        self.embedding.weight.data.uniform_(-1 / self.K, 1 / self.K)
        self.search_mode = L.VQ_AUTO            # MOVAE_VQ_AUTO | _EXACT | _TENSOR (include/movae_b200.h)
        self.last_usage_count: Optional[Tensor] = None   # int32 device scalar written by K5's bitmap

    def forward(self, latents: Tensor):
        quantized, commitment_loss, embedding_loss, encoding_inds, usage = _Quantize.apply(
            latents, self.embedding.weight, self.search_mode)
        self.last_usage_count = usage
        if getattr(self, "_summary_mode", False):
            return quantized
        return quantized, commitment_loss, embedding_loss, encoding_inds

    def embed_code(self, code: Tensor) -> Tensor:
        return self.embedding(code)

    def get_code_indices(self, latents: Tensor) -> Tensor:
        """int64 [B*H*W]; K4 only, no gather / losses (inference-mode code extraction)."""
        return code_indices(latents, self.embedding.weight, self.search_mode)

    def get_used_embeddings(self, latents: Tensor) -> Tensor:
        return torch.unique(self.get_code_indices(latents))

    def get_codebook_usage_percentage(self, latents: Tensor) -> float:
        return self.get_codebook_usage_percentage_from_indices(self.get_code_indices(latents))

    def get_codebook_usage_percentage_from_indices(self, encoding_inds: Tensor) -> float:
        num_used = int(codebook_usage_count(encoding_inds, self.K).item())      # python float by contract -> one D2H
        return float((num_used / self.K) * 100.0)

    def last_codebook_usage_percentage(self) -> float:
        """Usage of the most recent forward(), from the bitmap K5 already built (no extra kernel)."""
        if self.last_usage_count is None:
            raise RuntimeError("forward() has not run yet")
        return float(int(self.last_usage_count.item()) / self.K * 100.0)
