"""TEST INFRASTRUCTURE ONLY (parity oracle + timed CPU baseline) -- never imported by the
product path.

CPU restatement of the VQ nearest-codebook quantizer, following
/root/reference/models/vq_vae.py:27-64 (`VectorQuantizer.forward`), :79-93, :110-124 (usage
helpers) and the implied autograd backward.  Pinned against the
reference module itself executed behind oracle/ref_shim.py (tests/golden/vq_golden.npz).

The forward is written with the same float32 torch ops in the same order as the reference so that
on CPU it reproduces the reference's indices bit for bit; `exact_margins` adds a float64 view used
to classify rows whose two best codes are closer than a few float32 ulps of the distance ("fp32
distance ties", BASELINE.json north_star), where no implementation can be expected to agree.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F


def flatten_nchw(z: torch.Tensor) -> torch.Tensor:
    """[B,D,H,W] -> [BHW, D] rows in (b,h,w) order (vq_vae.py:28-31)."""
    return z.permute(0, 2, 3, 1).contiguous().view(-1, z.shape[1])


def distances_fp32(flat: torch.Tensor, E: torch.Tensor) -> torch.Tensor:
    """(sum z^2 [N,1] + sum E^2 [K]) - 2 (z @ E^T): two roundings after the GEMM (vq_vae.py:34-36)."""
    return torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(flat, E.t())


def code_indices(z: torch.Tensor, E: torch.Tensor) -> torch.Tensor:
    """int64 [BHW]; torch.argmin = first minimal index (vq_vae.py:39)."""
    return torch.argmin(distances_fp32(flatten_nchw(z), E), dim=1)


def quantize_forward(z: torch.Tensor, E: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Differentiable forward: (quantized[B,D,H,W], commitment_loss, embedding_loss, idx[BHW])."""
    zl = z.permute(0, 2, 3, 1).contiguous()
    flat = zl.view(-1, z.shape[1])
    idx = torch.argmin(distances_fp32(flat, E), dim=1)
    onehot = torch.zeros(idx.shape[0], E.shape[0], dtype=E.dtype)
    onehot.scatter_(1, idx.unsqueeze(1), 1)
    q = torch.matmul(onehot, E).view(zl.shape)
    commitment = F.mse_loss(q.detach(), zl)
    embedding = F.mse_loss(q, zl.detach())
    out = zl + (q - zl).detach()
    return out.permute(0, 3, 1, 2).contiguous(), commitment, embedding, idx


def quantize_backward(z, E, idx, d_out, g_commit: float, g_embed: float):
    """Closed-form backward in float64 then rounded:
        dz = d_out + g_commit * 2 (z - q) / (N D)
        dE[j] = sum_{n: idx_n = j} g_embed * 2 (q_n - z_n) / (N D)."""
    B, D, H, W = z.shape
    flat = flatten_nchw(z).to(torch.float64)
    q = E.to(torch.float64)[idx]
    n_el = flat.numel()
    dz_flat = g_commit * 2.0 * (flat - q) / n_el
    dz = d_out.to(torch.float64) + dz_flat.view(B, H, W, D).permute(0, 3, 1, 2)
    dE = torch.zeros(E.shape, dtype=torch.float64)
    dE.index_add_(0, idx, g_embed * 2.0 * (q - flat) / n_el)
    return dz.to(torch.float32), dE.to(torch.float32)


def exact_margins(z: torch.Tensor, E: torch.Tensor, chunk: int = 16384):
    """float64 distances: returns (best_idx, margin = d_2nd - d_best, dist_best) per row."""
    flat = flatten_nchw(z).to(torch.float64)
    E64 = E.to(torch.float64)
    e2 = (E64 ** 2).sum(1)
    best, margin, dbest = [], [], []
    for r0 in range(0, flat.shape[0], chunk):
        f = flat[r0:r0 + chunk]
        d = (f ** 2).sum(1, keepdim=True) + e2 - 2 * f @ E64.t()
        top2 = torch.topk(d, 2, dim=1, largest=False)
        best.append(top2.indices[:, 0])
        margin.append(top2.values[:, 1] - top2.values[:, 0])
        dbest.append(top2.values[:, 0])
    return torch.cat(best), torch.cat(margin), torch.cat(dbest)


def tie_rows(z: torch.Tensor, E: torch.Tensor, ulps: float = 8.0) -> torch.Tensor:
    """Boolean [N]: rows whose exact margin is below `ulps` float32 ulps of the distance magnitude
    |z|^2 + |E_j|^2 -- the rounding quantum of the reference's formula (vq_vae.py:34-36 adds the
    row constant |z|^2 BEFORE subtracting, so every distance is quantised to ulp(|z|^2+|E|^2))."""
    flat = flatten_nchw(z).to(torch.float64)
    _, margin, _ = exact_margins(z, E)
    mag = (flat ** 2).sum(1) + (E.to(torch.float64) ** 2).sum(1).max()
    ulp = torch.ldexp(torch.ones_like(mag), torch.floor(torch.log2(mag.clamp(min=1e-300))).to(torch.int32) - 23)
    return margin <= ulps * ulp


def codebook_usage_percentage(idx: torch.Tensor, K: int) -> float:
    """100 |unique(idx)| / K  (vq_vae.py:110-124)."""
    return float(torch.unique(idx).numel() / K * 100.0)
