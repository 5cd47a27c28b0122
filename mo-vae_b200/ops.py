"""Thin tensor-level wrappers over the C-ABI kernels (K1 gram, K2 solves, K3 recombine).
All tensors are caller-owned CUDA tensors; work is enqueued on the current stream, nothing syncs."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib as L

_workspaces: dict = {}


def _gram_workspace(device: torch.device, k: int, stream: int) -> torch.Tensor:
    key = (device.index, k, stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = L.lib().movae_gram_workspace_bytes(k)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)   # zero-filled once: the ticket self-resets
        _workspaces[key] = ws
    return ws


def check_jacobian(J: torch.Tensor) -> Tuple[int, int, int]:
    if J.dim() != 2:
        raise ValueError(f"Parameter `matrix` should be a 2-D tensor. Found `matrix.shape = {tuple(J.shape)}`.")
    L.require_cuda(J, "matrix")
    if J.dtype != torch.float32:
        raise TypeError(f"movae_b200: the Jacobian must be float32 (got {J.dtype})")
    k, P = J.shape
    if k < 1:
        raise ValueError("movae_b200: the Jacobian needs at least one row")
    if k > L.MAX_K:
        raise RuntimeError(f"movae_b200: k={k} objectives > {L.MAX_K} is not supported by this CUDA build")
    if P > 0 and J.stride(1) != 1:
        raise ValueError("movae_b200: Jacobian rows must be contiguous (stride(1) == 1)")
    ld = J.stride(0) if k > 1 else max(P, 1)
    if ld < P:
        raise ValueError("movae_b200: overlapping Jacobian rows")
    return k, P, ld


def gram(J: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """K1: float64 [k,k] Gramian of a float32 [k,P] Jacobian (one streaming pass over J)."""
    k, P, ld = check_jacobian(J)
    if out is None:
        out = torch.empty((k, k), dtype=torch.float64, device=J.device)
        accumulate = False
    stream = L.stream_of(J)
    ws = _gram_workspace(J.device, k, stream)
    with torch.cuda.device(J.device):
        L.check(L.lib().movae_gram_f32(L.ptr(J), k, P, ld, L.ptr(out), int(accumulate), L.ptr(ws), ws.numel(), stream),
                "gram_f32")
    return out


def aggregate(J: torch.Tensor, spec, vec: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, accumulate: bool = False, exchange=None, want_grad: bool = True):
    """The fused step (ONE launch): Gramian of J -> (multi-GPU: k x k exchange over peer memory, `exchange` = a
    parallel.P2PGramianExchange) -> solve `spec` -> out (=|+=) w @ J.  `want_grad=False` stops after the solve (what
    `aggregator.weighting(J)` returns).  Returns (w [k] -- [2k] for COMFORT: the blend, then the MGDA weights --,
    diag [8] float64, G [k,k] float64 rank-summed, out)."""
    k, P, ld = check_jacobian(J)
    dev = J.device
    n_w = 2 * k if spec.kind == L.SOLVE_COMFORT else k
    w = torch.empty(n_w, dtype=torch.float32, device=dev)
    diag = torch.empty(L.DIAG_DOUBLES, dtype=torch.float64, device=dev)
    G = torch.empty((k, k), dtype=torch.float64, device=dev)
    vec = _dev_f32(vec, dev, k, "pref_vector/losses")
    if aux is not None:
        L.require_cuda(aux, "aux")
        if aux.dtype != torch.float32 or not aux.is_contiguous():
            raise ValueError("`aux` must be a contiguous float32 CUDA tensor")
    if want_grad:
        if out is None:
            out = torch.empty(P, dtype=torch.float32, device=dev)
            accumulate = False
        elif out.dtype != torch.float32 or out.shape != (P,) or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"`out` must be a contiguous float32 CUDA tensor of shape ({P},)")
    else:
        out = None
    stream = L.stream_of(J)
    ws = _gram_workspace(dev, k, stream)
    ctx = ctypes.byref(exchange.ctx) if exchange is not None else None
    with torch.cuda.device(dev):
        L.check(L.lib().movae_aggregate_f32(L.ptr(J), k, P, ld, ctypes.byref(spec), L.ptr(vec), L.ptr(aux), L.ptr(out),
                                            int(accumulate), L.ptr(w), L.ptr(diag), L.ptr(G), L.ptr(ws), ws.numel(), ctx, stream),
                "aggregate_f32")
    return w, diag, G, out


def aggregate_segments(rows, numels, out_offsets, spec, vec: Optional[torch.Tensor], aux: Optional[torch.Tensor],
                       out: Optional[torch.Tensor], accumulate: bool = False, exchange=None):
    """The fused step over a SEGMENTED Jacobian (no flat J): `rows[i][s]` is the float32 CUDA tensor holding objective
    i's gradient of segment s (contiguous, `numels[s]` elements, 16-byte aligned -- `segments_ok` checks), `out_offsets[s]`
    the segment's offset (multiple of 4) in the flat gradient buffer `out`.  `out=None`: weights only.
    Returns (w, diag, G) like `aggregate`."""
    k, n_seg = len(rows), len(numels)
    if not (1 <= k <= L.MAX_K):
        raise RuntimeError(f"movae_b200: k={k} objectives outside 1..{L.MAX_K} is not supported by this CUDA build")
    if not (1 <= n_seg <= L.MAX_SEGMENTS):
        raise RuntimeError(f"movae_b200: {n_seg} Jacobian segments outside 1..{L.MAX_SEGMENTS}")
    dev = rows[0][0].device
    segs = L.JacSegments(n_segments=n_seg, k=k)
    for s in range(n_seg):
        segs.n[s] = int(numels[s])
        segs.out_off[s] = int(out_offsets[s])
        for i in range(k):
            segs.rows[s][i] = rows[i][s].data_ptr()
    n_w = 2 * k if spec.kind == L.SOLVE_COMFORT else k
    w = torch.empty(n_w, dtype=torch.float32, device=dev)
    diag = torch.empty(L.DIAG_DOUBLES, dtype=torch.float64, device=dev)
    G = torch.empty((k, k), dtype=torch.float64, device=dev)
    vec = _dev_f32(vec, dev, k, "pref_vector/losses")
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _gram_workspace(dev, k, stream)
    ctx = ctypes.byref(exchange.ctx) if exchange is not None else None
    with torch.cuda.device(dev):
        L.check(L.lib().movae_aggregate_segments_f32(ctypes.byref(segs), ctypes.byref(spec), L.ptr(vec), L.ptr(aux), L.ptr(out),
                                                     int(accumulate), L.ptr(w), L.ptr(diag), L.ptr(G), L.ptr(ws), ws.numel(), ctx,
                                                     stream), "aggregate_segments_f32")
    return w, diag, G


def segment_ok(t: torch.Tensor) -> bool:
    """A gradient tensor the segmented kernels can read in place."""
    return t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0


def current_workspace(device: torch.device, k: int) -> torch.Tensor:
    """The (cached) workspace the K1 / fused launches of the CURRENT stream of `device` use for this k."""
    return _gram_workspace(device, k, torch.cuda.current_stream(device).cuda_stream)


def aggregate_phase_times(ws: torch.Tensor) -> Tuple[float, float, float]:
    """(Gramian pass, combine + exchange + solve, recombine pass) in ms of the LAST fused launch that used the workspace
    `ws` (see current_workspace; under a CUDA graph: the workspace of the capturing stream), from the kernel's own
    globaltimer stamps.  Synchronises the current stream."""
    t0, t1, t2, t3, _, _ = aggregate_stamps(ws)
    return (t1 - t0) * 1e-6, (t2 - t1) * 1e-6, (t3 - t2) * 1e-6


def aggregate_stamps(ws: torch.Tensor):
    """The raw globaltimer stamps (ns) of the last fused launch on `ws`: start, all partials in, weights published, end,
    partials combined, exchange done."""
    stream = torch.cuda.current_stream(ws.device).cuda_stream
    stamps = (ctypes.c_uint64 * 6)()
    with torch.cuda.device(ws.device):
        L.check(L.lib().movae_aggregate_timestamps(L.ptr(ws), ctypes.byref(stamps), stream), "aggregate_timestamps")
    return tuple(int(x) for x in stamps)


def solve(G: torch.Tensor, spec, vec: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None):
    """K2 through the generic entry (any movae_solve_spec, incl. COMFORT / the PNUPGrad draw flag): (w, diag)."""
    k, w, diag, G = _solve_outputs(G)
    if spec.kind == L.SOLVE_COMFORT:
        w = torch.empty(2 * k, dtype=torch.float32, device=G.device)
    vec = _dev_f32(vec, G.device, k, "pref_vector/losses")
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_aux(L.ptr(G), k, ctypes.byref(spec), L.ptr(vec), L.ptr(aux), L.ptr(w), L.ptr(diag),
                                        L.stream_of(G)), "solve")
    return w, diag


def _solve_outputs(G: torch.Tensor):
    if G.dim() != 2 or G.shape[0] != G.shape[1]:
        raise ValueError(f"gramian must be square, got {tuple(G.shape)}")
    L.require_cuda(G, "gramian")
    if G.dtype != torch.float64 or not G.is_contiguous():
        G = G.to(torch.float64).contiguous()
    k = G.shape[0]
    w = torch.empty(k, dtype=torch.float32, device=G.device)
    diag = torch.empty(L.DIAG_DOUBLES, dtype=torch.float64, device=G.device)
    return k, w, diag, G


def _dev_f32(t: Optional[torch.Tensor], device, k: int, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dim() != 1 or t.shape[0] != k:
        raise ValueError(f"`{name}` must have shape ({k},), got {tuple(t.shape)}")
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def solve_constant(G: torch.Tensor, value: float):
    k, w, diag, G = _solve_outputs(G)
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_constant(L.ptr(G), k, float(value), L.ptr(w), L.ptr(diag), L.stream_of(G)),
                "solve_constant")
    return w, diag


def solve_upgrad(G: torch.Tensor, pref: Optional[torch.Tensor], norm_eps: float, reg_eps: float, norm_mode: str = "trace"):
    """UPGrad ("trace"), NUPGrad ("min_l2") or PNUPGrad's other branch ("l2"): k dual-cone QPs on the normalised Gramian."""
    k, w, diag, G = _solve_outputs(G)
    pref = _dev_f32(pref, G.device, k, "pref_vector")
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_nupgrad(L.ptr(G), k, L.ptr(pref), float(norm_eps), float(reg_eps),
                                            L.UPGRAD_NORM[norm_mode], L.ptr(w), L.ptr(diag), L.stream_of(G)), "solve_upgrad")
    return w, diag


def solve_dualproj(G: torch.Tensor, pref: Optional[torch.Tensor], norm_eps: float, reg_eps: float):
    """torchjd DualProj: one dual-cone QP with the whole preference vector as lower bound."""
    k, w, diag, G = _solve_outputs(G)
    pref = _dev_f32(pref, G.device, k, "pref_vector")
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_dualproj(L.ptr(G), k, L.ptr(pref), float(norm_eps), float(reg_eps), L.ptr(w), L.ptr(diag),
                                             L.stream_of(G)), "solve_dualproj")
    return w, diag


def solve_mgda(G: torch.Tensor, norm_type: str, losses: Optional[torch.Tensor], epsilon: float, max_iters: int,
               stable: bool, min_eigenvalue_eps: float):
    k, w, diag, G = _solve_outputs(G)
    losses = _dev_f32(losses, G.device, k, "losses")
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_mgda(L.ptr(G), k, L.MGDA_NORM[norm_type], L.ptr(losses), float(epsilon),
                                         int(max_iters), int(bool(stable)), float(min_eigenvalue_eps), L.ptr(w),
                                         L.ptr(diag), L.stream_of(G)), "solve_mgda")
    return w, diag


def solve_aligned_mtl(G: torch.Tensor, scale_mode: str, pref: Optional[torch.Tensor]):
    k, w, diag, G = _solve_outputs(G)
    pref = _dev_f32(pref, G.device, k, "pref_vector")
    with torch.cuda.device(G.device):
        L.check(L.lib().movae_solve_aligned_mtl(L.ptr(G), k, L.AMTL_SCALE[scale_mode], L.ptr(pref), L.ptr(w),
                                                L.ptr(diag), L.stream_of(G)), "solve_aligned_mtl")
    return w, diag


def recombine(J: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None,
              accumulate: bool = False) -> torch.Tensor:
    """K3: out (=|+=) w @ J, one streaming pass; `out` is the flat float32 buffer .grad views live in."""
    k, P, ld = check_jacobian(J)
    L.require_cuda(w, "weights")
    if w.dtype != torch.float32 or w.shape != (k,) or not w.is_contiguous():
        raise ValueError(f"weights must be a contiguous float32 tensor of shape ({k},)")
    if out is None:
        out = torch.empty(P, dtype=torch.float32, device=J.device)
        accumulate = False
    else:
        if out.dtype != torch.float32 or out.shape != (P,) or not out.is_contiguous() or out.device != J.device:
            raise ValueError(f"`out` must be a contiguous float32 CUDA tensor of shape ({P},)")
    with torch.cuda.device(J.device):
        L.check(L.lib().movae_recombine_f32(L.ptr(J), k, P, ld, L.ptr(w), L.ptr(out), int(accumulate),
                                            L.stream_of(J)), "recombine_f32")
    return out
