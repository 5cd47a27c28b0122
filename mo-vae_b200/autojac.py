"""Jacobian-descent engine entry points with the signatures the reference calls
(/root/reference/main.py:17, :189-196):

    backward(tensors, aggregator=..., inputs=None, retain_graph=False)
    mtl_backward(losses=, features=, aggregator=, tasks_params=None, shared_params=None, retain_graph=False)

They replace torchjd.autojac.{backward,mtl_backward} (un-vendored dependency, requirements.txt:58;
semantics restated in SURVEY.md App. A).  The k per-objective backward passes are plain torch
autograd (not ours); what changes is everything after them:

  * the rows are written into ONE flat float32 buffer J[k, ldJ] (ldJ = P rounded up to 4 so every row
    is 16-byte aligned for the float4 kernels) by a single multi-tensor copy -- no per-parameter
    reshape + `torch.cat`;
  * K1 -> K2 -> K3 run back to back on the current stream;
  * K3 writes the aggregated gradient into one flat buffer and the parameters' `.grad` become views
    of it (assign if `.grad is None`, `+=` otherwise: torchjd `Accumulate` semantics).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
from torch import Tensor

from .aggregation import Aggregator

_J_CACHE: dict = {}

# Jacobian rows by ONE batched (vmapped) backward pass over the k objectives, like torchjd's default
# (parallel_chunk_size=None); falls back to k sequential passes when an op in the graph has no batching rule.
BATCHED_JACOBIAN = True


def _leaves_of(roots: Sequence[Tensor], stop_at: Sequence[Tensor] = ()) -> List[Tensor]:
    """Leaf tensors requiring grad in the autograd graph of `roots`, in discovery order (DFS over
    grad_fn.next_functions).  Traversal does not descend below the grad_fns of `stop_at`."""
    stop = {t.grad_fn for t in stop_at if t.grad_fn is not None}
    seen, out, out_ids = set(), [], set()
    stack = [t.grad_fn for t in reversed(list(roots)) if t.grad_fn is not None]
    for t in roots:
        if t.grad_fn is None and t.requires_grad and id(t) not in out_ids:
            out.append(t)
            out_ids.add(id(t))
    while stack:
        fn = stack.pop()
        if fn is None or fn in seen:
            continue
        seen.add(fn)
        if hasattr(fn, "variable"):                       # AccumulateGrad node -> leaf
            v = fn.variable
            if v.requires_grad and id(v) not in out_ids:
                out.append(v)
                out_ids.add(id(v))
            continue
        if fn in stop:
            continue
        for nxt, _ in reversed(fn.next_functions):
            if nxt is not None and nxt not in seen:
                stack.append(nxt)
    return out


def _jacobian_buffer(k: int, P: int, device: torch.device) -> Tensor:
    ld = (P + 3) // 4 * 4
    key = (device, k, ld)
    buf = _J_CACHE.get(key)
    if buf is None:
        _J_CACHE.clear()                                  # one live model per process is the norm
        buf = torch.zeros((k, ld), dtype=torch.float32, device=device)
        _J_CACHE[key] = buf
    return buf[:, :P] if ld != P else buf


def _fill_row(J: Tensor, i: int, params: Sequence[Tensor], grads: Sequence[Optional[Tensor]]) -> None:
    dst, src = [], []
    off = 0
    for p, g in zip(params, grads):
        n = p.numel()
        view = J[i, off:off + n]
        if g is None:
            view.zero_()
        else:
            dst.append(view)
            src.append(g.reshape(-1))
        off += n
    if dst:
        torch._foreach_copy_(dst, src)


def _fill_rows_batched(J: Tensor, params: Sequence[Tensor], grads: Sequence[Optional[Tensor]]) -> None:
    """grads[p]: [k, *p.shape] (or None) -> columns off..off+numel of all k rows of J, one strided copy each."""
    k = J.shape[0]
    dst, src = [], []
    off = 0
    for p, g in zip(params, grads):
        n = p.numel()
        view = J[:, off:off + n]
        if g is None:
            view.zero_()
        else:
            dst.append(view)
            src.append(g.reshape(k, n))
        off += n
    if dst:
        torch._foreach_copy_(dst, src)


def _jacobian_rows(J: Tensor, outputs: Sequence[Tensor], params: Sequence[Tensor], grad_outputs_per_row: Sequence[Sequence[Tensor]],
                   retain_graph: bool) -> None:
    """Fills J[i] = d(sum_j <outputs[j], grad_outputs_per_row[i][j]>) / d params for every row i."""
    k = len(grad_outputs_per_row)
    if BATCHED_JACOBIAN and k > 1:
        try:
            stacked = [torch.stack([grad_outputs_per_row[i][j] for i in range(k)]) for j in range(len(outputs))]
            grads = torch.autograd.grad(outputs, params, grad_outputs=stacked, retain_graph=True, allow_unused=True,
                                        is_grads_batched=True)
            _fill_rows_batched(J, params, grads)
            if not retain_graph:
                pass          # the graph is released with the last reference; torchjd keeps the same contract
            return
        except RuntimeError as e:                         # no batching rule somewhere in the graph
            if "cuda" in str(e).lower() and "vmap" not in str(e).lower() and "batching" not in str(e).lower():
                raise
    for i in range(k):
        keep = retain_graph or i < k - 1
        grads = torch.autograd.grad(outputs, params, grad_outputs=list(grad_outputs_per_row[i]), retain_graph=keep, allow_unused=True)
        _fill_row(J, i, params, grads)


def _accumulate_flat(params: Sequence[Tensor], flat: Tensor) -> None:
    """torchjd Accumulate: `p.grad = g` if None else `p.grad += g`; g are views of the flat buffer."""
    off = 0
    add_dst, add_src = [], []
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view(p.shape)
        if p.grad is None:
            p.grad = g
        else:
            add_dst.append(p.grad)
            add_src.append(g)
        off += n
    if add_dst:
        torch._foreach_add_(add_dst, add_src)


def _check_aggregator(aggregator) -> None:
    if not isinstance(aggregator, Aggregator) and not hasattr(aggregator, "aggregate_into"):
        raise TypeError(f"aggregator must be a movae_b200 Aggregator, got {type(aggregator).__name__}")


def _aggregate_and_accumulate(J: Tensor, params: Sequence[Tensor], aggregator: Aggregator) -> Tensor:
    P = J.shape[1]
    flat = torch.empty(P, dtype=torch.float32, device=J.device)
    w = aggregator.aggregate_into(J, flat, accumulate=False)
    _accumulate_flat(params, flat)
    return w


def backward(tensors: Sequence[Tensor] | Tensor, aggregator: Aggregator, inputs: Optional[Iterable[Tensor]] = None,
             retain_graph: bool = False, parallel_chunk_size: Optional[int] = None) -> None:
    """Jacobian of `tensors` (k scalar losses) w.r.t. `inputs` (default: every leaf that requires
    grad in their graph), aggregated into `.grad` (main.py:196)."""
    _check_aggregator(aggregator)
    losses = [tensors] if isinstance(tensors, Tensor) else list(tensors)
    if len(losses) == 0:
        raise ValueError("`tensors` cannot be empty")
    for t in losses:
        if t.numel() != 1:
            raise ValueError("movae_b200.backward supports scalar objectives only (one Jacobian row each)")
    params = list(inputs) if inputs is not None else _leaves_of(losses)
    if not params:
        return
    k, P = len(losses), sum(p.numel() for p in params)
    J = _jacobian_buffer(k, P, params[0].device)
    stacked = torch.stack([t.reshape(()) for t in losses])
    eye = torch.eye(k, dtype=stacked.dtype, device=stacked.device)
    _jacobian_rows(J, [stacked], params, [[eye[i]] for i in range(k)], retain_graph)
    _aggregate_and_accumulate(J, params, aggregator)


def mtl_backward(losses: Sequence[Tensor], features: Sequence[Tensor] | Tensor, aggregator: Aggregator,
                 tasks_params: Optional[Sequence[Iterable[Tensor]]] = None,
                 shared_params: Optional[Iterable[Tensor]] = None, retain_graph: bool = False,
                 parallel_chunk_size: Optional[int] = None) -> None:
    """Multi-task variant (main.py:188-194): task-specific parameters receive the plain SUM of
    their losses' gradients; the shared parameters (those the features depend on) receive the
    aggregation of the k rows  d loss_i / d shared  obtained by back-propagating each loss's
    feature gradients through the shared part."""
    _check_aggregator(aggregator)
    losses = list(losses)
    feats = [features] if isinstance(features, Tensor) else list(features)
    if len(losses) == 0:
        raise ValueError("`losses` cannot be empty")
    if len(feats) == 0:
        raise ValueError("`features` cannot be empty")
    shared = list(shared_params) if shared_params is not None else _leaves_of(feats)
    shared_ids = {id(p) for p in shared}
    if tasks_params is None:
        tasks = [[p for p in _leaves_of([l], stop_at=feats) if id(p) not in shared_ids] for l in losses]
    else:
        tasks = [list(tp) for tp in tasks_params]
        if len(tasks) != len(losses):
            raise ValueError("`tasks_params` must have one entry per loss")
    k = len(losses)
    P = sum(p.numel() for p in shared)

    feat_grads = []
    for loss, tparams in zip(losses, tasks):
        outs = torch.autograd.grad(loss, feats + tparams, retain_graph=True, allow_unused=True)
        feat_grads.append([torch.zeros_like(f) if g is None else g for f, g in zip(feats, outs[:len(feats)])])
        for p, g in zip(tparams, outs[len(feats):]):
            if g is None:
                continue
            if p.grad is None:
                p.grad = g.clone() if g._base is not None else g
            else:
                p.grad += g
    if P == 0:
        return
    J = _jacobian_buffer(k, P, shared[0].device)
    _jacobian_rows(J, feats, shared, feat_grads, retain_graph)
    _aggregate_and_accumulate(J, shared, aggregator)
