"""Jacobian-descent engine entry points with the signatures the reference calls
(/root/reference/main.py:17, :189-196):

    backward(tensors, aggregator=..., inputs=None, retain_graph=False)
    mtl_backward(losses=, features=, aggregator=, tasks_params=None, shared_params=None, retain_graph=False)

They replace torchjd.autojac.{backward,mtl_backward} (un-vendored dependency, requirements.txt:58;
semantics restated in SURVEY.md App. A).  The k per-objective backward passes are plain torch
autograd (not ours); what changes is everything after them:

  * NO Jacobian matrix is built at all: the fused kernel reads every objective's gradient of every
    shared parameter tensor where autograd left it, through a table of k row pointers per tensor
    (`movae_aggregate_segments_f32`) -- no per-parameter reshape + `torch.cat`, not even one copy.  The flat
    J[k, ldJ] buffer (one multi-tensor copy per row) remains for the cases that need a matrix: a data-parallel plan
    (the rows are reduce-scattered), forward hooks on the weighting (they receive J), more than 32 shared tensors;
  * Gramian pass, solve and recombination are ONE launch on the current stream;
  * K3 writes the aggregated gradient into one flat buffer and the parameters' `.grad` become views
    of it (assign if `.grad is None`, `+=` otherwise: torchjd `Accumulate` semantics).  When the
    parameters live in a `movae_b200.FlatParameters` (optim.py) that buffer is the persistent flat
    gradient buffer the fused optimizer step (K7) reads, and J's columns follow its 16-byte aligned layout;
  * `mtl_backward` back-propagates only the objectives that reach the features at all: a loss whose
    feature gradients are all `None` (e.g. `embedding_loss`, vq_vae.py:52) is an identically zero row
    of J and costs one memset instead of a backward pass through the shared encoder.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
from torch import Tensor

from .aggregation import Aggregator
from .optim import flat_plan

_J_CACHE: dict = {}

# Jacobian rows by k sequential backward passes (default) or by ONE batched (vmapped) pass over the live objectives like
# torchjd's default (parallel_chunk_size=None).  Measured on B200 (tools/step_probe.py, BASELINE model configs, step under a
# CUDA graph): sequential 2.57 / 14.98 / 21.92 ms vs batched 2.64 / 15.97 / 23.11 ms per step (VQ-VAE / GG-VQ-VAE / VQ-VAE2)
# -- the vmapped convolution-backward costs more than the launches it saves.
BATCHED_JACOBIAN = False


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


def _jacobian_buffer(k: int, P: int, device: torch.device, layout: tuple = (), min_ld: int = 0) -> Tensor:
    """The flat J[k, ldJ] buffer, zero-filled at allocation (padding columns of an aligned layout are never
    written afterwards, so they stay zero) and reused every step for the same (k, layout).  `min_ld`: a data-parallel
    aggregation wants the row padded to world x shard columns (parallel.DataParallel)."""
    ld = max((P + 3) // 4 * 4, min_ld)
    key = (device, k, ld, layout)
    buf = _J_CACHE.get(key)
    if buf is None:
        _J_CACHE.clear()                                  # one live model per process is the norm
        buf = torch.zeros((k, ld), dtype=torch.float32, device=device)
        _J_CACHE[key] = buf
    return buf[:, :P] if ld != P else buf


def _dense_offsets(params: Sequence[Tensor]) -> List[int]:
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += p.numel()
    return offs


def _fill_row(J: Tensor, i: int, params: Sequence[Tensor], grads: Sequence[Optional[Tensor]], offsets: Sequence[int]) -> None:
    dst, src = [], []
    for p, g, off in zip(params, grads, offsets):
        view = J[i, off:off + p.numel()]
        if g is None:
            view.zero_()
        else:
            dst.append(view)
            src.append(g.reshape(-1))
    if dst:
        torch._foreach_copy_(dst, src)


def _fill_rows_batched(J: Tensor, rows: Sequence[int], params: Sequence[Tensor], grads: Sequence[Optional[Tensor]],
                       offsets: Sequence[int]) -> None:
    """grads[p]: [len(rows), *p.shape] (or None) -> columns off..off+numel of rows `rows` of J, one strided copy each."""
    r = len(rows)
    whole = list(rows) == list(range(J.shape[0]))
    Jr = J if whole else None
    dst, src = [], []
    for p, g, off in zip(params, grads, offsets):
        n = p.numel()
        if whole:
            view = Jr[:, off:off + n]
        elif list(rows) == list(range(rows[0], rows[0] + r)):
            view = J[rows[0]:rows[0] + r, off:off + n]
        else:
            view = None
        if view is not None:
            if g is None:
                view.zero_()
            else:
                dst.append(view)
                src.append(g.reshape(r, n))
        else:                                             # non-consecutive live rows: one copy per row
            for a, i in enumerate(rows):
                if g is None:
                    J[i, off:off + n].zero_()
                else:
                    dst.append(J[i, off:off + n])
                    src.append(g[a].reshape(n))
    if dst:
        torch._foreach_copy_(dst, src)


def _jacobian_rows(J: Tensor, outputs: Sequence[Tensor], params: Sequence[Tensor], grad_outputs_per_row: Sequence[Optional[Sequence[Tensor]]],
                   retain_graph: bool, offsets: Optional[Sequence[int]] = None, dp=None) -> None:
    """Fills J[i] = d(sum_j <outputs[j], grad_outputs_per_row[i][j]>) / d params for every row i; a row whose
    entry is None is identically zero (no backward pass).  With a data-parallel plan (`dp`) every finished row is
    handed to it at once, so its reduce-scatter overlaps the backward pass of the next row."""
    offsets = _dense_offsets(params) if offsets is None else offsets
    Jp = None
    if dp is not None:
        kk, P = J.shape
        padded = dp.padded_columns(P)
        Jp = J.as_strided((kk, padded), (J.stride(0) if kk > 1 else padded, 1))
        dp.begin_rows(kk, P, J.dtype, J.device)
    live = [i for i, g in enumerate(grad_outputs_per_row) if g is not None]
    for i, g in enumerate(grad_outputs_per_row):
        if g is None:
            J[i].zero_()
            if dp is not None:
                dp.row_zero(i)
    k = len(live)
    if k == 0:
        return
    if BATCHED_JACOBIAN and k > 1 and not isinstance(outputs[0], (list, tuple)):
        try:
            stacked = [torch.stack([grad_outputs_per_row[i][j] for i in live]) for j in range(len(outputs))]
            grads = torch.autograd.grad(outputs, params, grad_outputs=stacked, retain_graph=True, allow_unused=True,
                                        is_grads_batched=True)
            _fill_rows_batched(J, live, params, grads, offsets)
            if dp is not None:
                for i in live:
                    dp.row_ready(i, Jp[i])
            return            # the graph is released with the last reference; torchjd keeps the same contract
        except RuntimeError as e:                         # no batching rule somewhere in the graph
            if "cuda" in str(e).lower() and "vmap" not in str(e).lower() and "batching" not in str(e).lower():
                raise
    per_row = isinstance(outputs[0], (list, tuple))
    for a, i in enumerate(live):
        keep = retain_graph or a < k - 1
        grads = torch.autograd.grad(outputs[i] if per_row else outputs, params, grad_outputs=list(grad_outputs_per_row[i]),
                                    retain_graph=keep, allow_unused=True)
        _fill_row(J, i, params, grads, offsets)
        if dp is not None:
            dp.row_ready(i, Jp[i])


def _row_gradients(outputs: Sequence[Tensor], params: Sequence[Tensor], grad_outputs_per_row: Sequence[Optional[Sequence[Tensor]]],
                   retain_graph: bool) -> List[Optional[Sequence[Optional[Tensor]]]]:
    """Per row i: the tuple autograd returns for d(sum_j <outputs[j], grad_outputs_per_row[i][j]>) / d params (entries may be
    None), or None for an identically zero row (no backward pass).  `outputs[j]` may be per-row (a list of lists) when
    every row differentiates its own scalar (`backward`)."""
    live = [i for i, g in enumerate(grad_outputs_per_row) if g is not None]
    out: List[Optional[Sequence[Optional[Tensor]]]] = [None] * len(grad_outputs_per_row)
    for a, i in enumerate(live):
        keep = retain_graph or a < len(live) - 1
        outs_i = outputs[i] if isinstance(outputs[0], (list, tuple)) else outputs
        out[i] = torch.autograd.grad(outs_i, params, grad_outputs=list(grad_outputs_per_row[i]), retain_graph=keep, allow_unused=True)
    return out


_ZEROS: dict = {}


def _zero_row(device: torch.device, n: int) -> Tensor:
    """A shared all-zero buffer every identically zero Jacobian entry points at (never written)."""
    z = _ZEROS.get(device)
    if z is None or z.numel() < n:
        z = torch.zeros(max(n, 1 << 16), dtype=torch.float32, device=device)
        _ZEROS[device] = z
    return z


def _segments_possible(params: Sequence[Tensor], aggregator) -> bool:
    from . import _lib as L

    return (not BATCHED_JACOBIAN and _dp_of(aggregator) is None and len(params) <= L.MAX_SEGMENTS
            and hasattr(aggregator, "supports_segments") and aggregator.supports_segments())


def _aggregate_segments_and_accumulate(row_grads, params: Sequence[Tensor], cols: Sequence[int], aggregator, plan) -> Tensor:
    """The fused launch over the gradient tensors themselves; `.grad` semantics as _aggregate_and_accumulate."""
    from . import ops

    dev = params[0].device
    numels = [p.numel() for p in params]
    zero = _zero_row(dev, max(numels))
    rows = []
    for gs in row_grads:
        if gs is None:
            rows.append([zero] * len(params))
            continue
        r = []
        for g in gs:
            if g is None:
                r.append(zero)
            else:
                g = g.detach()
                r.append(g if ops.segment_ok(g) else g.to(torch.float32).contiguous().clone())     # rare: a strided / unaligned gradient
        rows.append(r)
    if plan is not None:
        owner, _, _, lo, hi = plan
        states = {owner.grad_state(p) for p in params}
        if states == {"none"}:
            w = aggregator.aggregate_segments_into(rows, numels, cols, owner.flat_grad[lo:hi], accumulate=False)
            for p in params:
                p.grad = owner.grad_view(p)
            return w
        if states == {"view"}:
            return aggregator.aggregate_segments_into(rows, numels, cols, owner.flat_grad[lo:hi], accumulate=True)
    # plain parameters (or mixed .grad states): a fresh flat buffer with every tensor on a 16-byte boundary
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += (n + 3) // 4 * 4
    flat = torch.empty(max(off, 4), dtype=torch.float32, device=dev)
    w = aggregator.aggregate_segments_into(rows, numels, offs, flat, accumulate=False)
    add_dst, add_src = [], []
    for p, o, n in zip(params, offs, numels):
        g = flat[o:o + n].view(p.shape)
        if p.grad is None:
            p.grad = g
        else:
            add_dst.append(p.grad)
            add_src.append(g)
    if add_dst:
        torch._foreach_add_(add_dst, add_src)
    return w


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


def _dp_of(aggregator):
    return getattr(aggregator, "data_parallel", None)


def _run_aggregation(J: Tensor, aggregator: Aggregator, out: Tensor, accumulate: bool) -> Tensor:
    """K1 -> K2 -> K3 into `out`; through the reduce-scatter / all-gather plan when the aggregator is data-parallel."""
    dp = _dp_of(aggregator)
    if dp is None:
        return aggregator.aggregate_into(J, out, accumulate=accumulate)
    return dp.aggregate_rows_into(J.shape[1], out, accumulate)     # the rows were handed over while they were built


def _aggregate_and_accumulate(J: Tensor, params: Sequence[Tensor], aggregator: Aggregator, plan=None) -> Tensor:
    if plan is not None:
        # the parameters are a run of a FlatParameters layout and J's columns follow it: K3 writes (or adds)
        # straight into the persistent flat gradient buffer the optimizer kernel reads
        owner, _, _, lo, hi = plan
        states = {owner.grad_state(p) for p in params}
        if states == {"none"}:
            w = _run_aggregation(J, aggregator, owner.flat_grad[lo:hi], False)
            for p in params:
                p.grad = owner.grad_view(p)
            return w
        if states == {"view"}:
            return _run_aggregation(J, aggregator, owner.flat_grad[lo:hi], True)
        cols = plan[2]
        flat = torch.empty(J.shape[1], dtype=torch.float32, device=J.device)
        w = _run_aggregation(J, aggregator, flat, False)
        for p, off in zip(params, cols):
            g = flat[off:off + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = g
            else:
                p.grad += g
        return w
    P = J.shape[1]
    flat = torch.empty(P, dtype=torch.float32, device=J.device)
    w = _run_aggregation(J, aggregator, flat, False)
    _accumulate_flat(params, flat)
    return w


def _layout(params: Sequence[Tensor]):
    """(params in J's column order, column offsets, P, plan, cache key)."""
    plan = flat_plan(params)
    if plan is not None:
        _, ordered, cols, lo, hi = plan
        return ordered, cols, hi - lo, plan, ("flat", plan[0].uid, lo, hi)
    params = list(params)
    return params, _dense_offsets(params), sum(p.numel() for p in params), None, ()


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
    k = len(losses)
    params, cols, P, plan, key = _layout(params)
    dp = _dp_of(aggregator)
    # every row differentiates ITS OWN loss: a stacked loss with one-hot cotangents would walk the union graph of all
    # objectives k times and push zero cotangents through the other objectives' private subgraphs (0 * inf = NaN)
    outputs = [[t.reshape(())] for t in losses]
    ones = [[torch.ones((), dtype=t.dtype, device=t.device)] for t in losses]
    if _segments_possible(params, aggregator):
        _aggregate_segments_and_accumulate(_row_gradients(outputs, params, ones, retain_graph), params, cols, aggregator, plan)
        return
    J = _jacobian_buffer(k, P, params[0].device, key, dp.padded_columns(P) if dp else 0)
    _jacobian_rows(J, outputs, params, ones, retain_graph, cols, dp)
    _aggregate_and_accumulate(J, params, aggregator, plan)


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

    dp = _dp_of(aggregator)
    feat_grads: List[Optional[List[Tensor]]] = []
    task_sum: dict = {}                                  # id(param) -> [param, this call's gradient summed over the tasks]
    for loss, tparams in zip(losses, tasks):
        outs = torch.autograd.grad(loss, feats + tparams, retain_graph=True, allow_unused=True)
        fg = outs[:len(feats)]
        if all(g is None for g in fg):
            feat_grads.append(None)                       # this objective never reaches the features: zero row of J
        else:
            feat_grads.append([torch.zeros_like(f) if g is None else g for f, g in zip(feats, fg)])
        for p, g in zip(tparams, outs[len(feats):]):
            if g is None:
                continue
            ent = task_sum.get(id(p))
            if ent is None:
                task_sum[id(p)] = [p, g]
            else:
                ent[1] = ent[1] + g
    # task-specific parameters: the plain SUM of their losses' gradients (averaged over the ranks when data-parallel),
    # assigned when .grad is None (into the flat gradient buffer if the parameter lives in one), else added
    fresh_flat: dict = {}
    for p, g in task_sum.values():
        if p.grad is None and getattr(p, "_movae_flat", None) is not None:
            ent = fresh_flat.setdefault(id(p._movae_flat[0]), [p._movae_flat[0], [], []])
            ent[1].append(p)
            ent[2].append(g)
        else:
            if dp is not None:
                g = g.contiguous() if g._base is None else g.clone()
                dp.average_([g])
            if p.grad is None:
                p.grad = g.clone() if g._base is not None else g
            else:
                p.grad += g
    for owner, ps, gs in fresh_flat.values():
        owner.adopt(ps, gs)
        if dp is not None:
            idx = sorted(p._movae_flat[1] for p in ps)
            runs, start, prev = [], idx[0], idx[0]
            for i in idx[1:] + [None]:
                if i is None or i != prev + 1:
                    runs.append(owner.flat_grad[owner.offsets[start]:owner.offsets[prev] + owner.padded_numel(prev)])
                    start = i
                prev = i if i is not None else prev
            dp.average_(runs)
    if not shared:
        return
    shared, cols, P, plan, key = _layout(shared)
    if P == 0:
        return
    if _segments_possible(shared, aggregator):
        _aggregate_segments_and_accumulate(_row_gradients(feats, shared, feat_grads, retain_graph), shared, cols, aggregator, plan)
        return
    J = _jacobian_buffer(k, P, shared[0].device, key, dp.padded_columns(P) if dp else 0)
    _jacobian_rows(J, feats, shared, feat_grads, retain_graph, cols, dp)
    _aggregate_and_accumulate(J, shared, aggregator, plan)
