"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-ins for the third-party names the reference's own files import, so that
`/root/reference/utils/torchmoo/{aligned_mtl,mgda}.py` and
`/root/reference/models/vq_vae.py::VectorQuantizer` can be executed UNMODIFIED in
the build container to produce golden vectors (tests/golden/make_golden.py).

The reference depends on `torchjd @ git+...@main` (requirements.txt:58, unpinned
branch), `qpsolvers==4.8.1`, `quadprog==0.1.13` (requirements.txt:46-47) and
`torchsummary`; none is installed here and there is no network.  Only the base
classes the two pure-torch files need are provided (aligned_mtl.py:33-36,
mgda.py:6-7).  `UPGrad` itself lives in torchjd and can NOT be obtained this way:
its parity is pinned only by the docstring known-answer vector (nupgrad.py:55-62).

`/root/reference` does not exist on the GPU box; nothing here runs there.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("MOVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "torchmoo", "mgda.py"))


class _Weighting(nn.Module):
    """torchjd.aggregation._weighting_bases.Weighting stand-in (generic-subscriptable)."""

    def __class_getitem__(cls, item):
        return cls

    def __lshift__(self, fn):
        return _Composed(self, fn)


class _Composed(_Weighting):
    def __init__(self, outer: nn.Module, fn):
        super().__init__()
        self.outer = outer
        self.fn = fn

    def forward(self, stat):
        return self.outer(self.fn(stat))


class _GramianWeightedAggregator(nn.Module):
    """weights = psd_weighting(J @ J.T); return weights @ J  (torchjd recall)."""

    def __init__(self, psd_weighting: nn.Module):
        super().__init__()
        self.psd_weighting = psd_weighting
        self.weighting = psd_weighting << (lambda m: m @ m.T)

    def forward(self, matrix):
        if matrix.dim() != 2:
            raise ValueError(f"expected a 2-D matrix, got shape {tuple(matrix.shape)}")
        return self.weighting(matrix) @ matrix


class _MeanWeighting(_Weighting):
    def forward(self, gramian):
        m = gramian.shape[0]
        return torch.full((m,), 1.0 / m, dtype=gramian.dtype, device=gramian.device)


class _ConstantWeighting(_Weighting):
    def __init__(self, w):
        super().__init__()
        self.w = w

    def forward(self, gramian):
        return self.w.to(dtype=gramian.dtype, device=gramian.device)


def _pref_vector_to_weighting(pref_vector, default):
    return default if pref_vector is None else _ConstantWeighting(pref_vector)


def _project_weights(U, G, solver):
    """[torchjd-recall] torchjd.aggregation._utils.dual_cone.project_weights: for each row u of U,
    argmin_{v >= u} v^T G v in float64 on the host (qpsolvers/quadprog in the real dependency; here the
    oracle's Goldfarb-Idnani restatement), stacked, cast back to G's dtype."""
    import numpy as np

    from oracle.aggregation import qp_lower_bounds_goldfarb_idnani

    H = G.detach().to(torch.float64).cpu().numpy()
    Un = U.detach().to(torch.float64).cpu().numpy()
    W = np.stack([qp_lower_bounds_goldfarb_idnani(H, Un[i]) for i in range(Un.shape[0])])
    return torch.from_numpy(W).to(dtype=G.dtype, device=G.device)


def _raise_non_differentiable_error(module, grad_output):
    raise RuntimeError(f"{module} is not differentiable")


def _pref_vector_to_str_suffix(pref_vector):
    return "" if pref_vector is None else f"([{', '.join(f'{float(v):g}' for v in pref_vector)}])"


def install() -> None:
    """Register fake `torchjd.*` / `torchsummary` modules in sys.modules (idempotent)."""
    if "torchjd" in sys.modules and getattr(sys.modules["torchjd"], "_movae_shim", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m._movae_shim = True
        sys.modules[name] = m
        return m

    mod("torchjd")
    mod("torchjd.aggregation")
    mod("torchjd.aggregation._aggregator_bases", GramianWeightedAggregator=_GramianWeightedAggregator)
    mod("torchjd.aggregation._weighting_bases", PSDMatrix=torch.Tensor, Weighting=_Weighting)
    mod("torchjd.aggregation._mean", MeanWeighting=_MeanWeighting)
    mod("torchjd.aggregation._utils")
    mod(
        "torchjd.aggregation._utils.pref_vector",
        pref_vector_to_weighting=_pref_vector_to_weighting,
        pref_vector_to_str_suffix=_pref_vector_to_str_suffix,
    )
    mod("torchjd.aggregation._utils.dual_cone", project_weights=_project_weights)
    mod("torchjd.aggregation._utils.non_differentiable", raise_non_differentiable_error=_raise_non_differentiable_error)
    mod("torchsummary", summary=lambda *a, **k: None)


def _load(name: str, relpath: str):
    install()
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def load_reference_aligned_mtl():
    return _load("_ref_aligned_mtl", "utils/torchmoo/aligned_mtl.py")


def load_reference_mgda():
    return _load("_ref_mgda", "utils/torchmoo/mgda.py")


def load_reference_nupgrad():
    """utils/torchmoo/nupgrad.py: its own code (normalisation, wrapper) runs unmodified; only
    `project_weights` (torchjd + quadprog) is the oracle's QP restatement."""
    return _load("_ref_nupgrad", "utils/torchmoo/nupgrad.py")


def load_reference_pnupgrad():
    return _load("_ref_pnupgrad", "utils/torchmoo/pnupgrad.py")


def load_reference_vq():
    """models/vq_vae.py imports `utils.objectives` -> put the reference root on sys.path."""
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        return _load("_ref_vq_vae", "models/vq_vae.py")
    finally:
        if REFERENCE_ROOT in sys.path:
            sys.path.remove(REFERENCE_ROOT)
