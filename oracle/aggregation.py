"""TEST INFRASTRUCTURE ONLY (parity oracle + timed CPU baseline) -- the product path
(`movae_b200/`) must never import this package; only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may.

CPU restatement of the multi-objective gradient-aggregation step of MO-VAE:

    J[k,P] --(a3)--> G = J J^T --(a4|a5|a6)--> w[k] --(a7)--> g = w @ J --(a8)--> .grad

Sources followed (all paths under /root/reference):
  * Aligned-MTL weights ........ utils/torchmoo/aligned_mtl.py:97-133
  * MGDA Frank-Wolfe ........... utils/torchmoo/mgda.py:221-272, normalisers :274-285, :319-367,
                                 eigen clamp :287-317
  * UPGrad ..................... NOT in /root/reference.  It lives in the un-vendored dependency
                                 `torchjd @ git+https://github.com/TorchJD/torchjd.git@main`
                                 (requirements.txt:58, no commit pin) which calls
                                 `qpsolvers==4.8.1` -> `quadprog==0.1.13` (requirements.txt:46-47).
                                 Restated from the published algorithm: UPGrad (Quinton & Rey,
                                 "Jacobian Descent for Multi-Objective Optimization", 2024, Sec. 4:
                                 project each row onto the dual cone of all rows and average) with
                                 torchjd's pipeline normalize(trace) -> regularize(+eps I) ->
                                 k QPs  argmin_{v >= u_i e_i} v^T G v  -> sum, and quadprog's
                                 solver = Goldfarb & Idnani (1983) dual active set.  The same
                                 Gram -> diag(u) -> regularize -> project_weights -> sum(dim=0)
                                 pattern is visible at utils/torchmoo/nupgrad.py:122-126.
  * Gramian / recombine / Sum .. torchjd `compute_gramian` (J @ J.T), `weights @ J`; call sites
                                 main.py:189-196, :1195-1224.

PINNING STATUS
  * aligned_mtl / mgda: pinned against the reference's own files executed behind
    oracle/ref_shim.py (tests/golden/aggregation_golden.json, made by tests/golden/make_golden.py)
    and against the docstring known-answer vectors mgda.py:57-86.
  * upgrad: **parity unpinned by reference code** -- the only reference-held expectation is the
    4-digit docstring vector nupgrad.py:55-62 ([0.2929, 1.9004, 1.9004]); it is additionally
    cross-checked here by two independent exact solvers (Goldfarb-Idnani vs exhaustive active-set
    enumeration) and KKT residuals.

Precision policy: the arbiter Gramian is accumulated in float64 and rounded to
float32 once; the small solves then run exactly as the reference runs them (float32 torch ops for
MGDA / Aligned-MTL, float64 QP for UPGrad); the recombination is accumulated in float64 and
rounded to float32 at the end.
"""
from __future__ import annotations

import itertools
import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

FP32_EPS = float(torch.finfo(torch.float32).eps)  # aligned_mtl.py:109 `torch.finfo().eps`

MGDA_NORM_TYPES = ("none", "l2", "loss", "loss+")
AMTL_SCALE_MODES = ("min", "median", "rmse")


# --------------------------------------------------------------------------------------
# a3: Gramian
# --------------------------------------------------------------------------------------
def gramian_fp64(J: torch.Tensor, chunk: int = 1 << 22) -> np.ndarray:
    """float64-accumulated J J^T (the arbiter).  Column-chunked so P=1e8 does not need a
    float64 copy of J."""
    assert J.dim() == 2
    k, P = J.shape
    G = np.zeros((k, k), dtype=np.float64)
    for c0 in range(0, P, chunk):
        blk = J[:, c0:c0 + chunk].to(torch.float64)
        G += (blk @ blk.T).numpy()
    return G


def gramian_reference_fp32(J: torch.Tensor) -> torch.Tensor:
    """The literal expression torchjd executes: float32 `J @ J.T` (library accumulation order)."""
    return J @ J.T


def arbiter_gramian(J: torch.Tensor) -> torch.Tensor:
    """float64 accumulation, one rounding to float32 -- what the kernels are compared with."""
    return torch.from_numpy(gramian_fp64(J)).to(torch.float32)


# --------------------------------------------------------------------------------------
# strictly convex QP  min 1/2 x^T H x  s.t. x >= lo   (two independent exact solvers)
# --------------------------------------------------------------------------------------
def qp_lower_bounds_goldfarb_idnani(H: np.ndarray, lo: np.ndarray, max_steps: int = 1000) -> np.ndarray:
    """Dual active-set method of Goldfarb & Idnani (Math. Prog. 27, 1983) specialised to
    `min 1/2 x^T H x  s.t.  x_j >= lo_j` (constraint normals are unit vectors), dense float64
    algebra instead of the factor updates quadprog keeps -- same pivoting rule (most violated
    constraint first), same iterates up to rounding.  H must be positive definite."""
    H = np.asarray(H, dtype=np.float64)
    lo = np.asarray(lo, dtype=np.float64)
    n = H.shape[0]
    Hinv = np.linalg.inv(H)
    x = np.zeros(n)                 # unconstrained minimiser (no linear term)
    active: list[int] = []          # indices of active constraints
    mult = np.zeros(0)              # their multipliers
    for _ in range(max_steps):
        slack = x - lo
        p = int(np.argmin(slack))
        if slack[p] >= -1e-15 * max(1.0, float(np.abs(lo).max())):
            return x
        u_plus = np.append(mult, 0.0)
        while True:
            if active:
                N = np.eye(n)[:, active]
                M = N.T @ Hinv @ N
                Nstar = np.linalg.solve(M, N.T @ Hinv)        # (N^T H^-1 N)^-1 N^T H^-1
                z = Hinv[:, p] - Hinv @ N @ Nstar[:, p]       # primal step direction
                r = Nstar[:, p]                                # dual step direction
            else:
                z = Hinv[:, p].copy()
                r = np.zeros(0)
            # partial step length (keeps multipliers >= 0)
            t1, drop = math.inf, -1
            for j, rj in enumerate(r):
                if rj > 0.0:
                    cand = u_plus[j] / rj
                    if cand < t1:
                        t1, drop = cand, j
            # full step length (makes constraint p tight)
            zp = z[p]
            t2 = -(x[p] - lo[p]) / zp if zp > 1e-300 else math.inf
            t = min(t1, t2)
            if math.isinf(t):
                raise ValueError("QP infeasible")  # torchjd raises ValueError when the solver returns None
            if math.isinf(t2):
                u_plus[:-1] -= t * r
                u_plus[-1] += t
                active.pop(drop)
                u_plus = np.delete(u_plus, drop)
                continue
            x = x + t * z
            u_plus[:-1] -= t * r
            u_plus[-1] += t
            if t == t2:
                active.append(p)
                mult = u_plus
                break
            active.pop(drop)
            u_plus = np.delete(u_plus, drop)
    raise ValueError("QP did not terminate")


def qp_lower_bounds_enumerate(H: np.ndarray, lo: np.ndarray) -> np.ndarray:
    """Exhaustive active-set enumeration (<= 2^n KKT systems).  The optimum of a strictly convex
    QP is unique, so the candidate with the smallest KKT violation is it."""
    H = np.asarray(H, dtype=np.float64)
    lo = np.asarray(lo, dtype=np.float64)
    n = H.shape[0]
    best, best_viol = None, math.inf
    for mask in itertools.product((False, True), repeat=n):
        act = np.array(mask)
        free = ~act
        x = np.where(act, lo, 0.0)
        if free.any():
            rhs = -H[np.ix_(free, act)] @ lo[act] if act.any() else np.zeros(int(free.sum()))
            x[free] = np.linalg.solve(H[np.ix_(free, free)], rhs)
        grad = H @ x
        viol = 0.0
        if free.any():
            viol = max(viol, float(np.max(lo[free] - x[free])))
        if act.any():
            viol = max(viol, float(np.max(-grad[act])))
        if viol < best_viol:
            best, best_viol = x, viol
    return best


# --------------------------------------------------------------------------------------
# a4: UPGrad weights  (torchjd recall)
# --------------------------------------------------------------------------------------
def upgrad_prepare(G: torch.Tensor, norm_eps: float, reg_eps: float) -> torch.Tensor:
    """normalize (divide by trace, zeros if trace < norm_eps) then regularize (+ reg_eps I),
    in the Gramian's own dtype (float32) as torchjd does before handing over to numpy."""
    tr = G.diagonal().sum()
    Gn = torch.zeros_like(G) if bool(tr < norm_eps) else G / tr
    return Gn + reg_eps * torch.eye(G.shape[0], dtype=G.dtype)


def upgrad_weights(
    G: torch.Tensor,
    norm_eps: float = 1e-4,
    reg_eps: float = 1e-4,
    pref_vector: Optional[torch.Tensor] = None,
    solver: str = "goldfarb_idnani",
) -> torch.Tensor:
    k = G.shape[0]
    u = torch.full((k,), 1.0 / k, dtype=G.dtype) if pref_vector is None else pref_vector.to(G.dtype)
    H = upgrad_prepare(G, norm_eps, reg_eps).to(torch.float64).numpy()
    U = np.diag(u.to(torch.float64).numpy())
    qp = qp_lower_bounds_goldfarb_idnani if solver == "goldfarb_idnani" else qp_lower_bounds_enumerate
    W = np.stack([qp(H, U[i]) for i in range(k)])
    return torch.from_numpy(W).to(G.dtype).sum(dim=0)


# --------------------------------------------------------------------------------------
# 8f: NUPGrad / PNUPGrad (utils/torchmoo/nupgrad.py:122-158, pnupgrad.py:127-134) and COMFORT (comfort.py)
# --------------------------------------------------------------------------------------
def dualproj_weights(G: torch.Tensor, pref: Optional[torch.Tensor] = None, norm_eps: float = 1e-4, reg_eps: float = 1e-4,
                     solver: str = "goldfarb_idnani") -> torch.Tensor:
    """torchjd `DualProj` [torchjd-recall; selectable at /root/reference/main.py:1221-1222]: u = preference weights (1/k),
    G' = regularize(normalize(G)), w = project_weights(u, G') = argmin_{v >= u} v^T G' v -- ONE QP (UPGrad solves k of them,
    one per row u_i e_i, and sums)."""
    k = G.shape[0]
    H = upgrad_prepare(G, norm_eps, reg_eps).double().numpy()
    lo = np.full(k, 1.0 / k) if pref is None else pref.double().numpy()
    x = qp_lower_bounds_goldfarb_idnani(H, lo) if solver == "goldfarb_idnani" else qp_lower_bounds_enumerate(H, lo)
    return torch.from_numpy(x).to(torch.float32)


def normalize_by_min_l2_norm(G: torch.Tensor, eps: float) -> torch.Tensor:
    """nupgrad.py:129-158, same float32 torch ops."""
    l2 = torch.sqrt(torch.clamp(G.diagonal(), min=eps))
    mask = l2 > eps
    if not bool(mask.any()):
        return torch.zeros_like(G)
    mn = l2[mask].min()
    sf = torch.where(mask, mn / l2, torch.zeros_like(l2))
    return G * (sf.unsqueeze(1) * sf.unsqueeze(0))


def normalize_l2(G: torch.Tensor, eps: float) -> torch.Tensor:
    """nupgrad.py:14-24 / pnupgrad.py `normalize`: G / (|g_i| |g_j|)."""
    n = torch.sqrt(torch.diag(G).clamp(min=eps))
    return G / (n.unsqueeze(1) * n.unsqueeze(0))


def nupgrad_weights(G: torch.Tensor, norm_eps: float = 1e-4, reg_eps: float = 1e-4, mode: str = "min_l2",
                    pref_vector: Optional[torch.Tensor] = None, solver: str = "goldfarb_idnani") -> torch.Tensor:
    """_NUPGradWrapper.forward (nupgrad.py:122-126) / _PNUPGradWrapper.forward (pnupgrad.py:127-134, `mode`
    = which branch the random draw took): U = diag(u); G' = regularize(normalise(G)); sum_i QP_i."""
    k = G.shape[0]
    u = torch.full((k,), 1.0 / k, dtype=G.dtype) if pref_vector is None else pref_vector.to(G.dtype)
    Gn = normalize_by_min_l2_norm(G, norm_eps) if mode == "min_l2" else normalize_l2(G, norm_eps)
    H = (Gn + torch.eye(k, dtype=G.dtype) * reg_eps).to(torch.float64).numpy()
    U = np.diag(u.to(torch.float64).numpy())
    qp = qp_lower_bounds_goldfarb_idnani if solver == "goldfarb_idnani" else qp_lower_bounds_enumerate
    W = np.stack([qp(H, U[i]) for i in range(k)])
    return torch.from_numpy(W).to(G.dtype).sum(dim=0)


def comfort_beta(epoch: int, total_epochs: int, k: float = 1.0, a: float = 1.0, l: float = 0.01, u: float = 1.0) -> float:  # noqa: E741
    """comfort.py:26-65."""
    if total_epochs <= 1:
        return u
    progress = (epoch - 1) / (total_epochs - 1)
    progress = min(1.0, max(0.0, progress)) ** a
    f = progress if k <= 0 else (1.0 - math.exp(-k * progress)) / (1.0 - math.exp(-k))
    return float(min(u, max(l, l + (u - l) * f)))


# --------------------------------------------------------------------------------------
# a5: MGDA weights (float32 torch ops, same op sequence as the reference loop)
# --------------------------------------------------------------------------------------
def mgda_normalize(G: torch.Tensor, norm_type: str, losses: Optional[torch.Tensor]) -> torch.Tensor:
    if norm_type not in MGDA_NORM_TYPES:
        raise ValueError(f"bad norm_type {norm_type!r}")
    if norm_type == "none":
        return G
    k = G.shape[0]
    if norm_type in ("loss", "loss+"):
        if losses is None:
            raise RuntimeError("losses must be set for norm_type 'loss'/'loss+'")
        if losses.dim() != 1:
            raise ValueError("losses must be 1-D")
        if losses.shape[0] != k:
            raise ValueError(f"{losses.shape[0]} losses for a {k}x{k} gramian")
        ell = losses.detach().to(G.dtype).clamp(min=1e-20)
    if norm_type == "l2":
        s = torch.sqrt(torch.diag(G).clamp(min=1e-20))
    elif norm_type == "loss":
        s = ell
    else:
        s = ell * torch.sqrt(torch.diag(G).clamp(min=1e-20))
    return G / (s.unsqueeze(1) * s.unsqueeze(0))


def mgda_eigen_clamp(G: torch.Tensor, min_eigenvalue_eps: float) -> torch.Tensor:
    lam, V = torch.linalg.eigh(G)
    return V @ (lam.clamp(min=min_eigenvalue_eps).unsqueeze(1) * V.T)


def mgda_weights(
    G: torch.Tensor,
    norm_type: str = "none",
    losses: Optional[torch.Tensor] = None,
    epsilon: float = 1e-5,
    max_iters: int = 250,
    stable: bool = False,
    min_eigenvalue_eps: float = 1e-10,
) -> Tuple[torch.Tensor, int, float]:
    """Returns (alpha, convergence_count, gamma).  alpha is NOT renormalised on exit and is
    meant to be applied to the un-normalised J (mgda.py:272 returns alpha as is)."""
    R = mgda_normalize(G, norm_type, losses)
    if stable:
        R = mgda_eigen_clamp(R, min_eigenvalue_eps)
    k = R.shape[0]
    alpha = torch.ones(k, dtype=R.dtype) / k
    gamma = 0.0
    it = -1
    for it in range(max_iters):
        t = torch.argmin(R @ alpha)            # first minimal index
        e_t = torch.zeros(k, dtype=R.dtype)
        e_t[t] = 1.0
        a = alpha @ (R @ e_t)
        b = alpha @ (R @ alpha)
        c = e_t @ (R @ e_t)
        if c <= a:
            gamma = 1.0
        elif b <= a:
            gamma = 0.0
        else:
            gamma = (b - a) / (b + c - 2 * a)
        alpha = (1 - gamma) * alpha + gamma * e_t
        if gamma < epsilon:
            break
    return alpha, it + 1, float(gamma)


# --------------------------------------------------------------------------------------
# a6: Aligned-MTL weights
# --------------------------------------------------------------------------------------
def aligned_mtl_weights(
    G: torch.Tensor,
    scale_mode: str = "min",
    pref_vector: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float32,
) -> Tuple[torch.Tensor, int]:
    """Returns (alpha, rank).  `dtype=float64` runs the same algorithm in double on the same
    (float32-valued) Gramian; the rank threshold keeps the reference's float32 eps."""
    if scale_mode not in AMTL_SCALE_MODES:
        raise ValueError(f"Invalid scale_mode={scale_mode!r}. Expected 'min', 'median', or 'rmse'.")
    M = G.to(dtype)
    k = M.shape[0]
    w0 = torch.full((k,), 1.0 / k, dtype=dtype) if pref_vector is None else pref_vector.to(dtype)
    lam, V = torch.linalg.eigh(M, UPLO="U")
    tol = lam.max() * k * FP32_EPS
    rank = int((lam > tol).sum())
    if rank == 0:
        return w0.to(G.dtype), 0
    order = torch.argsort(lam, descending=True)
    lam, V = lam[order][:rank], V[:, order][:, :rank]
    if scale_mode == "min":
        scale = lam[-1]
    elif scale_mode == "median":
        scale = torch.median(lam)          # lower middle
    else:
        scale = lam.mean()
    B = scale.sqrt() * V @ torch.diag(1 / lam.sqrt()) @ V.T
    return (B @ w0).to(G.dtype), rank


def aligned_mtl_conditioning(G: torch.Tensor) -> dict:
    """Well-posedness report: does the float32 rank decision agree with float64,
    and how close is any eigenvalue to the threshold?"""
    k = G.shape[0]
    lam64 = torch.linalg.eigvalsh(G.to(torch.float64), UPLO="U")
    lam32 = torch.linalg.eigvalsh(G.to(torch.float32), UPLO="U")
    tol64 = float(lam64.max()) * k * FP32_EPS
    tol32 = float(lam32.max()) * k * FP32_EPS
    r64 = int((lam64 > tol64).sum())
    r32 = int((lam32 > tol32).sum())
    lmax = max(float(lam64.max()), 1e-300)
    pos = lam64[lam64 > tol64]
    return {
        "rank_fp64": r64,
        "rank_fp32": r32,
        "rank_agrees": r64 == r32,
        "cond_kept": float(lmax / float(pos.min())) if len(pos) else float("inf"),
        "min_rel_gap_to_tol": float(((lam64 - tol64).abs() / lmax).min()),
    }


# --------------------------------------------------------------------------------------
# Sum / Mean, a7 recombine, a9 similarity hook
# --------------------------------------------------------------------------------------
def sum_weights(k: int) -> torch.Tensor:
    return torch.ones(k, dtype=torch.float32)


def mean_weights(k: int) -> torch.Tensor:
    return torch.full((k,), 1.0 / k, dtype=torch.float32)


def recombine_fp64(w: torch.Tensor, J: torch.Tensor, chunk: int = 1 << 22) -> torch.Tensor:
    """g = w @ J with float64 accumulation, rounded once to float32."""
    k, P = J.shape
    out = torch.empty(P, dtype=torch.float32)
    w64 = w.to(torch.float64)
    for c0 in range(0, P, chunk):
        out[c0:c0 + chunk] = (w64 @ J[:, c0:c0 + chunk].to(torch.float64)).to(torch.float32)
    return out


def recombine_reference_fp32(w: torch.Tensor, J: torch.Tensor) -> torch.Tensor:
    return w @ J


def gradient_similarity_from_gramian(G: np.ndarray, w: Sequence[float]) -> float:
    """cos(J^T w, mean_rows(J)) from the Gramian alone (replaces the two extra J passes of the
    hook main.py:94-122):  <J^T w, J^T 1/k> / (|J^T w| |J^T 1/k|)."""
    G = np.asarray(G, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    k = G.shape[0]
    m = np.full(k, 1.0 / k)
    num = w @ G @ m
    den = math.sqrt(max(w @ G @ w, 0.0)) * math.sqrt(max(m @ G @ m, 0.0))
    return float(num / max(den, 1e-8))   # F.cosine_similarity eps=1e-8 on the product of norms


# --------------------------------------------------------------------------------------
# whole step
# --------------------------------------------------------------------------------------
AGGREGATOR_NAMES = (
    "sum", "mean", "upgrad", "dualproj",
    "aligned_mtl", "aligned_mtl_median", "aligned_mtl_rmse",
    "mgda", "mgda_ln", "mgda_gn", "mgda_lgn",
)
_MGDA_NORM_OF = {"mgda": "none", "mgda_ln": "l2", "mgda_gn": "loss", "mgda_lgn": "loss+"}
_AMTL_MODE_OF = {"aligned_mtl": "min", "aligned_mtl_median": "median", "aligned_mtl_rmse": "rmse"}


def weights_from_gramian(name: str, G: torch.Tensor, losses: Optional[torch.Tensor] = None, **kw) -> Tuple[torch.Tensor, dict]:
    """Name map of main.py:1194-1230 onto the weightings above.  G is float32 [k,k]."""
    name = name.lower()
    k = G.shape[0]
    if name in ("sum", "jd_sum"):
        return sum_weights(k), {}
    if name == "mean":
        return mean_weights(k), {}
    if name == "upgrad":
        return upgrad_weights(G, kw.get("norm_eps", 1e-4), kw.get("reg_eps", 1e-4), kw.get("pref_vector")), {}
    if name == "dualproj":
        return dualproj_weights(G, kw.get("pref_vector"), kw.get("norm_eps", 1e-4), kw.get("reg_eps", 1e-4)), {}
    if name in _AMTL_MODE_OF or name in ("amtl", "amtl_min", "aligned_mtl_min"):
        # amtl_dtype=float64 is the ARBITER (same algorithm in double on the float32-valued Gramian);
        # the default float32 is the reference's literal computation (LAPACK ssyevd), reported beside it
        w, rank = aligned_mtl_weights(G, _AMTL_MODE_OF.get(name, "min"), kw.get("pref_vector"),
                                      dtype=kw.get("amtl_dtype", torch.float32))
        return w, {"rank": rank}
    if name in _MGDA_NORM_OF:
        w, count, gamma = mgda_weights(
            G, _MGDA_NORM_OF[name], losses, kw.get("epsilon", 1e-5), kw.get("max_iters", 250),
            kw.get("stable", False), kw.get("min_eigenvalue_eps", 1e-10))
        return w, {"convergence_count": count, "gamma": gamma}
    raise ValueError(f"Aggregator {name} not supported")


def aggregate(name: str, J: torch.Tensor, losses: Optional[torch.Tensor] = None, **kw):
    """(G, w, g) of the whole step under the precision policy in the module header."""
    if J.dim() != 2:
        raise ValueError(f"expected a 2-D matrix, got shape {tuple(J.shape)}")
    G = arbiter_gramian(J)
    w, info = weights_from_gramian(name, G, losses, **kw)
    return G, w, recombine_fp64(w, J), info


def aggregate_reference_fp32(name: str, J: torch.Tensor, losses: Optional[torch.Tensor] = None, **kw):
    """Same step with the reference's own float32 `J @ J.T` and `w @ J` -- this is the path that
    is TIMED as the CPU baseline (bench.py) and reported beside the arbiter in the parity tables."""
    G = gramian_reference_fp32(J)
    w, info = weights_from_gramian(name, G, losses, **kw)
    return G, w, recombine_reference_fp32(w, J), info
