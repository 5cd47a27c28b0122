"""TEST INFRASTRUCTURE -- CPU restatement (numpy float32) of the optimizer step the reference takes after the aggregation
(`optimizer.step()` /root/reference/main.py:214 on `optim.SGD | Adam | AdamW | RMSprop` built at main.py:1169-1176, preceded by
`clip_grad_norm_` main.py:211-212).  The algorithm lives in PyTorch itself (torch/optim/{sgd,adam,adamw,rmsprop}.py,
`_single_tensor_*`; torch/nn/utils/clip_grad.py), which IS installed here: tests/test_oracle_optim.py pins every function
below against torch.optim on the CPU, so the restatement is pinned by the reference's own dependency.  K7
(mo-vae_b200/csrc/optim.cu) implements exactly these formulas; only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def clip_coefficient(grads, max_norm: float) -> np.float32:
    """torch.nn.utils.clip_grad_norm_: total = || (||g_i||_2)_i ||_2 in float32, coef = max_norm / (total + 1e-6) clamped to 1."""
    total = F(np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads)))
    coef = F(max_norm) / (total + F(1e-6))
    return F(min(coef, F(1.0)))


def adam_step(p, g, m, v, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
              weight_decay: float = 0.0, decoupled: bool = False):
    """torch/optim/adam.py `_single_tensor_adam` (amsgrad=False, maximize=False); `step` counts from 1.
    Hyper-parameters are Python floats (float64): constants such as 1 - beta2 are formed in float64, THEN rounded."""
    p, g, m, v = (np.asarray(a, dtype=F) for a in (p, g, m, v))
    if decoupled:
        p = p * F(1.0 - lr * weight_decay)                       # AdamW: param.mul_(1 - lr * weight_decay)
    elif weight_decay != 0.0:
        g = g + F(weight_decay) * p                              # grad.add(param, alpha=weight_decay)
    m = m + (g - m) * F(1.0 - beta1)                             # exp_avg.lerp_(grad, 1 - beta1)
    v = v * F(beta2) + F(1.0 - beta2) * g * g                    # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / F(np.sqrt(bc2)) + F(eps)
    p = p - F(lr / bc1) * (m / denom)                            # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(F), m.astype(F), v.astype(F)


def sgd_step(p, g, buf, lr: float, momentum: float = 0.0, weight_decay: float = 0.0):
    """torch/optim/sgd.py `_single_tensor_sgd` (dampening 0, nesterov False); a zero `buf` reproduces the first step
    (torch clones the gradient into the buffer; momentum * 0 + g == g exactly)."""
    p, g, buf = (np.asarray(a, dtype=F) for a in (p, g, buf))
    if weight_decay != 0.0:
        g = g + F(weight_decay) * p
    if momentum != 0.0:
        buf = buf * F(momentum) + g
        g = buf
    return (p - F(lr) * g).astype(F), buf.astype(F)


def rmsprop_step(p, g, sq, lr: float, alpha: float = 0.99, eps: float = 1e-8, weight_decay: float = 0.0):
    """torch/optim/rmsprop.py `_single_tensor_rmsprop` (momentum 0, centered False)."""
    p, g, sq = (np.asarray(a, dtype=F) for a in (p, g, sq))
    if weight_decay != 0.0:
        g = g + F(weight_decay) * p
    sq = sq * F(alpha) + F(1.0 - alpha) * g * g
    avg = np.sqrt(sq) + F(eps)
    return (p - F(lr) * (g / avg)).astype(F), sq.astype(F)
