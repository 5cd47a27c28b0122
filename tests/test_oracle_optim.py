"""CPU tests: the optimizer restatement (oracle/optim.py, the formulas K7 implements) against torch.optim itself -- the
reference's own implementation of this step (main.py:1169-1176, :211-214).  float32 roundings agree to a few ulps
(numpy and ATen contract multiply-adds differently): rtol 1e-6 / atol 1e-7 after 5 steps."""
import numpy as np
import pytest
import torch

from oracle import optim as oo

SHAPES = [(7, 3), (5,), (33,), (2, 3, 4)]


def _data(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in SHAPES]


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (False, 0.05), (True, 0.01)])
def test_adam_restatement_matches_torch(decoupled, wd):
    ps = [torch.nn.Parameter(t.clone()) for t in _data(0)]
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    opt = cls(ps, lr=1e-2, weight_decay=wd)
    mine = [(t.numpy().copy(), np.zeros(t.shape, np.float32), np.zeros(t.shape, np.float32)) for t in _data(0)]
    for step in range(1, 6):
        gs = _data(100 + step)
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        mine = [oo.adam_step(p, g.numpy(), m, v, step, 1e-2, weight_decay=wd, decoupled=decoupled) for (p, m, v), g in zip(mine, gs)]
    for p, (q, m, v) in zip(ps, mine):
        np.testing.assert_allclose(q, p.detach().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m, opt.state[p]["exp_avg"].numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(v, opt.state[p]["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("momentum,wd", [(0.0, 0.0), (0.9, 0.0), (0.9, 0.01)])
def test_sgd_restatement_matches_torch(momentum, wd):
    ps = [torch.nn.Parameter(t.clone()) for t in _data(1)]
    opt = torch.optim.SGD(ps, lr=1e-2, momentum=momentum, weight_decay=wd)
    mine = [(t.numpy().copy(), np.zeros(t.shape, np.float32)) for t in _data(1)]
    for step in range(1, 6):
        gs = _data(200 + step)
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        mine = [oo.sgd_step(p, g.numpy(), b, 1e-2, momentum, wd) for (p, b), g in zip(mine, gs)]
    for p, (q, _) in zip(ps, mine):
        np.testing.assert_allclose(q, p.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_rmsprop_restatement_matches_torch():
    ps = [torch.nn.Parameter(t.clone()) for t in _data(2)]
    opt = torch.optim.RMSprop(ps, lr=1e-3, weight_decay=0.01)
    mine = [(t.numpy().copy(), np.zeros(t.shape, np.float32)) for t in _data(2)]
    for step in range(1, 6):
        gs = _data(300 + step)
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        mine = [oo.rmsprop_step(p, g.numpy(), s, 1e-3, weight_decay=0.01) for (p, s), g in zip(mine, gs)]
    for p, (q, _) in zip(ps, mine):
        np.testing.assert_allclose(q, p.detach().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("max_norm", [0.5, 100.0])
def test_clip_coefficient_matches_clip_grad_norm(max_norm):
    gs = _data(3)
    ps = [torch.nn.Parameter(torch.zeros_like(g)) for g in gs]
    for p, g in zip(ps, gs):
        p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_(ps, max_norm)
    coef = oo.clip_coefficient([g.numpy() for g in gs], max_norm)
    for p, g in zip(ps, gs):
        np.testing.assert_allclose(p.grad.numpy(), g.numpy() * coef, rtol=1e-6, atol=1e-8)
