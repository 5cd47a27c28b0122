"""GPU parity tests of K7 (fused optimizer step on the flat buffers), the FlatParameters layout the aggregation
path writes into, and CUDA-graph capture of a whole train step.

Oracle: the reference calls torch.optim directly (/root/reference/main.py:1169-1176, :211-214), so the checker is
torch.optim.{SGD,Adam,AdamW,RMSprop} + torch.nn.utils.clip_grad_norm_ themselves, run on the CPU in float32 on the
same parameters and gradients.  Tolerance: parameters within rtol 2e-6 / atol 2e-7 after 6 steps (parameters and
per-step updates are O(1), so one float32 rounding of a term is ~1e-7 absolute; roundings happen
in a different FMA contraction order), moment buffers within rtol 1e-5 / atol 1e-7 (gradients are O(0.1 .. 10)).
"""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

SHAPES = [(7, 3, 3, 3), (7,), (33, 5), (1,), (64, 64), (130,), (2, 3, 5, 7)]      # numels mostly NOT multiples of 4


@pytest.fixture(scope="module")
def mv():
    import movae_b200
    return movae_b200


def make_params(seed, device):
    g = torch.Generator().manual_seed(seed)
    return [nn.Parameter(torch.randn(s, generator=g).to(device)) for s in SHAPES]


def grads_for(step, seed=77):
    g = torch.Generator().manual_seed(seed + step)
    return [torch.randn(s, generator=g) * (10.0 ** ((i % 3) - 1)) for i, s in enumerate(SHAPES)]


def run_pair(mv, name, kwargs, torch_cls, torch_kwargs, steps=6, max_norm=None, skip=None):
    dev = torch.device("cuda")
    ours = make_params(5, dev)
    ref = make_params(5, "cpu")
    opt = getattr(mv, name)(ours, max_grad_norm=max_norm, **kwargs)
    ropt = torch_cls(ref, **torch_kwargs)
    for s in range(steps):
        gs = grads_for(s)
        opt.zero_grad()
        ropt.zero_grad()
        for i, (p, r, g) in enumerate(zip(ours, ref, gs)):
            if skip is not None and i == skip:
                continue
            p.grad = g.to(dev)
            r.grad = g.clone()
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt.step()
        ropt.step()
    return opt, ropt, ours, ref


def check_params(ours, ref):
    for p, r in zip(ours, ref):
        torch.testing.assert_close(p.detach().cpu(), r.detach(), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("wd", [0.0, 0.05])
def test_adam_matches_torch(mv, wd):
    opt, ropt, ours, ref = run_pair(mv, "Adam", dict(lr=1e-2, weight_decay=wd), torch.optim.Adam, dict(lr=1e-2, weight_decay=wd))
    check_params(ours, ref)
    for p, r in zip(ours, ref):
        torch.testing.assert_close(opt.state[p]["exp_avg"].cpu(), ropt.state[r]["exp_avg"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(opt.state[p]["exp_avg_sq"].cpu(), ropt.state[r]["exp_avg_sq"], rtol=1e-5, atol=1e-7)
    assert opt.step_count == 6 and opt.kernel_launches == 1


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adamw_matches_torch(mv, wd):
    _, _, ours, ref = run_pair(mv, "AdamW", dict(lr=1e-2, weight_decay=wd), torch.optim.AdamW, dict(lr=1e-2, weight_decay=wd))
    check_params(ours, ref)


@pytest.mark.parametrize("momentum,wd", [(0.0, 0.0), (0.9, 0.0), (0.9, 0.01)])
def test_sgd_matches_torch(mv, momentum, wd):
    _, _, ours, ref = run_pair(mv, "SGD", dict(lr=1e-2, momentum=momentum, weight_decay=wd), torch.optim.SGD,
                               dict(lr=1e-2, momentum=momentum, weight_decay=wd))
    check_params(ours, ref)


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_rmsprop_matches_torch(mv, wd):
    _, _, ours, ref = run_pair(mv, "RMSprop", dict(lr=1e-3, weight_decay=wd), torch.optim.RMSprop, dict(lr=1e-3, weight_decay=wd))
    check_params(ours, ref)


@pytest.mark.parametrize("max_norm", [0.5, 1e6])
def test_clipping_matches_clip_grad_norm(mv, max_norm):
    """main.py:211-212: clip_grad_norm_ then step; here the norm is one K1 launch and the scaling is fused into K7."""
    opt, _, ours, ref = run_pair(mv, "Adam", dict(lr=1e-2), torch.optim.Adam, dict(lr=1e-2), max_norm=max_norm)
    check_params(ours, ref)
    gs = grads_for(5)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs))
    torch.testing.assert_close(opt.global_grad_norm_sq().cpu().sqrt().reshape(()), total, rtol=1e-6, atol=0)


def test_parameters_without_grad_are_skipped(mv):
    """torch optimizers skip parameters whose .grad is None: two launches (one per run), ONE optimizer step."""
    opt, _, ours, ref = run_pair(mv, "Adam", dict(lr=1e-2), torch.optim.Adam, dict(lr=1e-2), skip=2)
    check_params(ours, ref)
    assert opt.kernel_launches == 2 and opt.step_count == 6
    torch.testing.assert_close(ours[2].detach().cpu(), make_params(5, "cpu")[2].detach(), rtol=0, atol=0)


def test_lr_scheduler_contract(mv):
    """main.py:1179-1188 schedulers write param_groups[0]['lr']; the kernel reads the device copy."""
    dev = torch.device("cuda")
    ours, ref = make_params(9, dev), make_params(9, "cpu")
    opt, ropt = mv.SGD(ours, lr=0.1), torch.optim.SGD(ref, lr=0.1)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    rsched = torch.optim.lr_scheduler.ExponentialLR(ropt, gamma=0.5)
    for s in range(4):
        for p, r, g in zip(ours, ref, grads_for(s)):
            p.grad, r.grad = g.to(dev), g.clone()
        opt.step()
        ropt.step()
        sched.step()
        rsched.step()
    check_params(ours, ref)


def test_state_dict_round_trip_and_torch_compatibility(mv):
    """main.py:1407 checkpoints optimizer.state_dict(): ours must reload into a fresh fused optimizer AND be interchangeable with
    torch.optim.Adam's (same keys), continuing the trajectory in both directions."""
    dev = torch.device("cuda")
    ours, ref = make_params(21, dev), make_params(21, dev)
    opt, ropt = mv.Adam(ours, lr=1e-2, weight_decay=0.01), torch.optim.Adam(ref, lr=1e-2, weight_decay=0.01)
    for s in range(3):
        for p, r, g in zip(ours, ref, grads_for(s)):
            p.grad, r.grad = g.to(dev), g.to(dev)
        opt.step()
        ropt.step()
    # deep copies stand in for torch.save / torch.load (state_dict() returns references, and torch's load_state_dict keeps
    # tensors whose dtype and device already match: without the copy the two optimizers would share moment buffers)
    sd_ours, sd_torch = copy.deepcopy(opt.state_dict()), copy.deepcopy(ropt.state_dict())
    assert set(sd_ours["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    # (a) torch's checkpoint into a fresh fused optimizer, (b) ours into a fresh torch optimizer; then 3 more steps each
    ours2 = [nn.Parameter(r.detach().clone()) for r in ref]
    opt2 = mv.Adam(ours2, lr=1e-2, weight_decay=0.01)
    opt2.load_state_dict(sd_torch)
    ref2 = [nn.Parameter(p.detach().clone()) for p in ours]
    ropt2 = torch.optim.Adam(ref2, lr=1e-2, weight_decay=0.01)
    ropt2.load_state_dict(sd_ours)
    assert opt2.step_count == 3
    for s in range(3, 6):
        for p, r, p2, r2, g in zip(ours, ref, ours2, ref2, grads_for(s)):
            p.grad, r.grad, p2.grad, r2.grad = g.to(dev), g.to(dev), g.to(dev), g.to(dev)
        for o in (opt, ropt, opt2, ropt2):
            o.step()
    for p, r, p2, r2 in zip(ours, ref, ours2, ref2):
        torch.testing.assert_close(p2.detach(), r.detach(), rtol=2e-6, atol=2e-7)
        torch.testing.assert_close(r2.detach(), p.detach(), rtol=2e-6, atol=2e-7)
        torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-6, atol=2e-7)
    assert opt2.step_count == 6 and opt.flat.grad_state(ours[0]) == "view"


def test_flat_layout_and_module_views(mv):
    dev = torch.device("cuda")
    net = nn.Sequential(nn.Conv2d(3, 5, 3), nn.Linear(7, 3)).to(dev)
    before = [p.detach().clone() for p in net.parameters()]
    flat = mv.FlatParameters(net.parameters())
    assert all(o % 4 == 0 for o in flat.offsets) and flat.total % 4 == 0
    for p, b, o in zip(net.parameters(), before, flat.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == flat.flat_param.data_ptr() + 4 * o
    with pytest.raises(RuntimeError):
        mv.FlatParameters(net.parameters())                   # a parameter belongs to one layout only
    with pytest.raises(RuntimeError, match="CUDA"):
        mv.FlatParameters(nn.Linear(2, 2).parameters())       # no CPU fallback
    # plain loss.backward() gradients are adopted into the flat buffer by the optimizer
    opt = mv.SGD(flat, lr=0.5)
    x = torch.randn(2, 3, 9, 9, device=dev)
    net[0](x).square().mean().backward()
    g0 = net[0].weight.grad.clone()
    opt.step()
    assert flat.grad_state(net[0].weight) == "view"
    torch.testing.assert_close(net[0].weight.detach(), before[0] - 0.5 * g0, rtol=1e-6, atol=1e-7)
    assert net[1].weight.grad is None and torch.equal(net[1].weight.detach(), before[2])


class TinyVQ(nn.Module):
    def __init__(self, mv):
        super().__init__()
        self.encoder = nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1), nn.LeakyReLU(), nn.Conv2d(16, 64, 3, 2, 1))
        self.vq_layer = mv.VectorQuantizer(512, 64)
        self.decoder = nn.Sequential(nn.ConvTranspose2d(64, 16, 4, 2, 1), nn.LeakyReLU(), nn.ConvTranspose2d(16, 3, 4, 2, 1))

    def forward(self, x):
        enc = self.encoder(x)
        q, commit, embed, _ = self.vq_layer(enc)
        rec = self.decoder(q)
        return enc, [torch.nn.functional.mse_loss(rec, x), embed, 0.25 * commit]


@pytest.mark.parametrize("agg", ["upgrad", "aligned_mtl", "mgda_lgn"])
def test_mtl_backward_into_flat_buffers_matches_plain_path(mv, agg):
    """Same model twice: (a) ordinary parameters + torch.optim.Adam, (b) FlatParameters + fused Adam.  J's column order
    differs (aligned flat layout vs discovery order), so Gramian sums round differently: gradients agree to rtol 1e-5."""
    dev = torch.device("cuda")
    torch.manual_seed(0)
    a = TinyVQ(mv).to(dev)
    b = copy.deepcopy(a)
    flat = mv.FlatParameters(b.parameters())
    x = torch.rand(8, 3, 16, 16, device=dev) * 2 - 1
    outs = []
    for net in (a, b):
        aggr = mv.make_aggregator(agg)
        enc, losses = net(x)
        if isinstance(aggr, mv.MGDA):
            aggr.set_losses(torch.stack([l.detach() for l in losses]))
        mv.mtl_backward(losses=losses, features=[enc], aggregator=aggr, retain_graph=True)
        outs.append([p.grad.clone() for p in net.parameters()])
    for p in b.parameters():
        assert flat.grad_state(p) == "view"                   # K3 / adopt wrote into the flat gradient buffer
    for ga, gb in zip(*outs):
        torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-7)
    # second call accumulates (torchjd Accumulate: `+=` when .grad exists)
    enc, losses = b(x)
    aggr = mv.make_aggregator(agg)
    if isinstance(aggr, mv.MGDA):
        aggr.set_losses(torch.stack([l.detach() for l in losses]))
    mv.mtl_backward(losses=losses, features=[enc], aggregator=aggr, retain_graph=True)
    for gb, p in zip(outs[1], b.parameters()):
        torch.testing.assert_close(p.grad, 2 * gb, rtol=1e-5, atol=1e-7)


def test_zero_row_is_not_backpropagated(mv):
    """embedding_loss never reaches the encoder (vq_vae.py:52): its Jacobian row is identically zero WITHOUT a backward
    pass -- every entry of that row in the segment table points at the shared zero buffer."""
    from movae_b200 import autojac

    dev = torch.device("cuda")
    torch.manual_seed(1)
    net = TinyVQ(mv).to(dev)
    x = torch.rand(4, 3, 16, 16, device=dev)
    enc, losses = net(x)
    agg = mv.UPGrad()
    seen = {}
    real = agg.aggregate_segments_into
    agg.aggregate_segments_into = lambda rows, *a, **k: (seen.update(rows=rows), real(rows, *a, **k))[1]
    mv.mtl_backward(losses=losses, features=[enc], aggregator=agg, retain_graph=True)
    zero = autojac._ZEROS[dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())]
    rows = seen["rows"]
    assert len(rows) == 3 and all(t.data_ptr() == zero.data_ptr() for t in rows[1])
    assert all(t.data_ptr() != zero.data_ptr() and float(t.abs().max()) > 0 for i in (0, 2) for t in rows[i])
    assert float(zero.abs().max()) == 0.0


def test_graphed_train_step_matches_eager(mv):
    """GraphedStep replays zero_grad -> forward -> mtl_backward -> fused Adam (with clipping) from one CUDA graph."""
    dev = torch.device("cuda")
    torch.manual_seed(3)
    nets = [TinyVQ(mv).to(dev)]
    nets.append(copy.deepcopy(nets[0]))
    x = torch.rand(8, 3, 16, 16, device=dev) * 2 - 1
    results = []
    for mode, net in zip(("eager", "graph"), nets):
        opt = mv.Adam(net.parameters(), lr=1e-3, max_grad_norm=1.0)
        aggr = mv.make_aggregator("upgrad")

        def step():
            opt.zero_grad()
            enc, losses = net(x)
            mv.mtl_backward(losses=losses, features=[enc], aggregator=aggr, retain_graph=True)
            opt.step()
            return torch.stack([l.detach() for l in losses])

        if mode == "eager":
            for _ in range(6):
                out = step()
        else:
            g = mv.GraphedStep(step, warmup=2)             # 2 eager warm-up steps + 1 captured (capture does not execute)
            for _ in range(4):
                out = g()
            assert g.replays == 4
        torch.cuda.synchronize()
        results.append(([p.detach().clone() for p in net.parameters()], out.clone(), opt.step_count))
    (pa, la, sa), (pb, lb, sb) = results
    assert sa == 6 and sb == 6
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-7)


def test_fused_step_bumps_autograd_versions():
    """K7 writes the parameters through raw pointers; a backward pass over a graph retained from before the step
    must raise torch's in-place-modification error exactly as it does after torch.optim's step."""
    import movae_b200 as mv

    lin = torch.nn.Sequential(torch.nn.Linear(8, 6), torch.nn.Tanh(), torch.nn.Linear(6, 4)).cuda()   # layer 2's weight is saved for backward
    opt = mv.SGD(lin.parameters(), lr=0.1)
    x = torch.randn(5, 8, device="cuda")
    y = lin(x).pow(2).sum()
    y.backward(retain_graph=True)
    v0 = [p._version for p in lin.parameters()]
    opt.step()
    assert all(p._version > v for p, v in zip(lin.parameters(), v0))
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        y.backward()
