"""Small fixed workload for ncu captures of K7 (fused optimizer step): n float32 parameters, Adam (+ clipping).
    python tools/optim_bench.py [--n 100000000] [--iters 5] [--opt adam|adamw|sgd|rmsprop] [--clip]"""
import argparse
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import movae_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_000)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--opt", default="adam")
ap.add_argument("--clip", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda")
p = torch.nn.Parameter(torch.randn(a.n, device=dev))
kw = dict(momentum=0.9) if a.opt == "sgd" else {}
opt = movae_b200.make_optimizer(a.opt, [p], lr=1e-4, max_grad_norm=1.0 if a.clip else None, **kw)
opt.flat.adopt([p], [torch.randn(a.n, device=dev)])
for _ in range(2):
    opt.step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(a.iters):
    opt.step()
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / a.iters
bpp = {"adam": 28, "adamw": 28, "sgd": 20, "rmsprop": 20}[a.opt] + (4 if a.clip else 0)
print(f"{a.opt}{'+clip' if a.clip else ''} n={a.n}: {ms:.4f} ms  {a.n * bpp / ms / 1e6:.1f} GB/s algorithmic ({bpp} B/param)")
