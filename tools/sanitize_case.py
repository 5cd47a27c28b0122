"""Smallest workload that touches every kernel once (for compute-sanitizer runs; one tool per gpurun call)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import movae_b200  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
J = torch.randn(3, 20_003, generator=g, device=dev)
for name in ("upgrad", "aligned_mtl", "mgda_lgn"):
    agg = movae_b200.make_aggregator(name)
    if isinstance(agg, movae_b200.MGDA):
        agg.set_losses(torch.tensor([0.3, 0.01, 0.2], device=dev))
    out = agg(J)
vq = movae_b200.VectorQuantizer(512, 64).to(dev)
z = (0.5 * torch.randn(3, 64, 10, 13, generator=g, device=dev)).requires_grad_(True)      # 390 rows: 4 tiles, ragged
q, c, e, idx = vq(z)
(q.sum() + c + e).backward()
vq2 = movae_b200.VectorQuantizer(100, 48).to(dev)                                           # exact + atomic paths
z2 = (0.5 * torch.randn(2, 48, 5, 7, generator=g, device=dev)).requires_grad_(True)
q2, c2, e2, idx2 = vq2(z2)
(q2.sum() + c2 + e2).backward()
torch.cuda.synchronize()
print("sanitize case ok", float(out.abs().sum()), int(idx.sum()), float(z.grad.abs().sum()), float(vq.embedding.weight.grad.abs().sum()))
