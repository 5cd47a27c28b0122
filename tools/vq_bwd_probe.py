"""Backward (K6a dz + K6b dE + reduce) timing at one shape (dev tool): python tools/vq_bwd_probe.py [B H W]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 128, 128)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
vq = movae_b200.VectorQuantizer(512, 64).to(dev)
with torch.no_grad():
    vq.embedding.weight.copy_(0.5 * torch.randn(512, 64, generator=g, device=dev))
z = (0.5 * torch.randn(B, 64, H, W, generator=g, device=dev)).requires_grad_(True)
go = torch.randn(B, 64, H, W, generator=g, device=dev)
one = torch.ones((), device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
tot = {"dE_only": 0.0, "dz+dE": 0.0}
iters = 5
for i in range(iters + 2):
    q, c, e, idx = vq(z)
    flush.fill_(float(i))
    a, b, d = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    gE, = torch.autograd.grad([e], [vq.embedding.weight], grad_outputs=[one], retain_graph=True)     # dE only
    b.record()
    gz, gE2 = torch.autograd.grad([q, c, e], [z, vq.embedding.weight], grad_outputs=[go, one, one])
    d.record()
    torch.cuda.synchronize()
    if i >= 2:
        tot["dE_only"] += a.elapsed_time(b)
        tot["dz+dE"] += b.elapsed_time(d)
N = B * H * W
print(f"N={N}: dE-only backward {tot['dE_only'] / iters:.4f} ms ({N * 264 / (tot['dE_only'] / iters) / 1e6:.0f} GB/s of 264 B/row), dz+dE {tot['dz+dE'] / iters:.4f} ms")
