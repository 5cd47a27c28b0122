"""Small fixed workload for ncu captures of the quantizer kernels: N = B*H*W code vectors, K=512, D=64.
    python tools/vq_bench.py [--B 256 --H 128 --W 128] [--iters 5] [--module]"""
import argparse
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import movae_b200  # noqa: E402
from movae_b200 import quantizer as Q  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--H", type=int, default=128)
ap.add_argument("--W", type=int, default=128)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--codebook", default="trained")
ap.add_argument("--module", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
z = 0.5 * torch.randn(a.B, 64, a.H, a.W, generator=g, device=dev)
E = 0.5 * torch.randn(512, 64, generator=g, device=dev) if a.codebook == "trained" else (torch.rand(512, 64, generator=g, device=dev) * 2 - 1) / 512
N = a.B * a.H * a.W
for _ in range(2):
    Q.code_indices(z, E, 2)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(a.iters):
    Q.code_indices(z, E, 2)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / a.iters
print(f"tensor search N={N}: {ms:.4f} ms  {N / ms / 1e6:.2f} Gcodes/s  {N * 65536 / ms / 1e9:.1f} TFLOP/s algorithmic  rechecked {Q.rechecked_rows(dev)}")
if a.module:
    vq = movae_b200.VectorQuantizer(512, 64).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    zz = z.clone().requires_grad_(True)
    for _ in range(2):
        q, c, e, i = vq(zz)
        (q.sum() + c + e).backward()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(a.iters):
        q, c, e, i = vq(zz)
        (q.sum() + c + e).backward()
    t1.record()
    torch.cuda.synchronize()
    print(f"module fwd+bwd N={N}: {t0.elapsed_time(t1) / a.iters:.4f} ms")
