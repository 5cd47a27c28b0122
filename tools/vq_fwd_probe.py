"""Forward split (search / gather) timing at N = 4.2 M (dev tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200
from movae_b200 import quantizer as Q
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
vq = movae_b200.VectorQuantizer(512, 64).to(dev)
with torch.no_grad():
    vq.embedding.weight.copy_(0.5 * torch.randn(512, 64, generator=g, device=dev))
z = 0.5 * torch.randn(256, 64, 128, 128, generator=g, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
ts = tf = 0.0
for i in range(7):
    flush.fill_(float(i))
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); Q.code_indices(z, vq.embedding.weight, 0); b.record()
    with torch.no_grad():
        vq(z)
    c.record(); torch.cuda.synchronize()
    if i >= 2:
        ts += a.elapsed_time(b); tf += b.elapsed_time(c)
N = z.shape[0] * z.shape[2] * z.shape[3]
print(f"N={N}: search {ts/5:.4f} ms; forward (search+gather) {tf/5:.4f} ms -> gather ~{(tf-ts)/5:.4f} ms = {N*520/((tf-ts)/5)/1e6:.0f} GB/s of 520 B/row")
