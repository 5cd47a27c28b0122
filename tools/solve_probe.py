import sys, torch
sys.path.insert(0, "/root/repo")
import movae_b200
from movae_b200 import ops
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "upgrad"
g = torch.Generator(device="cuda").manual_seed(0)
J = torch.randn(k, 100000, generator=g, device="cuda") * torch.logspace(0, -1, k, device="cuda")[:, None]
G = ops.gram(J)
agg = movae_b200.make_aggregator(name)
for _ in range(3):
    w = agg.weighting.from_gramian(G)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    w = agg.weighting.from_gramian(G)
b.record()
torch.cuda.synchronize()
print(name, k, "solve us:", a.elapsed_time(b) / 20 * 1e3, w.tolist())
