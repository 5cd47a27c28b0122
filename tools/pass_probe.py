"""Times the standalone K1 / K3 launches and the fused launch at one (k, P) with CUDA events (dev tool)."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200
from movae_b200 import ops

k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
P = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda")
J = torch.randn(k, P, device=dev)
out = torch.empty(P, device=dev)
G = torch.empty(k, k, dtype=torch.float64, device=dev)
agg = movae_b200.UPGrad()
w = agg.weighting(J)
def t(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t1 = t(lambda: ops.gram(J, out=G))
t3 = t(lambda: ops.recombine(J, w, out=out))
tf = t(lambda: agg.aggregate_into(J, out))
b1, b3 = 4.0 * k * P, 4.0 * (k + 1) * P
print(f"k={k} P={P}: K1 {t1:.4f} ms ({b1/t1/1e6/peak:.3f})  K3 {t3:.4f} ms ({b3/t3/1e6/peak:.3f})  fused {tf:.4f} ms ({(b1+b3)/tf/1e6/peak:.3f})")
