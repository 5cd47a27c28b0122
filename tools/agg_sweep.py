"""Aggregation microbench grid of BASELINE.json configs[4]: k in {2,3,8} x P in {1e7,1e8,1e9} (one GPU), every
aggregator of north_star; per-kernel CUDA-event timings -> markdown table (profiles/r1_agg_sweep.md).
    python tools/agg_sweep.py [--max-bytes 40e9]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200  # noqa: E402
from movae_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--max-bytes", type=float, default=40e9)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
LOSSES = [0.34, 1e-3, 2.5e-4, 0.17, 2.0]
print(f"| k | P | aggregator | K1 gram ms | K1 GB/s (frac of {peak:.0f}) | K2 solve ms | K3 recombine ms | K3 GB/s (frac) | step ms | step GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for k in (2, 3, 8):
    for P in (10_000_000, 100_000_000, 1_000_000_000):
        if 4.0 * k * P + 4.0 * P > a.max_bytes:
            print(f"| {k} | {P:.0e} | (skipped: {4.0 * (k + 1) * P / 1e9:.0f} GB > --max-bytes) | | | | | | | |")
            continue
        J = torch.empty((k, P), dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev).manual_seed(1234)
        s = torch.logspace(0, -1, k, device=dev)
        for c0 in range(0, P, 1 << 24):
            c = min(1 << 24, P - c0)
            g0 = torch.randn(c, generator=gen, device=dev)
            J[:, c0:c0 + c] = s[:, None] * (0.3 * g0[None] + 0.91 ** 0.5 * torch.randn(k, c, generator=gen, device=dev))
        out = torch.empty(P, dtype=torch.float32, device=dev)
        G = torch.empty((k, k), dtype=torch.float64, device=dev)
        losses = torch.tensor([LOSSES[i % 5] for i in range(k)], device=dev)
        for name in ("upgrad", "aligned_mtl", "aligned_mtl_median", "mgda_ln", "mgda_gn", "mgda_lgn", "jd_sum"):
            agg = movae_b200.make_aggregator(name)
            if isinstance(agg, movae_b200.MGDA):
                agg.set_losses(losses)
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.iters)]
            for i in range(a.iters + 3):
                e = ev[max(i - 3, 0)]
                e[0].record()
                ops.gram(J, out=G)
                e[1].record()
                w = agg.weighting.from_gramian(G)
                e[2].record()
                ops.recombine(J, w, out=out)
                e[3].record()
            torch.cuda.synchronize()
            t1 = sum(e[0].elapsed_time(e[1]) for e in ev) / a.iters
            t2 = sum(e[1].elapsed_time(e[2]) for e in ev) / a.iters
            t3 = sum(e[2].elapsed_time(e[3]) for e in ev) / a.iters
            tt = sum(e[0].elapsed_time(e[3]) for e in ev) / a.iters
            b1, b3 = 4.0 * k * P, 4.0 * (k + 1) * P
            print(f"| {k} | {P:.0e} | {name} | {t1:.4f} | {b1 / t1 / 1e6:.0f} ({b1 / t1 / 1e6 / peak:.3f}) | {t2:.4f} | {t3:.4f} | "
                  f"{b3 / t3 / 1e6:.0f} ({b3 / t3 / 1e6 / peak:.3f}) | {tt:.4f} | {(b1 + b3) / tt / 1e6:.0f} |", flush=True)
            if P >= 1_000_000_000 and name == "upgrad":
                break           # the streaming passes do not depend on the aggregator: one line per (k, P) at 1e9
        del J, out
        torch.cuda.empty_cache()
