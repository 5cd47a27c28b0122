"""Writes profiles/r2_parity.md: the three-deviation table BASELINE.md section 3 asks for (new vs float64 arbiter,
reference float32 vs float64 arbiter, new vs reference float32) per aggregation config, and per quantizer shape the
index mismatches against the reference formula with the fp32-tie rows counted.  Needs a GPU; the reference side is the
CPU oracle (reference files' expressions), so this lives under tests/ (the product never imports oracle/).

    python tests/parity_report.py > profiles/r2_parity.md
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200  # noqa: E402
from oracle import aggregation as oa  # noqa: E402
from oracle import vq as ov  # noqa: E402

LOSSES = [0.34, 1e-3, 2.5e-4, 0.17, 2.0]


def synthetic_J(k, P, seed=1234, decades=1.0, conflict=False):
    """SURVEY 8d generator: row i = s_i (c g0 + sqrt(1 - c^2) g_i), c = 0.3.  `conflict`: c = 0.6 and the shared component enters
    with alternating sign, so that G has negative off-diagonal entries and UPGrad's constraints are active (with the plain
    generator all rows agree and UPGrad answers w = 1 exactly)."""
    g = torch.Generator().manual_seed(seed)
    s = torch.logspace(0, -decades, k)
    cc = 0.6 if conflict else 0.3
    sign = torch.tensor([(-1.0) ** i for i in range(k)]) if conflict else torch.ones(k)
    J = torch.empty(k, P)
    for c0 in range(0, P, 1 << 22):
        c = min(1 << 22, P - c0)
        J[:, c0:c0 + c] = s[:, None] * (cc * sign[:, None] * torch.randn(c, generator=g)[None] + (1 - cc * cc) ** 0.5 * torch.randn(k, c, generator=g))
    return J


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


print("# r2 parity report (tests/parity_report.py, run on B200)\n")
print("Deviations are max |a - b| / max |b|.  *arbiter* = float64-accumulated Gramian / recombination rounded once to float32, small")
print("solve as the reference runs it (oracle/aggregation.py); *reference fp32* = the reference's own float32 `J @ J.T` and")
print("`w @ J` on the CPU (torch, all threads); *new* = the fused CUDA launch.  Contract: new vs arbiter <= rtol 1e-5 on the parity-gated")
print("tier A rows; tiers B and C (row scales over 2 and 3 decades: cond(G) ~ 1e4, 1e6 = the Aligned-MTL rank boundary) are report-only.\n")
print("| config | k | P | aggregator | quantity | new vs arbiter | reference fp32 vs arbiter | new vs reference fp32 |")
print("|---|---|---|---|---|---|---|---|")
cases = [("VAE CIFAR-10 (configs[0])", 2, 1_701_888, "upgrad", 1.0, None), ("VQ-VAE CIFAR-10 (configs[1])", 3, 2_448_064, "aligned_mtl", 1.0, None),
         ("GG-VQ-VAE CelebA (configs[2])", 4, 2_448_064, "mgda_lgn", 1.0, None), ("VQ-VAE2 CelebA-HQ (configs[3])", 3, 651_392, "upgrad", 1.0, None),
         ("microbench (configs[4])", 3, 10_000_000, "upgrad", 1.0, None), ("microbench (configs[4])", 8, 10_000_000, "aligned_mtl", 1.0, None),
         ("microbench, P unaligned", 3, 10_000_003, "aligned_mtl_median", 1.0, None), ("microbench, P unaligned", 2, 10_000_003, "mgda_ln", 1.0, None),
         ("microbench, row 1 all zero", 3, 10_000_000, "upgrad", 1.0, 1), ("microbench, row 1 all zero", 3, 10_000_000, "mgda_gn", 1.0, 1),
         ("microbench tier B (cond 1e4)", 3, 10_000_000, "upgrad", 2.0, None), ("microbench tier B (cond 1e4)", 8, 10_000_000, "aligned_mtl", 2.0, None),
         ("microbench tier C (cond 1e6)", 3, 10_000_000, "upgrad", 3.0, None), ("microbench tier C (cond 1e6)", 3, 10_000_000, "aligned_mtl", 3.0, None),
         ("microbench, conflicting rows", 3, 10_000_000, "upgrad", 1.0, None), ("microbench, conflicting rows", 8, 10_000_000, "upgrad", 1.0, None),
         ("microbench, conflicting rows", 4, 10_000_000, "dualproj", 1.0, None), ("microbench, conflicting rows", 3, 10_000_000, "mgda_gn", 1.0, None)]
for tag, k, P, name, decades, zero_row in cases:
    J = synthetic_J(k, P, decades=decades, conflict="conflicting" in tag)
    if zero_row is not None:
        J[zero_row] = 0.0
    losses = torch.tensor([LOSSES[i % 5] for i in range(k)])
    G_a, w_a, g_a, _ = oa.aggregate(name, J, losses, amtl_dtype=torch.float64)
    G_r, w_r, g_r, _ = oa.aggregate_reference_fp32(name, J, losses)
    agg = movae_b200.make_aggregator(name)
    if isinstance(agg, movae_b200.MGDA):
        agg.set_losses(losses.cuda())
    Jd = J.cuda()
    g_n = agg(Jd).cpu()
    w_n = agg.weighting(Jd).cpu()
    G_n = agg.weighting.last_gramian.cpu()
    for q, n, a, r in (("Gramian", G_n, G_a, G_r), ("weights", w_n, w_a, w_r), ("aggregated gradient", g_n, g_a, g_r)):
        print(f"| {tag} | {k} | {P} | {name} | {q} | {rel(n, a):.2e} | {rel(r, a):.2e} | {rel(n, r):.2e} |")

print("\n## Quantizer: code indices against the reference formula (oracle/vq.py = vq_vae.py:27-39 on the CPU)\n")
print("`fp32-tie rows` = rows whose two best codes are within 8 float32 ulps of the distance magnitude (the reference's own result there")
print("depends on its BLAS accumulation order); mismatches are only tolerated on those rows.\n")
print("| codebook | N | mismatches | of which on fp32-tie rows | fp32-tie rows | rows re-checked exactly in-kernel |")
print("|---|---|---|---|---|---|")
for cb in ("init U(-1/K, 1/K)", "trained-like 0.5 N(0,1)"):
    for (B, H, W) in ((128, 8, 8), (256, 16, 16), (64, 64, 64)):
        g = torch.Generator().manual_seed(B + H)
        z = 0.5 * torch.randn(B, 64, H, W, generator=g)
        E = (torch.rand(512, 64, generator=g) * 2 - 1) / 512 if cb.startswith("init") else 0.5 * torch.randn(512, 64, generator=g)
        idx = movae_b200.code_indices(z.cuda(), E.cuda(), 0).cpu()
        nre = movae_b200.quantizer.rechecked_rows(torch.device("cuda"))
        ref = ov.code_indices(z, E)
        ties = ov.tie_rows(z, E)
        mism = idx != ref
        print(f"| {cb} | {B * H * W} | {int(mism.sum())} | {int((mism & ties).sum())} | {int(ties.sum())} | {nre} |")
