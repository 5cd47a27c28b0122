"""In-graph time of the quantizer backward pieces at the BASELINE shapes (dev tool): 20 calls inside one CUDA graph."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
for (B, H, W) in ((128, 8, 8), (256, 16, 16), (64, 64, 64)):
    vq = movae_b200.VectorQuantizer(512, 64).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_(0.5 * torch.randn(512, 64, generator=g, device=dev))
    z = (0.5 * torch.randn(B, 64, H, W, generator=g, device=dev)).requires_grad_(True)
    go = torch.randn(B, 64, H, W, generator=g, device=dev)
    one = torch.ones((), device=dev)
    res = {}
    for name in ("fwd", "fwd+dE", "fwd+dz+dE"):
        def fn():
            out = []
            for _ in range(20):
                q, c, e, idx = vq(z)
                if name == "fwd+dE":
                    out.append(torch.autograd.grad([e], [vq.embedding.weight], grad_outputs=[one]))
                elif name == "fwd+dz+dE":
                    out.append(torch.autograd.grad([q, c, e], [z, vq.embedding.weight], grad_outputs=[go, one, one]))
                else:
                    out.append(q)
            return out
        gs = movae_b200.GraphedStep(fn, warmup=2)
        for _ in range(3):
            gs()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gs()
        b.record()
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 100 * 1e3
        del gs
    print(f"N={B * H * W}: forward {res['fwd']:.1f} us, dE (K6b + reduce) {res['fwd+dE'] - res['fwd']:.1f} us, dz + dE {res['fwd+dz+dE'] - res['fwd']:.1f} us")
