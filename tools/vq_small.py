"""Search latency of the BASELINE quantizer shapes (CUDA-graph replay = GPU time of the kernels) and the big-N rate (dev tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movae_b200
from movae_b200 import quantizer as Q

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(4321)
for name, E in (("trained", 0.5 * torch.randn(512, 64, generator=g, device=dev)),
                ("init", (torch.rand(512, 64, generator=g, device=dev) * 2 - 1) / 512)):
    for (B, H, W) in ((128, 8, 8), (256, 16, 16), (64, 64, 64), (256, 128, 128)):
        N = B * H * W
        z = 0.5 * torch.randn(B, 64, H, W, generator=g, device=dev)
        # 20 searches inside ONE graph: kernel time without the per-replay launch latency (~9 us per graph launch)
        gs = movae_b200.GraphedStep(lambda: [Q.code_indices(z, E, 0) for _ in range(20)], warmup=2)
        for _ in range(3):
            gs()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gs()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 100
        print(f"{name:8s} N={N:8d}: search {ms * 1e3:8.1f} us  {N / ms / 1e6:7.2f} Gcodes/s  {N * 65536 / ms / 1e9:7.1f} TFLOP/s alg  rechecked {Q.rechecked_rows(dev)}")
        del gs
