"""Quick GPU probe of the quantizer kernels (run under `timeout`): tensor path vs exact path vs the
reference formula evaluated by torch on the same GPU; prints score error and timing."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import movae_b200  # noqa: E402
from movae_b200 import quantizer as Q  # noqa: E402


def ref_idx(z, E):
    flat = z.permute(0, 2, 3, 1).contiguous().view(-1, z.shape[1])
    dist = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(flat, E.t())
    return torch.argmin(dist, dim=1)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    for name, B, H, W in (("tiny", 2, 8, 8), ("ragged", 3, 7, 9), ("N65536", 16, 64, 64)):
        for cb in ("trained", "init"):
            z = 0.5 * torch.randn(B, 64, H, W, generator=g, device=dev)
            E = 0.5 * torch.randn(512, 64, generator=g, device=dev) if cb == "trained" else \
                (torch.rand(512, 64, generator=g, device=dev) * 2 - 1) / 512
            N = B * H * W
            ie = Q.code_indices(z, E, 1)
            torch.cuda.synchronize()
            print(f"[{name}/{cb}] exact done", flush=True)
            dbg = torch.full((N, 512), float("nan"), device=dev) if N <= 8192 else None
            it = Q.code_indices(z, E, 2, debug_scores=dbg)
            torch.cuda.synchronize()
            nre = Q.rechecked_rows(dev)
            ir = ref_idx(z, E)
            print(f"[{name}/{cb}] N={N} tensor!=exact {(it != ie).sum().item()}  tensor!=torch {(it != ir).sum().item()}  "
                  f"exact!=torch {(ie != ir).sum().item()}  rechecked {nre} ({100.0 * nre / N:.2f}%)", flush=True)
            if dbg is not None:
                flat = z.permute(0, 2, 3, 1).reshape(N, 64).double()
                ex = (E.double() ** 2).sum(1)[None] - 2 * flat @ E.double().t()
                sc = flat.norm(dim=1)[:, None] * E.double().norm(dim=1)[None]
                print(f"    nan {torch.isnan(dbg).sum().item()}  max rel score err {((dbg.double() - ex).abs() / sc).max().item():.3e} "
                      f"(bound {2 ** -14:.3e})", flush=True)
    # timing
    for B, H, W in ((64, 64, 64), (256, 128, 128)):
        z = 0.5 * torch.randn(B, 64, H, W, generator=g, device=dev)
        E = 0.5 * torch.randn(512, 64, generator=g, device=dev)
        N = B * H * W
        for mode, nm in ((2, "tensor"), (1, "exact")):
            if mode == 1 and N > 300000:
                continue
            for _ in range(3):
                Q.code_indices(z, E, mode)
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                Q.code_indices(z, E, mode)
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 10
            print(f"[time] {nm} N={N}: {ms:.3f} ms  {N / ms / 1e6:.2f} Gcodes/s  {N * 65536 / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
        vq = movae_b200.VectorQuantizer(512, 64).to(dev)
        zz = z.clone().requires_grad_(True)
        for _ in range(3):
            q, c, e, i = vq(zz)
            (q.sum() + c + e).backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            q, c, e, i = vq(zz)
            (q.sum() + c + e).backward()
        torch.cuda.synchronize()
        print(f"[time] module fwd+bwd N={N}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
