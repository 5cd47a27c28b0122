"""PCIe probe (dev tool): H2D alone, D2H alone, both at once, plain pinned vs write-combined pinned source."""
import ctypes, os, sys, time
import torch
dev = torch.device("cuda")
rt = ctypes.CDLL("libcudart.so")
n = 1_200_000_000
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(400_000_000, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(400_000_000, dtype=torch.uint8).pin_memory()
h.fill_(1)
wc = ctypes.c_void_p()
assert rt.cudaHostAlloc(ctypes.byref(wc), ctypes.c_size_t(n), ctypes.c_uint(4)) == 0      # cudaHostAllocWriteCombined
ctypes.memset(wc, 1, n)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def h2d_wc():
    rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), wc, ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(s1.cuda_stream))
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def both_wc():
    h2d_wc(); d2h()
for name, fn, nb in (("H2D pinned", h2d, n), ("H2D write-combined", h2d_wc, n), ("D2H", d2h, 400_000_000), ("H2D + D2H concurrently", both, n), ("H2D(WC) + D2H concurrently", both_wc, n)):
    t = timed(fn)
    print(f"{name:28s} {t * 1e3:7.2f} ms  ->  {nb / t / 1e9:6.1f} GB/s (of the H2D bytes)" if "+" in name else f"{name:28s} {t * 1e3:7.2f} ms  ->  {nb / t / 1e9:6.1f} GB/s")
