#!/usr/bin/env python
"""bench.py -- headline benchmark of the MO-VAE hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl movae|reference] [--k 3] [--P 100000000] [--agg upgrad]

Workload (BASELINE.json configs[4], the one the metric is quoted on): aggregation microbench,
k=3 objectives x P=1e8 parameters per GPU, aggregator `upgrad`, synthetic Jacobian (SURVEY 8d recipe).
One "step" = one pass of the hot path over one resident Jacobian: K1 Gramian -> (k x k allreduce when
N > 1) -> K2 solve -> K3 recombine + write-back.  metric = aggregation GB/s = algorithmic bytes
4*P*(2k+1) per step / time, whole job over all N GPUs (weak scaling: P per GPU fixed).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "aggregation_GBps"
UNIT = "GB/s"
LOSSES = [0.34, 1e-3, 2.5e-4, 0.17, 2.0]


def algorithmic_bytes(k: int, P: int) -> dict:
    return {"gram": 4 * k * P, "recombine": 4 * k * P + 4 * P, "step": 4 * P * (2 * k + 1)}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this workload (profiles/r1_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


def synthetic_J_into(J: torch.Tensor, seed: int, chunk: int = 1 << 24) -> None:
    """SURVEY 8d tier A: row i = s_i (0.3 g0 + sqrt(0.91) g_i), s = logspace(0,-1,k); generated on J's device."""
    k, P = J.shape
    gen = torch.Generator(device=J.device).manual_seed(seed)
    s = torch.logspace(0, -1, k, device=J.device)
    for c0 in range(0, P, chunk):
        c = min(chunk, P - c0)
        g0 = torch.randn(c, generator=gen, device=J.device)
        rows = torch.randn(k, c, generator=gen, device=J.device)
        J[:, c0:c0 + c] = s[:, None] * (0.3 * g0[None, :] + 0.91 ** 0.5 * rows)


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (pynvml, ~every 5 ms)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.thread is not None:
            self.thread.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_numa_node(index: int):
    """Multi-GPU runs: pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's own NUMA node (8 ranks x 1.6 GB per step otherwise cross the
    socket interconnect).  Best effort: returns the CPU list or None when NVML / the cgroup does not allow it."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64 + 4
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(k: int, P: int, agg: str, steps: int, warmup: int, budget_s: float):
    """The reference's CPU implementation of the path (oracle port: `J @ J.T` -> solve -> `w @ J` with the
    reference's own float32 torch expressions), all host threads, on a bounded sample of the workload."""
    from oracle import aggregation as oa

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    losses = torch.tensor([LOSSES[i % len(LOSSES)] for i in range(k)])

    def make(Ps):
        s = torch.logspace(0, -1, k)
        J = torch.empty(k, Ps)
        for c0 in range(0, Ps, 1 << 24):
            c = min(1 << 24, Ps - c0)
            J[:, c0:c0 + c] = s[:, None] * (0.3 * torch.randn(c, generator=g)[None] + 0.91 ** 0.5 * torch.randn(k, c, generator=g))
        return J

    Ps = min(P, 10_000_000)
    J = make(Ps)
    t0 = time.perf_counter()
    oa.aggregate_reference_fp32(agg, J, losses)
    per_col = (time.perf_counter() - t0) / Ps
    # size the sample so that warmup+steps fit the budget
    Ps_fit = int(budget_s / max(per_col * (steps + warmup), 1e-12))
    Ps = max(1_000_000, min(P, Ps_fit))
    if Ps != J.shape[1]:
        J = make(Ps)
    for _ in range(warmup):
        oa.aggregate_reference_fp32(agg, J, losses)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oa.aggregate_reference_fp32(agg, J, losses)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    gbps = algorithmic_bytes(k, Ps)["step"] * steps / total / 1e9
    return {"value": gbps, "ms_per_step": 1e3 * total / steps, "cores": cores, "P_sample": Ps,
            "sample": f"k={k} P={Ps} of P={P} ({'full' if Ps == P else 'bounded'} workload), {steps} steps after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU float32, {cores} threads"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.k, args.P, args.agg, args.steps, max(args.warmup, 1), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"aggregation microbench k={args.k} P={args.P} agg={args.agg} (BASELINE.json configs[4])",
                   "timed_on": "host CPU", "sample_P": r["P_sample"]},
        "cpu_baseline": {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
def run_vq(dev, peaks: dict, with_cpu: bool) -> dict:
    """Quantizer leg (K=512, D=64): nearest-codebook search (K4 tcgen05 + exact re-check), gather/loss/STE (K5),
    backward (K6), per kernel group with CUDA events; L2 flushed between iterations for the BASELINE shape
    (its 67 MB of latents would otherwise sit in the 126 MB L2)."""
    import movae_b200
    from movae_b200 import quantizer as Q

    out = {}
    gen = torch.Generator(device=dev).manual_seed(4321)
    E = 0.5 * torch.randn(512, 64, generator=gen, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)       # 256 MB > L2
    vq = movae_b200.VectorQuantizer(512, 64).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    for tag, (B, H, W), iters in (("N8192 (VQ-VAE CIFAR 32x32 b128, BASELINE configs[1])", (128, 8, 8), 10),
                                  ("N65536 (GG-VQ-VAE CelebA 64x64 b256, configs[2]; VQ-VAE2 top codebook, configs[3])", (256, 16, 16), 10),
                                  ("N262144 (VQ-VAE2 256x256 b64 bottom codebook, BASELINE configs[3])", (64, 64, 64), 10),
                                  ("N4194304 (latents 1.07 GB > L2)", (256, 128, 128), 5)):
        N = B * H * W
        z = 0.5 * torch.randn(B, 64, H, W, generator=gen, device=dev)
        zz = z.clone().requires_grad_(True)
        go = torch.randn_like(z)
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        t_search = t_fwd = t_bwd = 0.0
        for i in range(iters + 2):
            flush.fill_(float(i))
            a, b, c, d = ev(), ev(), ev(), ev()
            a.record()
            Q.code_indices(z, E, 0)
            b.record()
            q, commit, embed, idx = vq(zz)
            c.record()
            torch.autograd.backward([q, commit, embed], [go, torch.ones_like(commit), torch.ones_like(embed)])
            d.record()
            torch.cuda.synchronize()
            zz.grad = None
            vq.embedding.weight.grad = None
            if i >= 2:
                t_search += a.elapsed_time(b)
                t_fwd += b.elapsed_time(c)
                t_bwd += c.elapsed_time(d)
        t_search, t_fwd, t_bwd = t_search / iters, t_fwd / iters, t_bwd / iters
        tf = N * 65536 / (t_search * 1e-3) / 1e12
        out[tag] = {
            "search_ms": round(t_search, 4), "codes_per_s": round(N / (t_search * 1e-3), 1),
            "search_tflops_algorithmic": round(tf, 1), "frac_of_bf16_peak_algorithmic": round(tf / peaks["bf16_tflops"], 4),
            "frac_of_bf16_peak_executed": round(tf * 15.0 / 4.0 / peaks["bf16_tflops"], 4),
            "rechecked_rows_frac": round(Q.rechecked_rows(dev) / N, 5),
            "forward_ms(search+gather+loss+ste)": round(t_fwd, 4), "backward_ms(dz+dE)": round(t_bwd, 4),
            "forward_GBps_algorithmic(776B/code)": round(N * 776 / (t_fwd * 1e-3) / 1e9, 1),
            "backward_GBps_algorithmic(776B/code)": round(N * 776 / (t_bwd * 1e-3) / 1e9, 1),
        }
        if N <= 262144:
            # the BASELINE shapes are launch-bound when driven eagerly from Python: the same forward + backward replayed from a
            # CUDA graph (how the train-step harness runs them) is the GPU time of the quantizer kernels themselves
            # fresh leaves: autograd ties a leaf's gradient accumulation to the stream it first ran on, and the tensors used
            # above first ran on the legacy default stream, which a capturing stream must not depend on
            one = torch.ones((), device=dev)
            zg = z.clone().requires_grad_(True)
            vqg = movae_b200.VectorQuantizer(512, 64).to(dev)
            with torch.no_grad():
                vqg.embedding.weight.copy_(E)

            def fwd_bwd():
                q_, c_, e_, _ = vqg(zg)
                return torch.autograd.grad([q_, c_, e_], [zg, vqg.embedding.weight], grad_outputs=[go, one, one])
            gstep = movae_b200.GraphedStep(fwd_bwd, warmup=2)
            for _ in range(3):
                gstep()
            a, b = ev(), ev()
            a.record()
            for _ in range(20):
                gstep()
            b.record()
            torch.cuda.synchronize()
            out[tag]["forward+backward_graph_replay_ms"] = round(a.elapsed_time(b) / 20, 4)
            del gstep, zg, vqg
            # context: the reference's torch expressions (vq_vae.py:28-47: permute, dist[N, K], argmin, one-hot GEMM) on this GPU
            def torch_fwd():
                lat = z.permute(0, 2, 3, 1).contiguous().view(-1, 64)
                dist = torch.sum(lat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(lat, E.t())
                inds = torch.argmin(dist, dim=1).unsqueeze(1)
                onehot = torch.zeros(inds.size(0), 512, device=dev)
                onehot.scatter_(1, inds, 1)
                return torch.matmul(onehot, E), inds
            for _ in range(2):
                torch_fwd()
            a, b = ev(), ev()
            a.record()
            for _ in range(5):
                torch_fwd()
            b.record()
            torch.cuda.synchronize()
            out[tag]["torch_gpu_context_forward_ms(search+one-hot gather, no losses)"] = round(a.elapsed_time(b) / 5, 4)
        del z, zz, go
    # bulk code extraction: search + narrow to int16 + usage bitmap + D2H to pinned memory, 8 batches of
    # N = 262,144 rows (VQ-VAE2 bottom shape), copies overlapped with the next batch's search; end to end incl. the D2H
    zb = [0.5 * torch.randn(64, 64, 64, 64, generator=gen, device=dev) for _ in range(2)]
    ex = movae_b200.CodeExtractor(vq)
    for i in range(2):
        ex.push(zb[i])
    ex.finish()
    ex = movae_b200.CodeExtractor(vq)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(8):
        ex.push(zb[i % 2])
    codes = ex.finish()
    dt = time.perf_counter() - t0
    # the reference's way on the same GPU: int64 indices, synchronous .cpu() per batch (vq_codes_lmdb.py:81-82)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for i in range(8):
        Q.code_indices(zb[i % 2], E, 0).cpu()
    dt_ref = time.perf_counter() - t1
    out["code_extraction"] = {"rows": int(codes.numel()), "batches": 8, "code_dtype": str(codes.dtype),
                              "codes_per_s_e2e": round(codes.numel() / dt, 1), "d2h_bytes_per_batch": 262144 * 2,
                              "usage_percent": round(ex.usage_percentage(), 2),
                              "same_search_int64_sync_cpu_per_batch_codes_per_s": round(codes.numel() / dt_ref, 1)}
    del zb
    out["roofline"] = {"bound": "tensor", "kernel": "vq_argmin_tc_kernel", "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                       "note": "algorithmic = 2*K*D flop per code vector; the kernel executes 15/4 of that (bf16x3 split + 3 key "
                               "steps); search_ms includes the exact re-check kernel"}
    if with_cpu:
        from oracle import vq as ov

        torch.set_num_threads(os.cpu_count() or 1)
        zc = 0.5 * torch.randn(128, 64, 8, 8)
        Ec = E.cpu()
        ov.quantize_forward(zc, Ec)
        t0 = time.perf_counter()
        for _ in range(3):
            ov.quantize_forward(zc, Ec)
        dt = (time.perf_counter() - t0) / 3
        out["cpu_baseline"] = {"value": round(8192 / dt, 1), "unit": "codes/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "N=8192 (VQ-VAE CIFAR b128), forward only, 3 runs, torch CPU float32"}
    return out


# ------------------------------------------------------------------------------------------------
def run_torch_gpu_context(J: torch.Tensor, w: torch.Tensor, flat_grad: torch.Tensor, nbytes: dict, iters: int = 5) -> dict:
    """Context only (SURVEY 8d "stronger bar"): the reference's own torch expressions for the two streaming passes on the
    SAME B200 -- `J @ J.T` (torchjd compute_gramian) and `weights @ J` + the `.grad` copy (WeightedAggregator + Accumulate),
    i.e. cuBLAS with M = N = k.  The small solve is excluded (the reference does it on the host)."""
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for _ in range(2):
        G = J @ J.T
        flat_grad.copy_(w @ J)
    t_g = t_r = 0.0
    for _ in range(iters):
        a, b, c = ev(), ev(), ev()
        a.record()
        G = J @ J.T                                   # noqa: F841
        b.record()
        flat_grad.copy_(w @ J)
        c.record()
        torch.cuda.synchronize()
        t_g += a.elapsed_time(b)
        t_r += b.elapsed_time(c)
    t_g, t_r = t_g / iters, t_r / iters
    return {"what": "torch float32 `J @ J.T` and `w @ J` + copy into the flat gradient on the same GPU (cuBLAS; context, not the baseline arm)",
            "gram_ms": round(t_g, 4), "gram_GBps": round(nbytes["gram"] / (t_g * 1e-3) / 1e9, 1),
            "recombine_ms": round(t_r, 4), "recombine_GBps": round(nbytes["recombine"] / (t_r * 1e-3) / 1e9, 1),
            "two_passes_GBps": round((nbytes["gram"] + nbytes["recombine"]) / ((t_g + t_r) * 1e-3) / 1e9, 1)}


# ------------------------------------------------------------------------------------------------
def run_optim(dev, peaks: dict, n: int = 100_000_000, iters: int = 20) -> dict:
    """K7 leg: fused optimizer step over flat float32 buffers of n parameters, per launch with CUDA
    events; 4 x 400 MB buffers > L2.  Algorithmic bytes per parameter: Adam 16 read + 12 written, with clipping one
    more 4-byte read of the gradients by K1 (k = 1 Gramian = squared norm)."""
    import movae_b200

    out = {}
    p = torch.nn.Parameter(torch.randn(n, device=dev))
    for name, ctor, bpp in (("adam", lambda: movae_b200.Adam([p], lr=1e-4), 28),
                            ("adam+clip_grad_norm", lambda: movae_b200.Adam([p], lr=1e-4, max_grad_norm=1.0), 32),
                            ("sgd_momentum", lambda: movae_b200.SGD([p], lr=1e-4, momentum=0.9), 20)):
        if getattr(p, "_movae_flat", None) is not None:
            del p._movae_flat
        opt = ctor()
        p.grad = None
        opt.flat.adopt([p], [torch.randn(n, device=dev)])
        for _ in range(3):
            opt.step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            opt.step()
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / iters
        gbps = n * bpp / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "algorithmic_bytes_per_param": bpp, "GBps": round(gbps, 1),
                     "frac_of_hbm_peak": round(gbps / peaks["hbm_gbs"], 4)}
        del opt
    out["n_params"] = n
    out["roofline"] = {"bound": "hbm", "kernel": "optim_step_kernel", "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    return out


# ------------------------------------------------------------------------------------------------
def run_movae(args) -> None:
    import torch.distributed as dist

    import movae_b200
    from movae_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (movae_b200 has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the data-parallel train step below captures NCCL collectives into a CUDA graph: the process-group watchdog's
        # asynchronous error handling must be off for that (PyTorch CUDA-graphs notes)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k, P, K, W = args.k, args.P, args.steps, max(args.warmup, 3)
    ld = (P + 3) // 4 * 4
    Jbuf = torch.empty((k, ld), dtype=torch.float32, device=dev)
    J = Jbuf[:, :P]
    synthetic_J_into(J, 1234 + rank)
    losses = torch.tensor([LOSSES[i % len(LOSSES)] for i in range(k)], device=dev)
    agg = movae_b200.make_aggregator(args.agg) if args.agg != "sum" else movae_b200.Sum()
    if isinstance(agg, movae_b200.MGDA):
        agg.set_losses(losses)
    G = torch.zeros((k, k), dtype=torch.float64, device=dev)
    flat_grad = torch.empty(P, dtype=torch.float32, device=dev)
    nbytes = algorithmic_bytes(k, P)

    exchange = None
    if world > 1 and args.exchange == "p2p":
        from movae_b200 import parallel as par
        exchange = par.P2PGramianExchange(dev)          # K1 tail publishes, K2 head gathers: no collective launch
    spec, vec = agg.weighting.solve_spec(k)

    def step(evs=None):
        if evs:
            evs[0].record()
        if exchange is not None:
            seq = exchange.next_seq()
            ops.gram(J, out=G, publish=(exchange.ctx, seq))
            if evs:
                evs[1].record()
            w, _, _ = ops.solve_p2p(exchange.ctx, seq, k, spec, vec, dev)
        else:
            ops.gram(J, out=G)
            if evs:
                evs[1].record()
            if world > 1:
                dist.all_reduce(G)
            w = agg.weighting.from_gramian(G)
        if evs:
            evs[2].record()
        ops.recombine(J, w, out=flat_grad)
        if evs:
            evs[3].record()

    # ---- parity gate before timing: rank-local Gramian and aggregated gradient against torch float64 on a slice
    step()
    torch.cuda.synchronize()
    sl = min(P, 4_000_000)
    Gs = ops.gram(J[:, :sl])
    ref = (J[:, :sl].double() @ J[:, :sl].double().T)
    if not torch.allclose(Gs, ref, rtol=1e-5, atol=1e-6):
        raise RuntimeError("bench parity gate failed: Gramian deviates from the float64 reference")
    if world == 1:
        w_now = agg.weighting.from_gramian(G)
        ref_g = (w_now.double() @ J[:, :sl].double()).float()
        if not torch.allclose(flat_grad[:sl], ref_g, rtol=1e-5, atol=1e-6):
            raise RuntimeError("bench parity gate failed: aggregated gradient deviates from the float64 reference")
    del Gs, ref

    for _ in range(W):
        step()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        t_beg.record()
        for i in range(K):
            step(evs[i])
        t_end.record()
        barrier()
    ms_total = t_beg.elapsed_time(t_end)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_gram = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    ms_solve = sum(e[1].elapsed_time(e[2]) for e in evs) / K
    ms_rec = sum(e[2].elapsed_time(e[3]) for e in evs) / K
    value = world * nbytes["step"] * K / (ms_total * 1e-3) / 1e9

    # ---- e2e: HOST buffers through the C-ABI host pipeline, copies inside the timed region ----
    h_J = torch.empty((k, P), dtype=torch.float32, pin_memory=True)
    h_J.copy_(J)
    h_out = torch.empty(P, dtype=torch.float32, pin_memory=True)
    plan = movae_b200.HostAggregationPlan(k, P, dev)
    from movae_b200 import parallel

    reducer = parallel.gramian_allreduce() if world > 1 else None
    e_steps, e_warm = max(2, min(K, args.e2e_steps)), 2
    for _ in range(e_warm):
        plan.run(h_J, agg, h_out, reducer)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        plan.run(h_J, agg, h_out, reducer)       # synchronous: returns when h_out is complete
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * nbytes["step"] * e_steps / float(te.item()) / 1e9
    # the e2e result must equal the resident-path result
    torch.cuda.synchronize()
    max_dev = float((h_out.to(dev) - flat_grad).abs().max())

    # ---- quantizer, batch-sharded: every rank searches its own shard of the rows, no collective ----
    vq_sharded = None
    if world > 1 and not args.no_vq:
        from movae_b200 import quantizer as Q

        gen = torch.Generator(device=dev).manual_seed(4321 + rank)
        E = 0.5 * torch.randn(512, 64, generator=torch.Generator(device=dev).manual_seed(4321), device=dev)   # replicated codebook
        zq = 0.5 * torch.randn(256, 64, 128, 128, generator=gen, device=dev)                                  # this rank's rows
        for _ in range(3):
            Q.code_indices(zq, E, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(10):
            Q.code_indices(zq, E, 0)
        b.record()
        barrier()
        tq = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        n_local = zq.shape[0] * zq.shape[2] * zq.shape[3]
        vq_sharded = {"rows_per_gpu": n_local, "search_ms_max_over_ranks": round(float(tq.item()), 4),
                      "codes_per_s_whole_job": round(world * n_local / (float(tq.item()) * 1e-3), 1),
                      "sharding": "rows (batch) split across ranks, codebook replicated, no collective on the forward"}
        del zq

    # ---- data-parallel train step (SURVEY 8e "real DP training"): all ranks ------------------------------------------
    train_dp = None
    if world > 1 and not args.no_vq:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from vqvae_harness import time_dp_train_steps

        train_dp = time_dp_train_steps(dev, rank, world)

    if rank == 0:
        peaks = measured_peaks()
        dominant = "recombine" if ms_rec >= ms_gram else "gram"
        dom_ms = ms_rec if dominant == "recombine" else ms_gram
        achieved = nbytes[dominant] / (dom_ms * 1e-3) / 1e9
        kernels = {
            "gram_kernel(K1)": {"ms": round(ms_gram, 4), "algorithmic_bytes": nbytes["gram"],
                                "GBps": round(nbytes["gram"] / (ms_gram * 1e-3) / 1e9, 1),
                                "frac": round(nbytes["gram"] / (ms_gram * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
            "solve_kernel(K2)" + ("+exchange" if world > 1 else ""): {"ms": round(ms_solve, 4)},
            "recombine_kernel(K3)": {"ms": round(ms_rec, 4), "algorithmic_bytes": nbytes["recombine"],
                                     "GBps": round(nbytes["recombine"] / (ms_rec * 1e-3) / 1e9, 1),
                                     "frac": round(nbytes["recombine"] / (ms_rec * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
        }
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"aggregation microbench k={k} P={P} per GPU agg={args.agg} (BASELINE.json configs[4])",
                       "global_P": world * P, "sharding": (f"P-sharded x{world}, k*k float64 Gramian exchange per step: " +
                                    ("fused into K1 tail / K2 head over NVLink peer memory" if exchange is not None else "NCCL all_reduce"))
                       if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2 ({nbytes['gram'] / 1e6:.0f} MB Jacobian + {4 * P / 1e6:.0f} MB output vs 126 MB L2), no flush needed",
                       "layout": f"J float32 [k, ldJ={ld}] resident in HBM, flat float32 grad [P]"},
            "roofline": {"bound": "hbm", "kernel": f"{dominant}_kernel", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": ncu_traffic(f"{dominant}_kernel"),
                         "peak_source": peaks["source"], "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                         "kernels": kernels},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": 4 * k * P, "d2h_bytes_per_step": 4 * P,
                    "steps": e_steps, "host_cpus_bound_to_gpu_numa_node": (len(numa) if numa else None),
                    "api": "movae_b200.HostAggregationPlan.run (movae_host_gram_f32 -> movae_solve -> movae_host_recombine_f32), pinned host buffers",
                    "max_abs_dev_vs_resident": max_dev},
            "gpu_launches": 3 * K,
            "clocks": clocks.summary(),
        }
        if vq_sharded is not None:
            line["vq_sharded"] = vq_sharded
        if train_dp is not None:
            line["train_step_data_parallel"] = {
                "workload": "VQ-VAE CIFAR-10 32x32 (BASELINE.json configs[1]) data-parallel over the GPUs, batch 128 per GPU, agg=aligned_mtl; "
                            "movae_b200.parallel.DataParallel: Jacobian rows reduce-scattered, K1/K3 on column shards, Gramian all_reduce, "
                            "aggregated gradient all-gathered; `graph` = the whole step incl. the NCCL collectives replayed from one CUDA graph",
                **train_dp}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(k, P, args.agg, steps=3, warmup=1, budget_s=20.0)
            line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        if not args.no_vq:
            if world == 1:
                line["torch_gpu_context"] = run_torch_gpu_context(J, agg.weighting.from_gramian(G), flat_grad, nbytes)
                step()                                # leave flat_grad as the product path wrote it
            line["optim"] = run_optim(dev, peaks)
            line["vq"] = run_vq(dev, peaks, with_cpu=(world == 1 and not args.no_cpu_baseline))
            # gpu_launches stays the count of OUR kernels inside the timed region (K1, K2, K3 per step); the legs below
            # (optimizer, quantizer, train steps) launch their own kernels outside it
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from vqvae_harness import (time_ggvqvae_train_steps, time_train_steps, time_vae_train_steps,
                                       time_vqvae2_train_steps)

            shell = ("model shell = tools/vqvae_harness.py (torch.nn convs, cuDNN); quantizer + mtl_backward + aggregator + "
                     "fused Adam = movae_b200; arms: eager launches / whole step replayed from one CUDA graph / graph + "
                     "per-step H2D of the batch and D2H of the losses; torch_sum_* = context (reference torch quantizer, "
                     "total_loss.backward(), torch Adam)")
            line["train_step"] = {
                "workload": "VQ-VAE CIFAR-10 32x32, K=512, D=64, hidden [128,256], agg=aligned_mtl, batch 128, synthetic data "
                            "(BASELINE.json configs[1]); " + shell,
                **time_train_steps(dev)}
            line["train_step_vae"] = {
                "workload": "VAE CIFAR-10 32x32, latent 128, hidden [32,64,128,256,512], agg=upgrad, batch 128, synthetic data "
                            "(BASELINE.json configs[0])",
                **time_vae_train_steps(dev)}
            line["train_step_ggvqvae"] = {
                "workload": "GG-VQ-VAE (v1, k=4) CelebA 64x64, agg=mgda_lgn, batch 256, synthetic data (BASELINE.json configs[2])",
                **time_ggvqvae_train_steps(dev)}
            line["train_step_vqvae2"] = {
                "workload": "VQ-VAE2 CelebA-HQ 256x256 (top + bottom codebooks), agg=upgrad, batch 64, synthetic data "
                            "(BASELINE.json configs[3])",
                **time_vqvae2_train_steps(dev)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL kernels are still alive: tearing the communicator down under them can block, and
        # there is nothing left to do -- synchronise, flush and leave without destroy_process_group()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="movae", choices=["movae", "reference"])
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--P", type=int, default=100_000_000)
    ap.add_argument("--agg", default="upgrad")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU Gramian exchange: fused peer-memory exchange (default) or a NCCL all_reduce launch")
    ap.add_argument("--no-vq", action="store_true", help="skip the quantizer leg (rank 0 only, after the aggregation timing)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_movae(args)


if __name__ == "__main__":
    main()
