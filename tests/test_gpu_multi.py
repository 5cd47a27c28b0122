"""Multi-GPU tests (need >= 2 CUDA devices; skipped on a single-GPU box): P-sharded aggregation through
(a) the NCCL all_reduce reducer (three launches) and (b) the exchange inside the fused aggregation kernel over peer
memory must both equal the single-GPU aggregation of the full Jacobian, with bit-identical weights on all ranks --
eagerly and replayed from a CUDA graph; the batch-sharded quantizer equals the single-GPU one; data-parallel
training (all aggregators incl. COMFORT / PNUPGrad) equals single-process training on the whole batch.
`profiles/r2_multi_gpu_tests.log` keeps the log of a passing 2-GPU run."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_J(k, P, seed=7):
    g = torch.Generator().manual_seed(seed)
    s = torch.logspace(0, -1, k)
    return s[:, None] * (0.3 * torch.randn(P, generator=g)[None] + 0.91 ** 0.5 * torch.randn(k, P, generator=g))


def _worker(rank, world, port, k, P, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import movae_b200
        from movae_b200 import parallel

        J = _make_J(k, P)
        lo, hi = parallel.shard_columns(P, rank, world)
        Jl = J[:, lo:hi].contiguous().to(dev)
        losses = torch.tensor(([0.34, 1e-3, 2.5e-4, 0.17, 2.0] * 2)[:k], device=dev)
        res = {}
        for name in ("upgrad", "aligned_mtl", "mgda_lgn"):
            for mode in ("nccl", "p2p"):
                agg = movae_b200.make_aggregator(name)
                if isinstance(agg, movae_b200.MGDA):
                    agg.set_losses(losses)
                ex = None
                if mode == "nccl":
                    parallel.install_gramian_allreduce(agg)
                else:
                    ex = parallel.install_p2p_gramian_exchange(agg, dev)
                for _ in range(3):                      # several steps: exercises the seq / parity double-buffering
                    g = agg(Jl)
                w = agg.weighting(Jl)
                assert parallel.check_replicated(w)
                assert float(agg.weighting.last_diag[4]) == 0.0
                res[f"{name}:{mode}"] = {"g": g.cpu(), "w": w.cpu(), "G": agg.weighting.last_gramian.cpu()}
                torch.cuda.synchronize()
                if ex is not None:
                    dist.barrier()
                    ex.close()
        torch.save({"lo": lo, "hi": hi, "res": res}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,P", [(3, 1_000_003), (8, 40_000)])
def test_sharded_aggregation_matches_single_gpu(tmp_path, k, P):
    import torch.multiprocessing as mp

    import movae_b200

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, P, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    J = _make_J(k, P).cuda()
    losses = torch.tensor(([0.34, 1e-3, 2.5e-4, 0.17, 2.0] * 2)[:k], device="cuda")
    for name in ("upgrad", "aligned_mtl", "mgda_lgn"):
        agg = movae_b200.make_aggregator(name)
        if isinstance(agg, movae_b200.MGDA):
            agg.set_losses(losses)
        g_ref = agg(J).cpu().numpy()
        w_ref = agg.weighting(J).cpu().numpy()
        for mode in ("nccl", "p2p"):
            key = f"{name}:{mode}"
            assert torch.equal(parts[0]["res"][key]["w"], parts[1]["res"][key]["w"])          # replicated solve
            assert torch.equal(parts[0]["res"][key]["G"], parts[1]["res"][key]["G"])
            np.testing.assert_allclose(parts[0]["res"][key]["w"].numpy(), w_ref, rtol=1e-5, atol=1e-6, err_msg=key)
            g = torch.cat([p["res"][key]["g"] for p in parts]).numpy()
            np.testing.assert_allclose(g, g_ref, rtol=1e-5, atol=1e-6, err_msg=key)
        # both exchanges sum the same two partials: identical Gramian
        np.testing.assert_allclose(parts[0]["res"][f"{name}:p2p"]["G"].numpy(), parts[0]["res"][f"{name}:nccl"]["G"].numpy(), rtol=1e-14)


def _graph_worker(rank, world, port, k, P, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import movae_b200
        from movae_b200 import parallel

        lo, hi = parallel.shard_columns(P, rank, world)
        Jl = _make_J(k, P, 7)[:, lo:hi].contiguous().to(dev)
        out = torch.zeros(hi - lo, device=dev)
        agg = movae_b200.make_aggregator("upgrad")
        ex = parallel.install_p2p_gramian_exchange(agg, dev)
        step = movae_b200.GraphedStep(lambda: agg.aggregate_into(Jl, out), warmup=2)     # exchange captured with the launch
        res = []
        for seed in (8, 9, 10, 11):
            Jl.copy_(_make_J(k, P, seed)[:, lo:hi])
            ex.barrier()                                                                # device-side barrier kernel
            w = step()
            torch.cuda.synchronize()
            assert parallel.check_replicated(w)
            res.append({"w": w.cpu().clone(), "g": out.cpu().clone(), "G": agg.weighting.last_gramian.cpu().clone(),
                        "status": float(agg.weighting.last_diag[4])})
        assert not ex.barrier_failed()
        torch.save(res, os.path.join(out_dir, f"g{rank}.pt"))
        dist.barrier()
        del step
        ex.close()
    finally:
        dist.destroy_process_group()


def test_p2p_exchange_replays_from_a_cuda_graph(tmp_path):
    """the exchange's sequence number was a host-incremented kernel argument, so a replayed graph summed
    stale peer slots.  It now lives in the exchange buffer: the captured launch must stay correct step after step."""
    import torch.multiprocessing as mp

    import movae_b200

    world, k, P = 2, 3, 600_004
    mp.spawn(_graph_worker, args=(world, _free_port(), k, P, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"g{r}.pt")) for r in range(world)]
    for i, seed in enumerate((8, 9, 10, 11)):
        J = _make_J(k, P, seed).cuda()
        agg = movae_b200.make_aggregator("upgrad")
        g_ref = agg(J).cpu().numpy()
        w_ref = agg.weighting(J).cpu().numpy()
        assert torch.equal(parts[0][i]["w"], parts[1][i]["w"]) and torch.equal(parts[0][i]["G"], parts[1][i]["G"])
        assert parts[0][i]["status"] == 0.0 and parts[1][i]["status"] == 0.0
        np.testing.assert_allclose(parts[0][i]["G"].numpy(), agg.weighting.last_gramian.cpu().numpy(), rtol=1e-8)   # other float32 chain boundaries
        np.testing.assert_allclose(parts[0][i]["w"].numpy(), w_ref, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(torch.cat([p[i]["g"] for p in parts]).numpy(), g_ref, rtol=1e-5, atol=1e-6)


def _vq_worker(rank, world, port, out_dir):
    torch.cuda.set_device(rank)
    from movae_b200 import quantizer as Q

    g = torch.Generator().manual_seed(3)
    E = 0.5 * torch.randn(512, 64, generator=g)
    z = 0.5 * torch.randn(64, 64, 16, 16, generator=g)
    per = z.shape[0] // world
    idx = Q.code_indices(z[rank * per:(rank + 1) * per].cuda(), E.cuda(), 0)            # this rank's rows, replicated codebook
    torch.save(idx.cpu(), os.path.join(out_dir, f"vq{rank}.pt"))


def test_batch_sharded_quantizer_equals_single_gpu(tmp_path):
    """rows are independent, the codebook is replicated, no collective on the forward."""
    import torch.multiprocessing as mp

    from movae_b200 import quantizer as Q

    world = 2
    mp.spawn(_vq_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(3)
    E = 0.5 * torch.randn(512, 64, generator=g)
    z = 0.5 * torch.randn(64, 64, 16, 16, generator=g)
    whole = Q.code_indices(z.cuda(), E.cuda(), 0).cpu()
    got = torch.cat([torch.load(os.path.join(tmp_path, f"vq{r}.pt")) for r in range(world)])
    assert torch.equal(got, whole)


# ------------------------------------------------------------------------------------- data-parallel training step
class _TinyVQNet(torch.nn.Module):
    def __init__(self, mv):
        super().__init__()
        nn = torch.nn
        self.encoder = nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1), nn.LeakyReLU(), nn.Conv2d(16, 64, 3, 2, 1))
        self.vq_layer = mv.VectorQuantizer(512, 64)
        self.decoder = nn.Sequential(nn.ConvTranspose2d(64, 16, 4, 2, 1), nn.LeakyReLU(), nn.ConvTranspose2d(16, 3, 4, 2, 1))

    def forward(self, x):
        enc = self.encoder(x)
        q, commit, embed, _ = self.vq_layer(enc)
        rec = self.decoder(q)
        return enc, [torch.nn.functional.mse_loss(rec, x), embed, 0.25 * commit]


def _dp_worker(rank, world, port, agg_name, flat, graphed, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"            # required to capture NCCL collectives into a CUDA graph
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import movae_b200
        from movae_b200 import parallel

        torch.manual_seed(0)
        net = _TinyVQNet(movae_b200).to(dev)                       # same initialisation on every rank
        x_all = torch.rand(16, 3, 16, 16, generator=torch.Generator().manual_seed(5)) * 2 - 1
        x = x_all[rank * 8:(rank + 1) * 8].to(dev)                  # this rank's slice of the batch
        agg = movae_b200.make_aggregator(agg_name)
        if agg_name == "comfort":
            agg.set_epoch(3, 10)
        parallel.DataParallel(agg)
        opt = movae_b200.SGD(net.parameters(), lr=0.0) if flat else None      # lr 0: keeps the gradients inspectable
        torch.manual_seed(100 + rank)                 # PNUPGrad: the ranks' host RNGs differ; rank 0's draw must win everywhere

        def step():
            if opt is not None:
                opt.zero_grad()
            else:
                net.zero_grad(set_to_none=True)
            enc, losses = net(x)
            if isinstance(agg, movae_b200.MGDA):
                lv = torch.stack([l.detach() for l in losses])
                dist.all_reduce(lv, op=dist.ReduceOp.AVG)
                agg.set_losses(lv)
            movae_b200.mtl_backward(losses=losses, features=[enc], aggregator=agg, retain_graph=True)
            if opt is not None:
                opt.step()

        if graphed:
            g = movae_b200.GraphedStep(step, warmup=2)
            g()
        else:
            step()
        torch.cuda.synchronize()
        torch.save({n: p.grad.detach().cpu() for n, p in net.named_parameters()}, os.path.join(out_dir, f"dp{rank}.pt"))
        if graphed:
            # a live CUDA graph holds NCCL kernels: destroy_process_group() blocks under it (observed); leave without it
            dist.barrier()
            os._exit(0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("agg_name,flat,graphed", [("upgrad", False, False), ("aligned_mtl", True, False), ("mgda_lgn", True, False),
                                                    ("upgrad", True, True), ("comfort", True, False), ("pnupgrad", True, False)])
def test_data_parallel_mtl_backward_equals_single_process_on_the_whole_batch(tmp_path, agg_name, flat, graphed):
    """parallel.DataParallel (reduce-scatter of the Jacobian rows, sharded K1 / K3, all-gather; task gradients all-reduced)
    on 2 GPUs with half the batch each == mtl_backward on one GPU with the whole batch.  Tolerance rtol 1e-4 / atol 1e-6:
    means of halves vs mean of the whole round differently, and cuDNN picks algorithms per batch size."""
    import torch.multiprocessing as mp

    import movae_b200

    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), agg_name, flat, graphed, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"dp{r}.pt")) for r in range(world)]
    torch.manual_seed(0)
    net = _TinyVQNet(movae_b200).cuda()
    x = (torch.rand(16, 3, 16, 16, generator=torch.Generator().manual_seed(5)) * 2 - 1).cuda()
    agg = movae_b200.make_aggregator(agg_name)
    if agg_name == "comfort":
        agg.set_epoch(3, 10)                          # beta = 0.15: both halves of the blend matter
    enc, losses = net(x)
    if isinstance(agg, movae_b200.MGDA):
        agg.set_losses(torch.stack([l.detach() for l in losses]))
    torch.manual_seed(100)                            # rank 0's RNG stream
    movae_b200.mtl_backward(losses=losses, features=[enc], aggregator=agg, retain_graph=True)
    for n, p in net.named_parameters():
        ref = p.grad.cpu().numpy()
        for r in range(world):
            np.testing.assert_allclose(parts[r][n].numpy(), ref, rtol=1e-4, atol=1e-6, err_msg=f"{n} rank {r}")
        assert torch.equal(parts[0][n], parts[1][n]), f"{n}: replicas diverged"
