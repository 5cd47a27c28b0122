"""Multi-GPU tests (need >= 2 CUDA devices; skipped on a single-GPU box): P-sharded aggregation through
(a) the NCCL all_reduce reducer and (b) the exchange fused into K1's tail / K2's head over peer memory
must both equal the single-GPU aggregation of the full Jacobian, with bit-identical weights on all ranks."""
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
