"""CPU tests of the multi-GPU host logic (mo-vae_b200/parallel.py) with world size 2 over gloo.

There is no GPU here, so the three kernels are stood in for by the CPU oracle (allowed in tests/): each
rank computes its local float64 Gramian partial with the oracle, the product's reducer sums it across
ranks, the oracle solves and recombines locally.  What is under test is the product's sharding and
reduction plumbing: P-sharded result == single-process result, identical weights on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_J(k, P, seed=7):
    g = torch.Generator().manual_seed(seed)
    s = torch.logspace(0, -1, k)
    return s[:, None] * (0.3 * torch.randn(P, generator=g)[None] + 0.91 ** 0.5 * torch.randn(k, P, generator=g))


def _worker(rank, world, port, k, P, agg_name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from movae_b200 import parallel
        from oracle import aggregation as oa

        J = _make_J(k, P)
        lo, hi = parallel.shard_columns(P, rank, world)
        Jl = J[:, lo:hi]
        G = torch.from_numpy(oa.gramian_fp64(Jl)) if hi > lo else torch.zeros(k, k, dtype=torch.float64)
        parallel.gramian_allreduce()(G)                       # the product's reducer (one k x k all_reduce)
        losses = torch.tensor([0.34, 1e-3, 2.5e-4, 0.17, 2.0][:k])
        w, _ = oa.weights_from_gramian(agg_name, G.to(torch.float32), losses=losses)
        assert parallel.check_replicated(w)
        g_local = (w.double() @ Jl.double()).float()
        torch.save({"lo": lo, "hi": hi, "w": w, "g": g_local, "G": G}, os.path.join(out_dir, f"r{rank}.pt"))
        with pytest.raises(TypeError):
            parallel.gramian_allreduce()(torch.zeros(2, 2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("agg_name", ["upgrad", "aligned_mtl", "mgda_lgn"])
@pytest.mark.parametrize("k,P", [(3, 10_003), (2, 6)])
def test_sharded_aggregation_equals_single_process(tmp_path, agg_name, k, P):
    from oracle import aggregation as oa

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, P, agg_name, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    J = _make_J(k, P)
    losses = torch.tensor([0.34, 1e-3, 2.5e-4, 0.17, 2.0][:k])
    G_ref = torch.from_numpy(oa.gramian_fp64(J))
    w_ref, _ = oa.weights_from_gramian(agg_name, G_ref.to(torch.float32), losses=losses)
    g_ref = (w_ref.double() @ J.double()).float()
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == P and parts[0]["hi"] == parts[1]["lo"]
    for p in parts:
        np.testing.assert_allclose(p["G"].numpy(), G_ref.numpy(), rtol=1e-12)
        assert torch.equal(p["w"], parts[0]["w"])                      # replicated solve: bit-identical weights
        np.testing.assert_allclose(p["w"].numpy(), w_ref.numpy(), rtol=1e-5, atol=1e-6)
    g = torch.cat([p["g"] for p in parts])
    np.testing.assert_allclose(g.numpy(), g_ref.numpy(), rtol=1e-5, atol=1e-6)


def test_shard_columns_properties():
    from movae_b200.parallel import all_shards, shard_columns

    for P in (0, 1, 3, 4, 17, 1000, 100_000_003):
        for world in (1, 2, 3, 4, 8):
            sh = all_shards(P, world)
            assert sh[0][0] == 0 and sh[-1][1] == P
            for (a, b), (c, d) in zip(sh[:-1], sh[1:]):
                assert b == c and a % 4 == 0 and b % 4 == 0 and b >= a
    with pytest.raises(ValueError):
        shard_columns(10, 2, 2)
