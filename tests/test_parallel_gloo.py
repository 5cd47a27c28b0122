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


def _dp_worker(rank, world, port, k, P, agg_name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import movae_b200
        from movae_b200 import parallel
        from oracle import aggregation as oa

        agg = movae_b200.make_aggregator(agg_name)
        dp = parallel.DataParallel(agg)
        assert agg.data_parallel is dp and agg.weighting.gramian_reducer is not None
        Ps = dp.shard_len(P)
        assert Ps % 4 == 0 and dp.padded_columns(P) == world * Ps >= P
        J_local = _make_J(k, P, seed=100 + rank)                      # this rank's rows (its batch slice)
        Jp = torch.zeros(k, world * Ps)
        Jp[:, :P] = J_local
        Jsh = dp.reduce_scatter_rows(Jp).clone()                      # averaged column shard of this rank
        # the row-wise protocol autojac drives (rows handed over as they are built; identically-zero rows stay off the wire)
        dp.begin_rows(k, P, Jp.dtype, Jp.device)
        for i in range(k):
            if i == 1:
                dp.row_zero(i)
            else:
                dp.row_ready(i, Jp[i])
        Jrw = dp.finish_rows()
        assert float(Jrw[1].abs().max()) == 0.0
        for i in range(k):
            if i != 1:
                assert torch.equal(Jrw[i], Jsh[i])
        Jsh = dp.reduce_scatter_rows(Jp)
        G = torch.from_numpy(oa.gramian_fp64(Jsh))                    # K1 stand-in (oracle; no GPU here)
        agg.weighting.gramian_reducer(G)
        losses = torch.tensor([0.34, 1e-3, 2.5e-4, 0.17, 2.0][:k])
        w, _ = oa.weights_from_gramian(agg_name, G.to(torch.float32), losses=losses)
        g_shard = (w.double() @ Jsh.double()).float()                 # K3 stand-in
        g = dp.all_gather_flat(g_shard)[:P].clone()
        t = [torch.full((5,), float(rank + 1)), torch.arange(3, dtype=torch.float32) * (rank + 1)]
        dp.average_(t)
        torch.save({"g": g, "w": w, "avg": t, "Jsh": Jsh.clone(), "Ps": Ps}, os.path.join(out_dir, f"dp{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("agg_name,k,P", [("upgrad", 3, 10_003), ("aligned_mtl", 2, 7), ("mgda_lgn", 4, 4096)])
def test_data_parallel_plan_equals_aggregating_the_mean_jacobian(tmp_path, agg_name, k, P):
    """parallel.DataParallel: reduce-scatter of the rows (averaged), Gramian all_reduce, all-gather of the aggregated
    gradient == aggregating the mean of the ranks' Jacobians in one process."""
    from oracle import aggregation as oa

    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), k, P, agg_name, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"dp{r}.pt")) for r in range(world)]
    J_mean = sum(_make_J(k, P, seed=100 + r) for r in range(world)) / world
    losses = torch.tensor([0.34, 1e-3, 2.5e-4, 0.17, 2.0][:k])
    G_ref = torch.from_numpy(oa.gramian_fp64(J_mean))
    w_ref, _ = oa.weights_from_gramian(agg_name, G_ref.to(torch.float32), losses=losses)
    g_ref = (w_ref.double() @ J_mean.double()).float()
    Ps = parts[0]["Ps"]
    for r, p in enumerate(parts):
        lo, hi = r * Ps, min(P, (r + 1) * Ps)
        np.testing.assert_allclose(p["Jsh"][:, :max(0, hi - lo)].numpy(), J_mean[:, lo:hi].numpy(), rtol=1e-6, atol=1e-7)
        assert torch.equal(p["w"], parts[0]["w"])
        np.testing.assert_allclose(p["g"].numpy(), g_ref.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(p["avg"][0].numpy(), np.full(5, 1.5))
        np.testing.assert_allclose(p["avg"][1].numpy(), np.arange(3) * 1.5)
