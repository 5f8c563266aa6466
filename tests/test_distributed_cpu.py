"""world_size-2 gloo test of the N>1 host logic: example sharding, Gram all-reduce, score all-gather.
The per-rank arithmetic is the numpy oracle (CPU); what is under test is the exchange plan of
gadm_b200.distributed / scoring."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, k, T, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gadm_b200.distributed import allgather_cat, allreduce_sum_, shard_range, world_and_rank

        assert world_and_rank() == (world, rank)
        rng = np.random.RandomState(0)
        train = rng.normal(size=(n_total, k)).astype(np.float32)
        gen = rng.normal(size=(T, k)).astype(np.float32)
        lo, hi = shard_range(n_total, world, rank)
        mine = train[lo:hi].astype(np.float64)
        gram = torch.from_numpy(mine.T @ mine + (0.5 / world) * np.eye(k))
        allreduce_sum_(gram)  # sum of per-rank Grams (+ lam split over ranks) == full regularised Gram
        kinv = np.linalg.inv(gram.numpy())
        local = torch.from_numpy((gen.astype(np.float64) @ kinv @ mine.T).mean(axis=0))  # this rank's score slice
        full = allgather_cat(local, n_total=n_total, dim=0)
        full2 = allgather_cat(local, dim=0)  # sizes discovered by an extra all-gather
        smat = allgather_cat(torch.from_numpy(gen.astype(np.float64) @ kinv @ mine.T), n_total=n_total, dim=1)
        if rank == 0:
            q.put((gram.numpy(), full.numpy(), full2.numpy(), smat.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    n_total, k, T, world = 101, 16, 5, 2  # odd N: unequal shards
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, k, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    gram, full, full2, smat = q.get()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    train = rng.normal(size=(n_total, k)).astype(np.float32).astype(np.float64)
    gen = rng.normal(size=(T, k)).astype(np.float32).astype(np.float64)
    K = train.T @ train + 0.5 * np.eye(k)
    want = gen @ np.linalg.inv(K) @ train.T
    np.testing.assert_allclose(gram, K, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(full, want.mean(axis=0), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(full, full2)
    np.testing.assert_allclose(smat, want, rtol=1e-9, atol=1e-12)
