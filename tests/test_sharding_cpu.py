"""CPU, world_size 2, gloo: clip -> rank partitioning and gather-by-index (no data-path collective)."""
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

from spittle_b200 import sharding


def test_round_robin_partition():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            parts = [sharding.owned_indices(n, w, r) for r in range(w)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    mine = sharding.owned_indices(n, world, rank)
    local = [f"clip-{i}-by-rank-{rank}" for i in mine]       # stands in for the per-rank engine output
    full = sharding.gather_by_index(local, n, world, rank, dist)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_by_index_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [f"clip-{i}-by-rank-{i % 2}" for i in range(11)]
    for _, full in outs:
        assert full == expect
