"""The N>1 path on CPU: two gloo ranks shard a batch by contiguous index ranges (no data-path
collective), each checks its shard against the oracle, and the reported time is the max over ranks —
the same partition and reduction bench.py uses under torchrun with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_batch():
    from libeddsa_b200.sharding import shard_range
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpu_ref import Oracle
    from libeddsa_b200.sharding import barrier, max_over_ranks, shard_range, sum_over_ranks
    n = 301
    rng = np.random.default_rng(5)                      # same synthetic batch on every rank
    sec = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    lo, hi = shard_range(n, rank, world)
    pub = Oracle().genpub(sec[lo:hi])                   # each rank computes only its own shard
    barrier()
    fake_ms = 10.0 * (rank + 1)
    slowest = max_over_ranks(fake_ms)
    total = sum_over_ranks(hi - lo)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, pub.tobytes()))     # test-only: collect the shards to compare
    if rank == 0:
        full = b"".join(g[2] for g in sorted(gathered))
        out.put((slowest, total, full == Oracle().genpub(sec).tobytes()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    slowest, total, same = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slowest == 20.0 and total == 301 and same
