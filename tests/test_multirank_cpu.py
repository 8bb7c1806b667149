"""world_size-2 `gloo` test of the multi-rank host logic of bench.py (no GPU): frame sharding by GLOBAL frame index and
the whole-job aggregation (time = max over ranks, work = sum over ranks). The data path itself has no collective."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import oraclelib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = 1000
    frame0, n = bench.shard(F, rank)
    # each rank's inputs depend only on the global frame index
    flips = ol.bsc_flips(7, frame0 + 3, 512, 0.1)
    ms, fi = bench.aggregate(10.0 + 5.0 * rank, 100.0 * (rank + 1), torch.device("cpu"), world)
    q.put((rank, frame0, n, flips.tolist(), ms, fi))
    dist.destroy_process_group()


def test_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, fl0, ms0, fi0), (r1, f1, n1, fl1, ms1, fi1) = res
    assert (f0, n0, f1, n1) == (0, 1000, 1000, 1000)          # contiguous, disjoint global ranges
    assert ms0 == ms1 == 15.0 and fi0 == fi1 == 300.0          # max of times, sum of work
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as ol
    assert fl0 == ol.bsc_flips(7, 3, 512, 0.1).tolist() and fl1 == ol.bsc_flips(7, 1003, 512, 0.1).tolist()
    assert fl0 != fl1
