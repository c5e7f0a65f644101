"""world_size-2 gloo test of bench.py's multi-rank host logic (sector sharding + exchange) on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # every rank builds its own sector: different seeds, same shape
    simList, scans = bench.make_scenario("cfg2_100_targets_1k_meas_N4", 2, seed_offset=rank)
    digest = float(np.sum(scans[0].measurements[:10]))
    t_dev, t_e2e, tracks = bench.exchange(dist, torch.device("cpu"), 1.0 + rank, 2.0 - rank, 100 + rank)
    out.put((rank, digest, t_dev, t_e2e, tracks, len(simList[0])))
    dist.destroy_process_group()


def test_sector_sharding_and_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, d0, tdev0, te0, tr0, n0), (r1, d1, tdev1, te1, tr1, n1) = res
    assert d0 != d1                       # independent sectors
    assert n0 == n1 == 100
    assert tdev0 == tdev1 == 2.0          # max over ranks
    assert te0 == te1 == 2.0
    assert tr0 == tr1 == [100, 101]       # every rank sees every sector's track count
