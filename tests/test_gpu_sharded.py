"""Tree-sharded forests on 2 GPUs (NCCL) must reproduce the single-forest tracks exactly: every rank owns a
contiguous slice of the trees of the cfg2 fixture, the column records are all-gathered, the global 0/1 program
is solved on the gathered columns (pymht_b200/sharded.py).  Skipped on a 1-GPU box."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, n_scans, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pymht_b200.sharded import ShardedTracker, shard_bounds
    from pymht_b200.models import pv
    from pymht_b200.pyTarget import Target
    from pymht_b200.utils.classDefinitions import MeasurementList
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    trk = ShardedTracker(pv, T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd, initiator=None, maxTargets=256, maxNodes=1 << 18,
                         maxParents=1 << 16, maxMeasurements=4096)
    trk.mergeThreshold = 0.0
    lo, hi = shard_bounds(len(g["init_x"]), world, rank)
    trk.trackIdCounter = lo
    for x in g["init_x"][lo:hi]:
        trk.initiateTarget(Target(float(g["init_time"]), None, x, pv.P0, status="preinitialized"))
    res = []
    for k in range(n_scans):
        pre = "s%d_" % k
        trk.addMeasurementList(MeasurementList(float(g[pre + "time"]), g[pre + "z"]))
        tracks = trk.gatherTracks()
        res.append((tracks, trk.exchangeLog[-1]["n_cols_global"], trk.scanInfo[-1]["certified"]))
    if rank == 0:
        out.put(res)
    trk.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,n_scans", [("cfg2", 8), ("cfg5_small", 10)])
def test_tree_sharded_equals_single_forest(name, n_scans):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, n_scans, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = golden(name)
    for k, (tracks, n_cols, certified) in enumerate(res):
        pre = "s%d_" % k
        assert certified == 1, (name, k)
        assert [t[0] for t in tracks] == list(g[pre + "ids"]), (name, k)                 # same live tracks
        H = g[pre + "hist"]
        assert [t[1] for t in tracks] == [int(H[i][np.sum(H[i] >= 0) - 1]) for i in range(len(tracks))], (name, k)
        np.testing.assert_allclose([t[2] for t in tracks], g[pre + "cnllr"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.array([t[3] for t in tracks]).reshape(-1, 4), g[pre + "x"], rtol=1e-5, atol=1e-5)


def _worker_births(rank, world, port, name, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pymht_b200.sharded import ShardedTracker
    from pymht_b200.models import pv
    from pymht_b200.utils.classDefinitions import MeasurementList
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    trk = ShardedTracker(pv, T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd, maxTargets=256, maxNodes=1 << 18,
                         maxParents=1 << 16, maxMeasurements=4096)          # the M-of-N initiator is live (the default)
    res = []
    for k in range(int(g["n_scans"])):
        pre = "s%d_" % k
        trk.addMeasurementList(MeasurementList(float(g[pre + "time"]), g[pre + "z"]))
        res.append((sorted(trk.gatherTracks()), len(trk.getTrackNodes())))
    if rank == 0:
        out.put(res)
    trk.close()
    dist.destroy_process_group()


def test_tree_sharded_births_equal_the_reference():
    """No pre-initialised track: every track is born by the M-of-N initiator (each rank runs it on the OR-reduced unused
    measurements), accepted / numbered / placed by ShardedTracker._births.  The union of the two ranks' tracks must be the
    reference's tracks of tests/golden/init_small.npz after every scan, and both ranks must hold some of them."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker_births, args=(r, 2, port, "init_small", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = golden("init_small")
    spread = False
    for k, (tracks, n_rank0) in enumerate(res):
        pre = "s%d_" % k
        ids = list(g[pre + "ids"])
        order = np.argsort(ids)
        assert [t[0] for t in tracks] == sorted(ids), k
        H = g[pre + "hist"]
        assert [t[1] for t in tracks] == [int(H[i][np.sum(H[i] >= 0) - 1]) if np.sum(H[i] >= 0) else int(t[1])
                                          for i, t in zip(order, tracks)], k
        np.testing.assert_allclose([t[2] for t in tracks], g[pre + "cnllr"][order], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(np.array([t[3] for t in tracks]).reshape(-1, 4), g[pre + "x"][order], rtol=1e-5, atol=1e-3)
        spread |= 0 < n_rank0 < len(tracks)
    assert spread
