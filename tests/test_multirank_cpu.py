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


def _shard_worker(rank, world, port, out):
    """The ragged column all-gather of pymht_b200.sharded on CPU tensors (gloo)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pymht_b200 import sharded as sh
    dev = torch.device("cpu")
    n_items = 7                                           # 7 trees over 2 ranks -> 4 + 3
    lo, hi = sh.shard_bounds(n_items, world, rank)
    n_cols = [5, 3][rank]                                 # ragged column counts
    cols, trees, col_off, tree_off = sh.exchange_counts(dist, torch, dev, n_cols, hi - lo)
    n_total, W = sum(cols), 3
    cost = torch.full((n_total,), -1.0, dtype=torch.float64)
    tree = torch.full((n_total,), -1, dtype=torch.int32)
    rows = torch.full((W, n_total), -9, dtype=torch.int32)
    o = col_off[rank]
    cost[o:o + n_cols] = torch.arange(n_cols, dtype=torch.float64) + 100 * rank
    tree[o:o + n_cols] = tree_off[rank] + torch.arange(n_cols, dtype=torch.int32) % (hi - lo)
    for w in range(W):
        rows[w, o:o + n_cols] = 1000 * rank + 10 * w + torch.arange(n_cols, dtype=torch.int32)
    sh.gather_slices(dist, [cost, tree, rows], cols, col_off)
    sel_global = torch.tensor([0, 2, -1, 4, 5, -1, 7], dtype=torch.int32)       # column per global tree
    loc = sh.local_selection(sel_global, tree_off[rank], hi - lo, col_off[rank])
    out.put((rank, (lo, hi), cols, trees, col_off, tree_off, cost.tolist(), tree.tolist(), rows.tolist(), loc.tolist()))
    dist.destroy_process_group()


def test_tree_sharded_column_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] == (0, 4) and r1[1] == (4, 7)                  # contiguous tree slices
    assert r0[2] == r1[2] == [5, 3] and r0[3] == r1[3] == [4, 3]
    assert r0[4] == r1[4] == [0, 5] and r0[5] == r1[5] == [0, 4]
    assert r0[6] == r1[6] == [0, 1, 2, 3, 4, 100, 101, 102]     # both ranks hold the same global columns
    assert r0[7] == r1[7] == [0, 1, 2, 3, 0, 4, 5, 6]
    assert r0[8] == r1[8]
    assert r0[8][1] == [10, 11, 12, 13, 14, 1010, 1011, 1012]
    assert r0[9] == [0, 2, -1, 4] and r1[9] == [0, -1, 2]       # global column -> local column


def _pack_worker(rank, world, port, out):
    """The ONE data-path collective of the tree-sharded tracker (padded packed records, all_gather_into_tensor) on gloo."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pymht_b200 import sharded as sh
    W = 3
    rec = np.dtype([("cost", "<f8"), ("tree", "<i4"), ("rows", "<i4", (W,))])      # 24 bytes = mht_record_bytes(3)
    n_cols = [5, 3][rank]
    cols, trees, col_off, tree_off = sh.exchange_counts(dist, torch, torch.device("cpu"), n_cols, [4, 3][rank])
    mx = max(cols)
    mine = np.zeros(mx, dtype=rec)
    mine["cost"][:n_cols] = np.arange(n_cols) + 100 * rank
    mine["tree"][:n_cols] = tree_off[rank] + np.arange(n_cols) % [4, 3][rank]
    mine["rows"][:n_cols] = 1000 * rank + np.arange(n_cols)[:, None] * 10 + np.arange(W)[None, :]
    send = torch.from_numpy(mine.view(np.uint8).copy())
    gathered = torch.empty(world * mx * rec.itemsize, dtype=torch.uint8)
    sh.pack_exchange(dist, torch, send, gathered)
    g = gathered.numpy().view(rec).reshape(world, mx)
    flat = np.concatenate([g[r, :cols[r]] for r in range(world)])                  # rank order = single-forest order
    out.put((rank, rec.itemsize, flat["cost"].tolist(), flat["tree"].tolist(), flat["rows"].tolist()))
    dist.destroy_process_group()


def test_packed_record_all_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + os.getpid() % 1000
    procs = [ctx.Process(target=_pack_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] == 24
    assert r0[2] == r1[2] == [0, 1, 2, 3, 4, 100, 101, 102]
    assert r0[3] == r1[3] == [0, 1, 2, 3, 0, 4, 5, 6]
    assert r0[4] == r1[4] and r0[4][5] == [1000, 1001, 1002]


def test_tracks_digest_is_order_independent():
    from pymht_b200 import sharded as sh
    a = [(3, 7, -1.25, None), (1, 0, 2.0, None)]
    assert sh.tracks_digest(a) == sh.tracks_digest(a[::-1])
    assert sh.tracks_digest(a) != sh.tracks_digest([(3, 7, -1.25, None), (1, 1, 2.0, None)])


def test_shard_bounds_cover_everything():
    from pymht_b200 import sharded as sh
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            b = [sh.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _births_worker(rank, world, port, out):
    """ShardedTracker._births on a stand-in forest (gloo): each rank only knows the leaves it holds; the accept / id / owner
    decisions must come out the same on every rank and follow tracker.py:147-160 for the union of the leaves."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from pymht_b200 import sharded as sh
    from pymht_b200.pyTarget import Target

    leaves = [np.array([[0.0, 0.0], [100.0, 0.0]]), np.array([[0.0, 100.0]])][rank]     # rank 0 holds two leaves, rank 1 one

    class FakeLib:
        def mht_forest_min_leaf_distance(self, forest, x, y, ref):
            ref._obj.value = float(np.min(np.linalg.norm(leaves - np.array([x, y]), axis=1)))
            return 0

    class Stub:
        pass
    trk = Stub()
    trk._torch, trk._dist, trk._group, trk._device = torch, dist, None, torch.device("cpu")
    trk.rank, trk.world = rank, world
    trk._lib, trk._forest, trk._slots = FakeLib(), None, list(range(len(leaves)))
    trk.mergeThreshold, trk.globalTrackCount, trk.trackIdCounter = 25.0, 3, 0
    born = []
    trk.initiateTarget = lambda tgt: born.append((trk.trackIdCounter, float(tgt.x_0[0]), float(tgt.x_0[1])))
    P = np.eye(4)
    cands = [Target(0.0, None, np.array([x, y, 0.0, 0.0]), P) for x, y in
             [(10.0, 0.0),      # 10 m from a leaf of rank 0 -> rejected everywhere
              (0.0, 90.0),      # 10 m from the leaf of rank 1 -> rejected everywhere
              (300.0, 300.0),   # free -> id 3, rank 3 % 2 = 1
              (310.0, 300.0),   # 10 m from the target accepted just before -> rejected
              (500.0, 0.0),     # free -> id 4, rank 0
              (0.0, 500.0)]]    # free -> id 5, rank 1
    sh.ShardedTracker._births(trk, cands)
    out.put((rank, born, trk.globalTrackCount, trk.mergeThreshold))
    dist.destroy_process_group()


def test_sharded_births_decide_globally_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_births_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, born0, n0, thr0), (_, born1, n1, thr1) = res
    assert n0 == n1 == 6 and thr0 == thr1 == 25.0                 # three accepted; the threshold is restored
    assert born0 == [(4, 500.0, 0.0)]
    assert born1 == [(3, 300.0, 300.0), (5, 0.0, 500.0)]
