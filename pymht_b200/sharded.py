"""Tree-sharded `Tracker`: N ranks of one `torch.distributed` group each own a contiguous
slice of the track trees of ONE surveillance region and are fed the SAME scans.

Per scan (reference pymht/tracker.py:194-259):
  1. grow     every rank gates its own trees              (tracker.py:207-209, independent per tree)
  2. exchange all ranks learn every rank's column count (16 bytes per rank); each rank packs its columns as
              records {f64 cost, i32 tree, i32 rows[N+1]} (40 B at N = 6) and ONE all-gather (NCCL over NVLink)
              of the padded record segments gives every rank every column; a kernel unpacks them into the
              structure-of-arrays columns in rank order.  All buffers persist across scans.
  3. solve    the global 0/1 program (tracker.py:228-236,979-1217) on the gathered columns, warm started from
              the previous scan's multipliers -- trees couple only through shared measurement rows
              (tracker.py:1042-1113); rank 0's selection is broadcast so every shard applies the same hypothesis
  4. select   every rank closes the scan for its trees: report, terminate, N-scan prune

Because rank r holds the trees [lo_r, hi_r) and the slices are concatenated in rank order, the global column
order equals the single-forest order, so the sharded run reproduces the one-GPU tracks exactly.

torch is used for what it is here for: device buffers and the process group.  The pure functions below
(`shard_bounds`, `exchange_counts`, `gather_slices`, `local_selection`) run on any backend (gloo in the CPU tests).
"""
import ctypes as C
import time

import numpy as np

from . import _lib
from .tracker import Tracker
from .pyTarget import Target, preinitializedTag


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_counts(dist, torch, device, n_cols, n_trees, group=None):
    """All ranks learn (columns, tree slots) of every rank.  Returns two int lists and the exclusive offsets."""
    mine = torch.tensor([int(n_cols), int(n_trees)], dtype=torch.int64, device=device)
    world = dist.get_world_size(group)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    cols = [int(o[0]) for o in out]
    trees = [int(o[1]) for o in out]
    col_off = [0] + list(np.cumsum(cols)[:-1].astype(int)) if world else [0]
    tree_off = [0] + list(np.cumsum(trees)[:-1].astype(int)) if world else [0]
    return cols, trees, [int(v) for v in col_off], [int(v) for v in tree_off]


def gather_slices(dist, tensors, counts, offsets, group=None):
    """Ragged all-gather in place: every tensor's last dimension is the global column axis; rank r has filled
    [offsets[r], offsets[r] + counts[r]) and receives the other ranks' slices by broadcast."""
    for r, (cnt, off) in enumerate(zip(counts, offsets)):
        if cnt == 0:
            continue
        for t in tensors:
            if t.dim() == 1:
                dist.broadcast(t[off:off + cnt], src=r, group=group)
            else:
                for w in range(t.shape[0]):      # a plane's slice is contiguous, the 2-D slice is not
                    dist.broadcast(t[w, off:off + cnt], src=r, group=group)


def local_selection(sel_global, tree_off, n_trees, col_off):
    """Slice of the global selection for this rank's tree slots, as LOCAL column indices (-1 stays -1)."""
    loc = sel_global[tree_off:tree_off + n_trees].clone()
    loc[loc >= 0] -= col_off
    return loc


def pack_exchange(dist, torch, send, gathered, group=None):
    """THE exchange of a scan: one all-gather of the packed column records (equal-sized, padded segments)."""
    dist.all_gather_into_tensor(gathered, send, group=group)


def tracks_digest(tracks):
    """Order-independent digest of (ID, measurementNumber, cNLLR) over all live tracks: equal for the sharded and the
    single-forest run of the same scans."""
    import hashlib
    h = hashlib.sha256()
    for tid, meas, cn in sorted((int(t[0]), int(t[1]), float(t[2])) for t in tracks):
        h.update(("%d:%d:%.6f;" % (tid, meas, cn)).encode())
    return h.hexdigest()[:16]


class ShardedTracker(Tracker):
    def __init__(self, model, radarPeriod, lambda_phi, lambda_nu, group=None, **kwargs):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self._group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        super().__init__(model, radarPeriod, lambda_phi, lambda_nu, **kwargs)
        self._recycle_slots = False    # slot s of rank r is global tree tree_off[r] + s on every rank: never renumbered
        self.exchangeLog = []          # per scan: dict(n_cols_global, bytes_gathered, ms_exchange, ms_solve)
        self._buf = {}                 # persistent device buffers (grown geometrically, never per scan)
        self._solve_geometry = None    # (cap_cols, n_tree_slots, n_rows) of the workspace holding warm multipliers
        self.globalTrackCount = 0      # tracks ever initiated on ANY rank: the next birth's Target.ID

    def _buffer(self, name, numel, dtype):
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = self._torch.empty(int(numel * 1.5) + 1024, dtype=dtype, device=self._device)
            self._buf[name] = t
        return t

    def preInitialize(self, simList):
        """This rank's contiguous slice of the ground-truth targets; track IDs stay global."""
        targets = list(simList[0])
        lo, hi = shard_bounds(len(targets), self.world, self.rank)
        self.trackIdCounter = lo
        for tgt in targets[lo:hi]:
            self.initiateTarget(Target(tgt.time, None, np.asarray(tgt.cartesianState(), dtype=np.float64), self.P_0,
                                       status=preinitializedTag))
        self.globalTrackCount = len(targets)

    def _births(self, new_targets):
        """tracker.py:147-160 / 266-277 for a forest spread over ranks: a new target is accepted when NO rank holds a leaf
        within mergeThreshold of it (one MIN all-reduce of the candidates' distances to the local leaves) and no target
        accepted earlier in this scan is that close; accepted targets get the next GLOBAL id and go to rank id % world.
        Every rank takes the same decisions from the same data."""
        if not new_targets:
            return
        torch, dist = self._torch, self._dist
        d = np.full(len(new_targets), np.inf)
        if self.mergeThreshold > 0 and len(self._slots) > 0:
            for k, tgt in enumerate(new_targets):
                dk = C.c_double()
                _lib.check(self._lib.mht_forest_min_leaf_distance(self._forest, float(tgt.x_0[0]), float(tgt.x_0[1]),
                                                                  C.byref(dk)))
                d[k] = dk.value
        d_t = torch.from_numpy(d).to(self._device)
        dist.all_reduce(d_t, op=dist.ReduceOp.MIN, group=self._group)
        d = d_t.cpu().numpy()
        accepted = []
        threshold, self.mergeThreshold = self.mergeThreshold, 0.0       # decided here, globally
        try:
            for k, tgt in enumerate(new_targets):
                p = np.asarray(tgt.x_0[0:2], dtype=np.float64)
                if threshold > 0 and (d[k] < threshold or any(np.linalg.norm(p - q) < threshold for q in accepted)):
                    continue
                gid = self.globalTrackCount
                self.globalTrackCount += 1
                accepted.append(p)
                if gid % self.world == self.rank:
                    self.trackIdCounter = gid
                    self.initiateTarget(tgt)
        finally:
            self.mergeThreshold = threshold

    def _agree(self, rc):
        """All ranks learn whether ANY rank failed before a collective is entered (a lone raise would hang the rest)."""
        torch, dist = self._torch, self._dist
        flag = torch.tensor([0 if rc == _lib.MHT_OK else 1], dtype=torch.int32, device=self._device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self._group)
        if rc != _lib.MHT_OK:
            _lib.check(rc)
        if int(flag[0]):
            raise RuntimeError("a peer rank failed in this scan (see its error); the scan was abandoned on every rank")

    def addMeasurementList(self, scanList, aisList=None, **kwargs):
        torch, dist, lib = self._torch, self._dist, self._lib
        if aisList is not None and len(aisList) > 0:
            raise NotImplementedError("AIS fusion (tracker.py:417-552) is outside the accelerated path")
        for kw in ("dynamicWindow", "pruneSimilar"):
            if kwargs.get(kw, False):
                raise NotImplementedError(kw + " is outside the accelerated path")
        self.tic.clear()
        self.toc.clear()
        t_total = time.time()
        z = np.ascontiguousarray(scanList.measurements, dtype=np.float64).reshape(-1, 2)
        nMeas = z.shape[0]
        used = np.zeros(max(nMeas, 1), dtype=np.uint8)
        ginfo = _lib.ScanInfo()
        self._agree(lib.mht_forest_grow(self._forest, nMeas, _lib.ptr(z), 0, float(scanList.time), C.byref(ginfo),
                                        _lib.ptr(used)))
        self.__scanHistory__.append(scanList)
        self.__aisHistory__.append(aisList if aisList is not None else [])
        n_slots = self._n_slots()
        t0 = time.time()
        cols, trees, col_off, tree_off = exchange_counts(dist, torch, self._device, ginfo.n_children, n_slots,
                                                         self._group)
        n_total, t_total_slots = sum(cols), sum(trees)
        W = self.N + 1
        rec = int(lib.mht_record_bytes(W))
        info = _lib.ScanInfo()
        if n_total == 0 or t_total_slots == 0:
            sel_local = torch.full((max(n_slots, 1),), -1, dtype=torch.int32, device=self._device)
            _lib.check(lib.mht_forest_select(self._forest, sel_local.data_ptr(), None, C.byref(info)))
            ext, ms_ex, ms_solve = None, 0.0, 0.0
        else:
            max_cols = max(cols)
            send = self._buffer("send", max_cols * rec, torch.uint8)[:max_cols * rec]
            gathered = self._buffer("gathered", self.world * max_cols * rec, torch.uint8)[:self.world * max_cols * rec]
            _lib.check(lib.mht_forest_export_records(self._forest, tree_off[self.rank], send.data_ptr(), max_cols))
            pack_exchange(dist, torch, send, gathered, self._group)          # the one data-path collective
            # persistent structure-of-arrays columns + workspace: the capacity only ever grows
            cost = self._buffer("cost", n_total, torch.float64)
            cap = cost.numel()                                   # column capacity = stride of the row planes
            tree = self._buffer("tree", cap, torch.int32)
            rows = self._buffer("rows", W * cap, torch.int32)
            scratch = self._buffer("scratch", 1024, torch.uint8)
            counts = np.ascontiguousarray(cols, dtype=np.int64)
            _lib.check(lib.mht_unpack_records(self.world, _lib.ptr(counts), max_cols, W, gathered.data_ptr(), cap,
                                              cost.data_ptr(), tree.data_ptr(), rows.data_ptr(), scratch.data_ptr(), None))
            torch.cuda.current_stream().synchronize()
            ms_ex = 1e3 * (time.time() - t0)
            t1 = time.time()
            n_rows = W * self.maxMeasurements
            geometry = (cap, t_total_slots, n_rows)
            wbytes = int(lib.mht_assoc_workspace(cap, t_total_slots, n_rows, W))
            work = self._buffer("work", wbytes, torch.uint8)
            warm = 1 if self._solve_geometry == (geometry, work.data_ptr()) else 0
            self._solve_geometry = (geometry, work.data_ptr())
            sel = self._buffer("sel", t_total_slots, torch.int32)[:t_total_slots]
            ext = np.zeros(8)
            plane = len(self.__scanHistory__) % W           # the plane this scan's rows recycle
            _lib.check(lib.mht_assoc_solve_warm(n_total, cap, t_total_slots, n_rows, W, cost.data_ptr(), tree.data_ptr(),
                                                rows.data_ptr(), sel.data_ptr(), _lib.ptr(ext), work.data_ptr(), None,
                                                warm, plane * self.maxMeasurements, (plane + 1) * self.maxMeasurements,
                                                float(self.exactBudgetMs if self.exactBudgetMs > 0 else 8.0), int(self.maxDualIterations)),
                       allow=(_lib.MHT_E_NOTOPTIMAL,))
            dist.broadcast(sel, src=0, group=self._group)     # one global hypothesis for every shard
            torch.cuda.current_stream().synchronize()
            ms_solve = 1e3 * (time.time() - t1)
            sel_local = local_selection(sel, tree_off[self.rank], max(n_slots, 0), col_off[self.rank]).contiguous()
            if n_slots == 0:
                sel_local = torch.full((1,), -1, dtype=torch.int32, device=self._device)
            _lib.check(lib.mht_forest_select(self._forest, sel_local.data_ptr(), _lib.ptr(ext), C.byref(info)))
        d = info.as_dict()
        d["ms_gate"] = ginfo.ms_gate
        d["ms_assoc"] = ms_solve
        self.scanInfo.append(d)
        self.exchangeLog.append({"n_cols_global": n_total, "bytes_gathered": self.world * max(cols + [0]) * rec,
                                 "record_bytes": rec, "ms_exchange": ms_ex, "ms_solve": ms_solve})
        self.toc["Process"] = ginfo.ms_gate * 1e-3
        self.toc["Cluster"] = 0.0
        self.toc["Optim"] = (ms_ex + ms_solve) * 1e-3
        self.toc["ILP-Prune"] = self.toc["DynN"] = 0.0
        t_term = time.time()
        self._collect_tracks(scanList)
        self.toc["Terminate"] = time.time() - t_term
        self.toc["N-Prune"] = info.ms_prune * 1e-3
        # births (tracker.py:266-277): the used mask is the OR over the shards; every rank runs the initiator on the
        # same unused measurements and takes the same global accept / id / owner decisions (_births)
        t_init = time.time()
        used_t = torch.from_numpy(used).to(self._device)
        dist.all_reduce(used_t, op=dist.ReduceOp.MAX, group=self._group)
        used = used_t.cpu().numpy()
        unused = scanList.filterUnused(used[:nMeas] == 0)
        self._births(self.initiator.processMeasurements(unused, []))
        self.toc["Init"] = time.time() - t_init
        self.toc["Total"] = time.time() - t_total
        for k, v in self.runtimeLog.items():
            if k in self.toc:
                v.append(self.toc[k])
        if kwargs.get("checkIntegrity", False):
            self._checkTrackerIntegrity()
        if kwargs.get("printTime", False):
            print(self.getTimeLogString())

    def _n_slots(self):
        """Tree slots of the local forest (live and dead): slot s holds global tree tree_off + s."""
        return len(self._slot_info)

    def gatherTracks(self):
        """Every rank's live tracks as (ID, measurementNumber, cumulativeNLLR, x_0), concatenated in rank order."""
        mine = [(int(n.ID), int(n.measurementNumber), float(n.cumulativeNLLR), np.asarray(n.x_0).tolist())
                for n in self.getTrackNodes()]
        out = [None] * self.world
        self._dist.all_gather_object(out, mine, group=self._group)
        return [t for part in out for t in part]
