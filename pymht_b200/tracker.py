"""`Tracker` with the reference's public surface (pymht/tracker.py:39-307) whose per-scan hot path
-- grow every track tree (gate + NLLR), cluster, maximise the global hypothesis, terminate,
N-scan prune -- runs on a B200 through libmht_b200 (include/mht_b200.h).

Replaced reference code: tracker.py:194-259 (steps 1-3, 6 and the N-scan prune of
addMeasurementList), i.e. _growTarget/_processLeafNodes (:309-351,:383-398,:804-889),
_findClustersFromSets (:961-974), _solveOptimumAssociation/_solveBLP_OR_TOOLS (:979-1217),
__analyzeTrackTermination/_terminateTracks (:891-916,:353-381), _nScanPruning (:1219-1231).
Kept call-compatible: __init__ kwargs, preInitialize, initiateTarget, addMeasurementList,
getTrackNodes, runtimeLog/toc keys, getRuntimeAverage.
Out of scope (fail loudly): AIS fusion, plotting, XML export, pruneSimilar.
Step 7 (tracker.py:266-277): `self.initiator` is the M-of-N initiator like in the reference (tracker.py:62-72), with its
assignment problems on the GPU (pymht_b200/initiators/m_of_n.py); `Tracker(..., initiator=None)` nulls it (what the
benchmark configurations of SURVEY.md 8d do), and any object with processMeasurements(unusedRadar, unusedAis) -> [Target]
can be plugged in.
"""
import ctypes as C
import logging
import os
import time

import numpy as np

from . import _lib
from .pyTarget import Target, STATUS_TAGS, preinitializedTag, backtrackMeasurementNumbers  # noqa: F401
from .utils.classDefinitions import AisMessageList


log = logging.getLogger(__name__)


class NullInitiator:
    def processMeasurements(self, unusedRadarMeasurements, unusedAisMeasurements=None):
        return []


class Tracker:
    def __init__(self, model, radarPeriod, lambda_phi, lambda_nu, **kwargs):
        self.position = np.asarray(kwargs.get("position", np.array([0.0, 0.0])), dtype=np.float64)
        self.radarRange = kwargs.get("radarRange", float("inf"))
        self.radarPeriod = radarPeriod
        self.default_P_d = kwargs.get("P_d", 0.8)
        assert 0 < self.default_P_d < 1, "Invalid P_d"
        self.model = model
        self.A = model.Phi(radarPeriod)
        self.C = model.C_RADAR
        self.P_0 = model.P0
        self.R_RADAR = model.R_RADAR()
        self.Q = model.Q(radarPeriod)
        self.mergeThreshold = 4 * (model.sigmaR_RADAR_tracker ** 2)
        # target initiator (tracker.py:62-72)
        self.maxSpeedMS = kwargs.get("maxSpeedMS", 20)
        self.M_required = kwargs.get("M_required", 2)
        self.N_checks = kwargs.get("N_checks", 3)
        which = kwargs.get("initiator", "m_of_n")
        if which is None or which is False:
            self.initiator = NullInitiator()
        elif which == "m_of_n" or which is True:
            from .initiators import m_of_n
            self.initiator = m_of_n.Initiator(self.M_required, self.N_checks, self.maxSpeedMS, self.C, self.R_RADAR,
                                              self.mergeThreshold)
        else:
            self.initiator = which

        self.__targetList__ = []            # root views, one per live track
        self.__targetWindowSize__ = []
        self.__scanHistory__ = []
        self.__trackNodes__ = np.empty(0, dtype=np.dtype(object))
        self.__terminatedTargets__ = []
        self.__clusterList__ = []
        self.__aisHistory__ = []
        self.trackIdCounter = 0
        self.runtimeLog = {k: [] for k in ("Total", "Process", "Cluster", "Optim", "ILP-Prune", "DynN",
                                           "N-Prune", "Terminate", "Init")}
        self.tic, self.toc = {}, {}
        self.nOptimSolved = 0
        self.scanInfo = []                  # mht_scan_info dict per scan

        self.lambda_phi, self.lambda_nu = lambda_phi, lambda_nu
        self.lambda_ex = lambda_phi + lambda_nu
        self.eta2 = kwargs.get("eta2", 5.99)
        N = int(kwargs.get("N", 5))
        self.N_max = self.N = N
        self.scoreUpperLimit = -np.log(1 - self.default_P_d) * 0.8
        self.clnnrUpperLimit = 3.0
        if kwargs.get("pruneSimilar", False):
            raise NotImplementedError("pruneSimilar is outside the accelerated path")
        self.targetSizeLimit = 3000                                  # tracker.py:118
        self.totalGrowTimeLimit = self.radarPeriod * 0.5             # tracker.py:47
        self.nodeGrowTimeLimit = 200e-3                              # tracker.py:48
        self._dynWindowOn = False

        # device forest capacities (HBM): hypotheses per level / live leaves per scan
        self.maxTargets = int(kwargs.get("maxTargets", 4096))
        self.maxMeasurements = int(kwargs.get("maxMeasurements", 65536))
        self.maxNodes = int(kwargs.get("maxNodes", 1 << 22))
        self.maxParents = int(kwargs.get("maxParents", max(1 << 16, self.maxNodes // 3)))
        self.maxDualIterations = int(kwargs.get("maxDualIterations", os.environ.get("MHT_DUAL_ITERS", 120)))
        # wall-clock budget (ms per scan) of the exact branch & bound behind the dual loop; strict=True raises
        # when a scan's global hypothesis could not be proven optimal (the reference only warns, tracker.py:1201-1204)
        self.exactBudgetMs = int(kwargs.get("exactBudgetMs", os.environ.get("MHT_EXACT_MS", 0)))
        self.strict = bool(kwargs.get("strict", False))
        self.nNotOptimal = 0                # scans whose association was not certified optimal
        self._lib = _lib.load()
        self._forest = None
        self._slots = []                    # forest slot of each live track (list order = reference order)
        self._slot_info = {}                # slot -> initial Target of the track
        self._last = None                   # per-track arrays of the last scan
        self._live_rows = None
        self._recycle_slots = True          # the tree-sharded tracker keeps slot numbers stable across ranks instead
        self._create_forest()

    # ------------------------------------------------------------------------------------------
    def _create_forest(self):
        cfg = _lib.ForestConfig()
        cfg.model = _lib.Model.from_arrays(self.A, self.Q, self.C, self.R_RADAR, self.eta2, self.lambda_ex)
        cfg.n_scan_window = self.N
        cfg.max_trees, cfg.max_meas = self.maxTargets, self.maxMeasurements
        cfg.max_nodes, cfg.max_parents = self.maxNodes, self.maxParents
        cfg.default_Pd = self.default_P_d
        cfg.score_upper, cfg.cnllr_upper = float(self.scoreUpperLimit), float(self.clnnrUpperLimit)
        cfg.radar_range = float(self.radarRange) if np.isfinite(self.radarRange) else 1e300
        cfg.position[0], cfg.position[1] = float(self.position[0]), float(self.position[1])
        cfg.max_dual_iters = self.maxDualIterations
        cfg.exact_ms = self.exactBudgetMs
        handle = C.c_void_p()
        _lib.check(self._lib.mht_forest_create(C.byref(cfg), C.byref(handle)))
        self._forest = handle

    def close(self):
        if self._forest is not None:
            self._lib.mht_forest_destroy(self._forest)
            self._forest = None
        closer = getattr(getattr(self, "initiator", None), "close", None)
        if callable(closer):
            closer()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def deviceBytes(self):
        return int(self._lib.mht_forest_bytes(self._forest))

    # ------------------------------------------------------------------------------------------
    def preInitialize(self, simList):
        """Seed one track per ground-truth target of simList[0] (tracker.py:139-145)."""
        for initialTarget in simList[0]:
            self.initiateTarget(Target(initialTarget.time, None, np.asarray(initialTarget.cartesianState(), dtype=np.float64),
                                       self.P_0, status=preinitializedTag))

    def initiateTarget(self, newTarget):
        """tracker.py:147-160: accept the target unless a live leaf is closer than mergeThreshold."""
        if self.mergeThreshold > 0 and len(self._slots) > 0:
            dist = C.c_double()
            _lib.check(self._lib.mht_forest_min_leaf_distance(self._forest, float(newTarget.x_0[0]),
                                                              float(newTarget.x_0[1]), C.byref(dist)))
            if dist.value < self.mergeThreshold:
                return
        x0 = np.ascontiguousarray(newTarget.x_0, dtype=np.float64)
        P0 = np.ascontiguousarray(newTarget.P_0, dtype=np.float32)
        slot = C.c_int32()
        _lib.check(self._lib.mht_forest_initiate(self._forest, _lib.ptr(x0), _lib.ptr(P0), float(self.default_P_d),
                                                 C.byref(slot)))
        node = Target(newTarget.time, len(self.__scanHistory__), x0.copy(), P0.copy(), ID=self.trackIdCounter,
                      P_d=self.default_P_d, status=newTarget.status, isRoot=True,
                      measurementNumber=newTarget.measurementNumber, measurement=newTarget.measurement)
        self.trackIdCounter += 1
        current = self.getTrackNodes()
        self._slots.append(slot.value)
        self._slot_info[slot.value] = node
        self.__targetList__.append(node)
        self.__trackNodes__ = np.append(current, node)
        self.__targetWindowSize__.append(self.N)

    # ------------------------------------------------------------------------------------------
    def addMeasurementList(self, scanList, aisList=None, **kwargs):
        if aisList is not None and len(aisList) > 0:
            raise NotImplementedError("AIS fusion (tracker.py:417-552) is outside the accelerated path")
        if kwargs.get("pruneSimilar", False):
            raise NotImplementedError("pruneSimilar is outside the accelerated path")
        dyn = bool(kwargs.get("dynamicWindow", False))
        if dyn or self._dynWindowOn:
            # tracker.py:244-248,918-950: the size criterion runs on the device inside this scan, before its pruning
            _lib.check(self._lib.mht_forest_set_dynamic_window(self._forest, int(dyn), int(self.targetSizeLimit),
                                                               int(self.N) if self.N < self.N_max else 0))
            self._dynWindowOn = dyn
        self.tic.clear()
        self.toc.clear()
        t_total = time.time()
        z = np.ascontiguousarray(scanList.measurements, dtype=np.float64).reshape(-1, 2)
        nMeas = z.shape[0]
        used = np.zeros(max(nMeas, 1), dtype=np.uint8)
        info = _lib.ScanInfo()
        # a refused scan (capacity) leaves the forest unchanged: the histories only grow once it succeeded
        _lib.check(self._lib.mht_forest_scan(self._forest, nMeas, _lib.ptr(z), float(scanList.time), C.byref(info),
                                             _lib.ptr(used)))
        self.__scanHistory__.append(scanList)
        self.__aisHistory__.append(aisList if aisList is not None else AisMessageList())
        self.scanInfo.append(info.as_dict())
        self.nOptimSolved = info.n_multi_clusters
        if not info.certified:
            # reference tracker.py:1201-1204: "Optim result NOT optimal" is a warning there, too
            self.nNotOptimal += 1
            msg = ("Optim result NOT optimal (scan %d): objective %.6f, lower bound %.6f, gap %.3g; %d open "
                   "component(s), largest %d trees" % (len(self.__scanHistory__), info.objective, info.lower_bound,
                                                       info.objective - info.lower_bound, info.open_components,
                                                       info.max_component))
            if info.repaired_trees:
                msg += "; %d track(s) moved to their miss hypothesis by the final feasibility check" % info.repaired_trees
            if self.strict:
                raise RuntimeError(msg)
            log.warning(msg)
        self.toc["Process"] = info.ms_gate * 1e-3
        self.toc["Cluster"] = info.ms_cluster * 1e-3
        self.toc["Optim"] = info.ms_assoc * 1e-3
        self.toc["ILP-Prune"] = 0.0
        t_dyn = time.time()
        if dyn:
            self._read_windows()
        self.toc["DynN"] = time.time() - t_dyn
        t_term = time.time()
        self._collect_tracks(scanList)
        self.toc["Terminate"] = time.time() - t_term
        self.toc["N-Prune"] = info.ms_prune * 1e-3

        t_init = time.time()
        unusedRadarMeasurements = scanList.filterUnused(used[:nMeas] == 0)
        for initial_target in self.initiator.processMeasurements(unusedRadarMeasurements, []):
            self.initiateTarget(initial_target)
        self.toc["Init"] = time.time() - t_init
        self.toc["Total"] = time.time() - t_total
        if dyn:
            # wall-clock part of __dynamicWindow (tracker.py:919-925,943-950): all trees grow in ONE batched device
            # pass, so a tree's share of the grow time is its share of the leaves; the roof on N follows the
            # reference's 0.8 x radarPeriod test.  Both act on the NEXT scan's pruning (the reference applies them
            # inside the same scan; its per-target clocks do not exist here).
            if info.ms_gate * 1e-3 > self.totalGrowTimeLimit or self.toc["Total"] > self.radarPeriod * 0.8:
                self.N = max(1, self.N - 1)
                log.warning("Iteration took too long (%.1f ms), reducing window size roof from %d to %d",
                            1e3 * self.toc["Total"], self.N + 1, self.N)
                self.__targetWindowSize__ = [min(e, self.N) for e in self.__targetWindowSize__]
        for k, v in self.runtimeLog.items():
            if k in self.toc:
                v.append(self.toc[k])
        if kwargs.get("checkIntegrity", False):      # reference tracker.py:261-262
            self._checkTrackerIntegrity()
        if kwargs.get("printTime", False):
            print(self.getTimeLogString())

    def _collect_tracks(self, scanList):
        """Selected hypothesis per track (tracker.py:228-236), termination (tracker.py:252-253).

        The per-track arrays are read back every scan; the `Target` objects for live tracks are built
        lazily by getTrackNodes(); terminated tracks keep a lazy parent chain too (the library copies the window
        records of the tracks that die in a scan to the host in one batched walk)."""
        cap = max(len(self._slots), 1)
        n = C.c_int32()
        slot = np.zeros(cap, dtype=np.int32)
        x = np.zeros((cap, 4), dtype=np.float64)
        P = np.zeros((cap, 4, 4), dtype=np.float32)
        cn = np.zeros(cap, dtype=np.float64)
        meas = np.zeros(cap, dtype=np.int32)
        status = np.zeros(cap, dtype=np.int32)
        _lib.check(self._lib.mht_forest_tracks(self._forest, cap, C.byref(n), _lib.ptr(slot), _lib.ptr(x), _lib.ptr(P),
                                               _lib.ptr(cn), _lib.ptr(meas), _lib.ptr(status)))
        k = n.value
        # the forest reports in slot order; tracks keep the reference's order (creation order, tracker.py:147-160),
        # which differs once a released slot has been reused
        assert k == len(self._slots) and sorted(slot[:k].tolist()) == sorted(self._slots), \
            "forest/track bookkeeping out of sync"
        if not np.array_equal(slot[:k], self._slots):
            row_of = {int(sl): r for r, sl in enumerate(slot[:k])}
            order = np.array([row_of[sl] for sl in self._slots], dtype=np.int64)
            x, P, cn, meas, status = x[order], P[order], cn[order], meas[order], status[order]
        self._last = (scanList, len(self.__scanHistory__), x, P, cn, meas, status)
        dead = np.flatnonzero(status[:k] != 0)
        hists = self._dead_histories([self._slots[i] for i in dead]) if (len(dead) and self._recycle_slots) else None
        for q, i in enumerate(dead):
            # the library keeps the window records of a track that died this scan (served from the host, no launch):
            # they were read in ONE call above; now the slot goes back so that a later initiation can reuse it
            hist = None
            if self._recycle_slots:
                hist = hists[q]                                        # raw arrays; Targets stay lazy
                _lib.check(self._lib.mht_forest_release(self._forest, int(self._slots[i])))
            self.__terminatedTargets__.append(self._make_node(int(i), self._slots[i], dead=True, hist=hist,
                                                              window=int(self.__targetWindowSize__[i])))
        if len(dead):
            keep = [i for i in range(k) if status[i] == 0]
            self._slots = [self._slots[i] for i in keep]
            self.__targetList__ = [self.__targetList__[i] for i in keep]
            self.__targetWindowSize__ = [self.__targetWindowSize__[i] for i in keep]
            self._live_rows = keep
        else:
            self._live_rows = None
        self.__trackNodes__ = None       # built on demand

    def _read_windows(self):
        """__targetWindowSize__ (tracker.py:83) after the device applied the size criterion of this scan."""
        cap = max(len(self._slots), 1)
        n = C.c_int32()
        slot = np.zeros(cap, dtype=np.int32)
        win = np.zeros(cap, dtype=np.int32)
        _lib.check(self._lib.mht_forest_windows(self._forest, cap, C.byref(n), _lib.ptr(slot), _lib.ptr(win)))
        by_slot = dict(zip(slot[:n.value].tolist(), win[:n.value].tolist()))
        self.__targetWindowSize__ = [by_slot.get(s, w) for s, w in zip(self._slots, self.__targetWindowSize__)]

    def _make_node(self, row, slot, dead=False, hist=None, window=None):
        scanList, scanNumber, x, P, cn, meas, status = self._last
        root = self._slot_info[slot]
        m = int(meas[row])
        return Target(scanList.time, scanNumber, x[row].copy(), P[row].copy(), ID=root.ID, P_d=root.P_d,
                      measurementNumber=m, measurement=(np.asarray(scanList.measurements)[m - 1] if m > 0 else None),
                      cumulativeNLLR=float(cn[row]), status=STATUS_TAGS[int(status[row])],
                      parent_loader=self._make_parent_loader(slot, dead, self._window_of(slot) if window is None else window,
                                                             hist))

    def _prefetch_histories(self):
        """Histories of ALL live tracks in a handful of launches (mht_forest_histories); valid until the next scan."""
        n_live = max(len(self._slots), 1)
        cap_len = 64
        for attempt in range(4):
            n = C.c_int32()
            slot = np.zeros(n_live, dtype=np.int32)
            ln = np.zeros(n_live, dtype=np.int32)
            meas = np.zeros((n_live, cap_len), dtype=np.int32)
            x = np.zeros((n_live, cap_len, 4), dtype=np.float64)
            cn = np.zeros((n_live, cap_len), dtype=np.float64)
            P = np.zeros((n_live, cap_len, 4, 4), dtype=np.float32)
            rc = self._lib.mht_forest_histories(self._forest, n_live, cap_len, C.byref(n), _lib.ptr(slot), _lib.ptr(ln),
                                                _lib.ptr(meas), _lib.ptr(x), _lib.ptr(cn), _lib.ptr(P))
            if rc == _lib.MHT_E_CAPACITY and attempt < 3:     # cap_tracks is exact, so the histories are longer
                cap_len = max(n.value, cap_len) + 8
                continue
            _lib.check(rc)
            break
        self._hist_cache = (len(self.__scanHistory__),
                            {int(s): (meas[i, :ln[i]], x[i, :ln[i]], cn[i, :ln[i]], P[i, :ln[i]])
                             for i, s in enumerate(slot[:n.value])})

    def _history(self, slot, prefetch=True):
        cache = getattr(self, "_hist_cache", None)
        if cache is not None and cache[0] == len(self.__scanHistory__) and slot in cache[1]:
            return cache[1][slot]
        if prefetch and slot in self._slots and len(self._slots) > 8:
            # a live track's history is wanted: fetch every live track's at once, the callers that walk parents
            # (helpFunctions.backtrackMeasurementNumbers, plotting) go over all tracks
            self._prefetch_histories()
            if slot in self._hist_cache[1]:
                return self._hist_cache[1][slot]
        cap = 64
        while True:
            n = C.c_int32()
            meas = np.zeros(cap, dtype=np.int32)
            x = np.zeros((cap, 4), dtype=np.float64)
            cn = np.zeros(cap, dtype=np.float64)
            P = np.zeros((cap, 4, 4), dtype=np.float32)
            rc = self._lib.mht_forest_history(self._forest, slot, cap, C.byref(n), _lib.ptr(meas), _lib.ptr(x),
                                              _lib.ptr(cn), _lib.ptr(P))
            if rc == _lib.MHT_E_CAPACITY:
                cap = n.value + 8
                continue
            _lib.check(rc)
            k = n.value
            return meas[:k], x[:k], cn[:k], P[:k]

    def _dead_histories(self, slots):
        """(meas, x, cnllr, P) of every given slot with one library call (mht_forest_histories_of)."""
        n, cap = len(slots), 64
        sl = np.asarray(slots, dtype=np.int32)
        while True:
            ln = np.zeros(n, dtype=np.int32)
            meas = np.zeros((n, cap), dtype=np.int32)
            x = np.zeros((n, cap, 4), dtype=np.float64)
            cn = np.zeros((n, cap), dtype=np.float64)
            P = np.zeros((n, cap, 4, 4), dtype=np.float32)
            rc = self._lib.mht_forest_histories_of(self._forest, n, _lib.ptr(sl), cap, _lib.ptr(ln), _lib.ptr(meas),
                                                   _lib.ptr(x), _lib.ptr(cn), _lib.ptr(P))
            if rc == _lib.MHT_E_CAPACITY:
                cap = int(ln[0]) + 8
                continue
            _lib.check(rc)
            return [(meas[i, :ln[i]], x[i, :ln[i]], cn[i, :ln[i]], P[i, :ln[i]]) for i in range(n)]

    def _window_of(self, slot):
        try:
            return int(self.__targetWindowSize__[self._slots.index(slot)])
        except ValueError:      # the track died this scan: it was not pruned
            return self.N

    def _make_parent_loader(self, slot, dead=False, window=None, hist=None):
        scan_at_creation = len(self.__scanHistory__)
        window = self.N if window is None else max(0, window)
        root = self._slot_info[slot]          # bound now: the slot may be reused by a later track

        def load(leaf):
            if not dead and len(self.__scanHistory__) != scan_at_creation:
                raise RuntimeError("Target.parent must be materialised before the next scan is added "
                                   "(the window nodes live on the device)")
            meas, x, cn, P = hist if hist is not None else self._history(slot)
            first_scan = root.scanNumber
            # current root of the tree: N scans above the leaf once the window is full (tracker.py:1219-1231);
            # a track terminated this scan was not pruned, its root is one scan older
            pruned_at = leaf.scanNumber - (0 if leaf.status == STATUS_TAGS[0] else 1)
            root_scan = max(first_scan, pruned_at - window)
            chain = Target(root.time, first_scan, x[0].copy(), P[0].copy(), ID=root.ID, P_d=root.P_d,
                           status=root.status, cumulativeNLLR=float(cn[0]), isRoot=(root_scan == first_scan))
            for k in range(1, len(meas) - 1):
                sc = first_scan + k
                scan = self.__scanHistory__[sc - 1]
                m = int(meas[k])
                chain = Target(scan.time, sc, x[k].copy(), P[k].copy(), ID=root.ID, P_d=root.P_d, parent=chain,
                               measurementNumber=m,
                               measurement=(np.asarray(scan.measurements)[m - 1] if m > 0 else None),
                               cumulativeNLLR=float(cn[k]), isRoot=(sc == root_scan))
            assert first_scan + len(meas) - 1 == leaf.scanNumber
            leaf._parent = chain
        return load

    # ------------------------------------------------------------------------------------------
    def getTrackNodes(self):
        """Selected leaf `Target` of every live track (tracker.py:976), in track order."""
        if self.__trackNodes__ is None:
            rows = self._live_rows if self._live_rows is not None else range(len(self._slots))
            nodes = np.empty(len(self._slots), dtype=np.dtype(object))
            for i, (row, slot) in enumerate(zip(rows, self._slots)):
                nodes[i] = self._make_node(row, slot)
            self.__trackNodes__ = nodes
        return self.__trackNodes__

    @property
    def __associatedMeasurements__(self):
        """Per live track: the set of (scanNumber, measurementNumber) used anywhere in its hypothesis tree below the root
        (reference tracker.py:83,331-332,1226-1227), read from the device on access."""
        out = []
        for slot in self._slots:
            cap = 4096
            while True:
                n = C.c_int32()
                sc = np.zeros(cap, dtype=np.int32)
                me = np.zeros(cap, dtype=np.int32)
                rc = self._lib.mht_forest_measurement_set(self._forest, slot, cap, C.byref(n), _lib.ptr(sc), _lib.ptr(me))
                if rc == _lib.MHT_E_CAPACITY:
                    cap = n.value + 16
                    continue
                _lib.check(rc)
                break
            out.append(set(zip(sc[:n.value].tolist(), me[:n.value].tolist())))
        return out

    def getLeafNodes(self, trackIndex):
        """Leaves of one track tree in the reference's DFS order (Target.getLeafNodes)."""
        slot = self._slots[trackIndex]
        cap = 1024
        while True:
            n = C.c_int64()
            x = np.zeros((cap, 4), dtype=np.float64)
            cn = np.zeros(cap, dtype=np.float64)
            meas = np.zeros(cap, dtype=np.int32)
            rc = self._lib.mht_forest_leaves(self._forest, slot, cap, C.byref(n), _lib.ptr(x), _lib.ptr(cn),
                                             _lib.ptr(meas))
            if rc == _lib.MHT_E_CAPACITY:
                cap = int(n.value)
                continue
            _lib.check(rc)
            k = int(n.value)
            return x[:k], cn[:k], meas[:k]

    def getRuntimeAverage(self):
        return {k: np.mean(np.array(v)) for k, v in self.runtimeLog.items()}

    def getTimeLogString(self):
        return " ".join("%s %.2fms" % (k, 1e3 * v) for k, v in self.toc.items())

    def _checkTrackerIntegrity(self):
        nodes = self.getTrackNodes()
        assert len(nodes) == len(self.__targetList__) == len(self._slots)
        assert len({n.ID for n in nodes}) == len(nodes)
        if len(nodes):
            assert len({n.scanNumber for n in nodes}) == 1
        for n in nodes:
            assert np.isfinite(n.cumulativeNLLR)
