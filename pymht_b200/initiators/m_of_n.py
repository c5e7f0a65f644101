"""M-of-N track initiator with the reference's interface (pymht/initiators/m_of_n.py:215-478) whose assignment
problems run on the GPU through libmht_b200 (include/mht_b200.h: mht_gnn_*).  There is no CPU path: without the library
or a B200 the constructor raises.

What the reference does per scan, and where it runs here:
  * _processPreliminaryTracks (:247-383): Kalman-predict every preliminary track, gate ALL measurements against it
    (dense n1 x n2 NIS matrix), global nearest neighbour assignment (dense padded Munkres, :24-104), update / count /
    confirm / drop.   -> predict/update/counting: vectorised float32 numpy on (n1,4) / (n1,4,4) arrays (same dtypes and
    operations as the reference's per-track loops); gate + assignment: mht_gnn_assign(mode 1).
  * _processInitiators (:385-402): dense distance matrix of last scan's leftover measurements against this scan's unused
    ones, gate v_max * dt, the same assignment.   -> mht_gnn_assign(mode 0).
  * __spawn_preliminary_tracks (:415-478): every assigned pair becomes a preliminary track unless its state is within
    NIS <= 1 of an existing one (all-pairs loop, each with a 4x4 inverse).   -> mht_gnn_similar + an O(conflicts) host pass
    for the order dependence (a rejected candidate does not block later ones).
  * _merge_similar_targets (:126-145): host, a handful of targets.
AIS measurements are outside the accelerated path: a non-empty AIS list raises.
"""
import ctypes as C
import logging

import numpy as np
from scipy.stats import chi2

from .. import _lib
from ..models import pv
from ..pyTarget import Target

tracking_parameters = {"gate_probability": 0.99}
tracking_parameters["gamma"] = chi2(df=2).ppf(tracking_parameters["gate_probability"])   # m_of_n.py:13-16

CONFIRMED, PRELIMINARY, DEAD = 1, 0, -1
R_AIS_LOW = np.float32(3.0 ** 2)          # models/ais.py:9-13, R(False), used by compareSimilarity (:198)

log = logging.getLogger(__name__)
_void_p = C.c_void_p        # Initiator.__init__ has a parameter called C (the observation matrix), like the reference


class PreliminaryTrack:
    """Read-only view of one row of the initiator's track arrays (m_of_n.py:147-201)."""

    def __init__(self, state, covariance, m, n, measurement_index):
        self.state, self.covariance, self.m, self.n, self.measurement_index = state, covariance, m, n, measurement_index
        self.mmsi = None

    def get_speed(self):
        return np.linalg.norm(self.state[2:4])

    def mn_analysis(self, M, N):
        if self.m >= M:
            return CONFIRMED
        if self.n >= N and self.m < M:
            return DEAD
        return PRELIMINARY


class Measurement:
    def __init__(self, value, timestamp):
        self.value, self.timestamp = value, timestamp


def _merge_targets(targets):
    """m_of_n.py:114-124."""
    if len(targets) == 1:
        return targets[0]
    x_0 = np.mean(np.array([t.x_0 for t in targets]), axis=0)
    P_0 = np.mean(np.array([t.P_0 for t in targets]), axis=0)
    return Target(targets[0].time, None, x_0, P_0, measurement=targets[0].measurement)


def _merge_similar_targets(initial_targets, threshold):
    """m_of_n.py:126-145, with the pairwise distances taken once (the reference recomputes a distance vector per target)."""
    if not initial_targets:
        return initial_targets
    pos = np.array([t.x_0[0:2] for t in initial_targets])
    close_matrix = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=2) < threshold
    lonely = close_matrix.sum(axis=1) == 1            # nothing but the target itself within the threshold
    targets, used = [], np.zeros(len(initial_targets), dtype=bool)
    for i, target in enumerate(initial_targets):
        if used[i]:
            continue
        if lonely[i]:
            targets.append(target)
            continue
        close = np.flatnonzero(close_matrix[i])
        selected = [initial_targets[j] for j in close if not used[j]]
        targets.append(_merge_targets(selected))
        used[close] = True
    return targets


class Initiator:
    def __init__(self, M, N, v_max, C, R, mergeThreshold=5, **kwargs):
        self.N, self.M, self.v_max = N, M, v_max
        self.C = np.asarray(C, dtype=np.float32)
        self.R = np.asarray(R, dtype=np.float32)
        if self.C.shape != (2, 4) or not np.array_equal(self.C, pv.C_RADAR):
            raise NotImplementedError("the device gate assumes the position observation matrix pv.C_RADAR")
        self.gamma = tracking_parameters["gamma"]
        self.last_timestamp = None
        self.merge_threshold = mergeThreshold
        self._lib = _lib.load()
        self._gnn = _void_p()
        self._cap = (0, 0, 0)
        n0 = int(kwargs.get("maxMeasurements", 4096))     # initial capacity only: the buffers grow with the problem
        self._ensure(n0, n0)
        self._state = np.zeros((0, 4), dtype=np.float32)
        self._cov = np.zeros((0, 4, 4), dtype=np.float32)
        self._m = np.zeros(0, dtype=np.int64)
        self._n = np.zeros(0, dtype=np.int64)
        self._midx = np.zeros(0, dtype=np.int64)
        self._init_z = np.zeros((0, 2), dtype=np.float32)
        self._init_time = None
        self.last_info = {}
        log.info("Initiator ready (%d/%d)", self.M, self.N)

    # -- reference-shaped views ------------------------------------------------------------------------------
    @property
    def preliminary_tracks(self):
        return [PreliminaryTrack(self._state[i], self._cov[i], int(self._m[i]), int(self._n[i]), int(self._midx[i]))
                for i in range(len(self._state))]

    @property
    def initiators(self):
        return [Measurement(v, self._init_time) for v in self._init_z]

    def close(self):
        """Free the device buffers (also done when the object is collected)."""
        gnn = getattr(self, "_gnn", None)
        if gnn is not None and gnn.value:
            self._lib.mht_gnn_destroy(gnn)
            self._gnn = _void_p()
            self._cap = (0, 0, 0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device plumbing -------------------------------------------------------------------------------------
    def _ensure(self, rows, cols):
        """(Re)create the device buffers when a problem outgrows them (powers of two; 64 gated pairs per row)."""
        r, c, _ = self._cap
        if rows <= r and cols <= c:
            return
        r = max(r, 1 << int(np.ceil(np.log2(max(rows, 256)))))
        c = max(c, 1 << int(np.ceil(np.log2(max(cols, 256)))))
        self._create(r, c, 64 * max(r, c))

    def _create(self, r, c, e):
        if self._gnn.value:
            self._lib.mht_gnn_destroy(self._gnn)
            self._gnn = _void_p()
        _lib.check(self._lib.mht_gnn_create(r, c, e, C.byref(self._gnn)))
        self._cap = (r, c, e)

    def _assign(self, mode, row_xy, row_sinv, col_xy, gate):
        """[(row, col)] in row order: _solve_global_nearest_neighbour (m_of_n.py:24-104) of the gated pairs."""
        n1, n2 = len(row_xy), len(col_xy)
        self._ensure(n1, n2)
        row_xy = np.ascontiguousarray(row_xy, dtype=np.float32)
        col_xy = np.ascontiguousarray(col_xy, dtype=np.float32)
        sinv = None if row_sinv is None else np.ascontiguousarray(row_sinv, dtype=np.float32)
        match = np.empty(n1, dtype=np.int32)
        info = _lib.GnnInfo()
        while True:
            rc = self._lib.mht_gnn_assign(self._gnn, mode, n1, _lib.ptr(row_xy), _lib.ptr(sinv), n2, _lib.ptr(col_xy),
                                          float(gate), _lib.ptr(match), C.byref(info))
            if rc == _lib.MHT_E_CAPACITY and self._cap[2] < (1 << 29):
                self._create(self._cap[0], self._cap[1], 4 * self._cap[2])      # more gated pairs than planned
                continue
            _lib.check(rc)
            break
        self.last_info["tracks" if mode else "initiators"] = info.as_dict()
        rows = np.nonzero(match >= 0)[0]
        return rows, match[rows].astype(np.int64)

    # -- the reference's entry point (m_of_n.py:233-245) ----------------------------------------------------------
    def processMeasurements(self, radar_measurement_list, ais_measurement_list=list()):
        if len(ais_measurement_list):
            raise NotImplementedError("AIS measurements are outside the accelerated path")
        unused_indices, initial_targets = self._processPreliminaryTracks(radar_measurement_list)
        unused_indices = self._processInitiators(unused_indices, radar_measurement_list)
        self._spawnInitiators(unused_indices, radar_measurement_list)
        self.last_timestamp = radar_measurement_list.time
        initial_targets = _merge_similar_targets(initial_targets, self.merge_threshold)
        log.info("new initial targets %d", len(initial_targets))
        return initial_targets

    def _processPreliminaryTracks(self, measurement_list):
        """m_of_n.py:247-383 (radar only)."""
        new_targets = []
        t = measurement_list.time
        z = np.array(measurement_list.measurements, dtype=np.float32).reshape(-1, 2)
        n1, n2 = len(self._state), z.shape[0]
        if self.last_timestamp is not None and n1:
            dt = t - self.last_timestamp
            F, Q = pv.Phi(dt), pv.Q(dt)
            pred = np.matmul(self._state, F.T)                                       # PreliminaryTrack.predict :176-178
            self._cov = np.matmul(np.matmul(F, self._cov), F.T) + Q
        else:
            assert n1 == 0, "Undefined situation"
            pred = self._state
        if n1 == 0 or n2 == 0:
            return np.arange(n2), new_targets
        Pb = self._cov
        S = Pb[:, :2, :2] + self.R                                                   # C P C^T + R with C = [I2 0]
        S_inv = np.linalg.inv(S)
        K = np.matmul(Pb[:, :, :2], S_inv)                                           # P C^T S^-1
        rows, cols = self._assign(1, pred[:, :2], S_inv.reshape(n1, 4), z, self.gamma)
        state = pred.copy()
        if len(rows):                                                                # :305-315
            dv = z[cols] - pred[rows, :2]
            state[rows] = pred[rows] + np.matmul(K[rows], dv[:, :, None])[:, :, 0]
            self._cov[rows] = Pb[rows] - np.matmul(K[rows], Pb[rows][:, :2, :])
            self._m[rows] += 1
            self._midx[rows] = cols
        self._state = state
        self._n += 1
        speed = np.linalg.norm(self._state[:, 2:4], axis=1)                          # :331-361
        too_fast = speed > self.v_max * 1.5
        confirmed = ~too_fast & (self._m >= self.M)
        dead = ~too_fast & ~confirmed & (self._n >= self.N)
        for i in np.nonzero(confirmed)[0]:
            mi = int(self._midx[i])
            new_targets.append(Target(t, None, np.array(self._state[i]), self._cov[i].copy(),
                                      measurementNumber=mi + 1, measurement=z[mi]))
        keep = ~(too_fast | confirmed | dead)
        if too_fast.any():
            log.warning("Removing %d TOO FAST preliminary tracks", int(too_fast.sum()))
        self._state, self._cov = self._state[keep], self._cov[keep]
        self._m, self._n, self._midx = self._m[keep], self._n[keep], self._midx[keep]
        used = np.zeros(n2, dtype=bool)
        used[cols] = True
        return np.nonzero(~used)[0], new_targets

    def _processInitiators(self, unused_indices, measurement_list):
        """m_of_n.py:385-402."""
        t = measurement_list.time
        z = np.array(measurement_list.measurements, ndmin=2, dtype=np.float32).reshape(-1, 2)
        n1, n2 = len(self._init_z), len(unused_indices)
        if n1 == 0 or n2 == 0:
            return unused_indices
        zu = z[unused_indices]
        dt = t - self._init_time
        rows, cols = self._assign(0, self._init_z, None, zu, self.v_max * dt)
        used = np.zeros(n2, dtype=bool)
        used[cols] = True
        self._spawn_preliminary_tracks(zu, rows, cols, dt)
        return np.asarray(unused_indices)[~used]

    def _spawnInitiators(self, unused_indices, measurement_list):
        """m_of_n.py:404-413."""
        z = np.array(measurement_list.measurements, dtype=np.float32).reshape(-1, 2)
        self._init_z = z[unused_indices].copy()
        self._init_time = measurement_list.time

    def _spawn_preliminary_tracks(self, zu, rows, cols, dt):
        """m_of_n.py:415-478: candidates in assignment (= initiator) order; one is dropped when an existing preliminary track
        or an earlier ACCEPTED candidate has it within NIS <= 1 under that track's covariance + R_ais."""
        nA = len(rows)
        if nA == 0:
            return
        vel = (zu[cols] - self._init_z[rows]) / dt
        speed = np.linalg.norm(vel, axis=1)
        if (speed > self.v_max * 1.5).any():
            log.warning("Initiator speed to high: %.1f m/s", float(speed.max()))
        cand = np.hstack((zu[cols], vel)).astype(np.float32)
        nP = len(self._state)
        eye = np.eye(4, dtype=np.float32) * R_AIS_LOW
        sinv = np.empty((nP + nA, 4, 4), dtype=np.float32)
        if nP:
            sinv[:nP] = np.linalg.inv(self._cov + eye)
        sinv[nP:] = np.linalg.inv(pv.P0 + eye)
        states = np.ascontiguousarray(np.vstack((self._state, cand)), dtype=np.float32)
        self._ensure(nP + nA, 1)
        cap = 16 * (self._cap[0] + self._cap[1])
        pairs = np.empty((cap, 2), dtype=np.int32)
        n_pairs = C.c_int64()
        _lib.check(self._lib.mht_gnn_similar(self._gnn, nP, nA, _lib.ptr(states), _lib.ptr(sinv), 1.0, _lib.ptr(pairs), cap,
                                             C.byref(n_pairs)))
        pairs = pairs[:n_pairs.value]
        accepted = np.ones(nA, dtype=bool)
        if len(pairs):
            accepted[pairs[pairs[:, 1] < nP, 0]] = False                             # similar to an existing track
            later = pairs[pairs[:, 1] >= nP]
            for k, e in later[np.lexsort((later[:, 1], later[:, 0]))]:               # candidate order
                if accepted[k] and accepted[e - nP]:
                    accepted[k] = False
        nK = int(accepted.sum())
        self._state = np.vstack((self._state, cand[accepted]))
        self._cov = np.concatenate((self._cov, np.broadcast_to(pv.P0, (nK, 4, 4)).astype(np.float32)))
        self._m = np.concatenate((self._m, np.zeros(nK, dtype=np.int64)))
        self._n = np.concatenate((self._n, np.zeros(nK, dtype=np.int64)))
        self._midx = np.concatenate((self._midx, -np.ones(nK, dtype=np.int64)))

