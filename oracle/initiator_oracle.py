"""CPU restatement of the reference's M-of-N track initiator (pymht/initiators/m_of_n.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(pymht_b200/initiators/m_of_n.py) runs the assignment problems on the GPU and never touches it.

Pinned by tests/golden/init_small.npz and init_dense.npz (oracle/gen_golden.py: the UNMODIFIED reference with its own
initiator live, all tracks born by it): tests/test_oracle_vs_golden.py replays the recorded unused-measurement lists and
compares initial targets, preliminary tracks and initiators scan by scan.

Third-party arithmetic: the reference's assignment solver is `munkres` (cython-munkres-wrapper, un-vendored, absent here).
It returns an optimal assignment of a square cost matrix; this restatement uses scipy.optimize.linear_sum_assignment on the
SAME padded matrix (m_of_n.py:56-63), which is what oracle/ref_shim.py gives the reference too.

Arrays are kept in the reference's dtypes: measurements, predicted states, model matrices and covariances float32.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.stats import chi2

GAMMA = chi2(df=2).ppf(0.99)            # m_of_n.py:13-16
R_AIS_LOW = np.float32(9.0)             # models/ais.py:9-13: R(False) = 3.0**2 * I4, float32


def phi(T):                             # models/pv.py Phi, float32
    A = np.eye(4, dtype=np.float32)
    A[0, 2] = A[1, 3] = T
    return A


def q_matrix(T, sigmaQ=1.0):            # models/pv.py:17-23, float32
    q = np.array([[T ** 4. / 4., 0., T ** 3. / 3., 0.], [0., T ** 4. / 4., 0., T ** 3. / 3.],
                  [T ** 3. / 3., 0., T ** 2., 0.], [0., T ** 3. / 3., 0., T ** 2.]], dtype=np.float32)
    return q * sigmaQ


P0 = np.array(np.diag([6.25, 6.25, 0.3 * 6.25, 0.3 * 6.25]), dtype=np.float32)     # models/pv.py:12-13


def solve_gnn(delta_matrix, gate_distance=np.inf):
    """m_of_n.py:24-104: gate, pad to a square matrix (invalid pairs bigM, padding 10 x max valid cost), optimal assignment,
    keep the valid pairs.  Returns [(row, col)] in row order."""
    cost = np.array(delta_matrix, dtype=delta_matrix.dtype, copy=True)
    cost[cost > gate_distance] = np.inf
    valid = cost < np.inf
    if not valid.any():
        return []
    bigM = np.power(10., 1.0 + np.ceil(np.log10(1. + np.sum(cost[valid]))))
    cost[~valid] = bigM
    vcol, vrow = valid.any(axis=0), valid.any(axis=1)
    nr, nc = int(vrow.sum()), int(vcol.sum())
    n = max(nr, nc)
    maxv = 10. * np.max(cost[valid])
    d = np.zeros((n, n)) + maxv
    d[:nr, :nc] = cost[np.ix_(vrow, vcol)]
    r, c = linear_sum_assignment(d.astype(np.double))
    ridx, cidx = np.where(vrow)[0], np.where(vcol)[0]
    out = []
    for a, b in zip(r, c):
        if a >= nr or b >= nc:
            continue
        if valid[ridx[a], cidx[b]]:
            out.append((int(ridx[a]), int(cidx[b])))
    return out


class PrelimTrack:
    __slots__ = ("state", "covariance", "n", "m", "measurement_index", "K", "pred")

    def __init__(self, state, covariance):
        self.state, self.covariance, self.n, self.m, self.measurement_index, self.K = state, covariance, 0, 0, None, None


def similarity(p, other_state):
    """PreliminaryTrack.compareSimilarity (m_of_n.py:196-201): NIS of the state difference under p's covariance + R_ais."""
    d = p.state - other_state
    S = p.covariance + np.eye(4, dtype=np.float32) * R_AIS_LOW
    return d.T.dot(np.linalg.inv(S)).dot(d)


class InitiatorOracle:
    """Radar-only restatement of m_of_n.Initiator (m_of_n.py:215-478)."""

    def __init__(self, M, N, v_max, C, R, mergeThreshold=5):
        self.M, self.N, self.v_max = M, N, v_max
        self.C, self.R = np.asarray(C, dtype=np.float32), np.asarray(R, dtype=np.float32)
        self.merge_threshold = mergeThreshold
        self.gamma = GAMMA
        self.initiators = np.zeros((0, 2), dtype=np.float32)
        self.initiator_time = None
        self.preliminary_tracks = []
        self.last_timestamp = None

    def processMeasurements(self, z, t):
        """m_of_n.py:233-245.  z: (n,2) float32 unused measurements, t: scan time.
        Returns [(x0 (4,), P0 (4,4), measurement (2,), measurement index)] after the merge of similar targets."""
        z = np.asarray(z, dtype=np.float32).reshape(-1, 2)
        unused, new = self._process_preliminary(z, t)
        unused = self._process_initiators(unused, z, t)
        self.initiators, self.initiator_time = z[unused].copy(), t          # _spawnInitiators, m_of_n.py:404-413
        self.last_timestamp = t
        return self._merge(new)

    def _process_preliminary(self, z, t):
        """m_of_n.py:247-383 without AIS."""
        new = []
        if self.last_timestamp is not None:
            dt = t - self.last_timestamp
            F, Q = phi(dt), q_matrix(dt)
            for p in self.preliminary_tracks:                                # PreliminaryTrack.predict :176-178
                p.pred = F.dot(p.state)
                p.covariance = F.dot(p.covariance).dot(F.T) + Q
        pred = np.array([p.pred for p in self.preliminary_tracks], ndmin=2, dtype=np.float32)
        n1, n2 = len(self.preliminary_tracks), z.shape[0]
        if n1 == 0 or n2 == 0:
            return list(range(n2)), new
        delta = np.ones((n1, n2), dtype=np.float32) * np.inf
        for i in range(n1):                                                  # :286-298
            zhat = self.C.dot(pred[i])
            dv = z - zhat
            dist = np.linalg.norm(dv, axis=1)
            Pb = self.preliminary_tracks[i].covariance
            S = self.C.dot(Pb).dot(self.C.T) + self.R
            Si = np.linalg.inv(S)
            self.preliminary_tracks[i].K = Pb.dot(self.C.T).dot(Si)
            nis = np.sum(np.matmul(dv, Si) * dv, axis=1)
            inside = nis <= self.gamma
            delta[i, inside] = dist[inside]
        assignments = solve_gnn(delta)
        for ti, mi in assignments:                                           # :305-315
            p = self.preliminary_tracks[ti]
            dv = z[mi] - self.C.dot(pred[ti])
            p.state = pred[ti] + p.K.dot(dv)
            p.covariance = p.covariance - p.K.dot(self.C).dot(p.covariance)
            p.m += 1
            p.measurement_index = mi
        assigned = {a[0] for a in assignments}
        for ti, p in enumerate(self.preliminary_tracks):                     # :317-327
            if ti not in assigned:
                p.state = pred[ti]
            p.n += 1
        keep = []
        for p in self.preliminary_tracks:                                    # :331-361
            speed = np.linalg.norm(p.state[2:4])
            if speed > self.v_max * 1.5:
                continue
            if p.m >= self.M:
                new.append((np.array(p.state), p.covariance, z[p.measurement_index], p.measurement_index))
                continue
            if p.n >= self.N:
                continue
            keep.append(p)
        self.preliminary_tracks = keep
        used = {a[1] for a in assignments}
        return [i for i in range(n2) if i not in used], new

    def _process_initiators(self, unused, z, t):
        """m_of_n.py:385-402 and __spawn_preliminary_tracks :415-478."""
        n1, n2 = len(self.initiators), len(unused)
        if n1 == 0 or n2 == 0:
            return unused
        zu = z[unused]
        dt = t - self.initiator_time
        delta = np.empty((n1, n2, 2))
        for i in range(n1):
            delta[i] = zu - self.initiators[i]
        dist = np.linalg.norm(delta, axis=2)
        assignments = solve_gnn(dist, self.v_max * dt)
        used = {unused[a[1]] for a in assignments}
        left = sorted(i for i in unused if i not in used)
        for ii, mi in assignments:
            dv = zu[mi] - self.initiators[ii]
            vel = dv / dt
            x0 = np.hstack((zu[mi], vel))
            cand = PrelimTrack(x0, P0)
            if not any(similarity(p, x0) <= 1.0 for p in self.preliminary_tracks):
                self.preliminary_tracks.append(cand)
        return left

    def _merge(self, new):
        """_merge_similar_targets / _merge_targets (m_of_n.py:117-153)."""
        if not new:
            return []
        out, used = [], set()
        for i, tgt in enumerate(new):
            if i in used:
                continue
            d = np.array([np.linalg.norm(tgt[0][0:2] - o[0][0:2]) for o in new])
            close = np.where(d < self.merge_threshold)[0]
            sel = [new[j] for j in close if j not in used]
            used.update(int(j) for j in close)
            if len(sel) == 1:
                out.append(sel[0])
            else:
                out.append((np.mean(np.array([s[0] for s in sel]), axis=0), np.mean(np.array([s[1] for s in sel]), axis=0),
                            sel[0][2], None))
        return out
