"""CPU oracle: a NumPy restatement of pyMHT's per-scan hot path.

TEST INFRASTRUCTURE ONLY -- never imported by pymht_b200/ (the product path).  Allowed users:
tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / `--impl reference` leg.

Parity status: PINNED against the live reference.  The reference's own tests hold no golden
vectors for this path, so the pin is (i) the KAT captured from the live
reference in SURVEY.md §8c and (ii) fixtures under tests/golden/ produced by running the
unmodified reference here through oracle/ref_shim.py (script: oracle/gen_golden.py).
tests/test_oracle_vs_golden.py replays them against this file.

What is restated (citations are into /root/reference):
  kalman_predict        pymht/utils/kalman.py:55-64
  kalman_precalc        pymht/utils/kalman.py:82-101
  innovations           pymht/utils/kalman.py:36-40
  nis                   pymht/utils/kalman.py:25-28
  nllr_radar            pymht/utils/kalman.py:14-22
  OracleTracker.grow    pymht/tracker.py:309-351,383-398,804-889 + pyTarget.py:227-258,319-328
  OracleTracker.cluster pymht/tracker.py:961-974
  OracleTracker.columns pymht/tracker.py:1029-1136 (A1/A2/c as sparse incidence) + pyTarget.py:124-125
  solve_blp             pymht/tracker.py:1155-1217 (formulation; HiGHS instead of un-vendored OR-Tools CBC)
  select singleton      pymht/pyTarget.py:446-459
  terminate             pymht/tracker.py:891-916,353-381
  n-scan prune          pymht/tracker.py:1219-1231 + pyTarget.py:330-356,414-430
Deliberate deviations from the reference's *implementation* (never from its results):
the ILP is assembled sparse (the reference builds dense bool A1/A2 and an O(rows*cols) Python
comprehension, tracker.py:1095,1178) and clusters are found from a sparse bipartite graph
(the reference fills a dense adjacency in a Python double loop, tracker.py:967-971).  Both make
this port a FASTER cpu baseline than the reference itself.

dtype_mode "reference": Phi,Q,C,R,P0 and the whole covariance chain are float32
(pymht/models/constants.py:2), states/innovations/NIS/NLLR are float64 -- exactly what the
reference holds.  "fp64" keeps everything in float64.
"""
import math
import time as _time

import numpy as np

ETA2_DEFAULT = 5.99  # reference tracker.py:110


# ------------------------------------------------------------------------------------------
# constant-velocity model constants (pymht/models/pv.py:7-34, constants.py:7-10)
# ------------------------------------------------------------------------------------------
def cv_model(T, sigmaQ=1.0, sigmaR=2.5, dtype=np.float32):
    Phi = np.eye(4, dtype=dtype)
    Phi[0, 2] = Phi[1, 3] = T
    q = np.zeros((4, 4))
    q[0, 0] = q[1, 1] = T ** 4 / 4.0
    q[0, 2] = q[2, 0] = q[1, 3] = q[3, 1] = T ** 3 / 3.0
    q[2, 2] = q[3, 3] = T ** 2
    Q = np.array(q, dtype=dtype) * sigmaQ
    C = np.zeros((2, 4), dtype=dtype)
    C[0, 0] = C[1, 1] = 1.0
    R = np.array(np.eye(2) * sigmaR ** 2, dtype=dtype)
    p = 2.5 ** 2
    P0 = np.array(np.diag([p, p, 0.3 * p, 0.3 * p]), dtype=dtype)
    return Phi, Q, C, R, P0


# ------------------------------------------------------------------------------------------
# batched Kalman operators
# ------------------------------------------------------------------------------------------
def kalman_predict(A, Q, x0, P0):
    """x_bar = A x ; P_bar = A P A^T + Q  (kalman.py:55-64).  x0 (L,4), P0 (L,4,4)."""
    x_bar = (A @ x0.T).T
    P_bar = np.matmul(np.matmul(A, P0), A.T) + Q
    return x_bar, P_bar


def kalman_precalc(C, R, x_bar, P_bar):
    """z_hat, S, S^-1, K, P_hat per leaf (kalman.py:82-101)."""
    z_hat = (C @ x_bar.T).T
    S = np.matmul(np.matmul(C, P_bar), C.T) + R
    S_inv = np.linalg.inv(S)
    K = np.matmul(np.matmul(P_bar, C.T), S_inv)
    P_hat = P_bar - np.matmul(K.dot(C), P_bar)
    return z_hat, S, S_inv, K, P_hat


def innovations(z, z_hat):
    """z_tilde[l,m] = z[m] - z_hat[l]  (kalman.py:36-40) -> (L,M,2)."""
    return z[None, :, :] - z_hat[:, None, :]


def nis(z_tilde, S_inv):
    """d2[l,m] = z_tilde^T S^-1 z_tilde  (kalman.py:25-28)."""
    return np.sum(np.matmul(z_tilde, S_inv) * z_tilde, axis=2)


def nllr_radar(lambda_ex, P_d, S, d2):
    """0.5 d2 + ln(lambda_ex sqrt(det(2 pi S)) / P_d)  (kalman.py:14-22).

    `2*np.pi*S` keeps S's dtype, so in reference mode the determinant, sqrt and log run in
    float32 and only the final sum is float64 -- restated as written."""
    lambda_ex, P_d = float(lambda_ex), float(P_d)   # the reference holds Python floats (weak scalars)
    if lambda_ex == 0:
        lambda_ex += 1e-20
    return 0.5 * d2 + np.log((lambda_ex * np.sqrt(np.linalg.det(2 * np.pi * S))) / P_d)


def miss_nllr(P_d):
    """-ln(1-P_d) added to the parent's score by the zero hypothesis (pyTarget.py:319-328)."""
    return -np.log(1 - float(P_d))


def gate_leaves(A, Q, C, R, x0, P0, z, eta2):
    """Everything _processLeafNodes computes for one batch of leaves (tracker.py:383-398,804-889).

    Returns x_bar, P_bar, P_hat, S, and per leaf: gated indices (ascending), d2, x_hat."""
    x_bar, P_bar = kalman_predict(A, Q, x0, P0)
    z_hat, S, S_inv, K, P_hat = kalman_precalc(C, R, x_bar, P_bar)
    zt = innovations(np.asarray(z), z_hat)
    d2 = nis(zt, S_inv)
    inside = d2 <= eta2
    idx = [np.nonzero(inside[i])[0] for i in range(x0.shape[0])]
    x_hat = [x_bar[i][None, :] + np.matmul(K[i], zt[i, idx[i]].T).T for i in range(x0.shape[0])]
    d2g = [d2[i, idx[i]] for i in range(x0.shape[0])]
    return x_bar, P_bar, P_hat, S, idx, d2g, x_hat


# ------------------------------------------------------------------------------------------
# exact 0/1 program  min c.tau  s.t. each measurement row <= 1, each tree row == 1
# ------------------------------------------------------------------------------------------
def solve_blp(cost, col_tree, col_rows_ptr, col_rows_idx, n_trees, n_rows):
    """Formulation of tracker.py:1155-1217 on sparse incidence; exact (HiGHS, gap 0).

    cost[n]; col_tree[n] in [0,n_trees); column j uses rows col_rows_idx[ptr[j]:ptr[j+1]].
    Returns ascending selected column indices (one per tree)."""
    from scipy.optimize import milp, LinearConstraint, Bounds
    from scipy.sparse import csr_matrix, vstack

    n = len(cost)
    cols_of_nnz = np.repeat(np.arange(n), np.diff(col_rows_ptr))
    A1 = csr_matrix((np.ones(len(col_rows_idx)), (col_rows_idx, cols_of_nnz)), shape=(n_rows, n))
    A2 = csr_matrix((np.ones(n), (col_tree, np.arange(n))), shape=(n_trees, n))
    cons = [LinearConstraint(A2, 1, 1)]
    if n_rows:
        cons.append(LinearConstraint(A1, -np.inf, 1))
    res = milp(np.asarray(cost, dtype=float), constraints=cons, integrality=np.ones(n),
               bounds=Bounds(0, 1), options={"mip_rel_gap": 0.0})
    if res.x is None:
        raise RuntimeError("oracle BLP infeasible: %s" % res.message)
    sel = np.flatnonzero(np.round(res.x) > 0)
    assert len(sel) == n_trees
    return sel, float(np.dot(cost, np.round(res.x)))


# ------------------------------------------------------------------------------------------
# hypothesis tree
# ------------------------------------------------------------------------------------------
class Node:
    """One hypothesis (a reference `Target` node, pyTarget.py:16-40) with only path fields."""
    __slots__ = ("tid", "time", "scan", "x", "P", "Pd", "parent", "children", "meas", "z",
                 "cnllr", "is_root", "status")

    def __init__(self, tid, time, scan, x, P, Pd, parent, meas, z, cnllr):
        self.tid, self.time, self.scan = tid, time, scan
        self.x, self.P, self.Pd = x, P, Pd
        self.parent, self.children = parent, None
        self.meas, self.z, self.cnllr = meas, z, cnllr
        self.is_root = False
        self.status = "Active"

    def root(self):
        n = self
        while not n.is_root:
            n = n.parent
        return n

    def score(self):
        return self.cnllr - self.root().cnllr  # pyTarget.py:124-125

    def meas_history(self, steps=None):
        """helpFunctions.backtrackMeasurementNumbers for one node (helpFunctions.py:66-83)."""
        out, n = [], self
        while n.parent is not None and (steps is None or steps > 0):
            out.append(int(n.meas))
            n = n.parent
            if steps is not None:
                steps -= 1
        return out[::-1]


def _leaves_under(node):
    out, stack = [], [node]
    while stack:
        n = stack.pop()
        if n.children is None:
            out.append(n)
        else:
            stack.extend(reversed(n.children))
    return out


class OracleTracker:
    """Steps 1-3 + terminate + N-scan prune of Tracker.addMeasurementList (tracker.py:162-307)
    with the initiator nulled, AIS empty and dynamicWindow/pruneSimilar off."""

    def __init__(self, radarPeriod, lambda_phi, lambda_nu, eta2=ETA2_DEFAULT, N=5, P_d=0.8,
                 dtype_mode="reference", sigmaQ=1.0, sigmaR=2.5, radarRange=float("inf"),
                 position=(0.0, 0.0)):
        dt = np.float32 if dtype_mode == "reference" else np.float64
        self.A, self.Q, self.C, self.R, self.P0 = cv_model(radarPeriod, sigmaQ, sigmaR, dt)
        self.lambda_ex = float(lambda_phi) + float(lambda_nu)
        self.eta2, self.N, self.P_d = float(eta2), int(N), float(P_d)
        self.radarRange, self.position = radarRange, np.asarray(position, dtype=float)
        self.score_upper = -np.log(1 - float(P_d)) * 0.8   # tracker.py:115
        self.cnllr_upper = 3.0                      # tracker.py:116
        self.roots, self.leaves, self.assoc, self.window, self.track = [], [], [], [], []
        self.terminated = []
        self.n_scans = 0
        self.next_id = 0
        self.toc = {}
        self.last_ilp = []   # (n_cols, n_rows, n_trees, objective) per multi-tree cluster

    # tracker.py:147-160 with mergeThreshold = 0
    def initiate(self, x0, time, P0=None):
        node = Node(self.next_id, time, self.n_scans, np.array(x0, dtype=np.float64),
                    np.array(self.P0 if P0 is None else P0, dtype=self.P0.dtype),
                    self.P_d, None, 0, None, 0.0)
        node.is_root = True
        node.status = "preinitialized"
        self.next_id += 1
        self.roots.append(node)
        self.leaves.append([node])
        self.assoc.append(set())
        self.window.append(self.N)
        self.track.append(node)
        return node

    # ---- step 1 --------------------------------------------------------------------------
    def _grow(self, z, time, scan):
        used = np.zeros(len(z), dtype=bool)
        n_pairs = 0
        for t, leaves in enumerate(self.leaves):
            x0 = np.array([n.x for n in leaves], ndmin=2)
            P0 = np.array([n.P for n in leaves], ndmin=3)
            x_bar, P_bar, P_hat, S, idx, d2g, x_hat = gate_leaves(
                self.A, self.Q, self.C, self.R, x0, P0, z, self.eta2)
            new_leaves = []
            for i, leaf in enumerate(leaves):
                score = nllr_radar(self.lambda_ex, leaf.Pd, S[i], d2g[i])
                kids = [Node(leaf.tid, time, scan, x_bar[i], P_bar[i], leaf.Pd, leaf, 0, None,
                             leaf.cnllr + miss_nllr(leaf.Pd))]
                for k, m in enumerate(idx[i]):
                    kids.append(Node(leaf.tid, time, scan, x_hat[i][k], P_hat[i], leaf.Pd, leaf,
                                     int(m) + 1, z[m], leaf.cnllr + score[k]))
                    self.assoc[t].add((scan, int(m) + 1))
                leaf.children = kids
                new_leaves.extend(kids)
                used[idx[i]] = True
                n_pairs += len(idx[i])
            self.leaves[t] = new_leaves
        return used, n_pairs

    # ---- step 2 --------------------------------------------------------------------------
    def _cluster(self):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        nT = len(self.assoc)
        keys = {}
        rows, cols = [], []
        for t, s in enumerate(self.assoc):
            for key in s:
                rows.append(t)
                cols.append(nT + keys.setdefault(key, len(keys)))
        n = nT + len(keys)
        g = coo_matrix((np.ones(len(rows), dtype=bool), (rows, cols)), shape=(n, n))
        n_cl, labels = connected_components(g, directed=False)
        lab = labels[:nT]
        order = np.argsort(lab, kind="stable")
        bounds = np.flatnonzero(np.diff(lab[order])) + 1
        return [c for c in np.split(order, bounds)] if nT else []

    # ---- step 3 --------------------------------------------------------------------------
    def _columns(self, cluster):
        """Sparse A1/A2/c of tracker.py:1042-1136: rows = distinct (scan, measNo) below the roots."""
        cost, col_tree, ptr, idx, nodes = [], [], [0], [], []
        rowid = {}
        for k, t in enumerate(cluster):
            root = self.roots[t]
            for leaf in self.leaves[t]:
                n = leaf
                while n is not root:
                    if n.meas:
                        idx.append(rowid.setdefault((n.scan, n.meas), len(rowid)))
                    n = n.parent
                ptr.append(len(idx))
                cost.append((leaf.cnllr - root.cnllr) / self.N)
                col_tree.append(k)
                nodes.append(leaf)
        return (np.array(cost), np.array(col_tree), np.array(ptr), np.array(idx, dtype=int),
                len(rowid), nodes)

    def _select(self, clusters):
        self.last_ilp = []
        for cl in clusters:
            if len(cl) == 1:
                best, best_score = None, float("inf")
                for leaf in self.leaves[cl[0]]:     # pyTarget.py:446-459 ('<=' => last wins)
                    if leaf.cnllr <= best_score:
                        best, best_score = leaf, leaf.cnllr
                self.track[cl[0]] = best
            else:
                cost, col_tree, ptr, idx, n_rows, nodes = self._columns(cl)
                sel, obj = solve_blp(cost, col_tree, ptr, idx, len(cl), n_rows)
                self.last_ilp.append((len(cost), n_rows, len(cl), obj))
                for j in sel:
                    self.track[cl[col_tree[j]]] = nodes[j]

    # ---- step 6 (tracker.py:891-916,353-381) -----------------------------------------------
    def _terminate(self):
        dead = []
        for t, n in enumerate(self.track):
            if np.linalg.norm(n.x[0:2] - self.position) > self.radarRange:
                n.status = "OutOfRange"
                dead.append(t)
            elif n.score() / (self.N + 1) > self.score_upper:
                n.status = "TooLowScore"
                dead.append(t)
            elif n.cnllr > self.cnllr_upper:
                n.status = "TooLowScore"
                dead.append(t)
        for t in sorted(dead, reverse=True):
            self.terminated.append(self.track[t])
            for lst in (self.roots, self.leaves, self.assoc, self.window, self.track):
                del lst[t]
        return dead

    # ---- N-scan prune (tracker.py:1219-1231, pyTarget.py:330-356,414-430) -------------------
    def _prune(self):
        for t, node in enumerate(self.track):
            steps, n = self.window[t], node
            while steps > 0 and n.parent is not None:
                n, steps = n.parent, steps - 1
            if n.parent is None or n is self.roots[t]:
                continue
            keep = n
            while keep.parent is not None:      # _pruneAllHypothesisExceptThis(backtrack=True)
                keep.parent.children = [keep]
                keep = keep.parent
            self.roots[t].is_root = False
            n.is_root = True
            self.roots[t] = n
            self.leaves[t] = _leaves_under(n)
            s = set()
            stack = list(n.children or [])
            while stack:                        # getMeasurementSet(root=True)
                c = stack.pop()
                if c.meas:
                    s.add((c.scan, c.meas))
                if c.children:
                    stack.extend(c.children)
            self.assoc[t] = s

    # ---- per-scan entry ----------------------------------------------------------------------
    def add_scan(self, z, time):
        z = np.asarray(z)
        self.n_scans += 1
        scan = self.n_scans
        t0 = _time.time()
        used, n_pairs = self._grow(z, time, scan)
        t1 = _time.time()
        clusters = self._cluster()
        t2 = _time.time()
        self._select(clusters)
        t3 = _time.time()
        n_leaves = sum(len(l) for l in self.leaves)
        dead = self._terminate()
        t4 = _time.time()
        self._prune()
        t5 = _time.time()
        self.toc = {"Process": t1 - t0, "Cluster": t2 - t1, "Optim": t3 - t2,
                    "Terminate": t4 - t3, "N-Prune": t5 - t4, "Total": t5 - t0}
        self.clusters = clusters
        return {"used": used, "n_pairs": n_pairs, "n_leaves": n_leaves, "dead": dead,
                "n_clusters": len(clusters)}

    def track_nodes(self):
        return list(self.track)
