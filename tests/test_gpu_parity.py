"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (north_star): Kalman states and NLLR within 1e-5 relative (atol 1e-6 where the value
crosses zero); gated sets, selected track ids and measurement histories bit-exact."""
import numpy as np
import pytest

from conftest import golden
from oracle import mht_oracle as mo
import gpu_util as gu
from pymht_b200 import _lib

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
XATOL = 1e-5   # states: 1e-5 relative, floor 1e-5 m (m/s) where a coordinate crosses zero (sigma_R is 2.5 m)


def assert_cov_close(got, want):
    """float32 covariances: 1e-5 relative to each matrix's scale.  P_hat = P_bar - K C P_bar cancels,
    so single elements differ from the reference's LAPACK-inverse chain by ~1 ulp OF THE OPERANDS."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    scale = np.abs(want).reshape(want.shape[0], -1).max(axis=1)[:, None, None]
    assert np.all(np.abs(got - want) <= RTOL * scale), float(np.max(np.abs(got - want) / scale))


def _check_gate(out, x_bar, P_bar, P_hat, S, idx, d2g, x_hat, cnllr, Pd, lam):
    L = len(idx)
    assert out["rc"] == 0
    np.testing.assert_allclose(out["x_bar"], x_bar, rtol=1e-12, atol=1e-9)
    assert_cov_close(out["P_bar"], P_bar)
    assert_cov_close(out["P_hat"], P_hat)
    np.testing.assert_allclose(out["miss"], cnllr + mo.miss_nllr(Pd), rtol=1e-12)
    off = out["off"]
    assert off[0] == 0 and off[L] == sum(len(i) for i in idx)
    for l in range(L):
        got = out["meas"][off[l]:off[l + 1]]
        assert list(got) == list(idx[l]), (l, got, idx[l])          # same set, ascending order
        want = cnllr[l] + mo.nllr_radar(lam, Pd, S[l], d2g[l])
        np.testing.assert_allclose(out["cnllr"][off[l]:off[l + 1]], want, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(out["xhat"][off[l]:off[l + 1]], x_hat[l], rtol=RTOL, atol=XATOL)
    used = np.zeros(len(out["used"]), bool)
    for i in idx:
        used[i] = True
    np.testing.assert_array_equal(out["used"].astype(bool)[:len(used)], used)


def test_gate_batch_reference_kat():
    """kalman_kat.npz holds outputs of the reference's own kalman.py on seeded random leaves."""
    g = golden("kalman_kat")
    T, lam, Pd, eta2 = [float(v) for v in g["params"]]
    model, (A, Q, C, R, _) = gu.model_from_oracle(mo, T, eta2, lam)
    L = g["x0"].shape[0]
    cn = np.linspace(-3, 3, L)
    out = gu.gate_batch_host(model, g["x0"], g["P0"], Pd, cn, g["z"])
    leaf, meas = g["pair_leaf"], g["pair_meas"]
    idx = [meas[leaf == l] for l in range(L)]
    off = np.concatenate([[0], np.cumsum([len(i) for i in idx])])
    d2g = [g["d2"][l, idx[l]] for l in range(L)]
    xh = [g["xhat"][off[l]:off[l + 1]] for l in range(L)]
    _check_gate(out, g["x_bar"], g["P_bar"], g["P_hat"], g["S"], idx, d2g, xh, cn, Pd, lam)
    # float32 covariance chain: how close to bit-identical with the reference's NumPy/OpenBLAS
    exact = np.mean(out["P_bar"] == g["P_bar"]), np.mean(out["P_hat"] == g["P_hat"])
    print("bit-identical fraction P_bar %.3f P_hat %.3f" % exact)
    assert exact[0] > 0.99


@pytest.mark.parametrize("L,M,seed", [(1, 1, 0), (7, 0, 1), (300, 40, 2), (5000, 3000, 3)])
def test_gate_batch_random_vs_oracle(L, M, seed):
    rng = np.random.RandomState(seed)
    T, lam, Pd, eta2 = 2.5, 1e-3 + 1e-9, 0.9, 5.99
    model, (A, Q, C, R, P0c) = gu.model_from_oracle(mo, T, eta2, lam)
    x0 = np.concatenate([rng.uniform(-400, 400, (L, 2)), rng.uniform(-12, 12, (L, 2))], axis=1)
    G = rng.normal(size=(L, 4, 4)) * np.array([2.5, 2.5, 1.0, 1.0])[None, :, None]
    P0 = (np.matmul(G, G.transpose(0, 2, 1)) + np.diag([6.25, 6.25, 1.9, 1.9])).astype(np.float32)
    P0[::3] = P0c
    z = rng.uniform(-450, 450, (M, 2)).astype(np.float32)
    if M:
        k = min(M, L)
        z[:k] = (x0[:k, :2] + T * x0[:k, 2:] + rng.normal(scale=4, size=(k, 2))).astype(np.float32)
    cn = rng.normal(size=L)
    out = gu.gate_batch_host(model, x0, P0, Pd, cn, z.astype(np.float64))
    x_bar, P_bar, P_hat, S, idx, d2g, x_hat = [], [], [], [], [], [], []
    for s in range(0, L, 500):   # the oracle materialises (L,M,2): chunk it
        r = mo.gate_leaves(A, Q, C, R, x0[s:s + 500], P0[s:s + 500], z, eta2)
        for dst, src in zip((x_bar, P_bar, P_hat, S), r[:4]):
            dst.append(src)
        idx += r[4]; d2g += r[5]; x_hat += r[6]
    _check_gate(out, np.concatenate(x_bar), np.concatenate(P_bar), np.concatenate(P_hat), np.concatenate(S),
                idx, d2g, x_hat, cn, Pd, lam)


def test_gate_batch_capacity_error():
    model, (A, Q, C, R, P0c) = gu.model_from_oracle(mo, 2.5, 5.99, 1e-4)
    x0 = np.zeros((4, 4))
    z = np.zeros((10, 2))
    out = gu.gate_batch_host(model, x0, np.broadcast_to(P0c, (4, 4, 4)).copy(), 0.9, np.zeros(4), z, cap=3)
    assert out["rc"] == _lib.MHT_E_CAPACITY
    assert out["off"][4] == 40


def _oracle_states(name, upto):
    """Yield (scan index, OracleTracker after _grow, golden) so column problems can be extracted."""
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = g["params"]
    trk = mo.OracleTracker(T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd)
    for x in g["init_x"]:
        trk.initiate(x, float(g["init_time"]))
    for k in range(upto):
        pre = "s%d_" % k
        trk.n_scans += 1
        trk._grow(g[pre + "z"], float(g[pre + "time"]), trk.n_scans)
        yield k, trk
        trk._select(trk._cluster())
        trk._terminate()
        trk._prune()


@pytest.mark.parametrize("name,upto", [("cfg1_crossing", 10), ("cfg2_small", 10), ("cfg5_small", 10), ("cfg2", 8)])
def test_cluster_and_assoc_vs_oracle(name, upto):
    for k, trk in _oracle_states(name, upto):
        cost, tree, RM, n_rows = gu.oracle_columns(trk)
        nT = len(trk.roots)
        if nT == 0:
            continue
        # clusters: same partition of the trees as the oracle's connected components
        lab = gu.cluster_device(tree, RM, nT, n_rows)
        want = np.empty(nT, dtype=int)
        for cl in trk._cluster():
            want[cl] = min(cl)
        np.testing.assert_array_equal(lab, want)
        # association: exact optimum (HiGHS, gap 0) on every multi-tree cluster + argmin on singletons
        rc, sel, info = gu.assoc_solve_device(cost, tree, RM, nT, n_rows)
        assert rc == 0, (name, k, rc, info)
        assert info[6] == 1
        exact_obj, exact_sel = 0.0, {}
        starts = np.searchsorted(tree, np.arange(nT))
        for cl in trk._cluster():
            if len(cl) == 1:
                t = cl[0]
                leaves = trk.leaves[t]
                best = max(i for i, l in enumerate(leaves) if l.cnllr == min(x.cnllr for x in leaves))
                exact_sel[t] = starts[t] + best
                exact_obj += cost[starts[t] + best]
            else:
                c2, ct2, ptr, idx, nr, nodes = trk._columns(cl)
                s2, obj = mo.solve_blp(c2 * trk.N, ct2, ptr, idx, len(cl), nr)
                exact_obj += obj
                for j in s2:
                    exact_sel[cl[ct2[j]]] = starts[cl[ct2[j]]] + trk.leaves[cl[ct2[j]]].index(nodes[j])
        assert abs(info[1] - exact_obj) <= 1e-9 * max(1.0, abs(exact_obj)), (name, k, info, exact_obj)
        assert [exact_sel[t] for t in range(nT)] == list(sel), (name, k)
        # feasibility: no measurement row used twice
        used = RM[:, sel][RM[:, sel] >= 0]
        assert len(used) == len(set(used.tolist()))


def _replay_tracker(name, n_scans=None, scan_kw=None, setup=None, **kw):
    from pymht_b200.tracker import Tracker, backtrackMeasurementNumbers
    from pymht_b200.models import pv
    from pymht_b200.pyTarget import Target
    from pymht_b200.utils.classDefinitions import MeasurementList
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    trk = Tracker(pv, T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd, initiator=kw.pop("initiator", None), **kw)
    trk.mergeThreshold = 0.0
    if setup:
        setup(trk)
    for x in g["init_x"]:
        trk.initiateTarget(Target(float(g["init_time"]), None, x, pv.P0, status="preinitialized"))
    stats = []
    for k in range(int(g["n_scans"]) if n_scans is None else n_scans):
        pre = "s%d_" % k
        trk.addMeasurementList(MeasurementList(float(g[pre + "time"]), g[pre + "z"]), **(scan_kw or {}))
        nodes = list(trk.getTrackNodes())
        info = trk.scanInfo[-1]
        stats.append(info)
        yield k, g, pre, trk, nodes, backtrackMeasurementNumbers(nodes), info
    trk.close()


@pytest.mark.parametrize("name", ["cfg1_crossing", "cfg2_small", "cfg5_small", "cfg2", "cfg5_n8", "cfg2_long"])
def test_tracker_replays_reference_golden(name):
    """Whole addMeasurementList sequences against what the unmodified reference produced."""
    for k, g, pre, trk, nodes, hist, info in _replay_tracker(name):
        assert info["certified"] == 1, (name, k, info)
        assert [n.ID for n in nodes] == list(g[pre + "ids"]), (name, k)
        H = g[pre + "hist"]
        for i, h in enumerate(hist):
            assert h == list(H[i, :len(h)]), (name, k, i, h, H[i])
            assert len(h) == np.sum(H[i] >= 0)
        np.testing.assert_allclose(np.array([n.x_0 for n in nodes]).reshape(-1, 4), g[pre + "x"], rtol=RTOL, atol=XATOL)
        if nodes:
            assert_cov_close(np.array([n.P_0 for n in nodes]).reshape(-1, 4, 4), g[pre + "P"])
        np.testing.assert_allclose([n.cumulativeNLLR for n in nodes], g[pre + "cnllr"], rtol=RTOL, atol=ATOL)
        nleaves = [len(trk.getLeafNodes(i)[1]) for i in range(len(nodes))]
        assert nleaves == list(g[pre + "nleaves"]), (name, k)
        assert info["n_clusters"] == int(g[pre + "nclusters"])
        assert info["n_multi_clusters"] == int(g[pre + "n_ilp"])
        trk._checkTrackerIntegrity()


def test_tracker_cfg3_scan3_vs_reference():
    """BASELINE config 3 one scan deeper than cfg3_head (own scenario draw: 3 scans): scan 3 holds 8.7e4 leaves and ONE
    cluster of ~950 trees (114 s of the reference's time).  Scans 1-2 must be certified and identical; scan 3 is
    reported: certified + identical, or -- like cfg3_lowclutter scan 3 -- a feasible near-optimum with the gap printed."""
    for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg3_scan3", maxTargets=1024, maxNodes=1 << 21,
                                                             maxParents=1 << 19, exactBudgetMs=30000):
        ids, want = [n.ID for n in nodes], list(g[pre + "ids"])
        H = g[pre + "hist"]
        common = [i for i in ids if i in set(want)]
        same = sum(hist[ids.index(i)] == list(H[want.index(i), :len(hist[ids.index(i)])]) for i in common)
        print("cfg3_scan3 scan", k + 1, {kk: info[kk] for kk in ("n_parents", "n_children", "n_clusters", "certified",
                                                                 "n_candidates", "max_component", "bb_nodes",
                                                                 "lower_bound", "objective", "ms_gate", "ms_assoc")},
              "identical histories %d / %d" % (same, len(want)))
        if k < 2 or info["certified"]:
            assert info["certified"] == 1, (k, info)
            assert ids == want and same == len(want), (k, same, len(want))
        else:
            print("  NOT CERTIFIED (one ~950-tree cluster): objective %.4f, bound %.4f" % (info["objective"], info["lower_bound"]))
            assert len(set(ids) ^ set(want)) <= 0.02 * len(want)
            assert same >= 0.94 * len(common)
            assert info["objective"] - info["lower_bound"] <= 5e-3 * abs(info["lower_bound"]) + 1e-9


def test_tracker_cfg5_full_vs_reference():
    """BASELINE config 5 at full size: 500 targets, 50 pairs crossing at 90 degrees on the same scan (less than a
    gate radius apart for +-2 scans), 400 background targets, lambda = 1e-4, N = 8 -- the six scans the reference
    finishes (scan 6: 8.6e4 leaves in ONE 490-tree cluster, 65 s of the reference's time).  Every scan must be
    certified and identical to the reference."""
    for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg5_full", maxTargets=1024, maxNodes=1 << 21,
                                                             maxParents=1 << 19, exactBudgetMs=30000):
        print("cfg5_full scan", k + 1, {kk: info[kk] for kk in ("n_parents", "n_children", "n_clusters", "certified",
                                                                "n_candidates", "max_component", "bb_nodes", "lower_bound",
                                                                "objective", "ms_gate", "ms_assoc")})
        assert info["certified"] == 1, (k, info)
        assert [n.ID for n in nodes] == list(g[pre + "ids"]), k
        H = g[pre + "hist"]
        for i, h in enumerate(hist):
            assert h == list(H[i, :len(h)]), (k, i, h, H[i])
        np.testing.assert_allclose([n.cumulativeNLLR for n in nodes], g[pre + "cnllr"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(np.array([n.x_0 for n in nodes]).reshape(-1, 4), g[pre + "x"], rtol=RTOL, atol=XATOL)
        nleaves = [len(trk.getLeafNodes(i)[1]) for i in range(len(nodes))]
        assert nleaves == list(g[pre + "nleaves"]), k
        assert info["n_clusters"] == int(g[pre + "nclusters"])


def test_dynamic_window_matches_reference():
    """addMeasurementList(dynamicWindow=True) (tracker.py:244-248,918-950): the reference ran this fixture with
    targetSizeLimit = 60 and its wall-clock criteria out of reach, so only the size criterion fires -- per-tree windows
    shrink 4 -> 3 -> 2, roots advance several levels in one scan.  Windows, tracks, histories and leaf counts must
    follow the reference scan by scan."""
    def setup(trk):
        trk.targetSizeLimit = 60
        trk.totalGrowTimeLimit = trk.nodeGrowTimeLimit = 1e9
        trk.radarPeriod = 1e9
    for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg2_dynwin", scan_kw={"dynamicWindow": True}, setup=setup):
        assert info["certified"] == 1, (k, info)
        assert [n.ID for n in nodes] == list(g[pre + "ids"]), k
        assert list(trk.__targetWindowSize__) == list(g[pre + "window"]), (k, trk.__targetWindowSize__)
        H = g[pre + "hist"]
        for i, h in enumerate(hist):
            assert h == list(H[i, :len(h)]), (k, i, h, H[i])
            assert len(h) == np.sum(H[i] >= 0)
        np.testing.assert_allclose([n.cumulativeNLLR for n in nodes], g[pre + "cnllr"], rtol=RTOL, atol=ATOL)
        nleaves = [len(trk.getLeafNodes(i)[1]) for i in range(len(nodes))]
        assert nleaves == list(g[pre + "nleaves"]), k
        assert info["n_clusters"] == int(g[pre + "nclusters"])


@pytest.mark.parametrize("name", ["cfg3_head", "cfg3_lowclutter"])
def test_tracker_cfg3_vs_reference(name):
    """1k targets: the scans the reference can still finish (cfg3_head: 5k measurements, lambda=1e-3, 2 scans;
    cfg3_lowclutter: lambda=1e-4, 3 scans).  The global hypothesis must be PROVEN optimal and the tracks IDENTICAL to
    the reference's -- ids, measurement histories -- on every scan, including the clusters whose LP relaxation has
    a gap (cfg3_head scan 2: 279 trees, gap 0.74; closed by the branch & bound of csrc/bb_core.h).
    One exception, stated, not hidden: cfg3_lowclutter scan 3 holds a 787-tree cluster with an LP gap of 6.0 that
    HiGHS itself needs ~100 s (cuts + strong branching) to close; the exact search's budget here is 30 s.  If it does
    not finish, the scan must still be a feasible near-optimum (reported uncertified, bound gap printed) and the
    test says so -- see DESIGN.md section 4."""
    for k, g, pre, trk, nodes, hist, info in _replay_tracker(name, maxTargets=1024, maxNodes=1 << 20,
                                                             exactBudgetMs=30000):
        ids, want = [n.ID for n in nodes], list(g[pre + "ids"])
        H = g[pre + "hist"]
        common = [i for i in ids if i in set(want)]
        same = sum(hist[ids.index(i)] == list(H[want.index(i), :len(hist[ids.index(i)])]) for i in common)
        print(name, "scan", k + 1, {kk: info[kk] for kk in ("n_parents", "n_children", "n_clusters", "certified",
                                                           "dual_iters", "n_candidates", "max_component", "bb_nodes",
                                                           "lower_bound", "objective", "ms_gate", "ms_assoc")})
        print("  identical measurement histories: %d / %d common tracks (%d ours, %d reference)" % (
            same, len(common), len(ids), len(want)))
        hard = name == "cfg3_lowclutter" and k == 2
        if info["certified"] or not hard:
            assert info["certified"] == 1, (name, k, info)
            assert ids == want
            assert same == len(want)
        else:
            print("  NOT CERTIFIED (787-tree cluster, LP gap 6.0): objective %.4f, bound %.4f" % (
                info["objective"], info["lower_bound"]))
            assert len(set(ids) ^ set(want)) <= 0.02 * len(want)
            assert same >= 0.94 * len(common)
            assert info["objective"] - info["lower_bound"] <= 5e-3 * abs(info["lower_bound"]) + 1e-9


def test_large_random_forest_properties():
    """Full-size property checks (no oracle at this size): every track selects exactly one leaf, no
    measurement is shared between selected hypotheses' current-scan associations, the bound brackets
    the objective, leaves stay sorted, and capacity errors are reported, not hidden."""
    from pymht_b200.tracker import Tracker, backtrackMeasurementNumbers
    from pymht_b200.models import pv
    import pymht_b200.utils.simulator as sim
    sim.seed_simulator(7)
    R, lam, nT = 1142.0, 1e-3, 1000
    init = sim.generateInitialTargets(nT, np.zeros(2), R, 0.9, 1.0)
    simList = sim.simulateTargets(init, 5 * 2.5, 2.5, pv)
    scans = sim.simulateScans(simList, 2.5, pv.C_RADAR, pv.R_RADAR(), lam, R, np.zeros(2), preInitialized=True)
    trk = Tracker(pv, 2.5, lam, 1e-9, N=6, P_d=0.9, initiator=None, maxTargets=1024, maxNodes=1 << 23, maxParents=1 << 21)
    trk.mergeThreshold = 0.0
    trk.preInitialize(simList)
    for scan in scans[:5]:
        trk.addMeasurementList(scan)
        info = trk.scanInfo[-1]
        nodes = trk.getTrackNodes()
        used = [n.measurementNumber for n in nodes if n.measurementNumber > 0]
        assert len(used) == len(set(used)), "a measurement was assigned to two tracks"
        hist = backtrackMeasurementNumbers(nodes)
        for back in range(1, 8):          # ... nor a measurement of any earlier scan of the window
            old = [h[-back] for h in hist if len(h) >= back and h[-back] > 0]
            assert len(old) == len(set(old)), ("two tracks share a measurement %d scans back" % (back - 1), info)
        assert info["lower_bound"] <= info["objective"] + 1e-6, info
        if info["repaired_trees"]:
            print("scan with %d repaired trees:" % info["repaired_trees"], info)
        assert len(nodes) + info["n_dead"] == info["n_trees"]
        assert info["n_children"] == info["n_parents"] + info["n_pairs"]
    trk.close()


def _small_tracker(**kw):
    from pymht_b200.tracker import Tracker
    from pymht_b200.models import pv
    args = dict(N=3, P_d=0.9, maxTargets=8, maxNodes=1 << 12, maxParents=1 << 10, maxMeasurements=256)
    args.update(kw)
    trk = Tracker(pv, 2.5, 1e-4, 1e-9, initiator=None, **args)
    trk.mergeThreshold = 0.0
    return trk, pv


def test_edge_cases_empty_scan_no_targets_and_parent_chain():
    """Edge cases the reference handles: a scan before any target exists, an empty scan (all tracks
    coast on the miss hypothesis), and Target.parent chains that reproduce the history."""
    from pymht_b200.pyTarget import Target
    from pymht_b200.utils.classDefinitions import MeasurementList
    from pymht_b200.tracker import backtrackMeasurementNumbers
    trk, pv = _small_tracker()
    trk.addMeasurementList(MeasurementList(1.0, np.zeros((3, 2), dtype=np.float32)))   # no targets yet
    assert len(trk.getTrackNodes()) == 0
    x0 = np.array([10.0, 20.0, 1.0, -1.0])
    trk.initiateTarget(Target(1.0, None, x0, pv.P0))
    assert trk.getTrackNodes()[0].scanNumber == 1 and trk.getTrackNodes()[0].isRoot
    trk.addMeasurementList(MeasurementList(3.5, np.zeros((0, 2), dtype=np.float32)))   # empty scan
    node = trk.getTrackNodes()[0]
    assert node.measurementNumber == 0 and node.scanNumber == 2
    A = pv.Phi(2.5).astype(float)
    np.testing.assert_allclose(node.x_0, A @ x0, rtol=1e-12)
    assert node.cumulativeNLLR == pytest.approx(-np.log(1 - 0.9), rel=1e-12)
    z = np.array([[15.2, 15.1], [400.0, 400.0]], dtype=np.float32)
    trk.addMeasurementList(MeasurementList(6.0, z))
    node = trk.getTrackNodes()[0]
    assert node.measurementNumber == 1 and np.allclose(node.measurement, z[0])
    # parent chain: leaf -> miss node -> initial node, consistent with backtrackMeasurementNumbers
    assert backtrackMeasurementNumbers([node]) == [[0, 1]]
    assert node.parent.measurementNumber == 0 and node.parent.scanNumber == 2
    assert node.parent.parent.parent is None and node.parent.parent.scanNumber == 1
    np.testing.assert_allclose(node.parent.parent.x_0, x0)
    assert node.getScore() == pytest.approx(node.cumulativeNLLR)
    trk.close()


def test_capacity_errors_are_loud_and_leave_the_forest_usable():
    from pymht_b200.pyTarget import Target
    from pymht_b200.utils.classDefinitions import MeasurementList
    trk, pv = _small_tracker(maxNodes=16, maxParents=16, maxMeasurements=64)
    trk.initiateTarget(Target(0.0, None, np.array([0.0, 0.0, 0.0, 0.0]), pv.P0))
    with pytest.raises(_lib.MhtError) as e:      # more measurements than max_meas
        trk.addMeasurementList(MeasurementList(2.5, np.zeros((65, 2), dtype=np.float32)))
    assert e.value.code == _lib.MHT_E_CAPACITY
    assert len(trk.__scanHistory__) == 0          # a refused scan does not enter the history
    z = np.random.RandomState(0).normal(scale=1.0, size=(40, 2)).astype(np.float32)   # 41 children > 16 nodes
    with pytest.raises(_lib.MhtError) as e:
        trk.addMeasurementList(MeasurementList(2.5, z))
    assert e.value.code == _lib.MHT_E_CAPACITY and "capacity" in str(e.value)
    assert len(trk.__scanHistory__) == 0
    trk.addMeasurementList(MeasurementList(2.5, z[:5]))   # still usable afterwards
    assert len(trk.getTrackNodes()) == 1 and trk.scanInfo[-1]["n_children"] == 6
    trk.close()
    # more live leaves than max_parents: refused at the start of the next scan, nothing is written out of bounds
    trk, pv = _small_tracker(N=5, maxNodes=1 << 12, maxParents=16, maxMeasurements=64)
    trk.initiateTarget(Target(0.0, None, np.array([0.0, 0.0, 0.0, 0.0]), pv.P0))
    z = np.random.RandomState(1).normal(scale=1.0, size=(30, 2)).astype(np.float32)
    trk.addMeasurementList(MeasurementList(2.5, z))        # 31 children from 1 parent: fine
    with pytest.raises(_lib.MhtError) as e:               # 31 live leaves > 16
        trk.addMeasurementList(MeasurementList(5.0, z))
    assert e.value.code == _lib.MHT_E_CAPACITY and "live leaves" in str(e.value)
    trk.close()


def test_initiate_target_respects_merge_threshold():
    """Tracker.initiateTarget drops a new target closer than mergeThreshold to any live leaf
    (reference tracker.py:147-160 / pyTarget.py:181-189)."""
    from pymht_b200.pyTarget import Target
    trk, pv = _small_tracker()
    trk.mergeThreshold = 25.0
    trk.initiateTarget(Target(0.0, None, np.array([0.0, 0.0, 1.0, 0.0]), pv.P0))
    trk.initiateTarget(Target(0.0, None, np.array([10.0, 0.0, 1.0, 0.0]), pv.P0))    # 10 m away: dropped
    trk.initiateTarget(Target(0.0, None, np.array([100.0, 0.0, 1.0, 0.0]), pv.P0))   # kept
    assert [n.ID for n in trk.getTrackNodes()] == [0, 1]
    assert [float(n.x_0[0]) for n in trk.getTrackNodes()] == [0.0, 100.0]
    trk.close()


def test_determinism_same_inputs_same_outputs():
    """Two runs of the same scenario give bit-identical selections and scores (fixed-point sums in the
    dual iteration make the result independent of atomic ordering)."""
    runs = []
    for _ in range(2):
        out = []
        for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg5_small"):
            out.append((hist, [n.cumulativeNLLR for n in nodes], info["objective"], info["lower_bound"]))
        runs.append(out)
    assert runs[0] == runs[1]


def test_dense_gate_heavy_paths_vs_oracle():
    """One target and a scan with 700 measurements inside its gate: exercises the warp-per-leaf gate kernel,
    the pooled (more than 16 gated) lists and the unstaged (more than 512 gated) path; the children must come
    out in ascending measurement order with the oracle's scores and states, also one scan later when every
    one of those children is a parent."""
    from pymht_b200.tracker import Tracker
    from pymht_b200.models import pv
    from pymht_b200.pyTarget import Target
    from pymht_b200.utils.classDefinitions import MeasurementList
    rng = np.random.RandomState(5)
    x0 = np.array([50.0, -20.0, 2.0, 1.0])
    trk = Tracker(pv, 2.5, 1e-3, 1e-9, N=3, P_d=0.9, initiator=None, maxTargets=8, maxNodes=1 << 20, maxParents=1 << 16,
                  maxMeasurements=4096)
    trk.mergeThreshold = 0.0
    orc = mo.OracleTracker(2.5, 1e-3, 1e-9, eta2=5.99, N=3, P_d=0.9)
    trk.initiateTarget(Target(0.0, None, x0, pv.P0))
    orc.initiate(x0, 0.0)
    centre = x0[:2] + 2.5 * x0[2:]
    for k, (n_in, spread) in enumerate([(700, 6.0), (40, 9.0)]):
        centre = centre + 2.5 * x0[2:] * k
        z = np.concatenate([centre + rng.uniform(-spread, spread, (n_in, 2)), rng.uniform(-3000, 3000, (300, 2))])
        z = z[rng.permutation(len(z))].astype(np.float32)
        trk.addMeasurementList(MeasurementList(2.5 * (k + 1), z))
        orc.n_scans += 1
        orc._grow(z, 2.5 * (k + 1), orc.n_scans)
        info = trk.scanInfo[-1]
        want = orc.leaves[0]
        assert info["n_children"] == len(want), (k, info["n_children"], len(want))
        orc._select(orc._cluster())
        orc._terminate()
        orc._prune()
        node, ref = trk.getTrackNodes()[0], orc.track_nodes()[0]
        assert node.measurementNumber == ref.meas
        np.testing.assert_allclose(node.x_0, ref.x, rtol=RTOL, atol=XATOL)
        np.testing.assert_allclose(node.cumulativeNLLR, ref.cnllr, rtol=RTOL, atol=ATOL)
        x, cn, meas = trk.getLeafNodes(0)
        live = orc.leaves[0]
        assert list(meas) == [l.meas for l in live], k
        np.testing.assert_allclose(cn, [l.cnllr for l in live], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(x, np.array([l.x for l in live]), rtol=RTOL, atol=XATOL)
    assert trk.scanInfo[0]["n_children"] > 513 and trk.scanInfo[1]["n_parents"] > 513
    trk.close()


def test_every_leaf_through_the_warp_per_leaf_gate():
    """MHT_HEAVY_ROWS=0 / MHT_HEAVY_CAND=0 send EVERY leaf through forest_gate_heavy_kernel (the thresholds are
    read once per process, hence the subprocess): the cfg2_small replay must still match the reference."""
    import os, subprocess, sys
    from conftest import ROOT
    env = dict(os.environ, MHT_HEAVY_ROWS="0", MHT_HEAVY_CAND="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                        "-m", "gpu", "-k", "replays and (cfg2_small or cfg5_small or cfg1)"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "3 passed" in r.stdout


def test_terminated_tracks_keep_their_history():
    """A track that dies keeps its full parent chain available AFTER later scans (the window nodes leave the
    device store; the library copies the records of dying tracks to the host in one batched walk): the chain
    must extend the reference's history of that track from the scan before it died (outside the N-scan window,
    inside which the selected hypothesis may still have switched)."""
    from pymht_b200.tracker import backtrackMeasurementNumbers
    prev, seen_dead, trk_ref = {}, 0, None
    for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg2"):
        trk_ref = trk
        for t in trk.__terminatedTargets__[seen_dead:]:
            t._died_after = dict(prev)
        seen_dead = len(trk.__terminatedTargets__)
        prev = {int(i): [int(v) for v in h[h >= 0]] for i, h in zip(g[pre + "ids"], g[pre + "hist"])}
        if k == int(g["n_scans"]) - 1:      # only now touch the chains of everything that died along the way
            assert seen_dead > 0
            for t in trk.__terminatedTargets__:
                h = backtrackMeasurementNumbers([t])[0]
                want = t._died_after.get(t.ID, [])
                # the hypothesis may still switch inside the N-scan window; everything older is fixed
                fixed = max(0, len(h) - 1 - int(g["params"][3]))
                assert h[:fixed] == want[:fixed] and len(h) == len(want) + 1, (t.ID, h, want)
                assert h[-1] == t.measurementNumber
                n, depth = t, 0
                while n.parent is not None:
                    n, depth = n.parent, depth + 1
                assert depth == len(h) and n.scanNumber + depth == t.scanNumber


def test_tree_slots_are_recycled():
    """max_trees must bound the LIVE tracks, not the tracks ever initiated.  8 slots, 40 births: tracks
    born outside the radar range die in their first scan (OutOfRange, tracker.py:894-899), their slots are released
    and reused; one long-lived track keeps its identity, order and history throughout."""
    from pymht_b200.pyTarget import Target, backtrackMeasurementNumbers
    from pymht_b200.utils.classDefinitions import MeasurementList
    trk, pv = _small_tracker(maxTargets=8, radarRange=500.0)
    trk.initiateTarget(Target(0.0, None, np.array([10.0, 10.0, 1.0, 0.0]), pv.P0))
    born = 1
    for k in range(10):
        for q in range(4):
            trk.initiateTarget(Target(2.5 * k, None, np.array([900.0 + 10 * q, 900.0, 0.0, 0.0]), pv.P0))
            born += 1
        z = np.array([[10.0 + 2.5 * (k + 1), 10.0]], dtype=np.float32)
        trk.addMeasurementList(MeasurementList(2.5 * (k + 1), z))
        nodes = trk.getTrackNodes()
        assert [n.ID for n in nodes] == [0], (k, [n.ID for n in nodes])
        assert nodes[0].measurementNumber == 1
    assert born == 41 and len(trk.__terminatedTargets__) == 40
    assert backtrackMeasurementNumbers(list(trk.getTrackNodes())) == [[1] * 10]
    dead = trk.__terminatedTargets__[7]
    assert dead.status == "OutOfRange" and dead.parent is not None and dead.parent.parent is None
    trk.close()


def test_histories_of_all_tracks_in_one_call():
    """backtrackMeasurementNumbers over all tracks goes through ONE batched walk (mht_forest_histories: a launch per
    256 tracks) instead of a single-thread launch + sync per track, and returns the reference's histories."""
    from pymht_b200.pyTarget import backtrackMeasurementNumbers
    checked = 0
    for k, g, pre, trk, nodes, hist, info in _replay_tracker("cfg2", n_scans=6):
        trk._hist_cache = None
        trk.__trackNodes__ = None                       # fresh Target views with unloaded parents
        l0 = trk._lib.mht_launch_count()
        again = backtrackMeasurementNumbers(list(trk.getTrackNodes()))
        launches = trk._lib.mht_launch_count() - l0
        H = g[pre + "hist"]
        for i, h in enumerate(again):
            assert h == list(H[i, :len(h)]), (k, i)
        assert len(nodes) > 50 and launches <= 2, (k, len(nodes), launches)
        checked += 1
    assert checked == 6


@pytest.mark.parametrize("name", ["cfg5_n8", "cfg2_small"])
def test_associated_measurements_match_oracle(name):
    """Tracker.__associatedMeasurements__ (tracker.py:83; recomputed from the new root after every N-scan prune,
    tracker.py:1226-1227 / pyTarget.py:414-430) against the oracle's sets, scan by scan, through window fill and pruning."""
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    orc = mo.OracleTracker(T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd)
    for x in g["init_x"]:
        orc.initiate(x, float(g["init_time"]))
    for k, g, pre, trk, nodes, hist, info in _replay_tracker(name):
        orc.add_scan(g[pre + "z"], float(g[pre + "time"]))
        got = trk.__associatedMeasurements__
        assert len(got) == len(orc.assoc), k
        for i, (a, b) in enumerate(zip(got, orc.assoc)):
            assert a == {(int(s), int(m)) for s, m in b}, (name, k, i, sorted(a ^ set(b))[:6])


_LOOP_PROBE = r"""
import json, sys
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import numpy as np
from test_gpu_parity import _replay_tracker
out = []
for name, kw in (("cfg2", {}), ("cfg3_head", dict(maxTargets=1024, maxNodes=1 << 20, exactBudgetMs=30000))):
    for k, g, pre, trk, nodes, hist, info in _replay_tracker(name, **kw):
        out.append([name, k, info["dual_iters"], repr(info["lower_bound"]), repr(info["objective"]), info["certified"],
                    [n.ID for n in nodes], hist])
print("PROBE" + json.dumps(out))
"""


def test_cluster_loop_is_bit_identical_to_the_grid_loop():
    """dual_loop_cluster_kernel (thread-block cluster, shared-memory resident columns, hoisted loads) against the
    cooperative grid version of the same loop (MHT_NO_CLUSTER_LOOP=1, read once per process -> two subprocesses): same
    iteration counts, bit-identical bounds and objectives, same tracks on every scan of cfg2 and cfg3_head."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = _LOOP_PROBE % (ROOT, os.path.join(ROOT, "tests"))
    res = []
    for extra in ({}, {"MHT_NO_CLUSTER_LOOP": "1"}):
        env = dict(os.environ, **extra)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("PROBE")][-1]
        res.append(json.loads(line[5:]))
    assert len(res[0]) == len(res[1]) > 0
    for a, b in zip(res[0], res[1]):
        assert a == b, (a[:6], b[:6])
