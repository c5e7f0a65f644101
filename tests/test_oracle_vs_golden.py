"""The oracle (oracle/mht_oracle.py) replayed against fixtures the UNMODIFIED reference produced
(oracle/gen_golden.py).  This is what pins the oracle; the CUDA path is then checked against it."""
import numpy as np
import pytest

from oracle import mht_oracle as mo
from conftest import golden

RTOL = 1e-5   # north_star: Kalman states and NLLR within 1e-5 relative


def test_survey_kat():
    """KAT captured from the live reference in SURVEY.md 8c."""
    A, Q, C, R, P0 = mo.cv_model(2.5)
    x0 = np.array([[100.0, 200.0, 3.0, -4.0]])
    z = np.array([[110, 188], [107.5, 190], [140, 190]], dtype=np.float32)
    x_bar, P_bar, P_hat, S, idx, d2, x_hat = mo.gate_leaves(A, Q, C, R, x0, P0[None], z, 5.99)
    np.testing.assert_allclose(x_bar[0], [107.5, 190, 3, -4])
    np.testing.assert_allclose([P_bar[0, 0, 0], P_bar[0, 0, 2], P_bar[0, 2, 2]], [27.734375, 9.895834, 8.125], rtol=1e-6)
    np.testing.assert_allclose(S[0], 33.984375 * np.eye(2), rtol=1e-7)
    np.testing.assert_allclose([P_hat[0, 0, 0], P_hat[0, 0, 2], P_hat[0, 2, 2]], [5.1005745, 1.8199234, 5.243454], rtol=1e-6)
    assert list(idx[0]) == [0, 1]
    np.testing.assert_allclose(d2[0], [0.301609198, 0.0], atol=1e-8)
    nl = mo.nllr_radar(1e-4, 0.9, S[0], d2[0])
    np.testing.assert_allclose(nl, [-3.590397279, -3.741201878], rtol=1e-7)
    np.testing.assert_allclose(x_hat[0][0], [109.540229887, 188.367816091, 3.727969408, -4.582375526], rtol=1e-8)
    assert mo.miss_nllr(0.9) == pytest.approx(2.302585092994046, rel=1e-15)


def test_kalman_kat_bitexact():
    g = golden("kalman_kat")
    T, lam, Pd, eta2 = g["params"]
    A, Q, C, R, _ = mo.cv_model(T)
    x_bar, P_bar = mo.kalman_predict(A, Q, g["x0"], g["P0"])
    z_hat, S, S_inv, K, P_hat = mo.kalman_precalc(C, R, x_bar, P_bar)
    d2 = mo.nis(mo.innovations(g["z"], z_hat), S_inv)
    for name, mine in (("x_bar", x_bar), ("P_bar", P_bar), ("z_hat", z_hat), ("S", S), ("S_inv", S_inv),
                       ("K", K), ("P_hat", P_hat), ("d2", d2)):
        assert mine.dtype == g[name].dtype, name
        np.testing.assert_array_equal(mine, g[name], err_msg=name)
    leaf, meas = np.nonzero(d2 <= eta2)
    np.testing.assert_array_equal(leaf, g["pair_leaf"])
    np.testing.assert_array_equal(meas, g["pair_meas"])
    _, _, _, _, idx, d2g, x_hat = mo.gate_leaves(A, Q, C, R, g["x0"], g["P0"], g["z"], eta2)
    nl = np.concatenate([mo.nllr_radar(lam, Pd, S[i], d2g[i]) for i in range(len(idx))])
    np.testing.assert_array_equal(nl, g["nllr"])
    np.testing.assert_array_equal(np.concatenate(x_hat), g["xhat"])


def replay(name, n_scans=None, check=None):
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = g["params"]
    trk = mo.OracleTracker(T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd)
    for x in g["init_x"]:
        trk.initiate(x, float(g["init_time"]))
    for k in range(int(g["n_scans"]) if n_scans is None else n_scans):
        pre = "s%d_" % k
        info = trk.add_scan(g[pre + "z"], float(g[pre + "time"]))
        nodes = trk.track_nodes()
        assert [n.tid for n in nodes] == list(g[pre + "ids"]), (name, k)
        hist = g[pre + "hist"]
        for i, n in enumerate(nodes):
            h = n.meas_history()
            assert h == list(hist[i, :len(h)]), (name, k, i)
        np.testing.assert_allclose(np.array([n.x for n in nodes]).reshape(-1, 4), g[pre + "x"], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(np.array([n.P for n in nodes], dtype=float).reshape(-1, 4, 4), g[pre + "P"], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose([n.cnllr for n in nodes], g[pre + "cnllr"], rtol=RTOL, atol=1e-7)
        assert [len(l) for l in trk.leaves] == list(g[pre + "nleaves"]), (name, k)
        assert info["n_clusters"] == int(g[pre + "nclusters"])
        assert len(trk.last_ilp) == int(g[pre + "n_ilp"])
    return trk


@pytest.mark.parametrize("name", ["cfg1_crossing", "cfg2_small", "cfg5_small"])
def test_replay_small(name):
    replay(name)


def test_replay_cfg5_n8():
    """BASELINE config 5's window (N = 8) on the crossing layout: the window fills at scan 8, then N-scan pruning runs."""
    replay("cfg5_n8", n_scans=9)


def test_replay_cfg2():
    replay("cfg2", n_scans=6)


def test_replay_cfg3_head():
    replay("cfg3_head", n_scans=2)


def test_replay_cfg3_lowclutter():
    """1000 targets, 10x less clutter: scans 1-2 (scan 3 needs ~3 min of HiGHS; the GPU test replays it)."""
    replay("cfg3_lowclutter", n_scans=2)


def test_replay_cfg5_full():
    """BASELINE config 5 at full size (500 targets, 50 scripted 90-degree crossings, N = 8): scans 1-4 (scans 5-6
    hold 3.4e4 / 8.6e4 leaves in one cluster -- minutes of HiGHS; the GPU test replays all six)."""
    replay("cfg5_full", n_scans=4)


def test_replay_cfg2_long():
    """BASELINE config 2 for 30 scans: 25 scans of steady state (N-scan pruning, terminations, re-clustering)."""
    replay("cfg2_long", n_scans=30)


@pytest.mark.parametrize("name", ["init_small", "init_dense"])
def test_initiator_oracle_vs_reference_fixture(name):
    """oracle/initiator_oracle.py (restatement of pymht/initiators/m_of_n.py:233-478) against what the unmodified reference's
    initiator returned and held after every scan (oracle/gen_golden.py run_reference_initiator)."""
    from oracle.initiator_oracle import InitiatorOracle
    g = golden(name)
    M, N, vmax, thr, gamma = g["init_params"]
    C = np.zeros((2, 4), np.float32)
    C[0, 0] = C[1, 1] = 1
    o = InitiatorOracle(int(M), int(N), vmax, C, (np.eye(2) * 6.25).astype(np.float32), thr)
    assert abs(o.gamma - gamma) < 1e-12
    for k in range(int(g["n_scans"])):
        pre = "s%d_" % k
        new = o.processMeasurements(g[pre + "ini_z"], float(g[pre + "ini_time"]))
        nx = np.array([n[0] for n in new]).reshape(-1, 4)
        assert nx.shape == g[pre + "new_x"].shape, (name, k)
        assert np.allclose(nx, g[pre + "new_x"], rtol=1e-6, atol=1e-6)
        assert np.allclose(np.array([n[1] for n in new]).reshape(-1, 4, 4), g[pre + "new_P"], rtol=1e-6, atol=1e-6)
        st = np.array([p.state for p in o.preliminary_tracks]).reshape(-1, 4)
        assert st.shape == g[pre + "pt_state"].shape, (name, k)
        assert np.allclose(st, g[pre + "pt_state"], rtol=1e-6, atol=1e-6)
        assert np.array_equal(np.array([[p.m, p.n] for p in o.preliminary_tracks]).reshape(-1, 2), g[pre + "pt_mn"])
        assert np.array_equal(o.initiators, g[pre + "initiators"])


def test_host_merge_of_similar_targets_equals_oracle():
    """pymht_b200.initiators.m_of_n._merge_similar_targets (host part of the product's initiator: pairwise distances taken
    once) against the oracle's restatement of m_of_n.py:117-145 on random clustered initial targets."""
    from oracle.initiator_oracle import InitiatorOracle
    from pymht_b200.initiators import m_of_n
    from pymht_b200.pyTarget import Target
    rng = np.random.RandomState(0)
    for trial in range(100):
        k = rng.randint(1, 40)
        cent = rng.uniform(0, 200, (6, 2))
        x = np.hstack([cent[rng.randint(0, 6, k)] + rng.normal(scale=12, size=(k, 2)), rng.normal(size=(k, 2))])
        tg = [Target(0.0, None, x[i], np.eye(4) * (i + 1), measurement=x[i, :2]) for i in range(k)]
        got = m_of_n._merge_similar_targets(tg, 25.0)
        o = InitiatorOracle(2, 3, 20, np.eye(2, 4), np.eye(2), 25.0)
        want = o._merge([(x[i], np.eye(4) * (i + 1), x[i, :2], i) for i in range(k)])
        assert len(got) == len(want)
        for g_, w_ in zip(got, want):
            assert np.allclose(g_.x_0, w_[0]) and np.allclose(g_.P_0, w_[1])
