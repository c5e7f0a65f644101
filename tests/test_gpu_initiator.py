"""GPU parity of the M-of-N initiator (reference pymht/initiators/m_of_n.py:233-478).

* mht_gnn_assign against the reference's formulation (_solve_global_nearest_neighbour, m_of_n.py:24-104: dense padded
  matrix + optimal assignment, restated in oracle/initiator_oracle.py) on random gated point sets, small components to one
  giant component, both gate modes;
* pymht_b200.initiators.m_of_n.Initiator replaying what the UNMODIFIED reference's initiator was fed and returned
  (tests/golden/init_small.npz, init_dense.npz): initial targets, preliminary tracks and initiators after every scan;
* the whole Tracker with the initiator live against the reference's tracks on the same fixtures (all tracks born by it).
Integer results (assignments, counts, indices) must be identical; float32 states within 1e-5 relative.
"""
import ctypes as C
import time

import numpy as np
import pytest

from conftest import golden
from oracle import initiator_oracle as io

pytestmark = pytest.mark.gpu


def _gnn(rows, cols, edges):
    from pymht_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.mht_gnn_create(rows, cols, edges, C.byref(h)))
    return lib, h


def _assign(lib, h, mode, a, sinv, b, gate):
    from pymht_b200 import _lib
    match = np.empty(len(a), dtype=np.int32)
    info = _lib.GnnInfo()
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    sinv = None if sinv is None else np.ascontiguousarray(sinv, dtype=np.float32)
    _lib.check(lib.mht_gnn_assign(h, mode, len(a), _lib.ptr(a), _lib.ptr(sinv), len(b), _lib.ptr(b), float(gate),
                                  _lib.ptr(match), C.byref(info)))
    return [(i, int(match[i])) for i in range(len(a)) if match[i] >= 0], info.as_dict()


@pytest.mark.parametrize("n1,n2,side,seed", [(40, 40, 400.0, 1), (60, 35, 300.0, 2), (35, 60, 300.0, 3), (300, 280, 530.0, 5),
                                             (280, 300, 300.0, 6), (200, 200, 150.0, 7), (1, 5, 50.0, 8), (5, 1, 50.0, 9),
                                             (50, 50, 5000.0, 10), (400, 120, 250.0, 11), (1500, 1450, 700.0, 12)])
def test_distance_gated_assignment_equals_reference(n1, n2, side, seed):
    """mode 0 = _processInitiators (m_of_n.py:385-401): float64 norm of float32 differences, gate v_max * dt."""
    rng = np.random.RandomState(seed)
    a = rng.uniform(0, side, (n1, 2)).astype(np.float32)
    b = rng.uniform(0, side, (n2, 2)).astype(np.float32)
    delta = np.empty((n1, n2, 2))
    for i in range(n1):
        delta[i] = b - a[i]
    d = np.linalg.norm(delta, axis=2)
    want = io.solve_gnn(d, 50.0)
    lib, h = _gnn(2048, 2048, 1 << 17)
    got, info = _assign(lib, h, 0, a, None, b, 50.0)
    lib.mht_gnn_destroy(h)
    assert info["n_edges"] == int((d <= 50.0).sum())
    assert got == sorted(want), info


@pytest.mark.parametrize("n1,n2,seed", [(50, 80, 21), (300, 300, 22), (700, 500, 23)])
def test_nis_gated_assignment_equals_reference(n1, n2, seed):
    """mode 1 = _processPreliminaryTracks (m_of_n.py:284-303): float32 NIS gate, float32 distance as cost."""
    rng = np.random.RandomState(seed)
    side = 18.0 * np.sqrt(n2)
    zhat = rng.uniform(0, side, (n1, 2)).astype(np.float32)
    z = rng.uniform(0, side, (n2, 2)).astype(np.float32)
    G = rng.normal(size=(n1, 2, 2)).astype(np.float32) * 6
    S = (np.matmul(G, G.transpose(0, 2, 1)) + np.eye(2, dtype=np.float32) * 30).astype(np.float32)
    Si = np.linalg.inv(S)
    delta = np.ones((n1, n2), dtype=np.float32) * np.inf
    for i in range(n1):
        dv = z - zhat[i]
        dist = np.linalg.norm(dv, axis=1)
        nis = np.sum(np.matmul(dv, Si[i]) * dv, axis=1)
        inside = nis <= io.GAMMA
        delta[i, inside] = dist[inside]
    want = io.solve_gnn(delta)
    lib, h = _gnn(1024, 1024, 1 << 16)
    got, info = _assign(lib, h, 1, zhat, Si.reshape(n1, 4), z, io.GAMMA)
    lib.mht_gnn_destroy(h)
    assert info["n_edges"] == int(np.isfinite(delta).sum()), info
    assert got == sorted(want), info


def test_capacity_error_reports_and_recovers():
    from pymht_b200 import _lib
    rng = np.random.RandomState(3)
    a = rng.uniform(0, 100, (200, 2)).astype(np.float32)
    lib, h = _gnn(256, 256, 64)
    match = np.empty(200, dtype=np.int32)
    rc = lib.mht_gnn_assign(h, 0, 200, _lib.ptr(a), None, 200, _lib.ptr(a), 50.0, _lib.ptr(match), None)
    assert rc == _lib.MHT_E_CAPACITY and b"max_edges" in lib.mht_last_error()
    rc = lib.mht_gnn_assign(h, 0, 300, _lib.ptr(a), None, 200, _lib.ptr(a), 50.0, _lib.ptr(match), None)
    assert rc == _lib.MHT_E_CAPACITY
    _lib.check(lib.mht_gnn_assign(h, 0, 50, _lib.ptr(a), None, 50, _lib.ptr(a), 0.0, _lib.ptr(match), None))
    assert list(match[:50]) == list(range(50))      # the handle still works: every point is its own neighbour at distance 0
    lib.mht_gnn_destroy(h)


def test_config3_scale_one_giant_component():
    """BASELINE config 3's clutter: ~4000 unused measurements per scan at 1e-3 / m^2 -> 7.9 candidates inside every 50 m
    gate, ONE connected component of ~4000 x 4000.  Must equal the dense optimum (SciPy on the padded 4000^2 matrix)."""
    rng = np.random.RandomState(4000)

    def disc(k, R=1142.0):
        r, th = R * np.sqrt(rng.uniform(size=k)), rng.uniform(0, 2 * np.pi, k)
        return np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
    a, b = disc(4000), disc(3963)
    delta = np.empty((len(a), len(b), 2))
    for i in range(len(a)):
        delta[i] = b - a[i]
    d = np.linalg.norm(delta, axis=2)
    t0 = time.time()
    want = io.solve_gnn(d, 50.0)
    t_ref = time.time() - t0
    lib, h = _gnn(4096, 4096, 1 << 18)
    got, info = _assign(lib, h, 0, a, None, b, 50.0)
    t0 = time.time()
    got2, info = _assign(lib, h, 0, a, None, b, 50.0)
    t_gpu = time.time() - t0
    lib.mht_gnn_destroy(h)
    print("config-3 scale GNN: %d pairs, largest component %d rows, %d searches / %d rounds; device gate %.2f ms + solve "
          "%.2f ms, call %.1f ms; dense SciPy on the reference's padded matrix %.0f ms" % (
              info["n_edges"], info["largest_component"], info["searches"], info["rounds"], info["ms_gate"],
              info["ms_solve"], 1e3 * t_gpu, 1e3 * t_ref))
    assert info["largest_component"] > 3000
    assert got == got2 == sorted(want)


def _device_initiator(g):
    from pymht_b200.initiators import m_of_n
    from pymht_b200.models import pv
    M, N, vmax, thr, gamma = g["init_params"]
    assert abs(gamma - m_of_n.tracking_parameters["gamma"]) < 1e-12
    return m_of_n.Initiator(int(M), int(N), float(vmax), pv.C_RADAR, pv.R_RADAR(), float(thr))


@pytest.mark.parametrize("name", ["init_small", "init_dense"])
def test_initiator_replays_reference_fixture(name):
    from pymht_b200.utils.classDefinitions import MeasurementList
    g = golden(name)
    ini = _device_initiator(g)
    for k in range(int(g["n_scans"])):
        pre = "s%d_" % k
        new = ini.processMeasurements(MeasurementList(float(g[pre + "ini_time"]), g[pre + "ini_z"]))
        nx = np.array([t.x_0 for t in new], dtype=np.float64).reshape(-1, 4)
        assert nx.shape == g[pre + "new_x"].shape, (name, k)
        np.testing.assert_allclose(nx, g[pre + "new_x"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(np.array([t.P_0 for t in new], dtype=np.float64).reshape(-1, 4, 4), g[pre + "new_P"],
                                   rtol=1e-5, atol=1e-4)
        assert np.array_equal(np.array([t.measurement for t in new], dtype=np.float64).reshape(-1, 2), g[pre + "new_meas"])
        pt = ini.preliminary_tracks
        assert len(pt) == len(g[pre + "pt_state"]), (name, k)
        np.testing.assert_allclose(np.array([p.state for p in pt]).reshape(-1, 4), g[pre + "pt_state"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(np.array([p.covariance for p in pt]).reshape(-1, 4, 4), g[pre + "pt_cov"], rtol=1e-5,
                                   atol=1e-4)
        assert np.array_equal(np.array([[p.m, p.n] for p in pt]).reshape(-1, 2), g[pre + "pt_mn"])
        assert np.array_equal(np.array([i.value for i in ini.initiators], dtype=np.float32).reshape(-1, 2),
                              g[pre + "initiators"])


@pytest.mark.parametrize("name", ["init_small", "init_dense"])
def test_tracker_with_live_initiator_vs_reference(name):
    """addMeasurementList end to end with NO pre-initialised track: every track is born by the initiator from the unused
    measurements the gate stage reports (tracker.py:266-277) -- ids, measurement histories and states of all tracks after
    every scan against the unmodified reference."""
    from pymht_b200.tracker import Tracker, backtrackMeasurementNumbers
    from pymht_b200.models import pv
    from pymht_b200.utils.classDefinitions import MeasurementList
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    trk = Tracker(pv, T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd, maxTargets=512, maxNodes=1 << 18, maxParents=1 << 16,
                  exactBudgetMs=2000)
    for k in range(int(g["n_scans"])):
        pre = "s%d_" % k
        trk.addMeasurementList(MeasurementList(float(g[pre + "time"]), g[pre + "z"]))
        nodes = list(trk.getTrackNodes())
        hist = backtrackMeasurementNumbers(nodes)
        assert trk.scanInfo[-1]["certified"] == 1 or not nodes, (name, k, trk.scanInfo[-1])
        assert [n.ID for n in nodes] == list(g[pre + "ids"]), (name, k)
        H = g[pre + "hist"]
        for i, h in enumerate(hist):
            assert h == list(H[i, :len(h)]) and len(h) == np.sum(H[i] >= 0), (name, k, i, h, H[i])
        np.testing.assert_allclose(np.array([n.x_0 for n in nodes]).reshape(-1, 4), g[pre + "x"], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose([n.cumulativeNLLR for n in nodes], g[pre + "cnllr"], rtol=1e-5, atol=1e-3)
    trk.close()


@pytest.mark.parametrize("seed", range(4))
def test_lattice_points_with_tied_distances(seed):
    """Points on an integer lattice: many pairs at exactly the same distance, so the optimal assignment is not unique -- the
    device result must still have the reference formulation's cardinality and total distance, be a matching inside the gate,
    and come out identical when the call is repeated (no dependence on thread timing)."""
    rng = np.random.RandomState(500 + seed)
    n1, n2 = 600 + 50 * seed, 640
    a = rng.randint(0, 60, (n1, 2)).astype(np.float32) * 10.0
    b = rng.randint(0, 60, (n2, 2)).astype(np.float32) * 10.0
    delta = np.empty((n1, n2, 2))
    for i in range(n1):
        delta[i] = b - a[i]
    d = np.linalg.norm(delta, axis=2)
    want = io.solve_gnn(d, 25.0)
    lib, h = _gnn(1024, 1024, 1 << 17)
    got, info = _assign(lib, h, 0, a, None, b, 25.0)
    got2, _ = _assign(lib, h, 0, a, None, b, 25.0)
    lib.mht_gnn_destroy(h)
    assert got == got2
    assert len(set(j for _, j in got)) == len(got) and all(d[i, j] <= 25.0 for i, j in got)
    assert len(got) == len(want), info
    assert abs(sum(d[i, j] for i, j in got) - sum(d[i, j] for i, j in want)) < 1e-9, info
