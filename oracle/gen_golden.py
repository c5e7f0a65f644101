"""Generate tests/golden/*.npz by running the UNMODIFIED reference under oracle/ref_shim.py.

Run here (the container that has /root/reference):   python -m oracle.gen_golden [names...]
The fixtures are what pins the oracle (oracle/mht_oracle.py) and, through it, the CUDA path.
TEST INFRASTRUCTURE ONLY.

Every fixture stores, per scan: the measurement array fed to the reference, the scan time, and
for every live track after `Tracker.addMeasurementList` returned: Target.ID, the selected leaf's
measurementNumber history over the whole track (helpFunctions.backtrackMeasurementNumbers),
x_0, P_0, cumulativeNLLR, plus leaf counts, cluster count and the reference's stage timers.
"""
import os
import sys
import time

import numpy as np

from oracle import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
T_RADAR = 2.5
T0 = 1000.0


def _scenario(name):
    """(initial states (T,4), n_scans, radarRange, lambda_phi, N, P_d, seed)."""
    rng = np.random.RandomState(99)
    if name == "cfg1_crossing":
        x0 = np.array([[-60.0, 1.0, 6.0, 0.0], [1.0, -60.0, 0.0, 6.0]])
        return x0, 14, 161.6, 1e-4, 3, 0.9, 5446
    if name == "cfg2_small":
        return None, 10, 760.0, 1e-4, 4, 0.9, 1234, 20
    if name == "cfg2":
        return None, 8, 1702.0, 1e-4, 4, 0.9, 1234, 100
    if name == "cfg2_long":
        # BASELINE config 2 run well into its steady state: 30 scans (window N = 4 full from scan 5 on; N-scan pruning,
        # terminations and re-clustering every scan) -- the longest sequence the reference replays in about a minute
        return None, 30, 1702.0, 1e-4, 4, 0.9, 2468, 100
    if name == "cfg3_head":
        return None, 2, 1142.0, 1e-3, 6, 0.9, 1234, 1000
    if name == "cfg3_lowclutter":
        # 1000 targets in the config-3 scene with 10x less clutter: 3 scans are what the reference
        # finishes in minutes (2.5 s, 5.6 s, 330 s); 140 + 140 + 32 multi-tree ILPs
        return None, 3, 1142.0, 1e-4, 6, 0.9, 4321, 1000
    if name == "cfg5_full":
        # BASELINE config 5 at full size: 500 targets = 50 pairs scripted to cross at 90 degrees on
        # the SAME scan (scan 5), less than one gate radius apart for +-2 scans, plus 400 background targets;
        # lambda = 1e-4, R = 1702 m, N = 8, seed 77
        rng5 = np.random.RandomState(77)
        xs = []
        for k in range(50):
            c = rng5.uniform(-1100, 1100, size=2)
            v = 4.0                               # 10 m per scan: within the 19.5 m gate radius for scans 3..7
            tc = 5 * T_RADAR
            xs.append([c[0] - v * tc, c[1] + 2.0, v, 0.0])
            xs.append([c[0] + 2.0, c[1] - v * tc, 0.0, v])
        for k in range(400):
            r = 1300.0 * np.sqrt(rng5.uniform())
            th = rng5.uniform(0, 2 * np.pi)
            a = rng5.uniform(0, 2 * np.pi)
            sp = rng5.choice([0.5, 5.0, 6.0, 7.5])
            xs.append([r * np.cos(th), r * np.sin(th), sp * np.cos(a), sp * np.sin(a)])
        return np.array(xs), 6, 1702.0, 1e-4, 8, 0.9, 77      # scan 7 (2e5 leaves, one cluster) is out of the reference's reach
    if name == "cfg2_dynwin":
        # cfg2_small's scene with addMeasurementList(dynamicWindow=True) and a small targetSizeLimit so that the
        # size criterion of Tracker.__dynamicWindow (tracker.py:918-950) fires; the wall-clock criteria are disabled
        # (limits set out of reach) because they depend on the host's speed
        return None, 12, 760.0, 2e-4, 4, 0.9, 4242, 20
    if name == "cfg3_scan3":
        return None, 3, 1142.0, 1e-3, 6, 0.9, 1234, 1000
    if name == "cfg5_small":
        # 6 crossing pairs + 12 background targets, N=5: dense ILP compatibility rows
        xs = []
        for k in range(6):
            c = rng.uniform(-500, 500, size=2)
            v = 8.0
            tc = 4 * T_RADAR                      # both reach c at scan 4
            xs.append([c[0] - v * tc, c[1] + 3.0, v, 0.0])
            xs.append([c[0] + 2.0, c[1] - v * tc, 0.0, v])
        for k in range(12):
            p = rng.uniform(-600, 600, size=2)
            a = rng.uniform(0, 2 * np.pi)
            xs.append([p[0], p[1], 5 * np.cos(a), 5 * np.sin(a)])
        return np.array(xs), 10, 900.0, 1e-4, 5, 0.9, 77
    if name == "cfg5_n8":
        # the same crossing layout with BASELINE config 5's window N=8 (W = 9 path planes: the wide emit kernel)
        # and enough scans for the window to fill and the N-scan pruning to run
        xs = []
        for k in range(6):
            c = rng.uniform(-500, 500, size=2)
            v = 8.0
            tc = 4 * T_RADAR
            xs.append([c[0] - v * tc, c[1] + 3.0, v, 0.0])
            xs.append([c[0] + 2.0, c[1] - v * tc, 0.0, v])
        for k in range(12):
            p = rng.uniform(-600, 600, size=2)
            a = rng.uniform(0, 2 * np.pi)
            xs.append([p[0], p[1], 5 * np.cos(a), 5 * np.sin(a)])
        return np.array(xs), 12, 900.0, 1e-4, 8, 0.9, 78
    raise KeyError(name)


def run_reference(name):
    ref_shim.install()
    import pymht.utils.simulator as sim
    import pymht.models.pv as pv
    from pymht.utils.classDefinitions import SimTargetCartesian
    import pymht.utils.helpFunctions as hpf

    sc = _scenario(name)
    x0, n_scans, R, lam, N, Pd, seed = sc[:7]
    sim.seed_simulator(seed)
    p0 = np.array([0.0, 0.0])
    if x0 is None:
        init = sim.generateInitialTargets(sc[7], p0, R, Pd, pv.sigmaQ_true)
        for tgt in init:
            tgt.time = T0
    else:
        init = [SimTargetCartesian(np.array(x, dtype=np.float32), T0, Pd, pv.sigmaQ_true) for x in x0]
    simList = sim.simulateTargets(init, n_scans * T_RADAR, T_RADAR, pv)
    scans = sim.simulateScans(simList, T_RADAR, pv.C_RADAR, pv.R_RADAR(pv.sigmaR_RADAR_true), lam, R, p0,
                              shuffle=True, localClutter=False, globalClutter=True, preInitialized=True)
    trk = ref_shim.make_reference_tracker(T_RADAR, lam, 1e-9, N=N, P_d=Pd)
    scan_kw = {}
    if name == "cfg2_dynwin":
        trk.targetSizeLimit = 60
        trk.totalGrowTimeLimit = trk.nodeGrowTimeLimit = 1e9
        trk.radarPeriod = 1e9            # only read by the window-roof test of __dynamicWindow (0.8 * radarPeriod)
        scan_kw = {"dynamicWindow": True}
    trk.preInitialize(simList)
    out = {"init_x": np.array([t.cartesianState() for t in simList[0]], dtype=np.float64),
           "init_time": np.float64(T0), "n_scans": np.int64(len(scans)),
           "params": np.array([T_RADAR, lam, 1e-9, N, Pd, 5.99, R])}
    for k, scan in enumerate(scans):
        t = time.time()
        trk.addMeasurementList(scan, **scan_kw)
        wall = time.time() - t
        nodes = list(trk.getTrackNodes())
        hist = hpf.backtrackMeasurementNumbers(nodes)
        width = max([len(h) for h in hist], default=0)
        H = -np.ones((len(nodes), width), dtype=np.int64)
        for i, h in enumerate(hist):
            H[i, :len(h)] = h
        pre = "s%d_" % k
        out[pre + "z"] = np.asarray(scan.measurements, dtype=np.float32)
        out[pre + "time"] = np.float64(scan.time)
        out[pre + "ids"] = np.array([n.ID for n in nodes], dtype=np.int64)
        out[pre + "hist"] = H
        out[pre + "x"] = np.array([n.x_0 for n in nodes], dtype=np.float64).reshape(len(nodes), 4)
        out[pre + "P"] = np.array([n.P_0 for n in nodes], dtype=np.float64).reshape(len(nodes), 4, 4)
        out[pre + "cnllr"] = np.array([n.cumulativeNLLR for n in nodes], dtype=np.float64)
        out[pre + "nleaves"] = np.array([len(t.getLeafNodes()) for t in trk.__targetList__], dtype=np.int64)
        out[pre + "nclusters"] = np.int64(len(trk.__clusterList__))
        out[pre + "n_ilp"] = np.int64(trk.nOptimSolved)
        out[pre + "toc"] = np.array([trk.toc[s] for s in ("Process", "Cluster", "Optim", "Terminate", "N-Prune")])
        out[pre + "window"] = np.array(trk.__targetWindowSize__, dtype=np.int64)
        print("%s scan %d: M=%d tracks=%d leaves=%d clusters=%d ilps=%d wall=%.2fs" % (
            name, k + 1, len(scan.measurements), len(nodes), int(out[pre + "nleaves"].sum()),
            len(trk.__clusterList__), trk.nOptimSolved, wall), flush=True)
    return out


def _init_scenario(name):
    """(n_targets, n_scans, radarRange, lambda_phi, N, P_d, seed) of the fixtures with the M-of-N initiator LIVE."""
    if name == "init_small":
        # 15 targets, ~80 clutter points per scan: small connected components in both assignment problems
        return 15, 14, 500.0, 1e-4, 4, 0.9, 911
    if name == "init_dense":
        # BASELINE config 3's clutter density (1e-3 / m^2) on a small disc: ~280 clutter points per scan, 7.9 of them
        # inside every initiator's 50 m gate -> the initiator's assignment problem is ONE giant component
        return 8, 10, 300.0, 1e-3, 3, 0.9, 912
    raise KeyError(name)


def run_reference_initiator(name):
    """The unmodified reference with its own M-of-N initiator (pymht/initiators/m_of_n.py) and NO pre-initialised tracks:
    every track is born by the initiator.  Stored per scan: what Initiator.processMeasurements was given (the unused
    measurements), what it returned (initial targets) and its state afterwards (preliminary tracks, initiators), plus the
    tracker's tracks like the other fixtures."""
    ref_shim.install()
    import pymht.utils.simulator as sim
    import pymht.models.pv as pv
    import pymht.tracker as rtracker
    import pymht.utils.helpFunctions as hpf

    nT, n_scans, R, lam, N, Pd, seed = _init_scenario(name)
    sim.seed_simulator(seed)
    p0 = np.array([0.0, 0.0])
    init = sim.generateInitialTargets(nT, p0, R * 0.8, Pd, pv.sigmaQ_true)
    for tgt in init:
        tgt.time = T0
    simList = sim.simulateTargets(init, n_scans * T_RADAR, T_RADAR, pv)
    scans = sim.simulateScans(simList, T_RADAR, pv.C_RADAR, pv.R_RADAR(pv.sigmaR_RADAR_true), lam, R, p0,
                              shuffle=True, localClutter=False, globalClutter=True, preInitialized=False)
    trk = rtracker.Tracker(pv, T_RADAR, lam, 1e-9, N=N, P_d=Pd)
    ini = trk.initiator
    rec = {}
    orig = ini.processMeasurements

    def spy(radar, ais=list()):
        rec["in_z"] = np.array(radar.measurements, dtype=np.float32).reshape(-1, 2)
        rec["in_time"] = radar.time
        out = orig(radar, ais)
        rec["out"] = out
        return out
    ini.processMeasurements = spy
    out = {"n_scans": np.int64(len(scans)), "params": np.array([T_RADAR, lam, 1e-9, N, Pd, 5.99, R]),
           "init_params": np.array([ini.M, ini.N, ini.v_max, ini.merge_threshold, ini.gamma])}
    for k, scan in enumerate(scans):
        t = time.time()
        trk.addMeasurementList(scan)
        wall = time.time() - t
        nodes = list(trk.getTrackNodes())
        hist = hpf.backtrackMeasurementNumbers(nodes)
        width = max([len(h) for h in hist], default=0)
        H = -np.ones((len(nodes), width), dtype=np.int64)
        for i, h in enumerate(hist):
            H[i, :len(h)] = h
        pre = "s%d_" % k
        out[pre + "z"] = np.asarray(scan.measurements, dtype=np.float32)
        out[pre + "time"] = np.float64(scan.time)
        out[pre + "ids"] = np.array([n.ID for n in nodes], dtype=np.int64)
        out[pre + "hist"] = H
        out[pre + "x"] = np.array([n.x_0 for n in nodes], dtype=np.float64).reshape(len(nodes), 4)
        out[pre + "cnllr"] = np.array([n.cumulativeNLLR for n in nodes], dtype=np.float64)
        out[pre + "ini_z"] = rec["in_z"]
        out[pre + "ini_time"] = np.float64(rec["in_time"])
        new = rec["out"]
        out[pre + "new_x"] = np.array([t_.x_0 for t_ in new], dtype=np.float64).reshape(len(new), 4)
        out[pre + "new_P"] = np.array([t_.P_0 for t_ in new], dtype=np.float64).reshape(len(new), 4, 4)
        out[pre + "new_meas"] = np.array([t_.measurement for t_ in new], dtype=np.float64).reshape(len(new), 2)
        pt = ini.preliminary_tracks
        out[pre + "pt_state"] = np.array([p_.state for p_ in pt], dtype=np.float64).reshape(len(pt), 4)
        out[pre + "pt_cov"] = np.array([p_.covariance for p_ in pt], dtype=np.float64).reshape(len(pt), 4, 4)
        out[pre + "pt_mn"] = np.array([[p_.m, p_.n] for p_ in pt], dtype=np.int64).reshape(len(pt), 2)
        out[pre + "initiators"] = np.array([i_.value for i_ in ini.initiators], dtype=np.float32).reshape(-1, 2)
        out[pre + "toc_init"] = np.float64(trk.toc["Init"])
        print("%s scan %d: M=%d unused=%d new=%d prelim=%d initiators=%d tracks=%d init=%.2fs wall=%.2fs" % (
            name, k + 1, len(scan.measurements), len(rec["in_z"]), len(new), len(pt), len(ini.initiators),
            len(nodes), trk.toc["Init"], wall), flush=True)
    return out


def kalman_kat():
    """Outputs of the reference's own kalman.py functions on seeded random leaves."""
    ref_shim.install()
    import pymht.utils.kalman as rk
    import pymht.models.pv as pv
    rng = np.random.RandomState(4242)
    L, M = 48, 300
    A, Q, C, R = pv.Phi(T_RADAR), pv.Q(T_RADAR), pv.C_RADAR, pv.R_RADAR()
    x0 = np.concatenate([rng.uniform(-200, 200, (L, 2)), rng.uniform(-10, 10, (L, 2))], axis=1)
    G = rng.normal(size=(L, 4, 4)) * np.array([2.0, 2.0, 1.0, 1.0])[None, :, None]
    P0 = (np.matmul(G, G.transpose(0, 2, 1)) + np.diag([6.25, 6.25, 1.9, 1.9])).astype(np.float32)
    P0[: L // 3] = pv.P0
    z = (x0[rng.randint(0, L, M), :2] + x0[rng.randint(0, L, M), 2:] * T_RADAR * rng.uniform(0, 2, (M, 1))
         + rng.normal(scale=6.0, size=(M, 2))).astype(np.float32)
    x_bar, P_bar = rk.predict(A, Q, x0, P0)
    z_hat, S, S_inv, K, P_hat = rk.precalc(C, R, x_bar, P_bar)
    zt = rk.z_tilde(z, z_hat, L, 2)
    d2 = rk.normalizedInnovationSquared(zt, S_inv)
    gate = d2 <= 5.99
    Pd, lam = 0.9, 1e-4 + 1e-9
    pair_leaf, pair_meas = np.nonzero(gate)
    nllr = np.concatenate([rk.nllr(lam, Pd, S[i], d2[i, gate[i]]) for i in range(L)])
    xhat = np.concatenate([rk.numpyFilter(x_bar[i], K[i], zt[i, gate[i]]) for i in range(L)])
    return dict(x0=x0, P0=P0, z=z, x_bar=x_bar, P_bar=P_bar, z_hat=z_hat, S=S, S_inv=S_inv, K=K, P_hat=P_hat,
                d2=d2, pair_leaf=pair_leaf, pair_meas=pair_meas, nllr=nllr, xhat=xhat,
                params=np.array([T_RADAR, lam, Pd, 5.99]))


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv or ["kalman_kat", "cfg1_crossing", "cfg2_small", "cfg5_small", "cfg2", "cfg3_head"]
    for name in names:
        if name == "kalman_kat":
            data = kalman_kat()
        elif name.startswith("init_"):
            data = run_reference_initiator(name)
        else:
            data = run_reference(name)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1:])
