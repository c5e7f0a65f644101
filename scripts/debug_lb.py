"""Repeat the scenario of tests/test_gpu_parity.py::test_large_random_forest_properties and report every scan whose lower
bound exceeds its objective, with a feasibility check of the selection over the whole window."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pymht_b200.tracker import Tracker, backtrackMeasurementNumbers
from pymht_b200.models import pv
import pymht_b200.utils.simulator as sim

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bad = 0
for rep in range(reps):
    sim.seed_simulator(7)
    R, lam, nT = 1142.0, 1e-3, 1000
    init = sim.generateInitialTargets(nT, np.zeros(2), R, 0.9, 1.0)
    simList = sim.simulateTargets(init, 5 * 2.5, 2.5, pv)
    scans = sim.simulateScans(simList, 2.5, pv.C_RADAR, pv.R_RADAR(), lam, R, np.zeros(2), preInitialized=True)
    trk = Tracker(pv, 2.5, lam, 1e-9, N=6, P_d=0.9, initiator=None, maxTargets=1024, maxNodes=1 << 23, maxParents=1 << 21)
    trk.mergeThreshold = 0.0
    trk.preInitialize(simList)
    for k, scan in enumerate(scans[:5]):
        trk.addMeasurementList(scan)
        info = trk.scanInfo[-1]
        nodes = trk.getTrackNodes()
        hist = backtrackMeasurementNumbers(nodes)
        conflicts = 0
        for back in range(1, 7):
            used = [h[-back] for h in hist if len(h) >= back and h[-back] > 0]
            conflicts += len(used) - len(set(used))
        flag = info["lower_bound"] > info["objective"] + 1e-6
        bad += flag
        if flag or conflicts or info["repaired_trees"] or rep == 0:
            print("rep %d scan %d: lb %.6f obj %.6f cert %d open %d comps %d maxc %d nodes %d iters %d cand %d conflicts %d repaired %d %s" % (
                rep, k + 1, info["lower_bound"], info["objective"], info["certified"], info["open_components"],
                info["n_components"], info["max_component"], info["bb_nodes"], info["dual_iters"], info["n_candidates"], conflicts,
                info["repaired_trees"], "<-- LB > OBJ" if flag else ""), flush=True)
    trk.close()
print("bad scans:", bad)
