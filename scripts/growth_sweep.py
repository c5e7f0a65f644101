"""Hypothesis-count distribution at config 3 over several seeds (capacity planning)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pymht_b200 import _lib
name = "cfg3_1k_targets_5k_meas_N6"
bench.WORKLOADS[name] = bench.WORKLOADS[name][:6] + (200 << 20, 48 << 20)
for off in range(int(sys.argv[1]), int(sys.argv[2])):
    simList, scans = bench.make_scenario(name, 24, seed_offset=off)
    trk = bench.make_tracker(name)
    trk.preInitialize(simList)
    ch, pa = [], []
    for s in scans:
        try:
            trk.addMeasurementList(s)
        except _lib.MhtError as e:
            print("seed", off, "ERR", e); break
        ch.append(trk.scanInfo[-1]["n_children"]); pa.append(trk.scanInfo[-1]["n_parents"])
    print("seed+%d children max %.3g  parents max %.3g  last8 children %s" % (off, max(ch), max(pa), ["%.2g" % c for c in ch[-8:]]), flush=True)
    trk.close()
