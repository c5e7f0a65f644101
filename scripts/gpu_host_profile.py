"""Where does the host time of Tracker.addMeasurementList go?  (run on the GPU box)"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
name = "cfg3_1k_targets_5k_meas_N6"
simList, scans = bench.make_scenario(name, 14)
trk = bench.make_tracker(name)
trk.preInitialize(simList)
for s in scans[:9]:
    trk.addMeasurementList(s)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in scans[9:]:
    trk.addMeasurementList(s)
pr.disable()
wall = time.perf_counter() - t0
dev = sum(d["ms_total"] for d in trk.scanInfo[9:]) * 1e-3
print("wall %.1f ms/scan, device %.1f ms/scan" % (1e3 * wall / 5, 1e3 * dev / 5))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
