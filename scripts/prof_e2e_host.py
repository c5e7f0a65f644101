"""cProfile of the host side of Tracker.addMeasurementList over the bench's timed scans (where the e2e leg's ~1 ms per scan on
top of the device time goes)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = "cfg3_1k_targets_5k_meas_N6"
n = 33
simList, scans = bench.make_scenario(name, n)
trk = bench.make_tracker(name, n)
trk.preInitialize(simList)
for s in scans[:13]:
    trk.addMeasurementList(s)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in scans[13:]:
    trk.addMeasurementList(s)
pr.disable()
dt = time.perf_counter() - t0
dev = sum(i["ms_total"] for i in trk.scanInfo[13:])
print("20 scans: wall %.1f ms, device %.1f ms, host overhead %.2f ms per scan" % (1e3 * dt, dev, (1e3 * dt - dev) / 20))
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
