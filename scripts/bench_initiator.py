"""Measure the M-of-N initiator at BASELINE config 3's scale on the GPU box, with the reference's own
initiator (baseline/_ref pymht/initiators/m_of_n.py under oracle/ref_shim.py, 1 host core) timed beside it on a bounded
sample.  Prints one JSON line (kept under profiles/initiator_r2.json).

Input of both: the measurements of every scan that NO track gated -- the `used` mask the gate stage returns
(tracker.py:266-277) -- of the cfg3 bench scenario (1000 targets, ~5000 measurements per scan from the reference's
simulator), i.e. ~4000 clutter points per scan at 1e-3 / m^2.  The initiator runs beside the tracker (its births are not
fed back, so the tracked scene is the bench's).
  * GPU arm: every scan, full input.
  * reference arm: its per-scan cost grows like n^3 (dense Munkres on the padded n x n matrix) plus an all-pairs Python loop
    with a 4x4 inverse per pair (m_of_n.py:462-470) -- minutes per scan at n = 4000 -- so it gets the FIRST `--sample`
    unused measurements of each of `--ref-scans` scans; the GPU arm is timed on that same sub-sample as well, and the
    outputs (initial targets, preliminary tracks) are compared.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def unused_lists(name, n_scans):
    from pymht_b200.utils.classDefinitions import MeasurementList
    simList, scans = bench.make_scenario(name, n_scans)
    trk = bench.make_tracker(name, n_scans)
    trk.preInitialize(simList)
    out = []
    for sc in scans:
        orig = trk.initiator.processMeasurements
        box = {}

        def spy(unusedRadar, ais, box=box):
            box["z"] = np.asarray(unusedRadar.measurements, dtype=np.float32).reshape(-1, 2)
            return []
        trk.initiator.processMeasurements = spy
        trk.addMeasurementList(sc)
        trk.initiator.processMeasurements = orig
        out.append(MeasurementList(sc.time, box["z"]))
    trk.close()
    return out


def run_gpu(lists):
    from pymht_b200.initiators import m_of_n
    from pymht_b200.models import pv
    ini = m_of_n.Initiator(2, 3, 20, pv.C_RADAR, pv.R_RADAR(), 25.0, maxMeasurements=8192)
    rows = []
    for ml in lists:
        t = time.perf_counter()
        new = ini.processMeasurements(ml)
        ms = 1e3 * (time.perf_counter() - t)
        rows.append({"unused": len(ml.measurements), "ms": round(ms, 3), "new_targets": len(new),
                     "preliminary": len(ini._state), "initiators": len(ini._init_z),
                     "tracks": dict(ini.last_info.get("tracks", {})), "initiators_gnn": dict(ini.last_info.get("initiators", {}))})
    return ini, rows


def run_reference(lists):
    from oracle import ref_shim
    ref_shim.install()
    import pymht.initiators.m_of_n as rm
    import pymht.models.pv as rpv
    from pymht.utils.classDefinitions import MeasurementList as RML
    ini = rm.Initiator(2, 3, 20, rpv.C_RADAR, rpv.R_RADAR(), 25.0)
    rows, outs = [], []
    for ml in lists:
        t = time.perf_counter()
        new = ini.processMeasurements(RML(ml.time, np.asarray(ml.measurements, dtype=np.float32)))
        rows.append({"unused": len(ml.measurements), "ms": round(1e3 * (time.perf_counter() - t), 1), "new_targets": len(new),
                     "preliminary": len(ini.preliminary_tracks)})
        outs.append((np.array([t_.x_0 for t_ in new]).reshape(-1, 4),
                     np.array([p.state for p in ini.preliminary_tracks]).reshape(-1, 4)))
    return rows, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=8)
    ap.add_argument("--ref-scans", type=int, default=3)
    ap.add_argument("--sample", type=int, default=1200)
    args = ap.parse_args()
    from pymht_b200.utils.classDefinitions import MeasurementList
    name = "cfg3_1k_targets_5k_meas_N6"
    lists = unused_lists(name, args.scans)
    run_gpu(lists[:2])                                      # warm-up: module load, buffers
    _, full = run_gpu(lists)
    sub = [MeasurementList(ml.time, np.asarray(ml.measurements)[:args.sample]) for ml in lists[:args.ref_scans]]
    _, gpu_sub = run_gpu(sub)
    ref_rows, ref_out = run_reference(sub)
    # same inputs, same results?
    from pymht_b200.initiators import m_of_n
    from pymht_b200.models import pv
    chk = m_of_n.Initiator(2, 3, 20, pv.C_RADAR, pv.R_RADAR(), 25.0, maxMeasurements=8192)
    same = True
    for ml, (rx, rp) in zip(sub, ref_out):
        new = chk.processMeasurements(ml)
        gx = np.array([t.x_0 for t in new]).reshape(-1, 4)
        same &= gx.shape == rx.shape and np.allclose(gx, rx, rtol=1e-5, atol=1e-3)
        same &= chk._state.shape == rp.shape and np.allclose(chk._state, rp, rtol=1e-5, atol=1e-3)
    steady = full[2:]
    line = {
        "what": "M-of-N initiator per scan, config 3 scale (unused measurements of the cfg3 bench scans)",
        "gpu_full": {"ms_per_scan_mean": round(float(np.mean([r["ms"] for r in steady])), 2),
                     "ms_per_scan_max": round(float(np.max([r["ms"] for r in steady])), 2),
                     "unused_per_scan": int(np.mean([r["unused"] for r in steady])), "scans": full},
        "sample": "first %d unused measurements of scans 1-%d" % (args.sample, args.ref_scans),
        "gpu_sample_ms": [r["ms"] for r in gpu_sub],
        "reference_sample_ms": [r["ms"] for r in ref_rows],
        "reference": "baseline/_ref pymht/initiators/m_of_n.py, munkres -> scipy.optimize.linear_sum_assignment (oracle/ref_shim.py), "
                     "1 core of %d" % len(os.sched_getaffinity(0)),
        "identical_results_on_sample": bool(same),
        "speedup_on_sample_last_scan": round(ref_rows[-1]["ms"] / max(gpu_sub[-1]["ms"], 1e-9), 1),
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
