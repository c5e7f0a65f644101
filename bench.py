#!/usr/bin/env python
"""bench.py -- scans/sec of the per-scan hot path at BASELINE config 3
(1k targets, ~5k measurements/scan, lambda=1e-3, N-scan=6, CV model), on N GPUs of one node.

A "step" is one Tracker.addMeasurementList hot path (gate+NLLR -> cluster -> global-hypothesis
0/1 program -> terminate -> N-scan prune) over one synthetic scan, at steady state (the forest is
pre-rolled N+2 scans so the hypothesis trees have reached their windowed size).

  value : device-timed (CUDA events on the forest's stream, inside libmht_b200), scan already in HBM
  e2e   : same scans through pymht_b200.Tracker.addMeasurementList with HOST measurement arrays
          (H2D of the scan and D2H of the per-track results inside the timed region), wall clock
  --impl reference : the reference's own CPU path -- the UNMODIFIED reference installed at baseline/_ref when it is
          there (driven through its public API under oracle/ref_shim.py), else the oracle port
          (oracle/mht_oracle.py) -- on a bounded sample (the first scans of the same scenario).
N > 1: by default every rank owns one independent surveillance sector (its own 1k-target forest, weak scaling);
the only exchange is an all_gather of the per-rank track summaries.  `--shard trees` instead shards the trees of
ONE region over the ranks and all-gathers the column records of the 0/1 program (pymht_b200/sharded.py).
stdout carries exactly one JSON line; everything else goes to stderr.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (targets, radarRange, lambda_phi, N, P_d, seed, maxNodes, maxParents)
    "cfg3_1k_targets_5k_meas_N6": (1000, 1142.0, 1e-3, 6, 0.9, 1234, 144 << 20, 32 << 20),
    "cfg2_100_targets_1k_meas_N4": (100, 1702.0, 1e-4, 4, 0.9, 1234, 1 << 22, 1 << 20),
    # BASELINE config 4 (CV model; the coordinated-turn variant has no reference): 10x config 3.  Its forest
    # peaks near 8e8 hypotheses per level while the window fills -- more than one GPU holds -- so it only runs tree
    # sharded (--shard trees, 8 GPUs); the per-rank capacities below are the whole-region ones, divided by the ranks
    "cfg4_10k_targets_50k_meas_N6": (10000, 3613.0, 1e-3, 6, 0.9, 1234, 1200 << 20, 280 << 20),
}
T_RADAR = 2.5


SCENARIO_SOURCE = "pymht_b200.utils.simulator (own generator)"


def make_scenario(name, n_scans, seed_offset=0):
    """Scans of the named workload.  BASELINE.md section 3: the reference's own simulator (seed_simulator(seed) ->
    generateInitialTargets -> simulateTargets -> simulateScans, pymht/utils/simulator.py:15-110), taken from the
    UNMODIFIED reference installed at baseline/_ref when it is there (same scans for both arms); the repo's own
    generator otherwise.  Returned in this repo's container classes."""
    global SCENARIO_SOURCE
    nT, R, lam, N, Pd, seed, _, _ = WORKLOADS[name]
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "pymht")) and not os.environ.get("MHT_BENCH_OWN_SIM"):
        os.environ["PYMHT_REFERENCE_ROOT"] = ref_root
        from oracle import ref_shim          # input generation only: nothing timed runs through oracle/
        ref_shim.install()
        import pymht.utils.simulator as rsim
        import pymht.models.pv as rpv
        from pymht_b200.utils.classDefinitions import MeasurementList, ScanList, SimList, SimTargetCartesian
        rsim.seed_simulator(seed + seed_offset)
        p0 = np.zeros(2)
        init = rsim.generateInitialTargets(nT, p0, R, Pd, rpv.sigmaQ_true)
        rl = rsim.simulateTargets(init, n_scans * T_RADAR, T_RADAR, rpv)
        rscans = rsim.simulateScans(rl, T_RADAR, rpv.C_RADAR, rpv.R_RADAR(rpv.sigmaR_RADAR_true), lam, R, p0,
                                    shuffle=True, localClutter=False, globalClutter=True, preInitialized=True)
        simList = SimList()
        for step in rl:
            simList.append([SimTargetCartesian(np.asarray(t.cartesianState(), dtype=np.float64), t.time, Pd,
                                               rpv.sigmaQ_true) for t in step])
        scans = ScanList()
        for sc in rscans[:n_scans]:
            scans.append(MeasurementList(sc.time, np.asarray(sc.measurements, dtype=np.float32).reshape(-1, 2)))
        SCENARIO_SOURCE = "reference simulator (baseline/_ref pymht/utils/simulator.py, seed %d)" % (seed + seed_offset)
        return simList, scans
    import pymht_b200.utils.simulator as sim
    from pymht_b200.models import pv
    sim.seed_simulator(seed + seed_offset)
    p0 = np.zeros(2)
    init = sim.generateInitialTargets(nT, p0, R, Pd, pv.sigmaQ_true)
    simList = sim.simulateTargets(init, n_scans * T_RADAR, T_RADAR, pv)
    scans = sim.simulateScans(simList, T_RADAR, pv.C_RADAR, pv.R_RADAR(pv.sigmaR_RADAR_true), lam, R, p0,
                              shuffle=True, globalClutter=True, preInitialized=True)
    return simList, scans[:n_scans]


def meas_capacity(name):
    nT, R, lam = WORKLOADS[name][:3]
    need = int(1.25 * (nT + lam * np.pi * R * R)) + 1024
    return 8192 if need <= 8192 else 1 << int(np.ceil(np.log2(need)))


def capacities(name, n_scans):
    """(maxNodes, maxParents): the steady-state sizes of the table, or what a short cold-start run needs (the forest
    grows ~x5 per scan from nT leaves until the window fills)."""
    nT, _, _, N, _, _, max_nodes, max_par = WORKLOADS[name]
    if n_scans <= N:
        need = int(nT * 3.2 * 7.5 ** max(n_scans - 1, 0) * 2.0)   # x5-7 per scan (the 10k-target scene is denser at its centre)
        return min(max_nodes, max(1 << 22, need)), min(max_par, max(1 << 20, need // 2))
    return max_nodes, max_par


def metric_name(name):
    nT, R, lam = WORKLOADS[name][:3]
    if name.startswith("cfg3"):
        return "scans/sec @ 1k targets, 5k meas/scan"
    return "scans/sec @ %dk targets, %dk meas/scan" % (nT // 1000, int(round((nT * 0.9 + lam * np.pi * R * R) / 1000.0))) \
        if nT >= 1000 else "scans/sec @ %d targets, %d meas/scan" % (nT, int(nT * 0.9 + lam * np.pi * R * R))


def make_tracker(name, n_scans=1 << 30):
    from pymht_b200.tracker import Tracker
    from pymht_b200.models import pv
    nT, R, lam, N, Pd, seed, max_nodes, max_par = WORKLOADS[name]
    max_nodes, max_par = capacities(name, n_scans)
    trk = Tracker(pv, T_RADAR, lam, 1e-9, N=N, P_d=Pd, initiator=None, maxTargets=max(1024, nT + 24), maxMeasurements=meas_capacity(name),
                  maxNodes=max_nodes, maxParents=max_par,
                  maxDualIterations=int(os.environ.get("MHT_DUAL_ITERS", "120")))
    trk.mergeThreshold = 0.0
    return trk


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gate_bytes(info):
    """B_gate = 280 L + 8 M + 48 G  (reference dtypes: f64 state, f32 covariance)."""
    return 280.0 * info["n_parents"] + 8.0 * info["n_meas"] + 48.0 * info["n_pairs"]


def run_device_leg(name, scans, simList, preroll, warmup, steps):
    """Timed with the library's CUDA events (first kernel -> results on host), scans resident in HBM."""
    import torch
    from pymht_b200 import _lib
    trk = make_tracker(name, len(scans))
    trk.preInitialize(simList)
    lib = trk._lib
    dev = torch.device("cuda", torch.cuda.current_device())
    d_scans = [torch.from_numpy(np.ascontiguousarray(s.measurements, dtype=np.float64)).to(dev) for s in scans]
    torch.cuda.synchronize()
    infos = []
    for k, (s, dz) in enumerate(zip(scans, d_scans)):
        info = _lib.ScanInfo()
        l0 = int(lib.mht_launch_count())
        _lib.check(lib.mht_forest_scan_device(trk._forest, dz.shape[0], dz.data_ptr(), float(s.time), C.byref(info)),
                   allow=(_lib.MHT_E_NOTOPTIMAL,))
        d = info.as_dict()
        d["n_meas"] = int(dz.shape[0])
        d["launches"] = int(lib.mht_launch_count()) - l0     # counted by the library at every launch site
        infos.append(d)
    timed = infos[preroll + warmup:preroll + warmup + steps]
    dev_bytes = trk.deviceBytes()
    trk.close()
    return timed, infos, dev_bytes


def run_e2e_leg(name, scans, simList, preroll, warmup, steps):
    trk = make_tracker(name, len(scans))
    trk.preInitialize(simList)
    t_steps = []
    for k, s in enumerate(scans):
        t0 = time.perf_counter()
        trk.addMeasurementList(s)
        t_steps.append(time.perf_counter() - t0)
    timed = t_steps[preroll + warmup:preroll + warmup + steps]
    h2d = int(np.mean([16 * len(s.measurements) for s in scans[preroll + warmup:]]))
    n_tracks = len(trk.getTrackNodes())
    trk.close()
    return timed, h2d, n_tracks, t_steps


def run_cpu_reference(name, max_scans=2, n_scans_scenario=None):
    """The reference's own CPU path on the first scans of the same workload, cold start.

    Preferred: the UNMODIFIED reference installed at baseline/_ref (pip --no-deps --target, see
    DESIGN.md) driven through its public API (Tracker.preInitialize / addMeasurementList) under
    oracle/ref_shim.py, which only stands in for absent third-party modules (matplotlib, termcolor,
    munkres; OR-Tools CBC -> SciPy HiGHS) and nulls the out-of-scope M-of-N initiator.  Fallback: the
    oracle port (oracle/mht_oracle.py).  Returns (kind, per-scan seconds of Process+Cluster+Optim+
    Terminate+N-Prune, leaves after each scan)."""
    nT, R, lam, N, Pd, seed, _, _ = WORKLOADS[name]
    # the scenario is generated at the GPU arm's length so that scans 1..max_scans are the SAME scans in both arms
    simList, scans = make_scenario(name, n_scans_scenario or max_scans)
    scans = scans[:max_scans]
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "pymht")):
        os.environ["PYMHT_REFERENCE_ROOT"] = ref_root
        from oracle import ref_shim
        trk = ref_shim.make_reference_tracker(T_RADAR, lam, 1e-9, N=N, P_d=Pd)
        from pymht.utils.classDefinitions import MeasurementList as RefScan
        trk.preInitialize(simList)
        times, leaves = [], []
        for s in scans:
            trk.addMeasurementList(RefScan(s.time, np.asarray(s.measurements)))
            times.append(sum(trk.toc[k] for k in ("Process", "Cluster", "Optim", "Terminate", "N-Prune")))
            leaves.append(int(sum(len(t.getLeafNodes()) for t in trk.__targetList__)))
        return "reference", times, leaves
    from oracle import mht_oracle as mo
    orc = mo.OracleTracker(T_RADAR, lam, 1e-9, N=N, P_d=Pd)
    for tgt in simList[0]:
        orc.initiate(np.asarray(tgt.cartesianState(), dtype=np.float64), tgt.time)
    times, leaves = [], []
    for s in scans:
        t0 = time.perf_counter()
        orc.add_scan(s.measurements, s.time)
        times.append(time.perf_counter() - t0)
        leaves.append(int(sum(len(l) for l in orc.leaves)))
    return "port", times, leaves


def cpu_sample_text(kind, times, leaves):
    return ("%s, scans 1-%d of the same scenario from a cold start (leaves after each scan: %s; seconds per scan: %s); "
            "later scans are out of CPU reach (the reference needs ~160 s for scan 3 and hours beyond)" % (
                "unmodified reference (baseline/_ref, HiGHS in place of OR-Tools CBC)" if kind == "reference"
                else "oracle port of the reference", len(times), leaves, [round(t, 2) for t in times]))


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner,
    library prints of the reference arm) is sent to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3_1k_targets_5k_meas_N6", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preroll", type=int, default=-1,
                    help="untimed scans before the warm-up (default N+2 = the window is full); 0 = time the cold start")
    ap.add_argument("--shard", default="sectors", choices=["sectors", "trees"],
                    help="N>1: 'sectors' = one independent region per rank (weak scaling, default); 'trees' = the "
                         "trees of ONE region sharded over the ranks, column records all-gathered (strong scaling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    nT, R, lam, N, Pd, seed, _, _ = WORKLOADS[name]
    preroll = N + 2 if args.preroll < 0 else args.preroll
    if name.startswith("cfg4") and not (args.shard == "trees" and world > 1) and preroll + args.warmup + args.steps > 4:
        raise SystemExit("cfg4 (10k targets) exceeds one GPU's HBM while the window fills: run it with --shard trees "
                         "under torchrun on 8 GPUs, or time the cold start (--preroll 0 --warmup 1 --steps 3)")
    config = {"workload": name, "targets": nT, "meas_per_scan": "~%d" % int(nT * Pd + lam * np.pi * R * R),
              "lambda_phi": lam, "n_scan": N, "P_d": Pd, "model": "CV (pv)", "radar_period_s": T_RADAR,
              "l2": "per-scan working set (hypothesis levels, GBs) exceeds the 126 MB L2; no explicit flush",
              "parallelism": "1 forest per GPU; %d independent sector(s)" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        kind, times, leaves = run_cpu_reference(name, max_scans=2, n_scans_scenario=preroll + args.warmup + args.steps)
        v = len(times) / sum(times)
        config["scans"] = SCENARIO_SOURCE
        line = {"metric": metric_name(name), "value": v, "unit": "scans/s", "n_gpus": args.gpus,
                "steps": len(times), "warmup": 0, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 state / f32 covariance", "data": "synthetic", "config": config,
                "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "scans/s", "cores": 1, "kind": kind,
                                 "sample": cpu_sample_text(kind, times, leaves)},
                "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its version banner / debug lines to stdout by default: send them to stderr so that stdout
        # carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_scans = preroll + args.warmup + args.steps
    config["preroll_scans"] = preroll
    if args.shard == "trees" and world > 1:
        return run_tree_sharded(args, name, config, dist, torch, rank, world, local_rank, preroll, n_scans)
    simList, scans = make_scenario(name, n_scans, seed_offset=rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    timed, infos, dev_bytes = run_device_leg(name, scans, simList, preroll, args.warmup, args.steps)
    barrier()
    if os.environ.get("MHT_BENCH_SKIP_E2E"):      # profiling runs (ncu) only: the line they print is never a bench value
        e2e_times, h2d, n_tracks, e2e_all = [1.0] * args.steps, 0, 0, [1.0] * n_scans
    else:
        e2e_times, h2d, n_tracks, e2e_all = run_e2e_leg(name, scans, simList, preroll, args.warmup, args.steps)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0 and os.environ.get("MHT_BENCH_VERBOSE"):
        for k, d in enumerate(infos):
            sys.stderr.write("scan %2d trees %4d L %8d G %9d cand %8d comps %3d maxc %4d open %2d cert %d iters %4d "
                             "nodes %6d | gate %.2f assoc %.2f dual %.2f exact %.2f total %.2f ms | gap %.3f\n" % (
                                 k + 1, d["n_trees"], d["n_parents"], d["n_pairs"], d["n_candidates"], d["n_components"],
                                 d["max_component"], d["open_components"], d["certified"], d["dual_iters"], d["bb_nodes"],
                                 d["ms_gate"], d["ms_assoc"], d["ms_dual"], d["ms_exact"], d["ms_total"],
                                 d["objective"] - d["lower_bound"]))
    t_dev = sum(d["ms_total"] for d in timed) * 1e-3
    t_e2e = sum(e2e_times)
    if dist is not None:
        t_dev, t_e2e, sector_tracks = exchange(dist, torch.device("cuda", local_rank), t_dev, t_e2e, n_tracks)
    if rank == 0:
        K = len(timed)
        value = world * K / t_dev
        e2e = world * K / t_e2e
        config["scans"] = SCENARIO_SOURCE
        ms_gate = float(np.mean([d["ms_gate"] for d in timed]))
        bytes_gate = float(np.mean([gate_bytes(d) for d in timed]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_gate / (ms_gate * 1e-3) / 1e9
        traffic = None   # measured DRAM bytes of the gate stage (ncu capture in profiles/), scaled by children
        for fn in ("traffic_r2.json", "traffic_r1.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", fn)))["gate_stage"]
                traffic = tr["dram_bytes"] / tr["n_children"] * float(np.mean([d["n_children"] for d in timed]))
                break
            except Exception:
                pass

        def pct(key, src=timed):
            v = np.array([d[key] for d in src], dtype=np.float64)
            return {"p50": float(np.percentile(v, 50)), "p95": float(np.percentile(v, 95)), "max": float(v.max()),
                    "min": float(v.min()), "mean": float(v.mean())}

        # ILP roofline: bytes per dual iteration = 8 nnz + 24 (cols + rows) on the columns the loop
        # iterates on, against the measured time per iteration of the persistent dual-loop kernel
        it_bytes = float(np.mean([8.0 * d["nnz_active"] + 24.0 * ((d["n_active"] or d["n_children"]) + d["rows_active"])
                                  for d in timed]))
        iters = float(np.mean([max(d["dual_iters"], 1) for d in timed]))
        ms_dual = float(np.mean([d["ms_dual"] for d in timed]))
        us_iter = 1e3 * ms_dual / iters
        ilp_ach = it_bytes / (us_iter * 1e-6) / 1e9 if us_iter > 0 else 0.0
        gaps = [d["objective"] - d["lower_bound"] for d in timed]
        line = {
            "metric": metric_name(name), "value": value, "unit": "scans/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 state / f32 covariance", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e, "unit": "scans/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(infos[-1]["n_trees"] * 152 + 128)},
            "gpu_launches": int(sum(d["launches"] for d in timed)),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "gate stage (forest_gate_kernel + forest_gate_heavy_kernel + forest_emit_kernel)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                         "bytes_per_launch": bytes_gate, "ms_per_launch": ms_gate},
            "roofline_ilp": {"bound": "hbm", "kernel": "dual_loop_cluster_kernel (one projected-subgradient iteration; grid fallback dual_loop_persistent_kernel)",
                             "bytes_per_iteration": it_bytes, "us_per_iteration": us_iter, "iterations_per_scan": iters,
                             "ms_dual_per_scan": ms_dual, "achieved": ilp_ach, "peak": peak, "unit": "GB/s",
                             "frac": ilp_ach / peak,
                             "note": "latency bound (dependent L2 round trips + atomics per cluster-wide phase, profiles/cluster_barrier_cost_r2.txt), far from HBM: the iteration's working set is L2 / shared-memory resident"},
            "stage_ms": {k: float(np.mean([d[k] for d in timed])) for k in
                         ("ms_gate", "ms_cluster", "ms_assoc", "ms_dual", "ms_exact", "ms_prune", "ms_total")},
            "scan_ms": {"ms_total": pct("ms_total"), "ms_assoc": pct("ms_assoc"), "ms_gate": pct("ms_gate")},
            "ilp": {"certified_scans": int(sum(1 for d in timed if d["certified"])), "scans": K,
                    "gap_mean": float(np.mean(gaps)), "gap_max": float(np.max(gaps)),
                    "gap_rel_mean": float(np.mean([g / max(abs(d["lower_bound"]), 1e-9) for g, d in zip(gaps, timed)])),
                    "open_components_mean": float(np.mean([d["open_components"] for d in timed])),
                    "repaired_trees": int(sum(d.get("repaired_trees", 0) for d in timed)),
                    "note": "certified = the exact search proved the selection optimal; otherwise a feasible "
                            "selection with the stated gap to the Lagrangian lower bound (the reference's CBC would "
                            "warn 'NOT optimal' under a time limit, tracker.py:1201-1204); repaired_trees = tracks the "
                            "final feasibility check of the selection had to move to their miss hypothesis (0 = every "
                            "selection left the solver conflict free)"},
            "scan_stats": {k: float(np.mean([d[k] for d in timed])) for k in
                           ("n_trees", "n_parents", "n_children", "n_pairs", "n_clusters", "n_multi_clusters", "dual_iters",
                            "n_candidates", "bb_nodes", "bb_iters", "certified", "lower_bound", "objective", "n_active",
                            "nnz_active", "rows_active")},
            "forest_hbm_bytes": dev_bytes,
        }
        if not args.no_cpu_baseline:
            kind, times, leaves = run_cpu_reference(name, max_scans=2, n_scans_scenario=n_scans)
            v = len(times) / sum(times)
            # like for like: the SAME first scans from a cold start on the GPU (device-timed and end to end), with
            # the leaf counts of both sides printed
            cold = infos[:len(times)]
            line["cpu_baseline"] = {"value": v, "unit": "scans/s", "cores": 1, "kind": kind,
                                    "sample": cpu_sample_text(kind, times, leaves) +
                                    "; the GPU headline is at steady state (%.2e live leaves per scan)"
                                    % line["scan_stats"]["n_parents"]}
            line["like_for_like"] = {
                "what": "scans 1-%d of the same scenario from a cold start, both sides" % len(times),
                "cpu_s": [float(t) for t in times], "cpu_leaves_after_scan": [int(l) for l in leaves],
                "gpu_ms_device": [float(d["ms_total"]) for d in cold],
                "gpu_ms_e2e": [1e3 * float(t) for t in e2e_all[:len(times)]],
                "gpu_leaves_after_scan": [int(d["n_children"]) for d in cold],
                "gpu_certified": [int(d["certified"]) for d in cold],
                "ratio_e2e": float(sum(times) / max(sum(e2e_all[:len(times)]), 1e-12)),
                "ratio_device": float(sum(times) / max(1e-3 * sum(d["ms_total"] for d in cold), 1e-12))}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def run_tree_sharded(args, name, config, dist, torch, rank, world, local_rank, preroll, n_scans):
    """ONE region, its trees sharded over the ranks (pymht_b200/sharded.py): every rank gates its own trees, ONE
    all-gather of the packed column records (NCCL) gives every rank every column, the global 0/1 program is solved on
    them (warm started; rank 0's selection is broadcast).  Timed end to end through ShardedTracker.addMeasurementList
    (host scan in, tracks out), max over ranks.  Before the timed run the sharded tracker and a single forest (rank 0)
    process the first scans of the scenario and their track digests are compared: parity where there is > 1 GPU."""
    from pymht_b200.sharded import ShardedTracker, tracks_digest
    from pymht_b200.tracker import Tracker
    from pymht_b200.models import pv
    nT, R, lam, N, Pd, seed, max_nodes, max_par = WORKLOADS[name]
    max_nodes, max_par = capacities(name, n_scans)
    simList, scans = make_scenario(name, n_scans, seed_offset=0)      # the same region on every rank
    dev = torch.device("cuda", local_rank)
    mcap = meas_capacity(name)

    # ---- parity leg: the first scans, sharded vs single forest.  A certified solve is the unique optimum on both sides;
    #      an uncertified one depends on which worker found which incumbent first, so only scans certified on BOTH
    #      sides are compared (and the tracks a later scan inherits are only comparable while that holds) ----
    n_verify = 2
    vt = ShardedTracker(pv, T_RADAR, lam, 1e-9, N=N, P_d=Pd, initiator=None, maxTargets=max(1024, nT + 24), maxMeasurements=mcap,
                        maxNodes=1 << 22, maxParents=1 << 20, exactBudgetMs=10000)
    vt.mergeThreshold = 0.0
    vt.preInitialize(simList)
    v_sharded = []
    for s in scans[:n_verify]:
        vt.addMeasurementList(s)
        v_sharded.append(tracks_digest(vt.gatherTracks()))
    v_cert = [int(d["certified"]) for d in vt.scanInfo]
    vt.close()
    v_single, v_cert1 = None, None
    if rank == 0:
        st = Tracker(pv, T_RADAR, lam, 1e-9, N=N, P_d=Pd, initiator=None, maxTargets=max(1024, nT + 24), maxMeasurements=mcap,
                     maxNodes=1 << 22, maxParents=1 << 20, exactBudgetMs=10000)
        st.mergeThreshold = 0.0
        st.preInitialize(simList)
        v_single = []
        for s in scans[:n_verify]:
            st.addMeasurementList(s)
            v_single.append(tracks_digest([(n.ID, n.measurementNumber, n.cumulativeNLLR) for n in st.getTrackNodes()]))
        v_cert1 = [int(d["certified"]) for d in st.scanInfo]
        st.close()

    per = 1.0 / world + 0.15
    trk = ShardedTracker(pv, T_RADAR, lam, 1e-9, N=N, P_d=Pd, initiator=None, maxTargets=max(1024, nT + 24), maxMeasurements=mcap,
                         maxNodes=int(max_nodes * per), maxParents=int(max_par * per),
                         maxDualIterations=int(os.environ.get("MHT_DUAL_ITERS", "120")))
    trk.mergeThreshold = 0.0
    trk.preInitialize(simList)
    times = []
    for s in scans:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        trk.addMeasurementList(s)
        times.append(time.perf_counter() - t0)
    timed = times[preroll + args.warmup:]
    t = torch.tensor([sum(timed)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    log = trk.exchangeLog[preroll + args.warmup:]
    infos = trk.scanInfo[preroll + args.warmup:]
    tracks = trk.gatherTracks()
    n_tracks = len(tracks)
    if rank == 0:
        K = len(timed)
        v = K / float(t[0])
        config = dict(config, scans=SCENARIO_SOURCE,
                      parallelism="trees of one region sharded over %d GPUs; ONE all-gather of packed %d-byte column "
                      "records per scan (NCCL) + replicated, warm-started global solve" % (world, log[-1]["record_bytes"]))
        line = {"metric": metric_name(name), "value": v, "unit": "scans/s", "n_gpus": world,
                "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64 state / f32 covariance", "data": "synthetic", "config": config,
                "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": int(np.mean([16 * len(s.measurements) for s in scans])),
                        "d2h_bytes_per_step": int(n_tracks * 152 + 128)},
                "stage_ms": {"ms_gate_rank0": float(np.mean([d["ms_gate"] for d in infos])),
                             "ms_exchange": float(np.mean([d["ms_exchange"] for d in log])),
                             "ms_solve": float(np.mean([d["ms_solve"] for d in log]))},
                "scan_ms": {"p50": float(np.percentile(timed, 50) * 1e3), "p95": float(np.percentile(timed, 95) * 1e3),
                            "max": float(np.max(timed) * 1e3)},
                "exchange": {"columns_global": float(np.mean([d["n_cols_global"] for d in log])),
                             "bytes_gathered_per_scan": float(np.mean([d["bytes_gathered"] for d in log])),
                             "collectives_per_scan": "1 all_gather_into_tensor of packed records (+ 16 B/rank of counts, "
                                                     "the selection broadcast and the used-mask all-reduce)"},
                "scan_stats": {"tracks": n_tracks, "tracks_hash": tracks_digest(tracks),
                               "certified": float(np.mean([d["certified"] for d in infos])),
                               "objective": float(np.mean([d["objective"] for d in infos])),
                               "lower_bound": float(np.mean([d["lower_bound"] for d in infos]))},
                "parity": parity_summary(v_sharded, v_single, v_cert, v_cert1)}
        emit(line)
    trk.close()
    dist.destroy_process_group()


def parity_summary(h_sharded, h_single, cert_sharded, cert_single):
    """Per-scan digests of the sharded and the single-forest run; `equal` covers the leading scans certified on both
    sides (at least one must be)."""
    n_cmp = 0
    for a, b in zip(cert_sharded, cert_single):
        if not (a and b):
            break
        n_cmp += 1
    return {"scans": len(h_sharded), "sharded_hash": h_sharded, "single_forest_hash": h_single,
            "certified_sharded": cert_sharded, "certified_single": cert_single, "scans_compared": n_cmp,
            "equal": bool(n_cmp > 0 and h_sharded[:n_cmp] == h_single[:n_cmp])}


def exchange(dist, device, t_dev, t_e2e, n_tracks):
    """Multi-rank reduction of one bench run: device/e2e times -> MAX over ranks (a step is as slow as
    the slowest sector), and the only exchange the sector-sharded forest needs: every rank learns every
    sector's live-track count (all_gather).  Backend-agnostic (NCCL on GPUs, gloo in the CPU test)."""
    import torch
    t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    summary = torch.tensor([n_tracks], dtype=torch.int64, device=device)
    gathered = [torch.zeros_like(summary) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, summary)
    return float(t[0]), float(t[1]), [int(g[0]) for g in gathered]


if __name__ == "__main__":
    main()
