# Round-2 profile recipe (run under gpurun from the repo root; outputs land in gpurun_out/, the summaries are copied
# to profiles/ by scripts/collect_profiles.py r2).  Everything is taken AT THE DRIVER'S ARGUMENTS:
#   bench.py --steps 20 --warmup 5   = 33 scans of the device leg (8 pre-roll + 5 warm-up + 20 timed)
# (MHT_BENCH_SKIP_E2E=1 drops the second, host-buffer leg so that ncu sees each scan once; the JSON these runs print
# is never a bench value).
export MHT_BENCH_SKIP_E2E=1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
# 1. every launch of the 20 timed scans with its device time: skip the launches of the 13 untimed scans
#    (scripts/summarize_launches.py splits scans at live_scan_kernel)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/launches_r2.log 2>&1
# 2. full captures of the top kernels at timed scan 8 of 20 (scan 21 of the run; one launch each)
cap() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_$3_r2 $B > gpurun_out/p_$3_r2.log 2>&1; }
cap dual_loop_cluster_kernel 41 dualloop      # 2 launches per scan (one per sifting round): the second of scan 21
cap forest_emit_kernel 20 emit
cap forest_gate_kernel 20 gate
cap bb_search_kernel 20 bbsearch
ls -la gpurun_out/*_r2.ncu-rep
