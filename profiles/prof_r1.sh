# Round-1 profile recipe (run under gpurun from the repo root; outputs land in gpurun_out/, the summaries are
# copied to profiles/ by scripts/collect_profiles.py).
# bench.py --steps 1 --warmup 3 = 12 scans of the device leg (~108 launches each) + 12 scans of the e2e leg.
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# 1. every launch of the device leg with its device time (scripts/summarize_launches.py --scan 12 picks the
#    timed steady-state scan: scans start at assoc_init_kernel ... live_scan_kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/launches_r1.log 2>&1
# 2. full captures of the top kernels at the steady-state scan (one launch each)
cap() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_$3_r1 $B > gpurun_out/p_$3.log 2>&1; }
cap forest_emit_kernel 11 emit
cap forest_gate_kernel 11 gate
cap forest_gate_heavy_kernel 11 heavy
cap dual_rc_kernel 56 dualrc
ls -la gpurun_out/*.ncu-rep
