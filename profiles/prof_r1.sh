# Round-1 profile recipe (run under gpurun from the repo root; outputs land in gpurun_out/).
# bench.py --steps 1 --warmup 3 = 12 scans of the device leg (51 launches each) + 12 scans of the e2e leg.
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# 1. every launch of the device leg with its device time (scripts/summarize_launches.py --scan 12 picks the
#    timed steady-state scan: scans start at live_scan_kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/launches_r1.log 2>&1
# 2. full captures of the top kernels at steady state (one launch each)
ncu --set full --clock-control none --import-source on -k regex:forest_emit_kernel -s 11 -c 1 -f -o gpurun_out/prof_emit_r1 $B > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forest_count_kernel -s 11 -c 1 -f -o gpurun_out/prof_count_r1 $B > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dual_rc_kernel -s 56 -c 1 -f -o gpurun_out/prof_dualrc_r1 $B > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:active_count_kernel -s 33 -c 1 -f -o gpurun_out/prof_active_r1 $B > gpurun_out/p4.log 2>&1
ls -la gpurun_out/*.ncu-rep
