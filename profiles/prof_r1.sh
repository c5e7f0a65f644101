set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 33600 -c 3100 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/launches_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forest_emit_kernel -s 10 -c 1 -f -o gpurun_out/prof_emit_r1 $B > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forest_count_kernel -s 10 -c 1 -f -o gpurun_out/prof_count_r1 $B > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dual_rc_kernel -s 1300 -c 1 -f -o gpurun_out/prof_dualrc_r1 $B > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dual_arg_kernel -s 1300 -c 1 -f -o gpurun_out/prof_dualarg_r1 $B > gpurun_out/p4.log 2>&1
tail -3 gpurun_out/p*.log
ls -la gpurun_out
