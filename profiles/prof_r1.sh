# ncu --set full of the top kernels at steady state (one launch each); bench shortened to 1 timed scan
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:forest_emit_kernel -s 10 -c 1 -f -o gpurun_out/prof_emit_r1 $B > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forest_count_kernel -s 10 -c 1 -f -o gpurun_out/prof_count_r1 $B > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uf_union_cols_kernel -s 10 -c 1 -f -o gpurun_out/prof_union_r1 $B > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:dual_rc_kernel<true>|dual_rc_kernel<\(bool\)1>" -s 40 -c 1 -f -o gpurun_out/prof_dualrc_r1 $B > gpurun_out/p4.log 2>&1
ls -la gpurun_out/*.ncu-rep
