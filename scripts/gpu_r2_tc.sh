mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_gate_experiment scripts/experiments/tc_gate_experiment.cu 2>/dev/null
/tmp/tc_gate_experiment > gpurun_out/tc_experiment_r2.txt 2>&1; cat gpurun_out/tc_experiment_r2.txt
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 6 -c 2 --csv /tmp/tc_gate_experiment 2>&1 | grep -E "^\"[0-9]" | awk -F'","' '{print $5, "|", $(NF-2), "|", $(NF-1), "|", $NF}' >> gpurun_out/tc_experiment_r2.txt
tail -16 gpurun_out/tc_experiment_r2.txt
# final bench lines at the driver's arguments (both arms) with the adopted defaults
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_reference.json
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; grep "mht\]" gpurun_out/bench_r2_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_default.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline','roofline_ilp','like_for_like'):
    print(k, d.get(k))
PY
