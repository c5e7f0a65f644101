mkdir -p gpurun_out
MHT_BENCH_VERBOSE=1 timeout 900 python bench.py --workload cfg4_10k_targets_50k_meas_N6 --preroll 0 --warmup 1 --steps 3 --no-cpu-baseline > gpurun_out/bench_r2_cfg4_n1.json 2> gpurun_out/bench_r2_cfg4_n1.err; grep "^scan" gpurun_out/bench_r2_cfg4_n1.err | cut -c1-260; tail -c 400 gpurun_out/bench_r2_cfg4_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_cfg4_n1.json'))
for k in ('metric','value','e2e','stage_ms','scan_ms','ilp','roofline'):
    print(k, d.get(k))
PY

