# round 2: per-scan table of the bench run + ncu launch list (device leg only) to find the association tail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; grep "^scan" gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp'):
    print(k, d.get(k))
PY
MHT_BENCH_SKIP_E2E=1 MHT_BENCH_VERBOSE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b_ncu.json 2> gpurun_out/bench_r2b_ncu.err
grep "^scan" gpurun_out/bench_r2b_ncu.err | tail -22
wc -l gpurun_out/launches_r2b.csv
