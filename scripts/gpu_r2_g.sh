mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | cut -c1-400 | head -20
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; grep "^scan\|mht\]" gpurun_out/bench_r2g.err | tail -12 | cut -c1-260
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline_ilp'):
    print(k, d.get(k))
PY
