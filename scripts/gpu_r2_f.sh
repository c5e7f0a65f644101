mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "cfg5_full|cfg3_low.*scan 3|NOT CERT|^E |passed|failed|Error" | cut -c1-600 | head -40
MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; grep "^scan" gpurun_out/bench_r2f.err | tail -21 | cut -c1-250
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2f.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline','roofline_ilp','like_for_like','cpu_baseline'):
    print(k, d.get(k))
PY
