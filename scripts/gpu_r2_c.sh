# round 2: parity suite + verbose bench (per-scan table)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; grep "^scan" gpurun_out/bench_r2c.err | tail -26
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline','roofline_ilp','like_for_like'):
    print(k, d.get(k))
PY
