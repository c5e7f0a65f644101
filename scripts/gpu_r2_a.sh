# round 2, first GPU pass: parity suite (cfg3 verbose), compute-sanitizer on the small exact-search fixture, bench 20/5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "cfg3|identical|NOT CERT|passed|failed|FAILED|^E |Error|error|Optim" | cut -c1-900 > gpurun_out/tests_r2a.log; tail -40 gpurun_out/tests_r2a.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline','roofline_ilp','like_for_like','scan_stats'):
    print(k, d.get(k))
PY
