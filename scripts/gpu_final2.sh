ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r1.csv --scan 12 | head -30
python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_default.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','stage_ms')})
PY
