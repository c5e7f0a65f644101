python -m pytest tests -q -m gpu -s 2>&1 | grep -E "^cfg3|identical|passed|failed|FAILED|^E |bit-identical" | cut -c1-500 > gpurun_out/tests_final.log; tail -3 gpurun_out/tests_final.log
python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log
python bench.py --impl reference > gpurun_out/bench_r1_reference.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_default.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','stage_ms')})
print(d['roofline']); print(d['scan_stats']); print(d['cpu_baseline']['value'])
PY
bash profiles/prof_r1.sh > gpurun_out/prof.log 2>&1; tail -5 gpurun_out/prof.log
