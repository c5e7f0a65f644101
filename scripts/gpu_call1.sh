# GPU call 1 of this session: state of HEAD -- parity suite, default bench line, reference arm, smoke, emit capture
python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "cfg3|identical|passed|failed|FAILED|^E |bit-identical" | cut -c1-600 > gpurun_out/tests.log; cat gpurun_out/tests.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','stage_ms')})
print(d['roofline']); print(d['scan_stats']); print(d['cpu_baseline']['value'])
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:forest_emit_kernel -s 11 -c 1 -f -o gpurun_out/prof_emit_a $B > gpurun_out/p1.log 2>&1
ls -la gpurun_out/
