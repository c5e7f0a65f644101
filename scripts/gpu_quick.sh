# quick GPU check: parity suite + default bench line (no reference arm) [+ optional ncu of the gate kernels: PROF=1]
python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "cfg3|identical|passed|failed|FAILED|^E |bit-identical|Error|error" | cut -c1-700 > gpurun_out/tests.log; cat gpurun_out/tests.log
python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','stage_ms')})
print(d['roofline']); print(d['scan_stats'])
PY
if [ -n "$PROF" ]; then
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:forest_emit_kernel -s 11 -c 1 -f -o gpurun_out/prof_emit_b $B > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forest_gate_kernel -s 11 -c 1 -f -o gpurun_out/prof_count_b $B > gpurun_out/p2.log 2>&1
fi
if [ -n "$SWEEP" ]; then
for rc in "4 20" "6 40" "8 64" "12 96" "100 100000"; do set -- $rc
MHT_HEAVY_ROWS=$1 MHT_HEAVY_CAND=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rows/cand $1 $2', d['stage_ms'])"
done
fi
