# final 1-GPU validation at the driver's arguments: smoke, full GPU suite, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_reference.json
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; grep "^scan\|mht\]" gpurun_out/bench_r2_default.err | tail -22 | cut -c1-250
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_default.json'))
for k in ('value','e2e','gpu_launches','clocks','stage_ms','scan_ms','ilp','roofline','roofline_ilp','like_for_like'):
    print(k, d.get(k))
PY
