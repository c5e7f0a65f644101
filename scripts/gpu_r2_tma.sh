mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
for t in 1 0 1 0; do
MHT_EMIT_TMA=$t MHT_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MHT_EMIT_TMA=$t value %.1f ms_gate %.4f (p50 %.4f) roofline %.4f' % (d['value'], d['stage_ms']['ms_gate'], d['scan_ms']['ms_gate']['p50'], d['roofline']['frac']))"
done
