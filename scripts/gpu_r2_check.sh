mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
MHT_LOOP_PROF=1 MHT_BENCH_SKIP_E2E=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; grep "mht\]" gpurun_out/bench_check.err
python -c "
import json; d=json.load(open('gpurun_out/bench_check.json')); print('value %.1f' % d['value'], d['scan_ms']['ms_total'], 'certified', d['ilp']['certified_scans'], 'gap %.2f' % d['ilp']['gap_mean'], d['stage_ms'])"
