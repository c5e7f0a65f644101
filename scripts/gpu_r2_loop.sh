#!/bin/bash
# cluster dual loop after the load hoisting: full GPU suite, phase clock, and the same bench through the grid version
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
for v in "" "MHT_NO_CLUSTER_LOOP=1"; do
  env $v MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 MHT_BENCH_SKIP_E2E=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err
  echo "== ${v:-cluster loop}"; grep "mht\]" gpurun_out/bench_loop.err | tail -3
  python -c "
import json; d=json.load(open('gpurun_out/bench_loop.json')); s=d['scan_stats']; print('value %.1f' % d['value'], d['scan_ms']['ms_total'], 'certified', d['ilp']['certified_scans'], 'gap %.4f' % d['ilp']['gap_mean'], 'lb %.9f obj %.9f iters %.1f' % (s['lower_bound'], s['objective'], s['dual_iters']), d['stage_ms'], d['roofline_ilp'])"
done
