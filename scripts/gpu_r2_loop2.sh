#!/bin/bash
mkdir -p gpurun_out
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 MHT_BENCH_SKIP_E2E=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err
grep "mht\]" gpurun_out/bench_loop.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/bench_loop.json')); s=d['scan_stats']; print('value %.1f' % d['value'], d['scan_ms']['ms_total'], 'certified', d['ilp']['certified_scans'], 'lb %.9f obj %.9f iters %.1f' % (s['lower_bound'], s['objective'], s['dual_iters']), 'ms_dual %.3f' % d['stage_ms']['ms_dual'])"
