#!/bin/bash
mkdir -p gpurun_out
for k in 1 2; do
timeout 1500 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "^E  |passed|failed|repaired" | cut -c1-400 | head -8
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['scan_ms']['ms_total'], d['ilp']['certified_scans'])"
