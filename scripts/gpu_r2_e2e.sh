#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
timeout 600 python scripts/prof_e2e_host.py 2>&1 | grep -v "NOT optimal" | head -16
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e2e.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['scan_ms']['ms_total'])"
