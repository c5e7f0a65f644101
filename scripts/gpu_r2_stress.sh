#!/bin/bash
for k in 1 2 3 4; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | grep -E "^E  |passed|failed" | cut -c1-700 | head -6
done
