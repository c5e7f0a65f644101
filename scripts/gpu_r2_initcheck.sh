#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -k "replays_reference_golden and cfg5_n8" > gpurun_out/initcheck_small.log 2>&1
grep -E "Uninitialized|at .*\(|ERROR SUMMARY|passed|failed" gpurun_out/initcheck_small.log | sort | uniq -c | sort -rn | head -30
timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -k "large_random" > gpurun_out/initcheck_large.log 2>&1
grep -E "Uninitialized|at .*\(|ERROR SUMMARY|passed|failed" gpurun_out/initcheck_large.log | sort | uniq -c | sort -rn | head -30
