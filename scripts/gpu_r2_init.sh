#!/bin/bash
# initiator (f3) on the GPU: parity tests, memcheck of the small ones, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_initiator.py -x -q -s > gpurun_out/init_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/init_tests.log
tail -25 gpurun_out/init_tests.log
MHT_GNN_SPEC=0 timeout 600 python -m pytest tests/test_gpu_initiator.py -x -q -s -k "giant" 2>&1 | grep -E "config-3|passed|failed"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_initiator.py -x -q -k "replays or capacity or distance" > gpurun_out/init_memcheck.log 2>&1
tail -4 gpurun_out/init_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_initiator.py -x -q -k "replays" > gpurun_out/init_racecheck.log 2>&1
tail -4 gpurun_out/init_racecheck.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_all.log 2>&1
tail -5 gpurun_out/tests_all.log
