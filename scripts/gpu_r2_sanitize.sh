# compute-sanitizer on the small fixtures (memcheck: every kernel; racecheck: shared-memory hazards of the cluster loop and
# the branch & bound)
mkdir -p gpurun_out
export MHT_BB_MS=2000
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and (cfg1 or cfg2_small or cfg5_small) or dynamic or recycled or edge" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -12
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and (cfg1 or cfg5_small)" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "cfg3_vs_reference and cfg3_head" > gpurun_out/sanitize_memcheck_cfg3.log 2>&1; echo "memcheck cfg3_head rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_memcheck_cfg3.log | head -8
