mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -30
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 MHT_BENCH_SKIP_E2E=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; grep "^scan\|mht\]" gpurun_out/bench_r2d.err | tail -12
