#!/bin/bash
# launch list of the final build at the driver's arguments (device leg only: ncu sees each scan once)
mkdir -p gpurun_out
export MHT_BENCH_SKIP_E2E=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/launches_r2.log 2>&1
tail -2 gpurun_out/launches_r2.log | cut -c1-200; wc -l gpurun_out/launches_r2.csv
