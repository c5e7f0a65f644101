# launch list of ~one steady-state scan of the device leg (sifting path: ~2.2k launches per scan)
ncu --metrics gpu__time_duration.sum --clock-control none -s 26000 -c 2300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
tail -c 300 gpurun_out/launches_r1.log
