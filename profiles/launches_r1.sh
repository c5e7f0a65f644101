# launch list of ONE steady-state scan (scan 12 of the device leg; 755 launches per scan)
ncu --metrics gpu__time_duration.sum --clock-control none -s 8305 -c 760 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
tail -c 600 gpurun_out/launches_r1.log
