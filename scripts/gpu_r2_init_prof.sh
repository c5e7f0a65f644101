#!/bin/bash
# ncu evidence for the initiator kernels at config 3's scale (tests/test_gpu_initiator.py::test_config3_scale_one_giant_component)
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_initiator.py -x -q -k giant"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_init_r2.csv $T > gpurun_out/launches_init_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gnn_spec_kernel -c 1 -f -o gpurun_out/prof_gnnspec_r2 $T > gpurun_out/p_gnnspec_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gnn_gate_kernel -c 2 -f -o gpurun_out/prof_gnngate_r2 $T > gpurun_out/p_gnngate_r2.log 2>&1
ls -la gpurun_out/*gnn*_r2.ncu-rep; tail -3 gpurun_out/launches_init_r2.log
