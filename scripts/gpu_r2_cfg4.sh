# BASELINE config 4 (10k targets, 50k meas/scan, CV model), cold start (scans 1-4: warm-up 1 + 3 timed), trees of the
# ONE region sharded over all visible GPUs; run with `gpurun --gpus 8`.  The same five scans on one GPU: gpu_r2_cfg4_n1.sh
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --shard trees --workload cfg4_10k_targets_50k_meas_N6 --preroll 0 --warmup 1 --steps 3 > gpurun_out/bench_r2_cfg4_n$N.json 2> gpurun_out/bench_r2_cfg4_n$N.err; tail -c 800 gpurun_out/bench_r2_cfg4_n$N.err; cut -c1-2500 gpurun_out/bench_r2_cfg4_n$N.json
