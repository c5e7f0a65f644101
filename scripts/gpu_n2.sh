T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
nvidia-smi -L
timeout 400 $T bench.py --gpus 2 --shard trees --steps 4 --warmup 3 > gpurun_out/bench_n2_trees.json 2> gpurun_out/n2_trees.err; echo "rc=$?"; tail -c 1500 gpurun_out/n2_trees.err; cat gpurun_out/bench_n2_trees.json
timeout 400 $T bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_sectors.json 2> gpurun_out/n2_sectors.err; echo "rc=$?"; tail -c 600 gpurun_out/n2_sectors.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_sectors.json')); print({k:d[k] for k in ('value','n_gpus','e2e','stage_ms')})"
