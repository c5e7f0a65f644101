# final 2-GPU validation: full GPU suite (tree-sharded parity included), the driver's default multi-GPU line (sectors),
# and the tree-sharded line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2_sectors.json 2> gpurun_out/bench_r2_n2_sectors.err; tail -c 300 gpurun_out/bench_r2_n2_sectors.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n2_sectors.json')); print({k:d[k] for k in ('value','n_gpus','scaling','e2e','scan_ms','ilp')})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --shard trees --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_trees.json 2> gpurun_out/bench_r2_n2_trees.err; tail -c 300 gpurun_out/bench_r2_n2_trees.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n2_trees.json')); print({k:d[k] for k in ('value','stage_ms','scan_ms','exchange','parity')})"
