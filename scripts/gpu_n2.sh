python -m pytest tests/test_gpu_sharded.py -q -m gpu -x 2>&1 | tail -2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29611 bench.py --gpus 2 --shard trees --steps 4 --warmup 3 > gpurun_out/bench_n2_trees.json 2> gpurun_out/n2_trees.err; echo "rc=$?"
timeout 300 $T --master-port 29655 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_sectors.json 2> gpurun_out/n2_sectors.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_n2_sectors.json','gpurun_out/bench_n2_trees.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, {k:d[k] for k in ('value','n_gpus','e2e','stage_ms')})
PY
