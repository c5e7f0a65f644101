# 2-GPU pass: full GPU suite (incl. the tree-sharded parity test) + tree-sharded bench at N=2 + 1-GPU bench for comparison
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E |passed|failed|Error" | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --shard trees --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_trees.json 2> gpurun_out/bench_r2_n2_trees.err; tail -c 600 gpurun_out/bench_r2_n2_trees.err; cat gpurun_out/bench_r2_n2_trees.json | cut -c1-3000
MHT_LOOP_PROF=1 MHT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; grep "^scan\|mht\]" gpurun_out/bench_r2e.err | tail -9
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
for k in ('value','e2e','gpu_launches','stage_ms','scan_ms','ilp','roofline_ilp'):
    print(k, d.get(k))
PY
