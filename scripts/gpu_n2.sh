python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; echo "N=1 rc=$? stdout lines: $(wc -l < gpurun_out/b1.json)"; head -c 100 gpurun_out/b1.json; echo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29655 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err; echo "N=2 rc=$? stdout lines: $(wc -l < gpurun_out/b2.json)"; head -c 100 gpurun_out/b2.json; echo; grep -c "NCCL version" gpurun_out/b2.err
