mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_barrier_cost scripts/experiments/cluster_barrier_cost.cu 2>/dev/null
/tmp/cluster_barrier_cost | tee gpurun_out/cluster_barrier_cost_r2.txt
