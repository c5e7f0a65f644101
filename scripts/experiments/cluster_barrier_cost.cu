// What does one cluster-wide phase of the dual loop cost at least?  16 CTAs x 1024 threads in ONE thread-block cluster
// (the geometry of dual_loop_cluster_kernel), 20 000 iterations of:
//   0: cg::cluster.sync()                      (barrier.cluster.arrive.release + wait.acquire: UCGABAR + MEMBAR.ALL.GPU + CCTL.IVALL)
//   1: arrive.relaxed + wait                   (no MEMBAR: what a barrier over distributed shared memory could get away with)
//   2: one L2 round trip per thread (volatile load of a line another SM wrote) + cg sync
//   3: one fire-and-forget global atomic per thread + cg sync
//   4: a dependent 2-hop chain (load index -> load value) + atomic + cg sync   (the shape of the loop's row phases)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_barrier_cost cluster_barrier_cost.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(1024, 1) phase_kernel(int mode, int iters, int *buf, int *idx, unsigned long long *out) {
    cg::cluster_group cluster = cg::this_cluster();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    unsigned long long t0 = 0, t1 = 0;
    int acc = 0;
    cluster.sync();
    if (gtid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters; ++it) {
        if (mode == 2) acc += ((volatile int *)buf)[(gtid + 1024 * (it & 15)) % nth];
        if (mode == 3) atomicAdd(&buf[(gtid * 7 + it) % nth], 1);
        if (mode == 4) {
            const int j = ((volatile int *)idx)[(gtid + 1024 * (it & 15)) % nth];
            acc += ((volatile int *)buf)[j];
            atomicAdd(&buf[(j * 7 + it) % nth], 1);
        }
        if (mode == 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n barrier.cluster.wait.aligned;\n" ::: "memory");
        else cluster.sync();
    }
    if (gtid == 0) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        out[mode] = t1 - t0;
    }
    if (acc == 123456789) buf[0] = acc;
}

int main() {
    int *buf, *idx; unsigned long long *out;
    const int nth = 16 * 1024, iters = 20000;
    cudaMalloc(&buf, nth * sizeof(int)); cudaMalloc(&idx, nth * sizeof(int)); cudaMalloc(&out, 8 * sizeof(unsigned long long));
    cudaMemset(buf, 0, nth * sizeof(int));
    int h[16 * 1024]; for (int i = 0; i < nth; ++i) h[i] = (i * 2654435761u) % nth;
    cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(phase_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const char *name[5] = {"cluster.sync() only", "arrive.relaxed + wait (no MEMBAR)", "L2 load + cluster.sync()", "global atomic + cluster.sync()",
                           "2-hop load chain + atomic + cluster.sync()"};
    for (int ctas : {16, 8}) {
        for (int mode = 0; mode < 5; ++mode) {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(1024);
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, phase_kernel, mode, iters, buf, idx, out);
            if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            unsigned long long ns; cudaMemcpy(&ns, out + mode, 8, cudaMemcpyDeviceToHost);
            printf("%2d CTAs x 1024 threads: %-44s %.3f us per phase\n", ctas, name[mode], ns / (double)iters * 1e-3);
        }
    }
    return 0;
}
