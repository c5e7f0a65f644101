// EXPERIMENTAL -- not part of libmht_b200.so and NOT yet run on a GPU (this round's GPU budget ended before it could
// be validated; pymht_b200/build.py does not compile this directory).  Device instantiation of lbb_core.h: one CTA
// per open component, the block-wide loops of lbb::Solver stride over columns / trees / rows, reductions go through
// the Work scratch.  The host build of the same core is checked against HiGHS by scripts/proto/lbb_host_check.py
// (279-tree cluster of cfg3 scan 2: optimum proven in 5 527 nodes).  Next round: (1) build Problem/Work per
// component from the candidate lists (comp_trees / cand_col, local row ids), (2) move u / usage / tmin / targ of
// small cores into shared memory, (3) replace the serial reductions of lbb::Solver by warp shuffles, (4) launch it
// after local_search_kernel for the components branch_bound_kernel gives up on.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -c lbb_repair.cu      (compile check)
#include <cuda_runtime.h>

#include "lbb_core.h"

namespace lbb {

struct DeviceCtx {
    __device__ int tid() const { return (int)threadIdx.x; }
    __device__ int nthr() const { return (int)blockDim.x; }
    __device__ void sync() { __syncthreads(); }
    __device__ void amin(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
    __device__ void amax(int *p, int v) { atomicMax(p, v); }
    __device__ void aadd(int *p, int v) { atomicAdd(p, v); }
};

// problems[i] / works[i] describe component i (device pointers inside); results in works[i].best / best_sel /
// proven / nodes.  Work::red and Work::redi need blockDim.x + 8 entries.
__global__ void __launch_bounds__(512) lbb_repair_kernel(const Problem *problems, Work *works, int n_components) {
    for (int i = blockIdx.x; i < n_components; i += gridDim.x) {
        DeviceCtx ctx;
        Solver<DeviceCtx> solver(problems[i], works[i], ctx);
        solver.run();
        __syncthreads();
    }
}

}  // namespace lbb
