// Gate stage: measurement grid + stateless mht_gate_batch operator.
// Replaces Tracker._processLeafNodes (reference pymht/tracker.py:383-398,804-889) and the
// kalman.py batch operators it calls (pymht/utils/kalman.py:14-101).
#include "common.cuh"

namespace mht {

// ------------------------------------------------------------------------------------------------
// Measurement grid: one CTA bins the scan (M <= ~10^5) into uniform cells by counting sort.
// The reference tests every leaf against every measurement (kalman.py:36-40 materialises an
// (L,M,2) tensor); the grid makes a leaf look at O(10) candidates and returns the identical set.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_reduce_minmax(double v, bool is_min, double *sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fmin(v, other) : fmax(v, other);
    }
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : sh[0];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_min ? fmin(v, other) : fmax(v, other);
        }
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    v = sh[0];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(1024, 1)
grid_build_kernel(const double2 *__restrict__ z, int M, GridDesc *grid, int *cell_start, int *cell_fill,
                  double2 *gz, int *gidx) {
    __shared__ double sh[32];
    __shared__ int carry;
    const int tid = threadIdx.x, nt = blockDim.x;
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = tid; i < M; i += nt) {
        const double2 p = z[i];
        xmin = fmin(xmin, p.x);
        xmax = fmax(xmax, p.x);
        ymin = fmin(ymin, p.y);
        ymax = fmax(ymax, p.y);
    }
    xmin = block_reduce_minmax(xmin, true, sh);
    xmax = block_reduce_minmax(xmax, false, sh);
    ymin = block_reduce_minmax(ymin, true, sh);
    ymax = block_reduce_minmax(ymax, false, sh);
    if (M == 0) xmin = xmax = ymin = ymax = 0.0;
    const double w = fmax(xmax - xmin, 1e-3), h = fmax(ymax - ymin, 1e-3);
    // ~1.5 measurements per cell on average, at most kGridMaxCells cells
    double cell = sqrt(1.5 * w * h / fmax((double)M, 1.0));
    cell = fmax(cell, sqrt(w * h / (0.9 * kGridMaxCells)));
    cell = fmax(cell, fmax(w, h) / 4096.0);
    const int nx = max(1, (int)floor(w / cell) + 1), ny = max(1, (int)floor(h / cell) + 1);
    const int ncell = nx * ny;
    const double inv_cell = 1.0 / cell;
    if (tid == 0) {
        grid->x0 = xmin;
        grid->y0 = ymin;
        grid->inv_cell = inv_cell;
        grid->nx = nx;
        grid->ny = ny;
        grid->n_meas = M;
        carry = 0;
    }
    for (int c = tid; c <= ncell; c += nt) cell_fill[c] = 0;
    __syncthreads();
    for (int i = tid; i < M; i += nt) {
        const double2 p = z[i];
        const int cx = min(nx - 1, max(0, (int)floor((p.x - xmin) * inv_cell)));
        const int cy = min(ny - 1, max(0, (int)floor((p.y - ymin) * inv_cell)));
        atomicAdd(&cell_fill[cy * nx + cx], 1);
    }
    __syncthreads();
    // exclusive scan of cell_fill[0..ncell) -> cell_start, chunked by blockDim
    __shared__ int wsum[32];
    for (int base = 0; base < ncell; base += nt) {
        const int c = base + tid;
        const int v = (c < ncell) ? cell_fill[c] : 0;
        int incl = v;
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < (nt >> 5)) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const int before = carry + (wid ? wsum[wid - 1] : 0) + incl - v;
        if (c < ncell) {
            cell_start[c] = before;
            cell_fill[c] = before;  // becomes the scatter cursor
        }
        __syncthreads();
        if (tid == nt - 1) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) cell_start[ncell] = M;
    __syncthreads();
    for (int i = tid; i < M; i += nt) {
        const double2 p = z[i];
        const int cx = min(nx - 1, max(0, (int)floor((p.x - xmin) * inv_cell)));
        const int cy = min(ny - 1, max(0, (int)floor((p.y - ymin) * inv_cell)));
        const int pos = atomicAdd(&cell_fill[cy * nx + cx], 1);
        gz[pos] = p;
        gidx[pos] = i;
    }
}

int64_t grid_workspace_bytes(int64_t M) {
    // GridDesc | cell_start[kGridMaxCells+2] | cell_fill[kGridMaxCells+2] | gz[M] | gidx[M]
    int64_t b = 256;
    b += 2 * (int64_t)(kGridMaxCells + 64) * sizeof(int);
    b += ((M + 16) * (int64_t)sizeof(double2));
    b += ((M + 16) * (int64_t)sizeof(int));
    return (b + 255) / 256 * 256;
}

int launch_grid_build(const double *d_z, int M, GridDesc *d_grid, int *d_cell_start, int *d_cell_fill,
                      double2 *d_gz, int *d_gidx, cudaStream_t s) {
    count_launch(), grid_build_kernel<<<1, 1024, 0, s>>>((const double2 *)d_z, M, d_grid, d_cell_start, d_cell_fill, d_gz,
                                         d_gidx);
    MHT_CUDA(cudaGetLastError());
    return MHT_OK;
}

// ------------------------------------------------------------------------------------------------
// Stateless operator kernels: one thread per leaf, two passes (count, emit) around a block scan.
// ------------------------------------------------------------------------------------------------
struct GateBatchArgs {
    mht_model model;
    int L;
    const double *x0;
    const float *P0;
    const double *Pd;
    const double *cnllr;
    const GridDesc *grid;
    const int *cell_start;
    const double2 *gz;
    const int *gidx;
    const double2 *z;  // original order
    int *count;      // [L]
    int *tile_sum;   // [ntiles+1]
    // outputs
    double *x_bar;
    float *P_bar, *P_hat;
    double *miss_cnllr;
    int *pair_off;
    int *pair_meas;
    double *pair_cnllr;
    double *pair_xhat;
    long long cap;
    unsigned char *meas_used;
};

__device__ __forceinline__ void load_leaf(const GateBatchArgs &a, int i, double x0[4], float P0[16]) {
    const double2 *xp = (const double2 *)(a.x0 + 4 * (size_t)i);
    const double2 x01 = xp[0], x23 = xp[1];
    x0[0] = x01.x;
    x0[1] = x01.y;
    x0[2] = x23.x;
    x0[3] = x23.y;
    const float4 *pp = (const float4 *)(a.P0 + 16 * (size_t)i);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float4 v = pp[r];
        P0[4 * r] = v.x;
        P0[4 * r + 1] = v.y;
        P0[4 * r + 2] = v.z;
        P0[4 * r + 3] = v.w;
    }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    const int excl = (wid ? wsum[wid - 1] : 0) + incl - v;
    if (total) *total = wsum[(blockDim.x >> 5) - 1];
    return excl;
}

__global__ void __launch_bounds__(kTile) gate_batch_count_kernel(GateBatchArgs a) {
    const int ntiles = (a.L + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        int cnt = 0;
        if (i < a.L) {
            double x0[4];
            float P0[16];
            load_leaf(a, i, x0, P0);
            LeafKF kf;
            leaf_kf<false>(a.model, x0, P0, a.Pd[i], kf);
            for_each_gated(*a.grid, a.cell_start, a.gz, a.gidx, kf, a.model.eta2,
                           [&](int, double, double, double) { ++cnt; });
            a.count[i] = cnt;
        }
        int total;
        block_exclusive_scan(cnt, &total);
        if (threadIdx.x == 0) a.tile_sum[tile] = total;
        __syncthreads();
    }
}

// single CTA: exclusive scan of tile sums in place; tile_sum[n] = grand total
__global__ void __launch_bounds__(1024, 1) scan_tiles_kernel(int *tile_sum, int n, int *total_out) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? tile_sum[i] : 0;
        int total;
        const int excl = block_exclusive_scan(v, &total);
        if (i < n) tile_sum[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_sum[n] = carry;
        if (total_out) *total_out = carry;
    }
}

// in-place insertion sort of a short int run in global memory (gated indices of one leaf)
__device__ __forceinline__ void sort_run(int *v, int n) {
    for (int i = 1; i < n; ++i) {
        const int key = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > key) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = key;
    }
}

__global__ void __launch_bounds__(kTile) gate_batch_emit_kernel(GateBatchArgs a) {
    const int ntiles = (a.L + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        const int cnt = (i < a.L) ? a.count[i] : 0;
        const int off = a.tile_sum[tile] + block_exclusive_scan(cnt, nullptr);
        if (i < a.L) {
            a.pair_off[i] = off;
            if (i == a.L - 1) a.pair_off[a.L] = off + cnt;
            double x0[4];
            float P0[16];
            load_leaf(a, i, x0, P0);
            LeafKF kf;
            const double Pd = a.Pd[i];
            leaf_kf<true>(a.model, x0, P0, Pd, kf);
            double2 *xb = (double2 *)(a.x_bar + 4 * (size_t)i);
            xb[0] = make_double2(kf.xbar[0], kf.xbar[1]);
            xb[1] = make_double2(kf.xbar[2], kf.xbar[3]);
            float4 *pb = (float4 *)(a.P_bar + 16 * (size_t)i), *ph = (float4 *)(a.P_hat + 16 * (size_t)i);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                pb[r] = make_float4(kf.Pbar[4 * r], kf.Pbar[4 * r + 1], kf.Pbar[4 * r + 2], kf.Pbar[4 * r + 3]);
                ph[r] = make_float4(kf.Phat[4 * r], kf.Phat[4 * r + 1], kf.Phat[4 * r + 2], kf.Phat[4 * r + 3]);
            }
            const double base = a.cnllr[i];
            a.miss_cnllr[i] = base - log(1.0 - Pd);
            if ((long long)off + cnt <= a.cap) {
                int k = 0;
                for_each_gated(*a.grid, a.cell_start, a.gz, a.gidx, kf, a.model.eta2,
                               [&](int m, double, double, double) { a.pair_meas[off + (k++)] = m; });
                sort_run(a.pair_meas + off, cnt);
                for (k = 0; k < cnt; ++k) {
                    const int m = a.pair_meas[off + k];
                    const double2 z = a.z[m];
                    const double v0 = z.x - kf.zhat[0], v1 = z.y - kf.zhat[1];
                    const double d2 = nis_f64(kf.si, v0, v1);
                    double xh[4];
                    filter_f64(kf, v0, v1, xh);
                    double2 *xo = (double2 *)(a.pair_xhat + 4 * (size_t)(off + k));
                    xo[0] = make_double2(xh[0], xh[1]);
                    xo[1] = make_double2(xh[2], xh[3]);
                    a.pair_cnllr[off + k] = base + (0.5 * d2 + kf.logterm);
                    a.meas_used[m] = 1;
                }
            }
        }
        __syncthreads();
    }
}

static inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace mht

using namespace mht;

extern "C" int64_t mht_gate_batch_workspace(int64_t L, int64_t M) {
    const int64_t ntiles = (L + kTile - 1) / kTile;
    return grid_workspace_bytes(M) + align256((L + 1) * 4) + align256((ntiles + 2) * 4) + 1024;
}

extern "C" int mht_gate_batch(const mht_model *model, int64_t L, int64_t M, const double *d_x0, const float *d_P0,
                              const double *d_Pd, const double *d_cnllr, const double *d_z, double *d_x_bar,
                              float *d_P_bar, float *d_P_hat, double *d_miss_cnllr, int32_t *d_pair_off,
                              int32_t *d_pair_meas, double *d_pair_cnllr, double *d_pair_xhat, int64_t cap,
                              uint8_t *d_meas_used, void *d_work, void *stream) {
    if (int rc = check_device()) return rc;
    if (!model || L < 0 || M < 0 || L > 0x7ffffff0ll || M > 0x7ffffff0ll || cap < 0 || !d_work) {
        set_error("mht_gate_batch: invalid argument (L=%lld M=%lld cap=%lld)", (long long)L, (long long)M,
                  (long long)cap);
        return MHT_E_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *w = (char *)d_work;
    GridDesc *grid = (GridDesc *)w;
    w += 256;
    int *cell_start = (int *)w;
    w += (int64_t)(kGridMaxCells + 64) * sizeof(int);
    int *cell_fill = (int *)w;
    w += (int64_t)(kGridMaxCells + 64) * sizeof(int);
    double2 *gz = (double2 *)w;
    w += (M + 16) * (int64_t)sizeof(double2);
    int *gidx = (int *)w;
    w = (char *)d_work + grid_workspace_bytes(M);
    int *count = (int *)w;
    w += align256((L + 1) * 4);
    int *tile_sum = (int *)w;
    const int ntiles = (int)((L + kTile - 1) / kTile);

    MHT_CUDA(cudaMemsetAsync(d_meas_used, 0, (size_t)M, s));
    if (int rc = launch_grid_build(d_z, (int)M, grid, cell_start, cell_fill, gz, gidx, s)) return rc;
    if (L == 0) {
        MHT_CUDA(cudaMemsetAsync(d_pair_off, 0, sizeof(int32_t), s));
        MHT_CUDA(cudaStreamSynchronize(s));
        return MHT_OK;
    }
    GateBatchArgs a;
    a.model = *model;
    a.L = (int)L;
    a.x0 = d_x0;
    a.P0 = d_P0;
    a.Pd = d_Pd;
    a.cnllr = d_cnllr;
    a.grid = grid;
    a.cell_start = cell_start;
    a.gz = gz;
    a.gidx = gidx;
    a.z = (const double2 *)d_z;
    a.count = count;
    a.tile_sum = tile_sum;
    a.x_bar = d_x_bar;
    a.P_bar = d_P_bar;
    a.P_hat = d_P_hat;
    a.miss_cnllr = d_miss_cnllr;
    a.pair_off = d_pair_off;
    a.pair_meas = d_pair_meas;
    a.pair_cnllr = d_pair_cnllr;
    a.pair_xhat = d_pair_xhat;
    a.cap = cap;
    a.meas_used = d_meas_used;
    const int grid_dim = ntiles < kSMs * 4 ? ntiles : kSMs * 4;
    count_launch(), gate_batch_count_kernel<<<grid_dim, kTile, 0, s>>>(a);
    MHT_CUDA(cudaGetLastError());
    count_launch(), scan_tiles_kernel<<<1, 1024, 0, s>>>(tile_sum, ntiles, nullptr);
    MHT_CUDA(cudaGetLastError());
    count_launch(), gate_batch_emit_kernel<<<grid_dim, kTile, 0, s>>>(a);
    MHT_CUDA(cudaGetLastError());
    int total = 0;
    MHT_CUDA(cudaMemcpyAsync(&total, tile_sum + ntiles, sizeof(int), cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    if (total > cap) {
        set_error("mht_gate_batch: %d gated pairs exceed capacity %lld", total, (long long)cap);
        return MHT_E_CAPACITY;
    }
    return MHT_OK;
}

extern "C" int mht_gate_batch_host(const mht_model *model, int64_t L, int64_t M, const double *h_x0,
                                   const float *h_P0, const double *h_Pd, const double *h_cnllr, const double *h_z,
                                   double *h_x_bar, float *h_P_bar, float *h_P_hat, double *h_miss_cnllr,
                                   int32_t *h_pair_off, int32_t *h_pair_meas, double *h_pair_cnllr,
                                   double *h_pair_xhat, int64_t cap, uint8_t *h_meas_used) {
    if (int rc = check_device()) return rc;
    if (!model || L < 0 || M < 0 || cap < 0) {
        set_error("mht_gate_batch_host: invalid argument");
        return MHT_E_INVALID;
    }
    const int64_t wbytes = mht_gate_batch_workspace(L, M);
    const int64_t Lp = L ? L : 1, Mp = M ? M : 1, cp = cap ? cap : 1;
    // one arena: inputs | outputs | workspace
    const int64_t sz[] = {Lp * 32, Lp * 64, Lp * 8, Lp * 8, Mp * 16,                  // x0 P0 Pd cnllr z
                          Lp * 32, Lp * 64, Lp * 64, Lp * 8, (Lp + 1) * 4, cp * 4, cp * 8, cp * 32, Mp,
                          wbytes};
    int64_t off[16], tot = 0;
    for (int i = 0; i < 15; ++i) {
        off[i] = tot;
        tot += align256(sz[i]);
    }
    char *d = nullptr;
    MHT_CUDA(cudaMalloc(&d, (size_t)tot));
    cudaStream_t s = 0;
    int rc = MHT_OK;
    do {
        if (cudaMemcpyAsync(d + off[0], h_x0, L * 32, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d + off[1], h_P0, L * 64, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d + off[2], h_Pd, L * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d + off[3], h_cnllr, L * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d + off[4], h_z, M * 16, cudaMemcpyHostToDevice, s) != cudaSuccess) {
            set_error("mht_gate_batch_host: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MHT_E_CUDA;
            break;
        }
        rc = mht_gate_batch(model, L, M, (double *)(d + off[0]), (float *)(d + off[1]), (double *)(d + off[2]),
                            (double *)(d + off[3]), (double *)(d + off[4]), (double *)(d + off[5]),
                            (float *)(d + off[6]), (float *)(d + off[7]), (double *)(d + off[8]),
                            (int32_t *)(d + off[9]), (int32_t *)(d + off[10]), (double *)(d + off[11]),
                            (double *)(d + off[12]), cap, (uint8_t *)(d + off[13]), d + off[14], s);
        if (rc != MHT_OK && rc != MHT_E_CAPACITY) break;
        cudaMemcpy(h_pair_off, d + off[9], (L + 1) * 4, cudaMemcpyDeviceToHost);
        if (rc == MHT_E_CAPACITY) break;
        const int64_t G = L ? h_pair_off[L] : 0;
        if (cudaMemcpy(h_x_bar, d + off[5], L * 32, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_P_bar, d + off[6], L * 64, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_P_hat, d + off[7], L * 64, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_miss_cnllr, d + off[8], L * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_pair_meas, d + off[10], G * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_pair_cnllr, d + off[11], G * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_pair_xhat, d + off[12], G * 32, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h_meas_used, d + off[13], M, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("mht_gate_batch_host: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = MHT_E_CUDA;
        }
    } while (0);
    cudaFree(d);
    return rc;
}
