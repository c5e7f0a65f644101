// Device-resident hypothesis forest: steps 1-3 + terminate + N-scan prune of
// Tracker.addMeasurementList (reference pymht/tracker.py:194-259) with every tree node in HBM.
//
// Layout.  One LEVEL per scan (ring of N+2 levels).  A level holds the hypotheses created by that
// scan as structure-of-arrays in the reference's DFS leaf order (Target.getLeafNodes,
// pymht/pyTarget.py:461-471): x[4] f64, cNLLR f64, measurementNumber i32, tree i32, parent link i32.
// Covariances are stored once per PARENT (P_bar for its miss child, P_hat shared by all its gated
// children, exactly the sharing of pyTarget.py:239-254) instead of once per child.
// Every leaf also carries its root->leaf measurement path as W = N+1 int32 "row planes"
// (row = plane*max_meas + measurement index, -1 = miss): the association columns, the clusters and the
// N-scan prune all stream these planes and never chase parent pointers.  Because children are written
// in ascending measurementNumber order, the leaves of a tree are sorted lexicographically by path, so
// N-scan pruning (Target.pruneDepth, pyTarget.py:343-356) keeps ONE contiguous range per tree found
// by binary search -- no compaction pass, the pruned level stays in place as the node store.
#include <new>
#include <vector>
#include <string.h>
#include <stdlib.h>

#include "assoc.cuh"

namespace mht {

struct Level {
    double2 *xa;     // [cap_nodes] (x, y)
    double2 *xb;     // [cap_nodes] (vx, vy)
    double *cnllr;   // [cap_nodes]
    int *meas;       // [cap_nodes]  measurementNumber (0 = miss / initial)
    int *pidx;       // [cap_nodes]  index into this level's Pbar/Phat tables
    int *tree;       // [cap_nodes]
    float *Pbar;     // [cap_ptab][16]
    float *Phat;     // [cap_ptab][16]
    int *tree_off;   // [T+1] children range per tree
    int *par_lo;     // [T+1] live range start (positions in the previous level) used to build this level
    int *par_off;    // [T+1] exclusive scan of live range lengths
};

struct TreeState {   // device arrays, one entry per tree slot
    int *root_scan, *init_scan, *alive, *window, *live_lo, *live_hi;
    double *root_cnllr, *Pd, *miss;
};

struct TrackOut {    // per-scan results, device + pinned host mirror
    int *pos, *status, *meas, *advanced, *root_meas;
    double *x, *cnllr, *root_x, *root_cnllr;
    float *P, *root_P;
};

struct ScanStatus {  // device status word copied back every scan
    int n_parents, n_children, overflow, n_dead;
    int assoc[kAssocInfo];
    unsigned long long bb_nodes;
    double lower_bound, objective;
};

struct TrunkNode {
    int scan, meas;
    double x[4], cnllr;
    float P[16];
};

}  // namespace mht

using namespace mht;

struct mht_forest {
    mht_forest_config cfg;
    int W, nslots, T;          // planes, levels in the ring, tree slots in use
    int scan;                  // scans processed so far (= current level number)
    int64_t cap_nodes, cap_par, cap_ptab;
    int64_t bytes;
    char *arena;
    Level lv[MHT_MAX_WINDOW + 2];
    int *rows[2];              // path planes, double buffered: [W][cap_nodes]
    TreeState ts;
    TrackOut out_d, out_h;
    char *out_h_base, *out_d_base;
    int64_t out_bytes;
    ScanStatus *status_d, *status_h;
    int *d_np, *d_nc;          // live parents / children of the scan in flight
    int *count, *tile_sum;
    char *grid_ws;
    double *z_d, *z_h;         // staged scan
    unsigned char *used_d, *used_h;
    void *assoc_ws;
    AssocWork aw;
    double *hist_d, *hist_h;   // history walk buffer
    cudaStream_t stream;
    cudaEvent_t ev[6];
    // host mirrors
    std::vector<int> h_alive, h_root_scan, h_init_scan, h_last_pos;
    std::vector<std::vector<TrunkNode>> trunk;
    int64_t h_level_nodes;     // nodes in the current level (for initiate)
    int64_t h_level_ptab;      // P-table entries in the current level
    std::vector<int> last_tracks;  // tree slots reported by the last scan
};

namespace mht {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
    mht_model model;
    Level prev, cur;
    TreeState ts;
    const int *rows_prev;
    int *rows_cur;
    long long stride;     // cap_nodes
    int W, T, scan, max_meas;
    const GridDesc *grid;
    const int *cell_start;
    const double2 *gz;
    const int *gidx;
    const double2 *z;
    int *count, *tile_sum;
    int *d_np, *d_nc;
    unsigned char *used;
    ScanStatus *status;
    long long cap_nodes, cap_par;
    int *uf, *row_owner, *row_multi;   // association union-find state (clusters are built while emitting)
};

// exclusive scan of live range lengths over tree slots (single CTA; T <= ~10^5)
__global__ void __launch_bounds__(1024, 1) live_scan_kernel(ScanArgs a) {
    __shared__ int part[1024];
    const int T = a.T;
    const int per = (T + blockDim.x - 1) / blockDim.x;
    const int lo = min(T, (int)threadIdx.x * per), hi = min(T, lo + per);
    int s = 0;
    for (int t = lo; t < hi; ++t) s += a.ts.alive[t] ? a.ts.live_hi[t] - a.ts.live_lo[t] : 0;
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) {
            const int v = part[i];
            part[i] = acc;
            acc += v;
        }
        *a.d_np = acc;
        a.status->n_parents = acc;
        a.status->overflow = acc > a.cap_par ? 1 : 0;
        a.status->n_dead = 0;
        if (acc > a.cap_par) *a.d_nc = 0;
        a.cur.par_off[T] = acc;
        a.cur.par_lo[T] = 0;
    }
    __syncthreads();
    int acc = part[threadIdx.x];
    for (int t = lo; t < hi; ++t) {
        a.cur.par_off[t] = acc;
        a.cur.par_lo[t] = a.ts.live_lo[t];
        acc += a.ts.alive[t] ? a.ts.live_hi[t] - a.ts.live_lo[t] : 0;
    }
}

// live index -> (tree, position in the previous level)
__device__ __forceinline__ void locate(const ScanArgs &a, int i, int &t, int &pos) {
    int lo = 0, hi = a.T;  // largest t with par_off[t] <= i and a non-empty range
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.cur.par_off[mid] <= i) lo = mid; else hi = mid;
    }
    t = lo;
    pos = a.cur.par_lo[t] + (i - a.cur.par_off[t]);
}

__device__ __forceinline__ void load_leaf(const ScanArgs &a, int pos, double x0[4], float P0[16]) {
    const double2 x01 = a.prev.xa[pos], x23 = a.prev.xb[pos];
    x0[0] = x01.x;
    x0[1] = x01.y;
    x0[2] = x23.x;
    x0[3] = x23.y;
    const float *tab = a.prev.meas[pos] ? a.prev.Phat : a.prev.Pbar;
    const float4 *pp = (const float4 *)(tab + 16 * (size_t)a.prev.pidx[pos]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float4 v = pp[r];
        P0[4 * r] = v.x;
        P0[4 * r + 1] = v.y;
        P0[4 * r + 2] = v.z;
        P0[4 * r + 3] = v.w;
    }
}

__device__ __forceinline__ int block_scan_excl(int v, int *total) {
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

// pass 1: children per live leaf (1 miss + gated) and per-tile sums
__global__ void __launch_bounds__(kTile) forest_count_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    const int np = *a.d_np;
    const int ntiles = (np + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        int cnt = 0;
        if (i < np) {
            int t, pos;
            locate(a, i, t, pos);
            double x0[4];
            float P0[16];
            load_leaf(a, pos, x0, P0);
            LeafKF kf;
            leaf_kf<false>(a.model, x0, P0, a.ts.Pd[t], kf);
            cnt = 1;
            for_each_gated(*a.grid, a.cell_start, a.gz, kf, a.model.eta2,
                           [&](int, double, double, double) { ++cnt; });
            a.count[i] = cnt;
        }
        int total;
        block_scan_excl(cnt, &total);
        if (threadIdx.x == 0) a.tile_sum[tile] = total;
        __syncthreads();
    }
}

// single CTA: exclusive scan of the tile sums; total children -> d_nc
__global__ void __launch_bounds__(1024, 1) forest_scan_tiles_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    const int np = *a.d_np;
    const int n = (np + kTile - 1) / kTile;
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? a.tile_sum[i] : 0;
        int total;
        const int excl = block_scan_excl(v, &total);
        if (i < n) a.tile_sum[i] = (int)min(carry + excl, (long long)0x7fffffff);
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool over = carry > a.cap_nodes;
        *a.d_nc = over ? 0 : (int)carry;
        a.status->n_children = (int)min(carry, (long long)0x7fffffff);
        if (over) a.status->overflow = 2;
    }
}

// pass 2: write the new level.  Child order per leaf = [miss, gated by ascending measurement index]
// (Target.spawnNewNodes, pyTarget.py:239-254).
//   phase A (one thread per live leaf): Kalman quantities -> shared memory, P_bar/P_hat -> HBM, gated
//            measurement indices (grid order) -> scratch at the leaf's child offset;
//   phase B (one thread per CHILD, consecutive threads = consecutive children): rank the measurement
//            among its siblings (ascending index), filter, score, and store every field coalesced.
constexpr int kEmitD = 13;  // doubles per leaf in smem: xbar[4] zhat[2] si[4] logterm miss_cnllr base_cnllr
__host__ __device__ inline size_t emit_smem_bytes(int W) {
    return (size_t)kTile * (kEmitD * 8 + 8 * 4 + 4 + 4 * W) + 264 * 4;
}

__global__ void __launch_bounds__(kTile) forest_emit_kernel(ScanArgs a, int *scratch) {
    if (a.status->overflow) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_d = (double *)smem_raw;                 // [kEmitD][kTile]
    float *s_K = (float *)(s_d + kEmitD * kTile);      // [8][kTile]
    int *s_tree = (int *)(s_K + 8 * kTile);            // [kTile]
    int *s_off = s_tree + kTile;                       // [kTile+1] child offsets inside the tile
    int *s_path = s_off + 264;                         // [W][kTile]
    const int np = *a.d_np;
    const int ntiles = (np + kTile - 1) / kTile;
    const int plane_cur = a.scan % a.W;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + tid;
        const int cnt = (i < np) ? a.count[i] : 0;
        int total;
        const int loc = block_scan_excl(cnt, &total);
        const int tile_base = a.tile_sum[tile];
        s_off[tid] = loc;
        if (tid == 0) s_off[kTile] = total;
        if (i < np) {
            a.count[i] = tile_base + loc;  // child offset, read back by tree_off_kernel
            int t, pos;
            locate(a, i, t, pos);
            double x0[4];
            float P0[16];
            load_leaf(a, pos, x0, P0);
            LeafKF kf;
            leaf_kf<true>(a.model, x0, P0, a.ts.Pd[t], kf);
            float4 *pb = (float4 *)(a.cur.Pbar + 16 * (size_t)i), *ph = (float4 *)(a.cur.Phat + 16 * (size_t)i);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                pb[r] = make_float4(kf.Pbar[4 * r], kf.Pbar[4 * r + 1], kf.Pbar[4 * r + 2], kf.Pbar[4 * r + 3]);
                ph[r] = make_float4(kf.Phat[4 * r], kf.Phat[4 * r + 1], kf.Phat[4 * r + 2], kf.Phat[4 * r + 3]);
            }
            const double base = a.prev.cnllr[pos];
#pragma unroll
            for (int q = 0; q < 4; ++q) s_d[q * kTile + tid] = kf.xbar[q];
            s_d[4 * kTile + tid] = kf.zhat[0];
            s_d[5 * kTile + tid] = kf.zhat[1];
#pragma unroll
            for (int q = 0; q < 4; ++q) s_d[(6 + q) * kTile + tid] = kf.si[q];
            s_d[10 * kTile + tid] = kf.logterm;
            s_d[11 * kTile + tid] = base + a.ts.miss[t];
            s_d[12 * kTile + tid] = base;
#pragma unroll
            for (int q = 0; q < 8; ++q) s_K[q * kTile + tid] = kf.K[q];
            s_tree[tid] = t;
            // inherited path planes (entries at or above the tree's root are dropped)
            const int root_scan = a.ts.root_scan[t];
            for (int w = 0; w < a.W; ++w) {
                const int back = ((a.scan - w) % a.W + a.W) % a.W;  // scans since plane w was written
                const int s_w = a.scan - back;
                const int r =
                    (w != plane_cur && s_w > root_scan) ? a.rows_prev[(long long)w * a.stride + pos] : -1;
                s_path[w * kTile + tid] = r;
                // cluster step (tracker.py:961-974): every measurement on the path links this tree to
                // the other trees using it; all children inherit these rows
                if (r >= 0) uf_touch_row(a.uf, a.row_owner, a.row_multi, r, t);
            }
            int k = 0;
            int *dst = scratch + tile_base + loc + 1;
            for_each_gated(*a.grid, a.cell_start, a.gz, kf, a.model.eta2,
                           [&](int p, double, double, double) { dst[k++] = a.gidx[p]; });
        }
        __syncthreads();
        for (int c = tid; c < total; c += kTile) {
            int lo = 0, hi = kTile;  // parent = largest p with s_off[p] <= c
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= c) lo = mid; else hi = mid;
            }
            const int p = lo;
            const int k = c - s_off[p];
            const int first = tile_base + s_off[p];
            const int t = s_tree[p];
            int g, row_new, mnum;
            double xo0, xo1, xo2, xo3, cn;
            if (k == 0) {
                g = first;
                row_new = -1;
                mnum = 0;
                xo0 = s_d[0 * kTile + p];
                xo1 = s_d[1 * kTile + p];
                xo2 = s_d[2 * kTile + p];
                xo3 = s_d[3 * kTile + p];
                cn = s_d[11 * kTile + p];
            } else {
                const int nsib = s_off[p + 1] - s_off[p] - 1;
                const int m = scratch[first + k];
                int rank = 0;
                for (int q = 1; q <= nsib; ++q) rank += scratch[first + q] < m;
                g = first + 1 + rank;
                row_new = plane_cur * a.max_meas + m;
                mnum = m + 1;
                const double2 z = a.z[m];
                const double v0 = z.x - s_d[4 * kTile + p], v1 = z.y - s_d[5 * kTile + p];
                const double si[4] = {s_d[6 * kTile + p], s_d[7 * kTile + p], s_d[8 * kTile + p], s_d[9 * kTile + p]};
                const double d2 = nis_f64(si, v0, v1);
                xo0 = s_d[0 * kTile + p] + fma((double)s_K[1 * kTile + p], v1, (double)s_K[0 * kTile + p] * v0);
                xo1 = s_d[1 * kTile + p] + fma((double)s_K[3 * kTile + p], v1, (double)s_K[2 * kTile + p] * v0);
                xo2 = s_d[2 * kTile + p] + fma((double)s_K[5 * kTile + p], v1, (double)s_K[4 * kTile + p] * v0);
                xo3 = s_d[3 * kTile + p] + fma((double)s_K[7 * kTile + p], v1, (double)s_K[6 * kTile + p] * v0);
                cn = s_d[12 * kTile + p] + (0.5 * d2 + s_d[10 * kTile + p]);
                a.used[m] = 1;
                uf_touch_row(a.uf, a.row_owner, a.row_multi, row_new, t);
            }
            a.cur.xa[g] = make_double2(xo0, xo1);
            a.cur.xb[g] = make_double2(xo2, xo3);
            a.cur.cnllr[g] = cn;
            a.cur.meas[g] = mnum;
            a.cur.pidx[g] = tile * kTile + p;
            a.cur.tree[g] = t;
            for (int w = 0; w < a.W; ++w)
                a.rows_cur[(long long)w * a.stride + g] = (w == plane_cur) ? row_new : s_path[w * kTile + p];
        }
        __syncthreads();
    }
}

__global__ void tree_off_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    const int np = *a.d_np, nc = *a.d_nc;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= a.T; t += gridDim.x * blockDim.x) {
        const int po = (t < a.T) ? a.cur.par_off[t] : np;
        a.cur.tree_off[t] = (po < np) ? a.count[po] : nc;
    }
}

struct UpdateArgs {
    Level lv[MHT_MAX_WINDOW + 2];
    int nslots, W, T, scan, N;
    long long stride;
    TreeState ts;
    const int *rows;        // planes of the current level
    const int *sel;         // selected column per tree
    TrackOut out;
    ScanStatus *status;
    const int *assoc_info;
    const unsigned long long *bb_nodes;
    const double *objective;
    mht_model model;
    double score_upper, cnllr_upper, radar_range, px, py;
};

__device__ __forceinline__ int path_cmp(const UpdateArgs &a, int p, int q, int s_lo, int s_hi) {
    for (int s = s_lo; s <= s_hi; ++s) {
        const long long pl = (long long)(s % a.W) * a.stride;
        const int vp = a.rows[pl + p], vq = a.rows[pl + q];
        if (vp != vq) return vp < vq ? -1 : 1;
    }
    return 0;
}

// per tree: report the selected hypothesis, terminate (tracker.py:891-916), N-scan prune
// (tracker.py:1219-1231): new root = ancestor N_t scans above the selected leaf; the surviving leaves
// are the contiguous range sharing the selected leaf's path prefix.
__global__ void track_update_kernel(UpdateArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        for (int i = 0; i < kAssocInfo; ++i) a.status->assoc[i] = a.assoc_info[i];
        a.status->bb_nodes = *a.bb_nodes;
        a.status->lower_bound = a.objective[0];
        a.status->objective = a.objective[1];
    }
    if (t >= a.T) return;
    a.out.pos[t] = -1;
    a.out.status[t] = -1;
    a.out.advanced[t] = 0;
    if (!a.ts.alive[t] || a.status->overflow) return;
    const Level &cur = a.lv[a.scan % a.nslots];
    const int sel = a.sel[t];
    const double cn = cur.cnllr[sel];
    double x[4];
    {
        const double2 p01 = cur.xa[sel], p23 = cur.xb[sel];
        x[0] = p01.x; x[1] = p01.y; x[2] = p23.x; x[3] = p23.y;
    }
    const int meas = cur.meas[sel];
    a.out.pos[t] = sel;
    a.out.meas[t] = meas;
    a.out.cnllr[t] = cn;
    for (int i = 0; i < 4; ++i) a.out.x[4 * t + i] = x[i];
    {
        const float *tab = (meas ? cur.Phat : cur.Pbar) + 16 * (size_t)cur.pidx[sel];
        for (int i = 0; i < 16; ++i) a.out.P[16 * t + i] = tab[i];
    }
    // termination tests, in the reference's order
    double zx = 0.0, zy = 0.0;
    for (int k = 0; k < 4; ++k) {
        zx = fma((double)a.model.C[k], x[k], zx);
        zy = fma((double)a.model.C[4 + k], x[k], zy);
    }
    const double dist = sqrt((zx - a.px) * (zx - a.px) + (zy - a.py) * (zy - a.py));
    int status = 0;
    if (dist > a.radar_range) status = 1;
    else if ((cn - a.ts.root_cnllr[t]) / (double)(a.N + 1) > a.score_upper) status = 2;
    else if (cn > a.cnllr_upper) status = 2;
    a.out.status[t] = status;
    if (status) {
        a.ts.alive[t] = 0;
        atomicAdd(&a.status->n_dead, 1);
        return;
    }
    const int root_old = a.ts.root_scan[t];
    const int root_new = max(root_old, a.scan - a.ts.window[t]);
    int lo = cur.tree_off[t], hi = cur.tree_off[t + 1];
    if (root_new > root_old) {
        // walk up from the selected leaf to the new root
        int pos = sel;
        for (int s = a.scan; s > root_new; --s) {
            const Level &L = a.lv[s % a.nslots];
            pos = L.par_lo[t] + (L.pidx[pos] - L.par_off[t]);
        }
        const Level &LR = a.lv[root_new % a.nslots];
        a.ts.root_cnllr[t] = LR.cnllr[pos];
        a.ts.root_scan[t] = root_new;
        a.out.advanced[t] = root_new - root_old;
        a.out.root_meas[t] = LR.meas[pos];
        a.out.root_cnllr[t] = LR.cnllr[pos];
        {
            const double2 p01 = LR.xa[pos], p23 = LR.xb[pos];
            a.out.root_x[4 * t] = p01.x; a.out.root_x[4 * t + 1] = p01.y;
            a.out.root_x[4 * t + 2] = p23.x; a.out.root_x[4 * t + 3] = p23.y;
        }
        const float *tab = (LR.meas[pos] ? LR.Phat : LR.Pbar) + 16 * (size_t)LR.pidx[pos];
        for (int i = 0; i < 16; ++i) a.out.root_P[16 * t + i] = tab[i];
        // contiguous range of leaves whose path agrees with the selected leaf on (root_old, root_new]
        int l = lo, h = hi;
        while (l < h) {  // lower bound
            const int mid = (l + h) >> 1;
            if (path_cmp(a, mid, sel, root_old + 1, root_new) < 0) l = mid + 1; else h = mid;
        }
        const int first = l;
        h = hi;
        while (l < h) {  // upper bound
            const int mid = (l + h) >> 1;
            if (path_cmp(a, mid, sel, root_old + 1, root_new) <= 0) l = mid + 1; else h = mid;
        }
        lo = first;
        hi = l;
    }
    a.ts.live_lo[t] = lo;
    a.ts.live_hi[t] = hi;
}

// window part of one track's history: nodes from the root (exclusive) down to position `pos`
constexpr int kHistRec = 24;  // doubles per history record: meas, cnllr, x[4], scan, pad, P[16]
__global__ void history_kernel(UpdateArgs a, int t, int pos, int scan_from, double *out) {
    if (threadIdx.x || blockIdx.x) return;
    int n = 0;
    const int root = a.ts.root_scan[t];
    for (int s = scan_from; s > root; --s) {
        const Level &L = a.lv[s % a.nslots];
        double *o = out + kHistRec + kHistRec * n++;
        o[0] = (double)L.meas[pos];
        o[1] = L.cnllr[pos];
        {
            const double2 p01 = L.xa[pos], p23 = L.xb[pos];
            o[2] = p01.x; o[3] = p01.y; o[4] = p23.x; o[5] = p23.y;
        }
        o[6] = (double)s;
        const float *tab = (L.meas[pos] ? L.Phat : L.Pbar) + 16 * (size_t)L.pidx[pos];
        for (int i = 0; i < 16; ++i) o[8 + i] = (double)tab[i];
        pos = L.par_lo[t] + (L.pidx[pos] - L.par_off[t]);
    }
    out[0] = (double)n;
}

// smallest distance from (px,py) to the position of any live leaf (Target.haveNoNeightbours,
// pymht/pyTarget.py:181-189)
__global__ void min_leaf_distance_kernel(Level cur, TreeState ts, int T, double px, double py,
                                         unsigned long long *out) {
    unsigned long long best = ~0ull;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        if (!ts.alive[t]) continue;
        for (int p = ts.live_lo[t] + threadIdx.x; p < ts.live_hi[t]; p += blockDim.x) {
            const double2 q = cur.xa[p];
            const double dx = q.x - px, dy = q.y - py;
            const unsigned long long k = f64_key(sqrt(dx * dx + dy * dy));
            best = k < best ? k : best;
        }
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(out, best);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static inline int64_t al(int64_t b) { return (b + 255) / 256 * 256; }

template <class T> static T *carve(char *&p, int64_t n) {
    T *r = (T *)p;
    p += al(n * (int64_t)sizeof(T));
    return r;
}

static void carve_out(char *&p, int T, TrackOut *o) {
    o->pos = carve<int>(p, T);
    o->status = carve<int>(p, T);
    o->meas = carve<int>(p, T);
    o->advanced = carve<int>(p, T);
    o->root_meas = carve<int>(p, T);
    o->x = carve<double>(p, 4 * (int64_t)T);
    o->cnllr = carve<double>(p, T);
    o->root_x = carve<double>(p, 4 * (int64_t)T);
    o->root_cnllr = carve<double>(p, T);
    o->P = carve<float>(p, 16 * (int64_t)T);
    o->root_P = carve<float>(p, 16 * (int64_t)T);
}

static int forest_layout(mht_forest *f, bool commit) {
    const int T = f->cfg.max_trees;
    const int64_t cn = f->cap_nodes, cp = f->cap_ptab;
    char *p = commit ? f->arena : nullptr;
    char *p0 = p;
    for (int s = 0; s < f->nslots; ++s) {
        Level &L = f->lv[s];
        L.xa = carve<double2>(p, cn);
        L.xb = carve<double2>(p, cn);
        L.cnllr = carve<double>(p, cn);
        L.meas = carve<int>(p, cn);
        L.pidx = carve<int>(p, cn);
        L.tree = carve<int>(p, cn);
        L.Pbar = carve<float>(p, 16 * cp);
        L.Phat = carve<float>(p, 16 * cp);
        L.tree_off = carve<int>(p, T + 1);
        L.par_lo = carve<int>(p, T + 1);
        L.par_off = carve<int>(p, T + 1);
    }
    for (int b = 0; b < 2; ++b) f->rows[b] = carve<int>(p, (int64_t)f->W * cn);
    f->ts.root_scan = carve<int>(p, T);
    f->ts.init_scan = carve<int>(p, T);
    f->ts.alive = carve<int>(p, T);
    f->ts.window = carve<int>(p, T);
    f->ts.live_lo = carve<int>(p, T);
    f->ts.live_hi = carve<int>(p, T);
    f->ts.root_cnllr = carve<double>(p, T);
    f->ts.Pd = carve<double>(p, T);
    f->ts.miss = carve<double>(p, T);
    f->out_d_base = p;
    carve_out(p, T, &f->out_d);
    f->out_bytes = p - f->out_d_base;
    f->status_d = carve<ScanStatus>(p, 1);
    f->d_np = carve<int>(p, 2);
    f->d_nc = f->d_np + 1;
    f->count = carve<int>(p, f->cap_par + 1);
    f->tile_sum = carve<int>(p, f->cap_par / kTile + 4);
    f->grid_ws = carve<char>(p, grid_workspace_bytes(f->cfg.max_meas));
    f->z_d = carve<double>(p, 2 * (int64_t)f->cfg.max_meas);
    f->used_d = carve<unsigned char>(p, f->cfg.max_meas);
    f->hist_d = carve<double>(p, kHistRec * (MHT_MAX_WINDOW + 4));
    const int64_t n_rows = (int64_t)f->W * f->cfg.max_meas;
    const int64_t cap_cand = cn < (int64_t)T * 256 ? cn : (int64_t)T * 256;
    f->assoc_ws = p;
    p += al(assoc_workspace_bytes(cn, T, n_rows, cap_cand));
    if (commit) assoc_carve(f->assoc_ws, cn, T, n_rows, cap_cand, &f->aw);
    f->bytes = p - p0;
    return MHT_OK;
}

static void fill_update_args(mht_forest *f, UpdateArgs *u) {
    for (int s = 0; s < f->nslots; ++s) u->lv[s] = f->lv[s];
    u->nslots = f->nslots;
    u->W = f->W;
    u->T = f->T;
    u->scan = f->scan;
    u->N = f->cfg.n_scan_window;
    u->stride = f->cap_nodes;
    u->ts = f->ts;
    u->rows = f->rows[f->scan & 1];
    u->sel = f->aw.sel;
    u->out = f->out_d;
    u->status = f->status_d;
    u->assoc_info = f->aw.info;
    u->bb_nodes = f->aw.bb_nodes;
    u->objective = f->aw.objective;
    u->model = f->cfg.model;
    u->score_upper = f->cfg.score_upper;
    u->cnllr_upper = f->cfg.cnllr_upper;
    u->radar_range = f->cfg.radar_range;
    u->px = f->cfg.position[0];
    u->py = f->cfg.position[1];
}

static int forest_scan_impl(mht_forest *f, int64_t M, const double *d_z, mht_scan_info *info,
                            unsigned char *h_used) {
    if (M < 0 || M > f->cfg.max_meas) {
        set_error("mht_forest_scan: %lld measurements exceed max_meas=%d", (long long)M, f->cfg.max_meas);
        return MHT_E_CAPACITY;
    }
    cudaStream_t s = f->stream;
    const int k = f->scan + 1;
    if (f->T == 0) {  // no trees yet: the scan only advances the clock (tracker.py:207 loops over nothing)
        f->scan = k;
        f->h_level_nodes = 0;
        f->h_level_ptab = 0;
        f->last_tracks.clear();
        if (info) memset(info, 0, sizeof(*info));
        if (h_used && M) memset(h_used, 0, (size_t)M);
        return MHT_OK;
    }
    ScanArgs a;
    a.model = f->cfg.model;
    a.prev = f->lv[f->scan % f->nslots];
    a.cur = f->lv[k % f->nslots];
    a.ts = f->ts;
    a.rows_prev = f->rows[f->scan & 1];
    a.rows_cur = f->rows[k & 1];
    a.stride = f->cap_nodes;
    a.W = f->W;
    a.T = f->T;
    a.scan = k;
    a.max_meas = f->cfg.max_meas;
    char *w = f->grid_ws;
    GridDesc *grid = (GridDesc *)w;
    w += 256;
    int *cell_start = (int *)w;
    w += (int64_t)(kGridMaxCells + 64) * sizeof(int);
    int *cell_fill = (int *)w;
    w += (int64_t)(kGridMaxCells + 64) * sizeof(int);
    double2 *gz = (double2 *)w;
    w += (M + 16) * (int64_t)sizeof(double2);
    int *gidx = (int *)w;
    a.grid = grid;
    a.cell_start = cell_start;
    a.gz = gz;
    a.gidx = gidx;
    a.z = (const double2 *)d_z;
    a.count = f->count;
    a.tile_sum = f->tile_sum;
    a.d_np = f->d_np;
    a.d_nc = f->d_nc;
    a.used = f->used_d;
    a.status = f->status_d;
    a.cap_nodes = f->cap_nodes;
    a.cap_par = f->cap_par;
    a.uf = f->aw.uf;
    a.row_owner = f->aw.row_owner;
    a.row_multi = f->aw.row_mark;

    const int grid_dim = kSMs * 8;
    MHT_CUDA(cudaEventRecord(f->ev[0], s));
    ColView c;
    c.n_ptr = f->d_nc;
    c.idx = nullptr;
    c.meas = a.cur.meas;
    c.plane_new = k % f->W;
    c.cost = a.cur.cnllr;
    c.tree_base = f->ts.root_cnllr;
    c.tree = a.cur.tree;
    c.rows = a.rows_cur;
    c.stride = f->cap_nodes;
    c.width = f->W;
    c.n_trees = f->T;
    c.n_rows = f->W * f->cfg.max_meas;
    assoc_carve(f->assoc_ws, f->cap_nodes, f->cfg.max_trees, (int64_t)f->W * f->cfg.max_meas, f->aw.cap_cand,
                &f->aw);
    // warm start: measurement rows keep their ids for W scans, so last scan's multipliers are a good
    // starting point; the plane being recycled for this scan starts from zero
    MHT_CUDA(cudaMemsetAsync(f->aw.u + (size_t)(k % f->W) * f->cfg.max_meas, 0, sizeof(double) * f->cfg.max_meas, s));
    if (int rc = assoc_begin(c, f->aw, grid_dim, s, k > 1)) return rc;
    MHT_CUDA(cudaMemsetAsync(f->used_d, 0, (size_t)(M ? M : 1), s));
    live_scan_kernel<<<1, 1024, 0, s>>>(a);
    if (int rc = launch_grid_build(d_z, (int)M, grid, cell_start, cell_fill, gz, gidx, s)) return rc;
    forest_count_kernel<<<grid_dim, kTile, 0, s>>>(a);
    forest_scan_tiles_kernel<<<1, 1024, 0, s>>>(a);
    forest_emit_kernel<<<grid_dim, kTile, emit_smem_bytes(f->W), s>>>(a, (int *)f->aw.rc);
    tree_off_kernel<<<(f->T + 256) / 256, 256, 0, s>>>(a);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaEventRecord(f->ev[1], s));

    f->scan = k;
    static const long long sift_min = getenv("MHT_SIFT_MIN") ? atoll(getenv("MHT_SIFT_MIN")) : 1000000;
    const bool sift = f->h_level_nodes > sift_min;  // last scan's hypothesis count is the size hint
    if (int rc = assoc_solve(c, f->aw, f->cfg.max_dual_iters, 400000, kSMs * 8, s, f->ev[5], f->scan > 1, sift, true))
        return rc;
    MHT_CUDA(cudaEventRecord(f->ev[2], s));

    UpdateArgs u;
    fill_update_args(f, &u);
    track_update_kernel<<<(f->T + 127) / 128, 128, 0, s>>>(u);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaEventRecord(f->ev[3], s));
    MHT_CUDA(cudaMemcpyAsync(f->out_h_base, f->out_d_base, (size_t)f->out_bytes, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(f->status_h, f->status_d, sizeof(ScanStatus), cudaMemcpyDeviceToHost, s));
    if (h_used) MHT_CUDA(cudaMemcpyAsync(f->used_h, f->used_d, (size_t)M, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaEventRecord(f->ev[4], s));
    MHT_CUDA(cudaStreamSynchronize(s));
    if (h_used && M) memcpy(h_used, f->used_h, (size_t)M);

    const ScanStatus &st = *f->status_h;
    if (st.overflow) {
        set_error("mht_forest_scan: capacity exceeded (%s: need %d, have %lld); the forest is unchanged "
                  "for this scan -- recreate it with larger max_nodes/max_parents",
                  st.overflow == 1 ? "live leaves" : "hypotheses", st.overflow == 1 ? st.n_parents : st.n_children,
                  (long long)(st.overflow == 1 ? f->cap_par : f->cap_nodes));
        f->scan = k - 1;
        return MHT_E_CAPACITY;
    }
    // host mirrors: trunk records, alive flags, reported tracks
    f->last_tracks.clear();
    int n_live_before = 0;
    for (int t = 0; t < f->T; ++t) {
        if (!f->h_alive[t]) continue;
        ++n_live_before;
        f->last_tracks.push_back(t);
        f->h_last_pos[t] = f->out_h.pos[t];
        if (f->out_h.status[t] > 0) {
            f->h_alive[t] = 0;
            continue;
        }
        if (f->out_h.advanced[t] > 0) {
            TrunkNode nd;
            f->h_root_scan[t] += f->out_h.advanced[t];
            nd.scan = f->h_root_scan[t];
            nd.meas = f->out_h.root_meas[t];
            nd.cnllr = f->out_h.root_cnllr[t];
            memcpy(nd.x, f->out_h.root_x + 4 * t, sizeof(nd.x));
            memcpy(nd.P, f->out_h.root_P + 16 * t, sizeof(nd.P));
            f->trunk[t].push_back(nd);
        }
    }
    f->h_level_nodes = st.n_children;
    f->h_level_ptab = st.n_parents;
    if (info) {
        memset(info, 0, sizeof(*info));
        info->n_parents = st.n_parents;
        info->n_children = st.n_children;
        info->n_pairs = (int64_t)st.n_children - st.n_parents;
        info->n_trees = n_live_before;
        info->n_clusters = st.assoc[7];
        info->n_multi_clusters = st.assoc[8];
        info->n_dead = st.n_dead;
        info->dual_iters = st.assoc[1];
        info->certified = st.assoc[10];
        info->n_candidates = st.assoc[3];
        info->bb_nodes = (int64_t)st.bb_nodes;
        info->lower_bound = st.lower_bound;
        info->objective = st.objective;
        info->n_active = sift ? st.assoc[12] : 0;
        info->max_component = st.assoc[9];
        info->n_components = st.assoc[4];
        cudaEventElapsedTime(&info->ms_gate, f->ev[0], f->ev[1]);
        cudaEventElapsedTime(&info->ms_cluster, f->ev[1], f->ev[5]);
        cudaEventElapsedTime(&info->ms_assoc, f->ev[5], f->ev[2]);
        cudaEventElapsedTime(&info->ms_prune, f->ev[2], f->ev[4]);
        cudaEventElapsedTime(&info->ms_total, f->ev[0], f->ev[4]);
    }
    return MHT_OK;
}

}  // namespace mht

extern "C" int mht_forest_create(const mht_forest_config *cfg, mht_forest **out) {
    if (int rc = check_device()) return rc;
    if (!cfg || !out || cfg->n_scan_window < 1 || cfg->n_scan_window + 1 > MHT_MAX_WINDOW || cfg->max_trees < 1 ||
        cfg->max_trees >= (1 << 24) || cfg->max_meas < 1 || cfg->max_nodes < 16 || cfg->max_parents < 16 ||
        cfg->max_nodes > 0x7ffffff0ll || (int64_t)(cfg->n_scan_window + 1) * cfg->max_meas > 0x7ffffff0ll) {
        set_error("mht_forest_create: invalid configuration");
        return MHT_E_INVALID;
    }
    mht_forest *f = new (std::nothrow) mht_forest();
    if (!f) return MHT_E_INVALID;
    f->cfg = *cfg;
    if (f->cfg.max_dual_iters <= 0) f->cfg.max_dual_iters = 120;
    f->W = cfg->n_scan_window + 1;
    f->nslots = cfg->n_scan_window + 2;
    f->T = 0;
    f->scan = 0;
    f->cap_nodes = cfg->max_nodes;
    f->cap_par = cfg->max_parents;
    f->cap_ptab = cfg->max_parents + cfg->max_trees;
    f->arena = nullptr;
    forest_layout(f, false);
    if (cudaMalloc(&f->arena, (size_t)f->bytes) != cudaSuccess) {
        set_error("mht_forest_create: cudaMalloc(%lld bytes) failed: %s", (long long)f->bytes,
                  cudaGetErrorString(cudaGetLastError()));
        delete f;
        return MHT_E_CUDA;
    }
    forest_layout(f, true);
    const int T = cfg->max_trees;
    cudaError_t e = cudaMallocHost(&f->out_h_base, (size_t)f->out_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&f->status_h, sizeof(ScanStatus));
    if (e == cudaSuccess) e = cudaMallocHost(&f->z_h, 16 * (size_t)cfg->max_meas);
    if (e == cudaSuccess) e = cudaMallocHost(&f->used_h, (size_t)cfg->max_meas);
    if (e == cudaSuccess) e = cudaMallocHost(&f->hist_h, kHistRec * sizeof(double) * (MHT_MAX_WINDOW + 4));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&f->ev[i]);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(forest_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)emit_smem_bytes(MHT_MAX_WINDOW));
    if (e == cudaSuccess) e = cudaMemsetAsync(f->ts.alive, 0, sizeof(int) * T, f->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->rows[0], 0xff, sizeof(int) * (size_t)f->W * f->cap_nodes, f->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
    if (e != cudaSuccess) {
        set_error("mht_forest_create: %s", cudaGetErrorString(e));
        cudaFree(f->arena);
        delete f;
        return MHT_E_CUDA;
    }
    char *p = f->out_h_base;
    carve_out(p, T, &f->out_h);
    f->h_alive.assign(T, 0);
    f->h_root_scan.assign(T, 0);
    f->h_init_scan.assign(T, 0);
    f->h_last_pos.assign(T, -1);
    f->trunk.resize(T);
    f->h_level_nodes = 0;
    f->h_level_ptab = 0;
    *out = f;
    return MHT_OK;
}

extern "C" void mht_forest_destroy(mht_forest *f) {
    if (!f) return;
    cudaStreamSynchronize(f->stream);
    for (int i = 0; i < 6; ++i) cudaEventDestroy(f->ev[i]);
    cudaStreamDestroy(f->stream);
    cudaFreeHost(f->out_h_base);
    cudaFreeHost(f->status_h);
    cudaFreeHost(f->z_h);
    cudaFreeHost(f->used_h);
    cudaFreeHost(f->hist_h);
    cudaFree(f->arena);
    delete f;
}

extern "C" int64_t mht_forest_bytes(const mht_forest *f) { return f ? f->bytes : 0; }

extern "C" int mht_forest_initiate(mht_forest *f, const double x0[4], const float P0[16], double Pd, int32_t *slot) {
    if (!f || !x0 || !P0 || !(Pd > 0.0 && Pd < 1.0)) {
        set_error("mht_forest_initiate: invalid argument");
        return MHT_E_INVALID;
    }
    if (f->T >= f->cfg.max_trees || f->h_level_nodes >= f->cap_nodes || f->h_level_ptab >= f->cap_ptab) {
        set_error("mht_forest_initiate: capacity exceeded (trees %d/%d)", f->T, f->cfg.max_trees);
        return MHT_E_CAPACITY;
    }
    cudaStream_t s = f->stream;
    const int t = f->T;
    const Level &L = f->lv[f->scan % f->nslots];
    const int pos = (int)f->h_level_nodes, pi = (int)f->h_level_ptab;
    const int zero = 0, one = 1, scan = f->scan, win = f->cfg.n_scan_window, hi = pos + 1;
    const double cn = 0.0, miss = -log(1.0 - Pd);
    int minus[MHT_MAX_WINDOW];
    for (int i = 0; i < MHT_MAX_WINDOW; ++i) minus[i] = -1;
    MHT_CUDA(cudaMemcpyAsync(L.xa + pos, x0, 16, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.xb + pos, x0 + 2, 16, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.cnllr + pos, &cn, 8, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.meas + pos, &zero, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.pidx + pos, &pi, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.tree + pos, &t, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(L.Pbar + 16 * (size_t)pi, P0, 64, cudaMemcpyHostToDevice, s));
    for (int w = 0; w < f->W; ++w)
        MHT_CUDA(cudaMemcpyAsync(f->rows[f->scan & 1] + (int64_t)w * f->cap_nodes + pos, minus, 4,
                                 cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.root_scan + t, &scan, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.init_scan + t, &scan, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.alive + t, &one, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.window + t, &win, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.live_lo + t, &pos, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.live_hi + t, &hi, 4, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.root_cnllr + t, &cn, 8, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.Pd + t, &Pd, 8, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.miss + t, &miss, 8, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    f->h_level_nodes += 1;
    f->h_level_ptab += 1;
    f->h_alive[t] = 1;
    f->h_root_scan[t] = f->h_init_scan[t] = f->scan;
    f->h_last_pos[t] = pos;
    TrunkNode nd;
    nd.scan = f->scan;
    nd.meas = 0;
    nd.cnllr = 0.0;
    memcpy(nd.x, x0, sizeof(nd.x));
    memcpy(nd.P, P0, sizeof(nd.P));
    f->trunk[t].clear();
    f->trunk[t].push_back(nd);
    f->T = t + 1;
    if (slot) *slot = t;
    return MHT_OK;
}

extern "C" int mht_forest_scan(mht_forest *f, int64_t M, const double *h_z, double scan_time, mht_scan_info *info,
                               uint8_t *h_meas_used) {
    (void)scan_time;
    if (!f || (M > 0 && !h_z)) {
        set_error("mht_forest_scan: invalid argument");
        return MHT_E_INVALID;
    }
    if (M > f->cfg.max_meas) {
        set_error("mht_forest_scan: %lld measurements exceed max_meas=%d", (long long)M, f->cfg.max_meas);
        return MHT_E_CAPACITY;
    }
    if (M) memcpy(f->z_h, h_z, 16 * (size_t)M);  // stage through pinned memory
    MHT_CUDA(cudaMemcpyAsync(f->z_d, f->z_h, 16 * (size_t)M, cudaMemcpyHostToDevice, f->stream));
    return forest_scan_impl(f, M, f->z_d, info, h_meas_used ? h_meas_used : nullptr);
}

extern "C" int mht_forest_scan_device(mht_forest *f, int64_t M, const double *d_z, double scan_time,
                                      mht_scan_info *info) {
    (void)scan_time;
    if (!f || (M > 0 && !d_z)) {
        set_error("mht_forest_scan_device: invalid argument");
        return MHT_E_INVALID;
    }
    return forest_scan_impl(f, M, d_z, info, nullptr);
}

extern "C" int mht_forest_tracks(mht_forest *f, int32_t cap, int32_t *n, int32_t *h_slot, double *h_x, float *h_P,
                                 double *h_cnllr, int32_t *h_meas, int32_t *h_status) {
    if (!f || !n) return MHT_E_INVALID;
    const int cnt = (int)f->last_tracks.size();
    *n = cnt;
    if (cnt > cap) {
        set_error("mht_forest_tracks: %d tracks exceed cap %d", cnt, cap);
        return MHT_E_CAPACITY;
    }
    for (int i = 0; i < cnt; ++i) {
        const int t = f->last_tracks[i];
        if (h_slot) h_slot[i] = t;
        if (h_x) memcpy(h_x + 4 * i, f->out_h.x + 4 * t, 32);
        if (h_P) memcpy(h_P + 16 * i, f->out_h.P + 16 * t, 64);
        if (h_cnllr) h_cnllr[i] = f->out_h.cnllr[t];
        if (h_meas) h_meas[i] = f->out_h.meas[t];
        if (h_status) h_status[i] = f->out_h.status[t];
    }
    return MHT_OK;
}

extern "C" int mht_forest_min_leaf_distance(mht_forest *f, double px, double py, double *dist) {
    if (!f || !dist) return MHT_E_INVALID;
    unsigned long long *d = (unsigned long long *)f->hist_d, h = ~0ull;
    MHT_CUDA(cudaMemcpyAsync(d, &h, 8, cudaMemcpyHostToDevice, f->stream));
    if (f->T > 0) {
        min_leaf_distance_kernel<<<kSMs * 2, 256, 0, f->stream>>>(f->lv[f->scan % f->nslots], f->ts, f->T, px, py, d);
        MHT_CUDA(cudaGetLastError());
    }
    MHT_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, f->stream));
    MHT_CUDA(cudaStreamSynchronize(f->stream));
    if (h == ~0ull) {
        *dist = INFINITY;
    } else {
        unsigned long long u = (h & 0x8000000000000000ull) ? (h & 0x7fffffffffffffffull) : ~h;
        memcpy(dist, &u, 8);
    }
    return MHT_OK;
}

extern "C" int mht_forest_history(mht_forest *f, int32_t slot, int32_t cap, int32_t *n, int32_t *h_meas, double *h_x,
                                  double *h_cnllr, float *h_P) {
    if (!f || !n || slot < 0 || slot >= f->T) {
        set_error("mht_forest_history: invalid argument");
        return MHT_E_INVALID;
    }
    const std::vector<TrunkNode> &tr = f->trunk[slot];
    int wn = 0;
    if (f->h_last_pos[slot] >= 0 && f->scan > f->h_root_scan[slot]) {
        UpdateArgs u;
        fill_update_args(f, &u);
        history_kernel<<<1, 1, 0, f->stream>>>(u, slot, f->h_last_pos[slot], f->scan, f->hist_d);
        MHT_CUDA(cudaGetLastError());
        MHT_CUDA(cudaMemcpyAsync(f->hist_h, f->hist_d, kHistRec * sizeof(double) * (MHT_MAX_WINDOW + 4),
                                 cudaMemcpyDeviceToHost, f->stream));
        MHT_CUDA(cudaStreamSynchronize(f->stream));
        wn = (int)f->hist_h[0];
    }
    const int total = (int)tr.size() + wn;
    *n = total;
    if (total > cap) {
        set_error("mht_forest_history: %d nodes exceed cap %d", total, cap);
        return MHT_E_CAPACITY;
    }
    int i = 0;
    for (const TrunkNode &nd : tr) {
        if (h_meas) h_meas[i] = nd.meas;
        if (h_x) memcpy(h_x + 4 * i, nd.x, 32);
        if (h_cnllr) h_cnllr[i] = nd.cnllr;
        if (h_P) memcpy(h_P + 16 * i, nd.P, 64);
        ++i;
    }
    for (int k = wn - 1; k >= 0; --k, ++i) {  // the walk is leaf -> root; report oldest first
        const double *o = f->hist_h + kHistRec + kHistRec * k;
        if (h_meas) h_meas[i] = (int)o[0];
        if (h_cnllr) h_cnllr[i] = o[1];
        if (h_x) memcpy(h_x + 4 * i, o + 2, 32);
        if (h_P)
            for (int q = 0; q < 16; ++q) h_P[16 * i + q] = (float)o[8 + q];
    }
    return MHT_OK;
}

extern "C" int mht_forest_leaves(mht_forest *f, int32_t slot, int64_t cap, int64_t *n, double *h_x, double *h_cnllr,
                                 int32_t *h_meas) {
    if (!f || !n || slot < 0 || slot >= f->T) {
        set_error("mht_forest_leaves: invalid argument");
        return MHT_E_INVALID;
    }
    int lohi[2] = {0, 0};
    MHT_CUDA(cudaMemcpyAsync(&lohi[0], f->ts.live_lo + slot, 4, cudaMemcpyDeviceToHost, f->stream));
    MHT_CUDA(cudaMemcpyAsync(&lohi[1], f->ts.live_hi + slot, 4, cudaMemcpyDeviceToHost, f->stream));
    MHT_CUDA(cudaStreamSynchronize(f->stream));
    const int64_t cnt = f->h_alive[slot] ? lohi[1] - lohi[0] : 0;
    *n = cnt;
    if (cnt > cap) {
        set_error("mht_forest_leaves: %lld leaves exceed cap %lld", (long long)cnt, (long long)cap);
        return MHT_E_CAPACITY;
    }
    if (cnt == 0) return MHT_OK;
    const Level &L = f->lv[f->scan % f->nslots];
    if (h_x) {  // two planes -> interleaved [n][4]
        MHT_CUDA(cudaMemcpy2DAsync(h_x, 32, L.xa + lohi[0], 16, 16, (size_t)cnt, cudaMemcpyDeviceToHost, f->stream));
        MHT_CUDA(cudaMemcpy2DAsync(h_x + 2, 32, L.xb + lohi[0], 16, 16, (size_t)cnt, cudaMemcpyDeviceToHost, f->stream));
    }
    if (h_cnllr) MHT_CUDA(cudaMemcpyAsync(h_cnllr, L.cnllr + lohi[0], 8 * cnt, cudaMemcpyDeviceToHost, f->stream));
    if (h_meas) MHT_CUDA(cudaMemcpyAsync(h_meas, L.meas + lohi[0], 4 * cnt, cudaMemcpyDeviceToHost, f->stream));
    MHT_CUDA(cudaStreamSynchronize(f->stream));
    return MHT_OK;
}
