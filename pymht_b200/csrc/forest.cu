// Device-resident hypothesis forest: steps 1-3 + terminate + N-scan prune of
// Tracker.addMeasurementList (reference pymht/tracker.py:194-259) with every tree node in HBM.
//
// Layout.  One LEVEL per scan (ring of N+2 levels).  A level holds the hypotheses created by that
// scan as structure-of-arrays in the reference's DFS leaf order (Target.getLeafNodes,
// pymht/pyTarget.py:461-471): x[4] f64, cNLLR f64, measurementNumber i32, tree i32, parent link i32.
// Covariances are NOT stored per node: with the reference's constant A, Q, C, R (one shared model for all
// nodes, kalman.py:55-64,82-101) the covariance chain of a node depends only on its tree's root covariance
// and on the hit/miss pattern of the path below the root, so every scan a small kernel evaluates the
// float32 chain once per (tree, pattern) -- at most 2^N entries per tree -- into an L2-resident table of
// S^-1, K, gate box and NLLR log term, and each leaf carries 16 bits of hit/miss history (`pat`).  The
// values are bit-identical to evaluating the same chain per leaf.
// Every leaf also carries its root->leaf measurement path as W = N+1 int32 "row planes"
// (row = plane*max_meas + measurement index, -1 = miss): the association columns, the clusters and the
// N-scan prune all stream these planes and never chase parent pointers.  Because children are written
// in ascending measurementNumber order, the leaves of a tree are sorted lexicographically by path, so
// N-scan pruning (Target.pruneDepth, pyTarget.py:343-356) keeps ONE contiguous range per tree found
// by binary search -- no compaction pass, the pruned level stays in place as the node store.
#include <algorithm>
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
    int *pidx;       // [cap_nodes]  live index of the parent in the scan that created the node
    int *tree;       // [cap_nodes]
    unsigned short *pat;  // [cap_nodes] hit(1)/miss(0) history of the path, bit 0 = this node's scan
    int *tree_off;   // [T+1] children range per tree
    int *par_lo;     // [T+1] live range start (positions in the previous level) used to build this level
    int *par_off;    // [T+1] exclusive scan of live range lengths
};

struct TreeState {   // device arrays, one entry per tree slot
    int *root_scan, *init_scan, *alive, *window, *live_lo, *live_hi;
    double *root_cnllr, *Pd, *miss;
    float *rootP;    // [T][16] covariance of the current root node
};

// Per (tree, hit/miss pattern) gate quantities, heap-indexed: h = (1 << depth) | pattern bits.
struct PatGate {
    float si[4];     // S^-1
    float K[8];      // Kalman gain
    float logterm;   // ln(lambda_ex sqrt(det(2 pi S)) / P_d)
    float pad[3];
    double hx, hy;   // half extents of the gate ellipse's bounding box
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
    int rows_active, pad;
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
    int64_t cap_nodes, cap_par;
    int PT;                    // pattern-table entries per tree = 2^(N+1)
    PatGate *pt_gate;          // [max_trees][PT]
    float *pt_P;               // [max_trees][PT][16] posterior covariance per pattern
    int *tile_tree;            // [cap_par/kTile + 4] tree of the first live leaf of every tile
    int *glist;                // [cap_par][kInline] inline gated lists
    int2 *heavy_list;          // [cap_par]
    unsigned *tm_bits;         // [max_trees][tm_words]
    int tm_words;
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
    double *histb_d, *histb_h; // [kDeadChunk][kHistStride] batched walks of the tracks that died this scan
    int *dead_d, *dead_h;      // [2][kDeadChunk] slot, position
    std::vector<std::vector<double>> dead_hist;   // per slot: window records (leaf -> root) kept for dead tracks
    cudaStream_t stream;
    cudaEvent_t ev[6];
    cudaEvent_t evx[10];       // dual-loop launches (4 pairs) + exact repair begin/end
    int n_dual_ev = 0;
    // host mirrors
    std::vector<int> h_alive, h_root_scan, h_init_scan, h_last_pos;
    std::vector<std::vector<TrunkNode>> trunk;
    int64_t h_level_nodes;     // nodes in the current level (for initiate)
    std::vector<int> last_tracks;  // tree slots reported by the last scan
    std::vector<int> free_slots;   // slots of dead tracks the host has released (mht_forest_release): reused by initiate
    bool open_scan = false;        // mht_forest_grow done, mht_forest_select pending
    // dynamic window (Tracker.__dynamicWindow, tracker.py:918-950): size criterion on device, roof from the host
    int dyn_window = 0, target_size_limit = 3000, window_roof = 0;
    double *adv_d = nullptr, *adv_h = nullptr;   // [T][kMaxAdv - 1][kHistRec] nodes a root skipped over in one scan
    int64_t open_M = 0, open_children = 0;
    int64_t h_level_nodes_open() const { return open_children; }
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
    int *count, *tile_sum, *tile_tree;
    int *glist;           // [cap_par][kInline] ascending gated measurement indices per live leaf
    int2 *heavy_list;            // worklist (live index, tree) of leaves gated by the warp-per-leaf kernel
    int *heavy_n;
    int heavy_rows, heavy_cand;  // thresholds: grid rows / candidate measurements under the gate box
    int use_tma;                 // emit kernel: stage the warps' gated-list tiles with cp.async.bulk + mbarrier
    unsigned *tm_bits;           // [max_trees][tm_words] (tree, measurement) pairs already linked this scan
    int tm_words;
    long long *pool_ctr;
    int *d_np, *d_nc;
    PatGate *pt_gate;
    float *pt_P;
    int PT, N;
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
        *a.heavy_n = 0;
        *a.pool_ctr = 0;
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
        const int len = a.ts.alive[t] ? a.ts.live_hi[t] - a.ts.live_lo[t] : 0;
        // tiles whose first live leaf belongs to this tree (the gate kernels start their tree search here)
        // (nothing past the table when the live leaves exceed max_parents: the scan is refused anyway)
        const int k_max = (int)(a.cap_par / kTile) + 2;
        for (int k = (acc + kTile - 1) / kTile; (long long)k * kTile < (long long)acc + len && k <= k_max; ++k)
            a.tile_tree[k] = t;
        acc += len;
    }
    if (threadIdx.x == 0)   // sentinel for the last tile
        a.tile_tree[min((*a.d_np + kTile - 1) / kTile, (int)(a.cap_par / kTile) + 3)] = T - 1;
}

// live index -> (tree, position in the previous level): largest t in [t_lo, t_hi] with par_off[t] <= i.
// Callers pass the tree range of their tile (one tree almost always), so the search is 0-2 steps.
__device__ __forceinline__ int locate_tree(const ScanArgs &a, int i, int t_lo, int t_hi) {
    int lo = t_lo, hi = t_hi + 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.cur.par_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Everything a leaf needs for the gate: float64 state prediction + its pattern-table entry.
__device__ __forceinline__ const PatGate *load_leaf(const ScanArgs &a, int t, int pos, LeafKF &kf, int &ent) {
    const double2 x01 = a.prev.xa[pos], x23 = a.prev.xb[pos];
    const double x0[4] = {x01.x, x01.y, x23.x, x23.y};
    state_kf(a.model, x0, kf);
    const int d = min(a.N, (a.scan - 1) - a.ts.root_scan[t]);        // depth of the leaf below its root
    const int h = (1 << d) | ((int)a.prev.pat[pos] & ((1 << d) - 1));
    ent = t * a.PT + h;
    const PatGate *e = a.pt_gate + ent;
    const float4 si = __ldg((const float4 *)e->si);
    kf.si[0] = (double)si.x;
    kf.si[1] = (double)si.y;
    kf.si[2] = (double)si.z;
    kf.si[3] = (double)si.w;
    kf.hx = __ldg(&e->hx);
    kf.hy = __ldg(&e->hy);
    return e;
}

// One CTA per tree: the float32 covariance chain for every hit/miss pattern of depth <= the tree's
// current depth, level by level from the root covariance (kalman.py:55-64 predict, :82-101 precalc).
__global__ void __launch_bounds__(128) pat_table_kernel(ScanArgs a) {
    const int t = blockIdx.x;
    if (a.status->overflow || t >= a.T || !a.ts.alive[t]) return;
    const int dmax = min(a.N, (a.scan - 1) - a.ts.root_scan[t]);
    float *P = a.pt_P + (size_t)t * a.PT * 16;
    PatGate *G = a.pt_gate + (size_t)t * a.PT;
    if (threadIdx.x < 16) P[16 + threadIdx.x] = a.ts.rootP[16 * t + threadIdx.x];
    __syncthreads();
    const double Pd = a.ts.Pd[t];
    for (int d = 0; d <= dmax; ++d) {
        for (int c = threadIdx.x; c < (1 << d); c += blockDim.x) {
            const int h = (1 << d) | c;
            float P0[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) P0[q] = P[16 * h + q];
            LeafKF kf;
            cov_kf<true>(a.model, P0, Pd, kf);
            PatGate g;
#pragma unroll
            for (int q = 0; q < 4; ++q) g.si[q] = (float)kf.si[q];
#pragma unroll
            for (int q = 0; q < 8; ++q) g.K[q] = kf.K[q];
            g.logterm = (float)kf.logterm;
            g.pad[0] = g.pad[1] = g.pad[2] = 0.0f;
            g.hx = kf.hx;
            g.hy = kf.hy;
            G[h] = g;
            if (d < dmax) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    P[16 * (2 * h) + q] = kf.Pbar[q];
                    P[16 * (2 * h + 1) + q] = kf.Phat[q];
                }
            }
        }
        __syncthreads();
    }
}

// Posterior covariance of the node `depth` scans below a root with covariance rootP whose path has the
// hit/miss bits `pat` (bit 0 = the node's own scan): the same chain pat_table_kernel evaluates.
__device__ void chain_P(const mht_model &m, const float *rootP, unsigned pat, int depth, double Pd, float *out) {
    float P[16];
    for (int q = 0; q < 16; ++q) P[q] = rootP[q];
    for (int j = depth - 1; j >= 0; --j) {
        LeafKF kf;
        cov_kf<true>(m, P, Pd, kf);
        const bool hit = (pat >> j) & 1u;
        for (int q = 0; q < 16; ++q) P[q] = hit ? kf.Phat[q] : kf.Pbar[q];
    }
    for (int q = 0; q < 16; ++q) out[q] = P[q];
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

// Bitonic sorting network on registers (compile-time indices only).
template <int N>
__device__ __forceinline__ void sort_net(int (&r)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const int x = r[i], y = r[l];
                    r[i] = up ? min(x, y) : max(x, y);
                    r[l] = up ? max(x, y) : min(x, y);
                }
            }
}

// tree t gates measurement m this scan: mark the measurement used and link the tree to the row's other
// users (tracker.py:331-332, :961-974) -- once per (tree, measurement): a bitmap filters the repeats
// (thousands of leaves of a tree gate the same measurement).
__device__ __forceinline__ void link_new_row(const ScanArgs &a, int plane_cur, int t, int m);

// pass 1 -- the gate proper, one thread per live leaf: children of the leaf (1 miss + gated) and its
// gated measurement indices, ASCENDING, in an inline list of kInline slots (collected in shared memory,
// sorted by a register network).  Leaves whose gate box holds very many candidate measurements (several
// recent misses blow the gate up) or more than kInline gated ones would make their whole warp wait: they
// go to a worklist for the warp-per-leaf kernel below.  The cluster links of the new measurement rows
// are made here, from registers.
constexpr int kInline = 16;
__global__ void __launch_bounds__(kTile, 4) forest_gate_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    __shared__ int s_lst[kInline][kTile];
    const int np = *a.d_np;
    const int ntiles = (np + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
    const int plane_cur = a.scan % a.W;
    const GridDesc g = *a.grid;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + tid;
        int t = -1, n = 0;
        bool heavy = false;
        if (i < np) {
            t = locate_tree(a, i, a.tile_tree[tile], a.tile_tree[tile + 1]);
            const int pos = a.cur.par_lo[t] + (i - a.cur.par_off[t]);
            LeafKF kf;
            int ent;
            load_leaf(a, t, pos, kf, ent);
            int cx0, cx1, cy0, cy1;
            if (gate_box(g, kf, cx0, cx1, cy0, cy1)) {
                heavy = cy1 - cy0 + 1 > a.heavy_rows;
                if (!heavy) {
                    int ncand = 0;
                    for (int cy = cy0; cy <= cy1; ++cy)
                        ncand += a.cell_start[cy * g.nx + cx1 + 1] - a.cell_start[cy * g.nx + cx0];
                    heavy = ncand > a.heavy_cand;
                }
                if (!heavy) {
                    for_each_gated_box(g, a.cell_start, a.gz, a.gidx, kf, a.model.eta2, cx0, cx1, cy0, cy1,
                                       [&](int m, double, double, double) {
                                           if (n < kInline) s_lst[n][tid] = m;
                                           ++n;
                                       });
                    heavy = n > kInline;
                }
            }
            if (!heavy) a.count[i] = n + 1;
        }
        // worklist of heavy leaves, one atomic per warp
        const unsigned hm = __ballot_sync(0xffffffffu, heavy);
        if (hm) {
            int base = 0;
            if (lane == __ffs(hm) - 1) base = atomicAdd(a.heavy_n, __popc(hm));
            base = __shfl_sync(0xffffffffu, base, __ffs(hm) - 1);
            if (heavy) a.heavy_list[base + __popc(hm & ((1u << lane) - 1))] = make_int2(i, t);
        }
        if (heavy) n = 0;
        // sort (network size chosen per warp), store, link
        const int t_up = __shfl_up_sync(0xffffffffu, t, 1);
        int4 *dst = (int4 *)(a.glist + (size_t)kInline * i);
        if (__any_sync(0xffffffffu, n > 8)) {
            int r[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) r[q] = q < n ? s_lst[q][tid] : 0x7fffffff;
            sort_net<16>(r);
#pragma unroll
            for (int q = 0; q < 16; q += 4)
                if (q < n) dst[q >> 2] = make_int4(r[q], r[q + 1], r[q + 2], r[q + 3]);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int m_up = __shfl_up_sync(0xffffffffu, r[q], 1);
                if (q < n && !(lane > 0 && m_up == r[q] && t_up == t)) link_new_row(a, plane_cur, t, r[q]);
            }
        } else {
            int r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = q < n ? s_lst[q][tid] : 0x7fffffff;
            sort_net<8>(r);
#pragma unroll
            for (int q = 0; q < 8; q += 4)
                if (q < n) dst[q >> 2] = make_int4(r[q], r[q + 1], r[q + 2], r[q + 3]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int m_up = __shfl_up_sync(0xffffffffu, r[q], 1);
                if (q < n && !(lane > 0 && m_up == r[q] && t_up == t)) link_new_row(a, plane_cur, t, r[q]);
            }
        }
    }
}

__device__ __forceinline__ void link_new_row(const ScanArgs &a, int plane_cur, int t, int m) {
    const int r = plane_cur * a.max_meas + m;
    unsigned *word = a.tm_bits + (size_t)t * a.tm_words + (r >> 5);
    const unsigned bit = 1u << (r & 31);
    if (*word & bit) return;
    if (atomicOr(word, bit) & bit) return;
    a.used[m] = 1;
    uf_touch_row(a.uf, a.row_owner, a.row_multi, r, t);
}
// same for a row inherited from the path (any plane)
__device__ __forceinline__ void link_old_row(const ScanArgs &a, int t, int r) {
    unsigned *word = a.tm_bits + (size_t)t * a.tm_words + (r >> 5);
    const unsigned bit = 1u << (r & 31);
    if (*word & bit) return;
    if (atomicOr(word, bit) & bit) return;
    uf_touch_row(a.uf, a.row_owner, a.row_multi, r, t);
}

// pass 1b -- heavy leaves, one WARP per leaf: lanes stride over the candidate measurements of every grid
// row under the gate box, the gated indices are ranked (ascending) and written to a run of the pool;
// glist[leaf][0] = start of the run.
constexpr int kHeavyStage = 512;  // gated indices staged in shared memory per warp
__global__ void __launch_bounds__(kTile) forest_gate_heavy_kernel(ScanArgs a, int *pool, long long pool_cap) {
    if (a.status->overflow) return;
    __shared__ int s_stage[kTile / 32][kHeavyStage];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nh = *a.heavy_n;
    const int plane_cur = a.scan % a.W;
    const GridDesc g = *a.grid;
    const unsigned lt = (1u << lane) - 1;
    for (int h = blockIdx.x * (kTile / 32) + wib; h < nh; h += gridDim.x * (kTile / 32)) {
        const int2 it = a.heavy_list[h];
        const int i = it.x, t = it.y;
        const int pos = a.cur.par_lo[t] + (i - a.cur.par_off[t]);
        LeafKF kf;
        int ent;
        load_leaf(a, t, pos, kf, ent);
        int cx0, cx1, cy0, cy1;
        gate_box(g, kf, cx0, cx1, cy0, cy1);   // true: the leaf was sent here because its box is large
        // pass A: count
        int n = 0;
        for (int cy = cy0; cy <= cy1; ++cy) {
            const int beg = a.cell_start[cy * g.nx + cx0], end = a.cell_start[cy * g.nx + cx1 + 1];
            for (int p0 = beg; p0 < end; p0 += 32) {
                const int p = p0 + lane;
                bool ok = false;
                if (p < end) {
                    const double2 z = a.gz[p];
                    ok = nis_f64(kf.si, z.x - kf.zhat[0], z.y - kf.zhat[1]) <= a.model.eta2;
                }
                n += __popc(__ballot_sync(0xffffffffu, ok));
            }
        }
        const bool staged = n <= kHeavyStage, inl = n <= kInline;
        long long off = 0;
        if (!inl) {
            if (lane == 0) off = atomicAdd((unsigned long long *)a.pool_ctr, (unsigned long long)(staged ? n : 2 * n));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (off + (staged ? n : 2 * n) > pool_cap) {   // more children than the level can hold anyway
                if (lane == 0) a.status->overflow = 2;
                continue;
            }
        }
        // pass B: collect (grid order), then rank
        int *buf = staged ? s_stage[wib] : pool + off + n;
        int *out = inl ? a.glist + (size_t)kInline * i : pool + off;   // few gated after all: inline list
        int k = 0;
        for (int cy = cy0; cy <= cy1; ++cy) {
            const int beg = a.cell_start[cy * g.nx + cx0], end = a.cell_start[cy * g.nx + cx1 + 1];
            for (int p0 = beg; p0 < end; p0 += 32) {
                const int p = p0 + lane;
                bool ok = false;
                int m = 0;
                if (p < end) {
                    const double2 z = a.gz[p];
                    m = a.gidx[p];
                    ok = nis_f64(kf.si, z.x - kf.zhat[0], z.y - kf.zhat[1]) <= a.model.eta2;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (ok) buf[k + __popc(mask & lt)] = m;
                k += __popc(mask);
            }
        }
        __syncwarp();
        for (int e = lane; e < n; e += 32) {
            const int m = buf[e];
            int rank = 0;
            for (int q = 0; q < n; ++q) rank += buf[q] < m;
            out[rank] = m;
            link_new_row(a, plane_cur, t, m);
        }
        if (lane == 0) {
            a.count[i] = n + 1;
            if (!inl) a.glist[(size_t)kInline * i] = (int)off;
        }
        __syncwarp();
    }
}

// pass 1c: children counts -> tile-local exclusive offsets + per-tile sums
__global__ void __launch_bounds__(kTile) forest_count_scan_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    const int np = *a.d_np;
    const int ntiles = (np + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        const int cnt = (i < np) ? a.count[i] : 0;
        int total;
        const int excl = block_scan_excl(cnt, &total);
        if (i < np) a.count[i] = excl;
        if (threadIdx.x == 0) a.tile_sum[tile] = total;
        __syncthreads();
    }
}

// single CTA: exclusive scan of the tile sums (tile_sum[n] = total); total children -> d_nc
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
        a.tile_sum[n] = (int)min(carry, (long long)0x7fffffff);
        a.status->n_children = (int)min(carry, (long long)0x7fffffff);
        if (over) a.status->overflow = 2;
    }
}

// pass 2: write the new level.  Child order per leaf = [miss, gated by ascending measurement index]
// (Target.spawnNewNodes, pyTarget.py:239-254).  Every WARP owns 32 consecutive live leaves and never
// synchronises with the rest of its CTA (child offsets come from pass 1):
//   phase A (lane = leaf): state prediction, pattern-table entry, inherited path planes and their
//            cluster links;
//   phase B (lane = CHILD, consecutive lanes = consecutive children): measurement from the parent's sorted
//            list (inline, or its run in the pool), filter, score, and every field stored coalesced.
constexpr int kEmitD = 8;    // doubles per leaf in smem: xbar[4] zhat[2] base_cnllr miss_cnllr
__host__ __device__ inline size_t emit_warp_bytes(int W) {
    return (size_t)32 * kEmitD * 8 + (size_t)(3 + W) * 32 * 4 + 36 * 4 + 32 * kInline * 4 + 16;   // + the warp's mbarrier
}

// ---- TMA (bulk async copy engine) staging of a warp's tile: one cp.async.bulk global -> shared, completion on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_tile(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    // generic-proxy reads of the previous tile are ordered before the async-proxy write (the warp synchronised already)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "MHT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra MHT_DONE;\n\t"
        "bra MHT_WAIT;\n\t"
        "MHT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase), "r"(0x989680u)
        : "memory");
}
__host__ __device__ inline size_t emit_smem_bytes(int W) { return (kTile / 32) * emit_warp_bytes(W); }

template <int WMAX>
__global__ void __launch_bounds__(kTile) forest_emit_kernel(ScanArgs a, const int *__restrict__ pool) {
    if (a.status->overflow) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)wib * emit_warp_bytes(a.W);
    double *s_d = (double *)wbase;             // [kEmitD][32]
    int *s_tree = (int *)(s_d + kEmitD * 32);  // [32]
    int *s_ent = s_tree + 32;                  // [32] pattern-table entry
    int *s_pat = s_ent + 32;                   // [32] hit/miss history of the leaf
    int *s_off = s_pat + 32;                   // [33] child offsets inside the warp (+3 pad)
    int *s_lst = s_off + 36;                   // [32][kInline] the leaves' sorted gated lists (16-byte aligned: TMA destination)
    int *s_path = s_lst + 32 * kInline;        // [W][32]
    unsigned long long *s_bar = (unsigned long long *)(s_path + a.W * 32);   // the warp's mbarrier
    unsigned bar_phase = 0;
    if (a.use_tma) {
        if (lane == 0) mbar_init(s_bar, 1);
        __syncwarp();
    }
    const int np = *a.d_np;
    const int nwt = (np + 31) >> 5;
    const int W = a.W;
    const int plane_cur = a.scan % W;
    const int warps_per_cta = kTile / 32;
    for (int wt = blockIdx.x * warps_per_cta + wib; wt < nwt; wt += gridDim.x * warps_per_cta) {
        const int i = wt * 32 + lane;
        const int tile = (wt * 32) / kTile;
        const bool valid = i < np;
        // the 32 leaves' inline lists are ONE contiguous, 16-byte aligned 2 KB block: the TMA engine stages it while the
        // lanes fetch the rest of phase A (waited for right before phase B)
        if (a.use_tma && lane == 0) bulk_load_tile(s_lst, a.glist + (size_t)kInline * (size_t)wt * 32, 32 * kInline * 4, s_bar);
        const int tile_base = a.tile_sum[tile], tile_end = a.tile_sum[tile + 1];
        const int off_i = valid ? tile_base + a.count[i] : tile_end;
        const int off_n = (valid && i + 1 < np && ((i + 1) & (kTile - 1))) ? tile_base + a.count[i + 1] : tile_end;
        const int warp_first = __shfl_sync(0xffffffffu, off_i, 0);
        const int total = __shfl_sync(0xffffffffu, off_n, 31) - warp_first;
        s_off[lane] = off_i - warp_first;
        if (lane == 0) s_off[32] = total;
        int t = -1, pos = 0, root_scan = 0x7fffffff;
        if (valid) {
            t = locate_tree(a, i, a.tile_tree[tile], a.tile_tree[tile + 1]);
            pos = a.cur.par_lo[t] + (i - a.cur.par_off[t]);
            root_scan = a.ts.root_scan[t];
        }
        // inherited path planes (entries at or above the tree's root are dropped): every load is issued
        // before the first use.  Plane w holds the rows of scan  a.scan - ((a.scan - w) mod W).
        int rr[WMAX];
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
            int back = plane_cur - w;        // (a.scan - w) mod W, given plane_cur = a.scan mod W
            back += back < 0 ? W : 0;
            const bool take = w < W && back != 0 && a.scan - back > root_scan;
            rr[w] = take ? a.rows_prev[(long long)w * a.stride + pos] : -1;
        }
        if (!a.use_tma) {   // the leaf's inline list -> shared memory (only the quarters that hold entries)
            const int nl = off_n - off_i - 1;
            if (valid && nl <= kInline) {
                const int4 *src = (const int4 *)(a.glist + (size_t)kInline * i);
                int4 *dst = (int4 *)(s_lst + kInline * lane);
#pragma unroll
                for (int q = 0; q < kInline / 4; ++q)
                    if (4 * q < nl) dst[q] = src[q];
            }
        }
        if (valid) {
            LeafKF kf;
            int ent;
            load_leaf(a, t, pos, kf, ent);
            const double base = a.prev.cnllr[pos];
#pragma unroll
            for (int q = 0; q < 4; ++q) s_d[q * 32 + lane] = kf.xbar[q];
            s_d[4 * 32 + lane] = kf.zhat[0];
            s_d[5 * 32 + lane] = kf.zhat[1];
            s_d[6 * 32 + lane] = base;
            s_d[7 * 32 + lane] = base + a.ts.miss[t];
            s_tree[lane] = t;
            s_ent[lane] = ent;
            s_pat[lane] = (int)a.prev.pat[pos];
        }
        // cluster step (tracker.py:961-974): every measurement on the path links this tree to the other
        // trees using it; leaves are sorted by path, so a lane repeats its left neighbour's (row, tree)
        // most of the time and skips the touch
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
                const int r = rr[w];
                s_path[w * 32 + lane] = r;
                const int r_up = __shfl_up_sync(0xffffffffu, r, 1), t_up = __shfl_up_sync(0xffffffffu, t, 1);
                if (r >= 0 && !(lane > 0 && r_up == r && t_up == t)) link_old_row(a, t, r);
            }
        }
        __syncwarp();
        if (a.use_tma) {
            mbar_wait(s_bar, bar_phase);
            bar_phase ^= 1u;
        }
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int c = c0 + lane;
            const bool live = c < total;
            int lo = 0, hi = 32;  // parent = largest p with s_off[p] <= c
#pragma unroll
            for (int step = 0; step < 5; ++step) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= c) lo = mid; else hi = mid;
            }
            const int p = lo;
            const int k = live ? c - s_off[p] : 0;
            const int first = warp_first + s_off[p];
            const int t_p = s_tree[p];
            const int ip = wt * 32 + p;      // live index of the parent
            int g = first, row_new = -1, m = -1;
            if (live && k > 0) {
                const int nsib = s_off[p + 1] - s_off[p] - 1;
                // heavy leaves: a run in the pool whose start is the first entry of the leaf's list
                m = (nsib <= kInline) ? s_lst[kInline * p + (k - 1)] : pool[a.glist[(size_t)kInline * ip] + (k - 1)];
                g = first + k;
                row_new = plane_cur * a.max_meas + m;
            }
            if (!live) continue;
            double xo0 = s_d[0 * 32 + p], xo1 = s_d[1 * 32 + p], xo2 = s_d[2 * 32 + p], xo3 = s_d[3 * 32 + p];
            double cn = s_d[7 * 32 + p];
            if (m >= 0) {
                const PatGate *e = a.pt_gate + s_ent[p];
                const float4 sif = __ldg((const float4 *)e->si);
                const float4 K0 = __ldg((const float4 *)e->K), K1 = __ldg((const float4 *)(e->K + 4));
                const double logterm = (double)__ldg(&e->logterm);
                const double2 z = a.z[m];
                const double v0 = z.x - s_d[4 * 32 + p], v1 = z.y - s_d[5 * 32 + p];
                const double si[4] = {(double)sif.x, (double)sif.y, (double)sif.z, (double)sif.w};
                const double d2 = nis_f64(si, v0, v1);
                xo0 = xo0 + fma((double)K0.y, v1, (double)K0.x * v0);
                xo1 = xo1 + fma((double)K0.w, v1, (double)K0.z * v0);
                xo2 = xo2 + fma((double)K1.y, v1, (double)K1.x * v0);
                xo3 = xo3 + fma((double)K1.w, v1, (double)K1.z * v0);
                cn = s_d[6 * 32 + p] + (0.5 * d2 + logterm);
            }
            a.cur.xa[g] = make_double2(xo0, xo1);
            a.cur.xb[g] = make_double2(xo2, xo3);
            a.cur.cnllr[g] = cn;
            a.cur.meas[g] = m + 1;
            a.cur.pidx[g] = ip;
            a.cur.tree[g] = t_p;
            a.cur.pat[g] = (unsigned short)(((s_pat[p] << 1) | (m >= 0)) & 0xffff);
            int *col = a.rows_cur + g;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) {
                if (w < W) *col = (w == plane_cur) ? row_new : s_path[w * 32 + p];
                col += a.stride;
            }
        }
        __syncwarp();
    }
}

__global__ void tree_off_kernel(ScanArgs a) {
    if (a.status->overflow) return;
    const int np = *a.d_np, nc = *a.d_nc;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= a.T; t += gridDim.x * blockDim.x) {
        const int po = (t < a.T) ? a.cur.par_off[t] : np;
        a.cur.tree_off[t] = (po < np) ? a.tile_sum[po / kTile] + a.count[po] : nc;
    }
}

constexpr int kMaxAdv = 4;      // levels a root may advance in one scan (N-scan pruning: 1; dynamic window: up to 3)
constexpr int kAdvRec = 24;     // = kHistRec: meas, cnllr, x[4], scan, pad, P[16]

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
    const int *row_n;
    mht_model model;
    double score_upper, cnllr_upper, radar_range, px, py;
    int dyn_window, target_size_limit, window_roof;
    double *adv;            // [T][kMaxAdv - 1][kHistRec]
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
        a.status->rows_active = *a.row_n;
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
    const int root_old = a.ts.root_scan[t];
    if (a.dyn_window) {
        // Tracker.__dynamicWindow (tracker.py:918-950), size criterion: a tree of more than targetSizeLimit nodes
        // (Target.getNumOfNodes, pyTarget.py:148-151: the root and everything below it, after this scan's growth)
        // loses one scan of window.  The surviving nodes of every level are the contiguous range between the
        // ancestors of the first and the last leaf.
        long long size = 1;
        int lo_p = cur.tree_off[t], hi_p = cur.tree_off[t + 1] - 1;
        for (int s2 = a.scan; s2 > root_old && hi_p >= lo_p; --s2) {
            size += hi_p - lo_p + 1;
            const Level &L = a.lv[s2 % a.nslots];
            lo_p = L.par_lo[t] + (L.pidx[lo_p] - L.par_off[t]);
            hi_p = L.par_lo[t] + (L.pidx[hi_p] - L.par_off[t]);
        }
        if (size > a.target_size_limit) a.ts.window[t] -= 1;
    }
    if (a.ts.window[t] > a.window_roof) a.ts.window[t] = a.window_roof;   // roof lowered by the host (tracker.py:943-950)
    const unsigned pat_sel = cur.pat[sel];
    const int depth_sel = min(a.scan - root_old, 16);
    const double Pd_t = a.ts.Pd[t];
    chain_P(a.model, a.ts.rootP + 16 * t, pat_sel, depth_sel, Pd_t, a.out.P + 16 * t);
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
    // Target.pruneDepth (pyTarget.py:343-356): the new root is window[t] scans above the selected leaf (the leaf
    // itself once the window is <= 0), never above the current root
    const int root_new = min(a.scan, max(root_old, a.scan - a.ts.window[t]));
    int lo = cur.tree_off[t], hi = cur.tree_off[t + 1];
    if (root_new > root_old) {
        // walk up from the selected leaf to the new root
        int pos = sel;
        for (int s = a.scan; s > root_new; --s) {
            const Level &L = a.lv[s % a.nslots];
            pos = L.par_lo[t] + (L.pidx[pos] - L.par_off[t]);
        }
        const Level &LR = a.lv[root_new % a.nslots];
        // nodes the root skips over when it advances more than one level (dynamic window): the host keeps them in
        // the track's trunk, oldest first at slot 0
        if (root_new - root_old > 1 && a.adv) {
            int p2 = pos;
            for (int s2 = root_new; s2 > root_old + 1; --s2) {
                const Level &L2 = a.lv[s2 % a.nslots];
                p2 = L2.par_lo[t] + (L2.pidx[p2] - L2.par_off[t]);
                const int lvl = s2 - 1;                       // p2 is the node of scan lvl
                const int slot = lvl - (root_old + 1);
                if (slot >= kMaxAdv - 1) continue;
                const Level &L3 = a.lv[lvl % a.nslots];
                double *o = a.adv + ((size_t)t * (kMaxAdv - 1) + slot) * kAdvRec;
                float Pn[16];
                chain_P(a.model, a.ts.rootP + 16 * t, pat_sel >> (depth_sel - (lvl - root_old)), lvl - root_old, Pd_t, Pn);
                o[0] = (double)L3.meas[p2];
                o[1] = L3.cnllr[p2];
                const double2 q01 = L3.xa[p2], q23 = L3.xb[p2];
                o[2] = q01.x; o[3] = q01.y; o[4] = q23.x; o[5] = q23.y;
                o[6] = (double)lvl;
                for (int i = 0; i < 16; ++i) o[8 + i] = (double)Pn[i];
            }
        }
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
        // covariance of the new root: the selected path's first (root_new - root_old) hit/miss bits
        float Pn[16];
        chain_P(a.model, a.ts.rootP + 16 * t, pat_sel >> (depth_sel - (root_new - root_old)), root_new - root_old,
                Pd_t, Pn);
        for (int i = 0; i < 16; ++i) {
            a.out.root_P[16 * t + i] = Pn[i];
            a.ts.rootP[16 * t + i] = Pn[i];
        }
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
__device__ void history_walk(const UpdateArgs &a, int t, int pos, int scan_from, double *out) {
    int n = 0;
    const int root = a.ts.root_scan[t];
    // covariances along the path: the chain from the root covariance over the leaf's hit/miss bits
    const unsigned pat_leaf = a.lv[scan_from % a.nslots].pat[pos];
    const double Pd_t = a.ts.Pd[t];
    for (int s = scan_from; s > root; --s) {
        const Level &L = a.lv[s % a.nslots];
        double *o = out + kHistRec + kHistRec * n++;
        float Pn[16];
        chain_P(a.model, a.ts.rootP + 16 * t, pat_leaf >> (scan_from - s), min(s - root, 16), Pd_t, Pn);
        o[0] = (double)L.meas[pos];
        o[1] = L.cnllr[pos];
        {
            const double2 p01 = L.xa[pos], p23 = L.xb[pos];
            o[2] = p01.x; o[3] = p01.y; o[4] = p23.x; o[5] = p23.y;
        }
        o[6] = (double)s;
        for (int i = 0; i < 16; ++i) o[8 + i] = (double)Pn[i];
        pos = L.par_lo[t] + (L.pidx[pos] - L.par_off[t]);
    }
    out[0] = (double)n;
}

__global__ void history_kernel(UpdateArgs a, int t, int pos, int scan_from, double *out) {
    if (threadIdx.x || blockIdx.x) return;
    history_walk(a, t, pos, scan_from, out);
}
// the same walk for every track that died this scan (one thread each): their window nodes leave the device
// store with the next scans, so the host keeps the records (mht_forest_history serves them from there)
constexpr int kHistStride = kHistRec * (MHT_MAX_WINDOW + 4);
constexpr int kDeadChunk = 256;
__global__ void history_batch_kernel(UpdateArgs a, const int *slot_pos, int n, int scan_from, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) history_walk(a, slot_pos[i], slot_pos[kDeadChunk + i], scan_from, out + (size_t)i * kHistStride);
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
    const int64_t cn = f->cap_nodes;
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
        L.pat = carve<unsigned short>(p, cn);
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
    f->ts.rootP = carve<float>(p, 16 * (int64_t)T);
    f->pt_gate = carve<PatGate>(p, (int64_t)T * f->PT);
    f->pt_P = carve<float>(p, 16 * (int64_t)T * f->PT);
    f->tile_tree = carve<int>(p, f->cap_par / kTile + 4);
    f->glist = carve<int>(p, 16 * (f->cap_par + 1));
    f->heavy_list = carve<int2>(p, f->cap_par + 1);
    f->tm_words = (int)(((int64_t)f->W * f->cfg.max_meas + 31) / 32);
    f->tm_bits = carve<unsigned>(p, (int64_t)T * f->tm_words);
    f->out_d_base = p;
    carve_out(p, T, &f->out_d);
    f->out_bytes = p - f->out_d_base;
    f->status_d = carve<ScanStatus>(p, 1);
    f->d_np = carve<int>(p, 8);   // d_np, d_nc, heavy_n, pad, pool_ctr (8 bytes)
    f->d_nc = f->d_np + 1;
    f->count = carve<int>(p, f->cap_par + 1);
    f->tile_sum = carve<int>(p, f->cap_par / kTile + 4);
    f->grid_ws = carve<char>(p, grid_workspace_bytes(f->cfg.max_meas));
    f->z_d = carve<double>(p, 2 * (int64_t)f->cfg.max_meas);
    f->used_d = carve<unsigned char>(p, f->cfg.max_meas);
    f->hist_d = carve<double>(p, kHistRec * (MHT_MAX_WINDOW + 4));
    f->histb_d = carve<double>(p, (int64_t)256 * kHistRec * (MHT_MAX_WINDOW + 4));
    f->dead_d = carve<int>(p, 2 * 256);
    f->adv_d = carve<double>(p, (int64_t)T * (kMaxAdv - 1) * kAdvRec);
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
    u->row_n = f->aw.row_n;
    u->model = f->cfg.model;
    u->score_upper = f->cfg.score_upper;
    u->cnllr_upper = f->cfg.cnllr_upper;
    u->radar_range = f->cfg.radar_range;
    u->px = f->cfg.position[0];
    u->py = f->cfg.position[1];
    u->dyn_window = f->dyn_window;
    u->target_size_limit = f->target_size_limit;
    u->window_roof = f->window_roof > 0 ? f->window_roof : f->cfg.n_scan_window;
    u->adv = f->adv_d;
}

// phase: 0 = whole scan; 1 = grow only (gate stage; the caller exchanges columns and calls the finish
// phase through mht_forest_select); 2 = finish only (selection already in aw.sel; ext = the 8 doubles
// mht_assoc_solve reports, or null).
static int forest_scan_impl(mht_forest *f, int64_t M, const double *d_z, mht_scan_info *info,
                            unsigned char *h_used, int phase = 0, const double *ext = nullptr) {
    if (M < 0 || M > f->cfg.max_meas) {
        set_error("mht_forest_scan: %lld measurements exceed max_meas=%d", (long long)M, f->cfg.max_meas);
        return MHT_E_CAPACITY;
    }
    cudaStream_t s = f->stream;
    const int k = phase == 2 ? f->scan : f->scan + 1;
    if (f->T == 0) {  // no trees yet: the scan only advances the clock (tracker.py:207 loops over nothing)
        if (phase != 2) f->scan = k;
        if (phase == 1) {
            f->open_scan = true;
            f->open_children = 0;
        }
        f->h_level_nodes = 0;
        f->last_tracks.clear();
        if (info) {
            memset(info, 0, sizeof(*info));
            info->certified = 1;   // nothing to associate
        }
        if (h_used && M) memset(h_used, 0, (size_t)M);
        return MHT_OK;
    }
    ScanArgs a;
    a.model = f->cfg.model;
    a.prev = f->lv[(k - 1) % f->nslots];
    a.cur = f->lv[k % f->nslots];
    a.ts = f->ts;
    a.rows_prev = f->rows[(k - 1) & 1];
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
    a.tile_tree = f->tile_tree;
    a.glist = f->glist;
    a.heavy_list = f->heavy_list;
    a.tm_bits = f->tm_bits;
    a.tm_words = f->tm_words;
    static const int heavy_rows = getenv("MHT_HEAVY_ROWS") ? atoi(getenv("MHT_HEAVY_ROWS")) : 6;
    static const int heavy_cand = getenv("MHT_HEAVY_CAND") ? atoi(getenv("MHT_HEAVY_CAND")) : 40;
    a.heavy_rows = heavy_rows;
    a.heavy_cand = heavy_cand;
    static const int emit_tma = getenv("MHT_EMIT_TMA") ? atoi(getenv("MHT_EMIT_TMA")) : 1;
    a.use_tma = emit_tma;
    a.heavy_n = f->d_np + 2;
    a.pool_ctr = (long long *)(f->d_np + 4);
    a.pt_gate = f->pt_gate;
    a.pt_P = f->pt_P;
    a.PT = f->PT;
    a.N = f->cfg.n_scan_window;
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
    if (phase != 2) MHT_CUDA(cudaEventRecord(f->ev[0], s));
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
    if (phase != 2) {
        // warm start: measurement rows keep their ids for W scans, so last scan's multipliers are a good
        // starting point; the plane being recycled for this scan starts from zero
        MHT_CUDA(cudaMemsetAsync(f->aw.u + (size_t)(k % f->W) * f->cfg.max_meas, 0, sizeof(double) * f->cfg.max_meas, s));
        if (int rc = assoc_begin(c, f->aw, grid_dim, s, k > 1)) return rc;
        MHT_CUDA(cudaMemsetAsync(f->used_d, 0, (size_t)(M ? M : 1), s));
        MHT_CUDA(cudaMemsetAsync(f->tm_bits, 0, sizeof(unsigned) * (size_t)f->T * f->tm_words, s));
        count_launch(), live_scan_kernel<<<1, 1024, 0, s>>>(a);
        count_launch(), pat_table_kernel<<<f->T, 128, 0, s>>>(a);
        if (int rc = launch_grid_build(d_z, (int)M, grid, cell_start, cell_fill, gz, gidx, s)) return rc;
        count_launch(), forest_gate_kernel<<<grid_dim, kTile, 0, s>>>(a);
        count_launch(), forest_gate_heavy_kernel<<<grid_dim, kTile, 0, s>>>(a, (int *)f->aw.rc, 2 * (long long)f->cap_nodes);
        count_launch(), forest_count_scan_kernel<<<grid_dim, kTile, 0, s>>>(a);
        count_launch(), forest_scan_tiles_kernel<<<1, 1024, 0, s>>>(a);
        if (f->W <= 8) count_launch(), forest_emit_kernel<8><<<grid_dim, kTile, emit_smem_bytes(f->W), s>>>(a, (int *)f->aw.rc);
        else count_launch(), forest_emit_kernel<MHT_MAX_WINDOW><<<grid_dim, kTile, emit_smem_bytes(f->W), s>>>(a, (int *)f->aw.rc);
        count_launch(), tree_off_kernel<<<(f->T + 256) / 256, 256, 0, s>>>(a);
    }
    MHT_CUDA(cudaGetLastError());
    if (phase != 2) MHT_CUDA(cudaEventRecord(f->ev[1], s));

    f->scan = k;
    static const long long sift_min = getenv("MHT_SIFT_MIN") ? atoll(getenv("MHT_SIFT_MIN")) : 1000000;
    const bool sift = f->h_level_nodes > sift_min;  // last scan's hypothesis count is the size hint
    if (phase == 1) {  // grow only: report the level's size, keep the scan open
        MHT_CUDA(cudaMemcpyAsync(f->status_h, f->status_d, sizeof(ScanStatus), cudaMemcpyDeviceToHost, s));
        if (h_used) MHT_CUDA(cudaMemcpyAsync(f->used_h, f->used_d, (size_t)M, cudaMemcpyDeviceToHost, s));
        MHT_CUDA(cudaStreamSynchronize(s));
        if (h_used && M) memcpy(h_used, f->used_h, (size_t)M);
        const ScanStatus &g = *f->status_h;
        if (g.overflow) {
            set_error("mht_forest_grow: capacity exceeded (%s: need %d, have %lld); the forest is unchanged "
                      "for this scan -- recreate it with larger max_nodes/max_parents",
                      g.overflow == 1 ? "live leaves" : "hypotheses", g.overflow == 1 ? g.n_parents : g.n_children,
                      (long long)(g.overflow == 1 ? f->cap_par : f->cap_nodes));
            f->scan = k - 1;
            return MHT_E_CAPACITY;
        }
        if (info) {
            memset(info, 0, sizeof(*info));
            info->n_parents = g.n_parents;
            info->n_children = g.n_children;
            info->n_pairs = (int64_t)g.n_children - g.n_parents;
            cudaEventElapsedTime(&info->ms_gate, f->ev[0], f->ev[1]);
        }
        f->open_scan = true;
        f->open_children = g.n_children;
        return MHT_OK;
    }
    if (phase == 0) {
        const double exact_ms = f->cfg.exact_ms == 0 ? 8.0 : (f->cfg.exact_ms < 0 ? 0.0 : (double)f->cfg.exact_ms);
        AssocEvents aev;
        aev.after_cluster = f->ev[5];
        for (int i = 0; i < 8; ++i) aev.dual[i] = f->evx[i];
        aev.exact_begin = f->evx[8];
        aev.exact_end = f->evx[9];
        if (int rc = assoc_solve(c, f->aw, f->cfg.max_dual_iters, 4096, kSMs * 8, s, &aev, f->scan > 1, sift, true,
                                 exact_ms))
            return rc;
        f->n_dual_ev = aev.n_dual;
    } else {
        MHT_CUDA(cudaEventRecord(f->ev[5], s));
    }
    MHT_CUDA(cudaEventRecord(f->ev[2], s));
    f->open_scan = false;

    UpdateArgs u;
    fill_update_args(f, &u);
    count_launch(), track_update_kernel<<<(f->T + 127) / 128, 128, 0, s>>>(u);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaEventRecord(f->ev[3], s));
    MHT_CUDA(cudaMemcpyAsync(f->out_h_base, f->out_d_base, (size_t)f->out_bytes, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(f->status_h, f->status_d, sizeof(ScanStatus), cudaMemcpyDeviceToHost, s));
    if (h_used) MHT_CUDA(cudaMemcpyAsync(f->used_h, f->used_d, (size_t)M, cudaMemcpyDeviceToHost, s));
    if (f->dyn_window)   // roots may have advanced several levels: the nodes they skipped
        MHT_CUDA(cudaMemcpyAsync(f->adv_h, f->adv_d, sizeof(double) * (size_t)f->T * (kMaxAdv - 1) * kAdvRec,
                                 cudaMemcpyDeviceToHost, s));
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
            for (int k = 0; k + 1 < f->out_h.advanced[t] && k < kMaxAdv - 1 && f->dyn_window; ++k) {
                const double *o = f->adv_h + ((size_t)t * (kMaxAdv - 1) + k) * kAdvRec;
                TrunkNode mid;
                mid.scan = (int)o[6];
                mid.meas = (int)o[0];
                mid.cnllr = o[1];
                memcpy(mid.x, o + 2, sizeof(mid.x));
                for (int q = 0; q < 16; ++q) mid.P[q] = (float)o[8 + q];
                f->trunk[t].push_back(mid);
            }
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
    {   // tracks that died this scan: keep their window records on the host
        std::vector<int> dead;
        for (int t : f->last_tracks)
            if (f->out_h.status[t] > 0 && f->h_last_pos[t] >= 0 && f->scan > f->h_root_scan[t]) dead.push_back(t);
        UpdateArgs ub;
        if (!dead.empty()) fill_update_args(f, &ub);
        for (size_t lo = 0; lo < dead.size(); lo += kDeadChunk) {
            const int n = (int)std::min<size_t>(kDeadChunk, dead.size() - lo);
            for (int i = 0; i < n; ++i) {
                f->dead_h[i] = dead[lo + i];
                f->dead_h[kDeadChunk + i] = f->h_last_pos[dead[lo + i]];
            }
            MHT_CUDA(cudaMemcpyAsync(f->dead_d, f->dead_h, sizeof(int) * 2 * kDeadChunk, cudaMemcpyHostToDevice, s));
            count_launch(), history_batch_kernel<<<(n + 63) / 64, 64, 0, s>>>(ub, f->dead_d, n, f->scan, f->histb_d);
            MHT_CUDA(cudaGetLastError());
            MHT_CUDA(cudaMemcpyAsync(f->histb_h, f->histb_d, sizeof(double) * (size_t)n * kHistStride,
                                     cudaMemcpyDeviceToHost, s));
            MHT_CUDA(cudaStreamSynchronize(s));
            for (int i = 0; i < n; ++i) {
                const double *o = f->histb_h + (size_t)i * kHistStride;
                const int wn = (int)o[0];
                f->dead_hist[dead[lo + i]].assign(o, o + kHistRec * (wn + 1));
            }
        }
    }
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
        info->open_components = st.assoc[5];
        info->repaired_trees = st.assoc[11];
        info->nnz_active = st.assoc[15];
        info->bb_iters = st.assoc[14];
        info->rows_active = st.rows_active;
        if (phase == 0) {
            for (int i = 0; i < f->n_dual_ev; ++i) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, f->evx[2 * i], f->evx[2 * i + 1]);
                info->ms_dual += ms;
            }
            cudaEventElapsedTime(&info->ms_exact, f->evx[8], f->evx[9]);
        }
        if (phase == 2) {  // the association was solved outside the forest (sharded trees)
            info->n_clusters = info->n_multi_clusters = info->n_active = 0;
            info->dual_iters = ext ? (int)ext[5] : 0;
            info->certified = ext ? (int)ext[6] : 0;
            info->n_candidates = ext ? (int64_t)ext[2] : 0;
            info->n_components = ext ? (int)ext[3] : 0;
            info->bb_nodes = ext ? (int64_t)ext[4] : 0;
            info->max_component = ext ? (int)ext[7] : 0;
            info->lower_bound = ext ? ext[0] : 0.0;
            info->objective = ext ? ext[1] : 0.0;
        }
    }
    return MHT_OK;
}

// this forest's columns (leaf hypotheses of the open scan) -> caller-owned global column arrays
__global__ void export_columns_kernel(Level cur, const int *rows_cur, long long stride_in, int W, const int *d_nc,
                                      const double *root_cnllr, int tree_offset, long long col_offset,
                                      long long stride_out, double *cost, int *tree, int *rows) {
    const int n = *d_nc;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int t = cur.tree[j];
        cost[col_offset + j] = cur.cnllr[j] - root_cnllr[t];
        tree[col_offset + j] = t + tree_offset;
        for (int w = 0; w < W; ++w) rows[w * stride_out + col_offset + j] = rows_cur[w * stride_in + j];
    }
}

// packed column records for the multi-GPU exchange: {f64 cost, i32 tree, i32 rows[W]} padded to a multiple of 8 bytes
__global__ void export_records_kernel(Level cur, const int *rows_cur, long long stride_in, int W, const int *d_nc,
                                      const double *root_cnllr, int tree_offset, int rec_bytes, unsigned char *out) {
    const int n = *d_nc;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int t = cur.tree[j];
        unsigned char *rec = out + (size_t)j * rec_bytes;
        *(double *)rec = cur.cnllr[j] - root_cnllr[t];
        int *ri = (int *)(rec + 8);
        ri[0] = t + tree_offset;
        for (int w = 0; w < W; ++w) ri[1 + w] = rows_cur[w * stride_in + j];
    }
}

// gathered records of every rank ([world][max_per_rank] records, rank r holds counts[r]) -> structure-of-arrays
// columns at the global offsets (rank order = single-forest column order)
__global__ void unpack_records_kernel(int world, const long long *counts, const long long *offsets, long long max_per_rank,
                                      int W, int rec_bytes, const unsigned char *in, long long stride_out, double *cost,
                                      int *tree, int *rows) {
    for (int r = 0; r < world; ++r) {
        const long long cnt = counts[r], off = offsets[r];
        const unsigned char *base = in + (size_t)r * max_per_rank * rec_bytes;
        for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < cnt; j += (long long)gridDim.x * blockDim.x) {
            const unsigned char *rec = base + (size_t)j * rec_bytes;
            cost[off + j] = *(const double *)rec;
            const int *ri = (const int *)(rec + 8);
            tree[off + j] = ri[0];
            for (int w = 0; w < W; ++w) rows[w * stride_out + off + j] = ri[1 + w];
        }
    }
}

}  // namespace mht

extern "C" int32_t mht_record_bytes(int32_t width) { return (12 + 4 * width + 7) / 8 * 8; }

extern "C" int mht_forest_export_records(mht_forest *f, int32_t tree_offset, void *d_records, int64_t cap_records) {
    if (!f || !d_records || cap_records < f->h_level_nodes_open()) {
        set_error("mht_forest_export_records: invalid argument (need room for %lld records)",
                  f ? (long long)f->h_level_nodes_open() : 0ll);
        return MHT_E_INVALID;
    }
    if (f->T == 0) return MHT_OK;
    count_launch(), export_records_kernel<<<kSMs * 8, 256, 0, f->stream>>>(
        f->lv[f->scan % f->nslots], f->rows[f->scan & 1], f->cap_nodes, f->W, f->d_nc, f->ts.root_cnllr, tree_offset,
        mht_record_bytes(f->W), (unsigned char *)d_records);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaStreamSynchronize(f->stream));
    return MHT_OK;
}

extern "C" int mht_unpack_records(int32_t world, const int64_t *h_counts, int64_t max_per_rank, int32_t width,
                                  const void *d_gathered, int64_t stride_out, double *d_cost, int32_t *d_tree,
                                  int32_t *d_rows, void *d_scratch, void *stream) {
    if (int rc = check_device()) return rc;
    if (world < 1 || world > 64 || !h_counts || !d_gathered || !d_cost || !d_tree || !d_rows || !d_scratch) {
        set_error("mht_unpack_records: invalid argument");
        return MHT_E_INVALID;
    }
    long long host[128];
    long long acc = 0;
    for (int r = 0; r < world; ++r) {
        host[r] = h_counts[r];
        host[64 + r] = acc;
        acc += h_counts[r];
    }
    cudaStream_t s = (cudaStream_t)stream;
    MHT_CUDA(cudaMemcpyAsync(d_scratch, host, sizeof(host), cudaMemcpyHostToDevice, s));
    count_launch(), unpack_records_kernel<<<kSMs * 8, 256, 0, s>>>(world, (const long long *)d_scratch,
                                                                   (const long long *)d_scratch + 64, max_per_rank, width,
                                                                   mht_record_bytes(width), (const unsigned char *)d_gathered,
                                                                   stride_out, d_cost, d_tree, d_rows);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaStreamSynchronize(s));   // `host` is a stack buffer
    return MHT_OK;
}

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
    f->PT = 1 << (cfg->n_scan_window + 1);
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
    if (e == cudaSuccess) e = cudaMallocHost(&f->histb_h, sizeof(double) * 256 * kHistRec * (MHT_MAX_WINDOW + 4));
    if (e == cudaSuccess) e = cudaMallocHost(&f->dead_h, sizeof(int) * 2 * 256);
    if (e == cudaSuccess) e = cudaMallocHost(&f->adv_h, sizeof(double) * (size_t)cfg->max_trees * (kMaxAdv - 1) * kAdvRec);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&f->ev[i]);
    for (int i = 0; i < 10 && e == cudaSuccess; ++i) e = cudaEventCreate(&f->evx[i]);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(forest_emit_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)emit_smem_bytes(8));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(forest_emit_kernel<MHT_MAX_WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)emit_smem_bytes(MHT_MAX_WINDOW));
    // the arena starts from zeros whatever a previous owner of the memory left in it: a freshly mapped allocation is zero
    // filled by the driver, a recycled one is not, and nothing here may depend on which of the two it got
    if (e == cudaSuccess) e = cudaMemsetAsync(f->arena, 0, (size_t)f->bytes, f->stream);
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
    f->dead_hist.resize(T);
    f->h_level_nodes = 0;
    *out = f;
    return MHT_OK;
}

extern "C" void mht_forest_destroy(mht_forest *f) {
    if (!f) return;
    cudaStreamSynchronize(f->stream);
    for (int i = 0; i < 6; ++i) cudaEventDestroy(f->ev[i]);
    for (int i = 0; i < 10; ++i) cudaEventDestroy(f->evx[i]);
    cudaStreamDestroy(f->stream);
    cudaFreeHost(f->out_h_base);
    cudaFreeHost(f->status_h);
    cudaFreeHost(f->z_h);
    cudaFreeHost(f->used_h);
    cudaFreeHost(f->hist_h);
    cudaFreeHost(f->histb_h);
    cudaFreeHost(f->dead_h);
    cudaFreeHost(f->adv_h);
    cudaFree(f->arena);
    delete f;
}

extern "C" int64_t mht_forest_bytes(const mht_forest *f) { return f ? f->bytes : 0; }

extern "C" int mht_forest_initiate(mht_forest *f, const double x0[4], const float P0[16], double Pd, int32_t *slot) {
    if (!f || !x0 || !P0 || !(Pd > 0.0 && Pd < 1.0)) {
        set_error("mht_forest_initiate: invalid argument");
        return MHT_E_INVALID;
    }
    if ((f->T >= f->cfg.max_trees && f->free_slots.empty()) || f->h_level_nodes >= f->cap_nodes) {
        set_error("mht_forest_initiate: capacity exceeded (trees %d/%d, no released slot)", f->T, f->cfg.max_trees);
        return MHT_E_CAPACITY;
    }
    cudaStream_t s = f->stream;
    // a released slot of a dead track first (lowest index: deterministic), a fresh one otherwise
    int t = f->T;
    if (!f->free_slots.empty()) {
        auto it = std::min_element(f->free_slots.begin(), f->free_slots.end());
        t = *it;
        f->free_slots.erase(it);
    }
    const Level &L = f->lv[f->scan % f->nslots];
    const int pos = (int)f->h_level_nodes, pi = 0;
    const unsigned short pat0 = 0;
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
    MHT_CUDA(cudaMemcpyAsync(L.pat + pos, &pat0, 2, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(f->ts.rootP + 16 * (size_t)t, P0, 64, cudaMemcpyHostToDevice, s));
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
    f->dead_hist[t].clear();
    if (t == f->T) f->T = t + 1;
    if (slot) *slot = t;
    return MHT_OK;
}

extern "C" int mht_forest_release(mht_forest *f, int32_t slot) {
    if (!f || slot < 0 || slot >= f->T || f->h_alive[slot]) {
        set_error("mht_forest_release: slot %d is not a dead track's slot", slot);
        return MHT_E_INVALID;
    }
    if (std::find(f->free_slots.begin(), f->free_slots.end(), slot) != f->free_slots.end()) return MHT_OK;
    f->trunk[slot].clear();
    f->trunk[slot].shrink_to_fit();
    f->dead_hist[slot].clear();
    f->dead_hist[slot].shrink_to_fit();
    f->free_slots.push_back(slot);
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

extern "C" int mht_forest_grow(mht_forest *f, int64_t M, const double *z, int32_t z_on_device, double scan_time,
                               mht_scan_info *info, uint8_t *h_meas_used) {
    (void)scan_time;
    if (!f || (M > 0 && !z) || f->open_scan) {
        set_error("mht_forest_grow: invalid argument or a grown scan is still waiting for mht_forest_select");
        return MHT_E_INVALID;
    }
    if (M < 0 || M > f->cfg.max_meas) {
        set_error("mht_forest_grow: %lld measurements exceed max_meas=%d", (long long)M, f->cfg.max_meas);
        return MHT_E_CAPACITY;
    }
    const double *d_z = z;
    if (!z_on_device) {
        if (M) memcpy(f->z_h, z, 16 * (size_t)M);
        MHT_CUDA(cudaMemcpyAsync(f->z_d, f->z_h, 16 * (size_t)M, cudaMemcpyHostToDevice, f->stream));
        d_z = f->z_d;
    }
    f->open_M = M;
    return forest_scan_impl(f, M, d_z, info, h_meas_used, 1);
}

extern "C" int mht_forest_export_columns(mht_forest *f, int32_t tree_offset, int64_t col_offset, int64_t stride,
                                         double *d_cost, int32_t *d_tree, int32_t *d_rows) {
    if (!f || !d_cost || !d_tree || !d_rows || col_offset < 0 || stride < col_offset + f->h_level_nodes_open()) {
        set_error("mht_forest_export_columns: invalid argument");
        return MHT_E_INVALID;
    }
    if (f->T == 0) return MHT_OK;
    count_launch(), export_columns_kernel<<<kSMs * 8, 256, 0, f->stream>>>(f->lv[f->scan % f->nslots], f->rows[f->scan & 1], f->cap_nodes,
                                                         f->W, f->d_nc, f->ts.root_cnllr, tree_offset, col_offset, stride,
                                                         d_cost, d_tree, d_rows);
    MHT_CUDA(cudaGetLastError());
    MHT_CUDA(cudaStreamSynchronize(f->stream));
    return MHT_OK;
}

extern "C" int mht_forest_select(mht_forest *f, const int32_t *d_selected_col, const double *h_assoc_info,
                                 mht_scan_info *info) {
    if (!f || !f->open_scan || (f->T > 0 && !d_selected_col)) {
        set_error("mht_forest_select: no grown scan is open (call mht_forest_grow first)");
        return MHT_E_INVALID;
    }
    if (f->T > 0)
        MHT_CUDA(cudaMemcpyAsync(f->aw.sel, d_selected_col, sizeof(int) * (size_t)f->T, cudaMemcpyDeviceToDevice,
                                 f->stream));
    const int rc = forest_scan_impl(f, f->open_M, f->z_d, info, nullptr, 2, h_assoc_info);
    f->open_scan = false;
    return rc;
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
        count_launch(), min_leaf_distance_kernel<<<kSMs * 2, 256, 0, f->stream>>>(f->lv[f->scan % f->nslots], f->ts, f->T, px, py, d);
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

extern "C" int mht_forest_set_dynamic_window(mht_forest *f, int32_t enabled, int32_t target_size_limit,
                                             int32_t window_roof) {
    if (!f || target_size_limit < 1 || window_roof < 0 || window_roof > f->cfg.n_scan_window) {
        set_error("mht_forest_set_dynamic_window: invalid argument");
        return MHT_E_INVALID;
    }
    f->dyn_window = enabled ? 1 : 0;
    f->target_size_limit = target_size_limit;
    f->window_roof = window_roof;
    return MHT_OK;
}

extern "C" int mht_forest_windows(mht_forest *f, int32_t cap, int32_t *n, int32_t *h_slot, int32_t *h_window) {
    if (!f || !n || !h_slot || !h_window) {
        set_error("mht_forest_windows: invalid argument");
        return MHT_E_INVALID;
    }
    std::vector<int> win(f->T > 0 ? f->T : 1);
    if (f->T > 0) {
        MHT_CUDA(cudaMemcpyAsync(win.data(), f->ts.window, sizeof(int) * (size_t)f->T, cudaMemcpyDeviceToHost, f->stream));
        MHT_CUDA(cudaStreamSynchronize(f->stream));
    }
    int k = 0;
    for (int t = 0; t < f->T; ++t) {
        if (!f->h_alive[t]) continue;
        if (k < cap) {
            h_slot[k] = t;
            h_window[k] = win[t];
        }
        ++k;
    }
    *n = k;
    if (k > cap) {
        set_error("mht_forest_windows: %d live trees exceed cap %d", k, cap);
        return MHT_E_CAPACITY;
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
    const bool cached = !f->h_alive[slot] && !f->dead_hist[slot].empty();
    if (cached) {
        memcpy(f->hist_h, f->dead_hist[slot].data(), sizeof(double) * f->dead_hist[slot].size());
        wn = (int)f->hist_h[0];
    } else if (f->h_last_pos[slot] >= 0 && f->scan > f->h_root_scan[slot]) {
        UpdateArgs u;
        fill_update_args(f, &u);
        count_launch(), history_kernel<<<1, 1, 0, f->stream>>>(u, slot, f->h_last_pos[slot], f->scan, f->hist_d);
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

// mht_forest_history for SEVERAL slots in one call (the tracks that died in a scan: Tracker reads their histories before it
// hands the slots back): row i of the [n][cap_len] outputs belongs to h_slots[i] and holds h_len[i] nodes, oldest first.
extern "C" int mht_forest_histories_of(mht_forest *f, int32_t n, const int32_t *h_slots, int32_t cap_len, int32_t *h_len,
                                       int32_t *h_meas, double *h_x, double *h_cnllr, float *h_P) {
    if (!f || n < 0 || (n && (!h_slots || !h_len)) || cap_len < 1) {
        set_error("mht_forest_histories_of: invalid argument");
        return MHT_E_INVALID;
    }
    int need = 0;
    for (int i = 0; i < n; ++i) {
        int k = 0;
        const int rc = mht_forest_history(f, h_slots[i], cap_len, &k, h_meas ? h_meas + (size_t)i * cap_len : nullptr,
                                          h_x ? h_x + (size_t)4 * i * cap_len : nullptr,
                                          h_cnllr ? h_cnllr + (size_t)i * cap_len : nullptr,
                                          h_P ? h_P + (size_t)16 * i * cap_len : nullptr);
        h_len[i] = k;
        if (rc == MHT_E_CAPACITY) {
            need = k > need ? k : need;
            continue;
        }
        if (rc != MHT_OK) return rc;
    }
    if (need) {
        h_len[0] = need;
        set_error("mht_forest_histories_of: a history of %d nodes exceeds cap_len %d", need, cap_len);
        return MHT_E_CAPACITY;
    }
    return MHT_OK;
}

// measurement rows on the root->leaf paths of one tree's live leaves -> bitmap (one bit per row)
__global__ void measurement_set_kernel(const int *rows, long long stride, int W, int lo, int hi, unsigned *bits) {
    const long long n = (long long)(hi - lo) * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i / (hi - lo)), leaf = lo + (int)(i % (hi - lo));
        const int r = rows[(long long)w * stride + leaf];
        if (r >= 0) atomicOr(&bits[r >> 5], 1u << (r & 31));
    }
}

// Tracker.__associatedMeasurements__[i] (tracker.py:83,331-332,1226-1227): the (scanNumber, measurementNumber) pairs
// of every node below the tree's root, i.e. Target.getMeasurementSet of the root (pyTarget.py:414-430), after the
// last scan's pruning.  MHT_E_CAPACITY: *n holds the required count.
extern "C" int mht_forest_measurement_set(mht_forest *f, int32_t slot, int32_t cap, int32_t *n, int32_t *h_scan,
                                          int32_t *h_meas) {
    if (!f || !n || slot < 0 || slot >= f->T || f->open_scan) {
        set_error("mht_forest_measurement_set: invalid argument (or a grown scan is still open)");
        return MHT_E_INVALID;
    }
    *n = 0;
    if (!f->h_alive[slot] || f->scan == 0) return MHT_OK;
    cudaStream_t s = f->stream;
    int range[2];
    MHT_CUDA(cudaMemcpyAsync(&range[0], f->ts.live_lo + slot, 4, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(&range[1], f->ts.live_hi + slot, 4, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    // the (tree, row) bitmap of the gate stage is rebuilt at the start of every scan: its first row is free here
    MHT_CUDA(cudaMemsetAsync(f->tm_bits, 0, sizeof(unsigned) * (size_t)f->tm_words, s));
    if (range[1] > range[0]) {
        count_launch(), measurement_set_kernel<<<kSMs, 256, 0, s>>>(f->rows[f->scan & 1], f->cap_nodes, f->W, range[0], range[1],
                                                    f->tm_bits);
        MHT_CUDA(cudaGetLastError());
    }
    std::vector<unsigned> bits(f->tm_words);
    MHT_CUDA(cudaMemcpyAsync(bits.data(), f->tm_bits, sizeof(unsigned) * (size_t)f->tm_words, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    int k = 0;
    for (int wd = 0; wd < f->tm_words; ++wd) {
        unsigned m = bits[wd];
        while (m) {
            const int b = __builtin_ctz(m);
            m &= m - 1;
            const int r = wd * 32 + b, plane = r / f->cfg.max_meas, idx = r % f->cfg.max_meas;
            int sc = f->scan - ((f->scan - plane) % f->W + f->W) % f->W;      // latest scan <= now stored in this plane
            if (sc <= f->h_root_scan[slot]) continue;                          // at or above the (new) root
            if (k < cap && h_scan && h_meas) {
                h_scan[k] = sc;
                h_meas[k] = idx + 1;
            }
            ++k;
        }
    }
    *n = k;
    if (k > cap) {
        set_error("mht_forest_measurement_set: %d pairs exceed cap %d", k, cap);
        return MHT_E_CAPACITY;
    }
    return MHT_OK;
}

// mht_forest_history for EVERY live track in one go: the window walks of all tracks share kDeadChunk-wide launches
// (4 launches for 1000 tracks instead of 1000 single-thread launches with a stream sync each).
extern "C" int mht_forest_histories(mht_forest *f, int32_t cap_tracks, int32_t cap_len, int32_t *n_tracks,
                                    int32_t *h_slot, int32_t *h_len, int32_t *h_meas, double *h_x, double *h_cnllr,
                                    float *h_P) {
    if (!f || !n_tracks || !h_slot || !h_len || cap_len < 1) {
        set_error("mht_forest_histories: invalid argument");
        return MHT_E_INVALID;
    }
    std::vector<int> live;
    for (int t = 0; t < f->T; ++t)
        if (f->h_alive[t]) live.push_back(t);
    *n_tracks = (int)live.size();
    if ((int)live.size() > cap_tracks) {
        set_error("mht_forest_histories: %d live tracks exceed cap %d", (int)live.size(), cap_tracks);
        return MHT_E_CAPACITY;
    }
    cudaStream_t s = f->stream;
    UpdateArgs ub;
    fill_update_args(f, &ub);
    int need = 0;
    for (size_t lo = 0; lo < live.size(); lo += kDeadChunk) {
        const int n = (int)std::min<size_t>(kDeadChunk, live.size() - lo);
        int n_walk = 0;
        for (int i = 0; i < n; ++i) {
            const int t = live[lo + i];
            f->dead_h[i] = t;
            f->dead_h[kDeadChunk + i] = f->h_last_pos[t];
            if (f->h_last_pos[t] >= 0 && f->scan > f->h_root_scan[t]) ++n_walk;
        }
        if (n_walk) {
            MHT_CUDA(cudaMemcpyAsync(f->dead_d, f->dead_h, sizeof(int) * 2 * kDeadChunk, cudaMemcpyHostToDevice, s));
            count_launch(), history_batch_kernel<<<(n + 63) / 64, 64, 0, s>>>(ub, f->dead_d, n, f->scan, f->histb_d);
            MHT_CUDA(cudaGetLastError());
            MHT_CUDA(cudaMemcpyAsync(f->histb_h, f->histb_d, sizeof(double) * (size_t)n * kHistStride,
                                     cudaMemcpyDeviceToHost, s));
            MHT_CUDA(cudaStreamSynchronize(s));
        }
        for (int i = 0; i < n; ++i) {
            const int t = live[lo + i];
            const size_t row = lo + i;
            const std::vector<TrunkNode> &tr = f->trunk[t];
            const bool walk = f->h_last_pos[t] >= 0 && f->scan > f->h_root_scan[t];
            const double *o0 = f->histb_h + (size_t)i * kHistStride;
            const int wn = walk ? (int)o0[0] : 0;
            const int total = (int)tr.size() + wn;
            h_slot[row] = t;
            h_len[row] = total;
            need = std::max(need, total);
            if (total > cap_len) continue;
            int k = 0;
            for (const TrunkNode &nd : tr) {
                const size_t e = row * cap_len + k;
                if (h_meas) h_meas[e] = nd.meas;
                if (h_x) memcpy(h_x + 4 * e, nd.x, 32);
                if (h_cnllr) h_cnllr[e] = nd.cnllr;
                if (h_P) memcpy(h_P + 16 * e, nd.P, 64);
                ++k;
            }
            for (int q = wn - 1; q >= 0; --q, ++k) {   // the walk is leaf -> root; report oldest first
                const double *o = o0 + kHistRec + kHistRec * q;
                const size_t e = row * cap_len + k;
                if (h_meas) h_meas[e] = (int)o[0];
                if (h_cnllr) h_cnllr[e] = o[1];
                if (h_x) memcpy(h_x + 4 * e, o + 2, 32);
                if (h_P)
                    for (int z = 0; z < 16; ++z) h_P[16 * e + z] = (float)o[8 + z];
            }
        }
    }
    if (need > cap_len) {
        set_error("mht_forest_histories: the longest history has %d nodes, cap_len is %d", need, cap_len);
        *n_tracks = need;      // the caller retries with this length
        return MHT_E_CAPACITY;
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
