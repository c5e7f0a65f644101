// Association stage on device.
//
//   mht_cluster      = Tracker._findClustersFromSets            (reference pymht/tracker.py:961-974)
//   mht_assoc_solve  = the cluster loop tracker.py:228-236: Target._selectBestHypothesis
//                      (pymht/pyTarget.py:446-459) + _solveOptimumAssociation/_solveBLP_OR_TOOLS
//                      (tracker.py:979-1217), every cluster at once.
//
// 0/1 program (tracker.py:1155-1217):  min sum_j c_j tau_j,  one column per tree (A2 tau = 1),
// every measurement row used at most once (A1 tau <= 1).  Columns are leaves; a column's rows are
// the <= N+1 measurements on its root->leaf path, so A1 is never materialised (the reference builds
// it dense: 680 MB at 86k x 7.9k).
//
// Algorithm (LP-relaxation primal-dual + integer repair):
//   1. union-find over (tree,row) incidences -> clusters (independent sub-problems).
//   2. Lagrangian dual of the row constraints, L(u) = -sum u_r + sum_t min_j (c_j + sum_{r in j} u_r):
//      projected subgradient ascent with a Polyak step PER CLUSTER; one streaming pass over the
//      columns per iteration (reduced cost + per-tree argmin).  A cluster whose argmins are
//      conflict-free and complementary is solved exactly (this covers every singleton cluster in
//      the first iteration = _selectBestHypothesis, ties -> last leaf like the reference's '<=').
//   3. primal: parallel greedy on reduced costs (bids on rows, lowest reduced cost wins).
//   4. reduced-cost fixing: only columns with rc_j - min_t <= UB_c - L_c can be optimal; the
//      survivors are split into connected components and each is searched exactly, depth first,
//      with the Lagrangian bound.  A finished search certifies optimality.
// Sums that steer the iteration are accumulated in 2^-34 fixed point so the result does not depend
// on atomic ordering.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "assoc.cuh"

namespace cg = cooperative_groups;

namespace mht {

constexpr double kFix = 17179869184.0;  // 2^34
constexpr int kPatience = 10;
constexpr double kShrink = 0.7;
constexpr int kMaxCandPerTree = 2048;   // per-tree candidate list limit (insertion sort, one thread)
constexpr int kDomMax = 64;              // sorting / dominance reduction / enumeration only for lists up to this length
constexpr int kGreedyRounds = 40;
constexpr int kGreedyEvery = 60;   // profiles/sweep_r2.txt: 60 is faster than 40 AND leaves a smaller gap
constexpr int kStallStop = 60;
// an improvement of the bound counts as progress when it exceeds this fraction of |L|.  (1e-6 made the loop
// leave a 285-tree cluster 0.015 short of its LP optimum -- |L| = 4.5e3 -- and the greedy primal 0.8 above it;
// run to 1e-4 of the optimum, the same greedy finds the optimal selection: scripts/search_proto.py)
constexpr double kBigGain = 1e-9;
constexpr int kSiftRounds = 2;   // MHT_SIFT_ROUNDS=3: 25 % slower scans for a ~0.6-point smaller bound gap (profiles/README.md)
constexpr unsigned long long kKeyInf = ~0ull;

__device__ __forceinline__ long long to_fix(double v) { return __double2ll_rn(v * kFix); }
__device__ __forceinline__ double from_fix(long long v) { return (double)v / kFix; }

__device__ __forceinline__ double col_cost(const ColView &c, int j, int t) {
    return c.tree_base ? c.cost[j] - c.tree_base[t] : c.cost[j];
}

// ------------------------------------------------------------------------------------------------
// union-find (link the larger root under the smaller: label = smallest tree index)
// ------------------------------------------------------------------------------------------------
__global__ void assoc_init_kernel(ColView c, AssocWork w, int *tstart, int *tend, int warm) {
    const int T = c.n_trees, R = c.n_rows;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(max(T + 1, R), kAssocInfo); i += gridDim.x * blockDim.x) {
        if (i < T) {
            tstart[i] = -1;
            tend[i] = -1;
        }
        if (i < T) {
            w.uf[i] = i;
            w.tmin[i] = kKeyInf;
            w.targ[i] = -1;
            w.sel[i] = -1;
            w.cl_best[i] = -1e300;
            w.cl_ub[i] = 1e300;
            w.cl_theta[i] = 1.0;
            w.cl_step[i] = 0.0;
            w.cl_stall[i] = 0;
            w.cl_done[i] = 0;
            w.cl_flag[i] = 0;
            w.tdone[i] = 0;
            w.cl_m[i] = 0;
            w.cl_u[i] = 0;
            w.cl_cost[i] = 0;
            w.cl_nrm[i] = 0;
        }
        if (i <= T) w.cand_cnt[i] = 0;
        if (i < R) {
            w.row_owner[i] = -1;
            w.row_mark[i] = 0;         // becomes 1 when two different trees touch the row
            if (!warm) w.u[i] = 0.0;   // warm start: keep last scan's multipliers (rows persist W scans)
            w.best_u[i] = warm ? w.u[i] : 0.0;
            w.usage[i] = 0;
            w.bbw.row_local[i] = -1;
        }
        if (i < 4) {
            w.stall_ctr[i] = 0;
            w.row_n[i] = 0;
        }
        if (i < kAssocInfo) w.info[i] = 0;
        if (i == 0) {
            *w.bb_nodes = 0ull;
            w.objective[0] = w.objective[1] = 0.0;
        }
    }
}

// per-column part of the reset: tree column ranges, argmin frequencies
__global__ void assoc_init_cols_kernel(ColView c, AssocWork w, int *tstart, int *tend) {
    const int n = *c.n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        w.freq[i] = 0;
        const int t = c.tree[i];
        if (i == 0 || c.tree[i - 1] != t) tstart[t] = i;
        if (i == n - 1 || c.tree[i + 1] != t) tend[t] = i + 1;
    }
}

__global__ void uf_union_cols_kernel(ColView c, int *uf, int *row_owner, int *row_multi) {
    const int n = *c.n_ptr;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int t = c.tree[j];
        // forest columns: siblings share every plane but the newest, so only the first sibling (the
        // miss child) handles the inherited planes and everybody handles the newest one
        const bool inherited = !c.meas || c.meas[j] == 0;
        for (int k = 0; k < c.width; ++k) {
            if (!inherited && k != c.plane_new) continue;
            const int r = c.rows[(long long)k * c.stride + j];
            if (r < 0) continue;
            // generic columns: the left neighbour (same tree, same row) already did this row
            if (!c.meas && (threadIdx.x & 31) && c.rows[(long long)k * c.stride + j - 1] == r && c.tree[j - 1] == t)
                continue;
            uf_touch_row(uf, row_owner, row_multi, r, t);
        }
    }
}

// warm start hygiene: a row only one tree can use needs no multiplier
__global__ void warm_fix_kernel(int R, const int *row_multi, double *u, double *best_u) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x)
        if (!row_multi[r]) u[r] = best_u[r] = 0.0;
}

__global__ void row_list_kernel(int R, const int *row_owner, int *row_list, int *row_n) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x)
        if (row_owner[r] >= 0) row_list[atomicAdd(row_n, 1)] = r;
}

__global__ void uf_flatten_kernel(int T, int *uf) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) uf[t] = uf_find(uf, t);
}

// single CTA: cluster statistics (n clusters, clusters with > 1 tree); cl_nrm used as size scratch
__global__ void __launch_bounds__(1024, 1)
cluster_stats_kernel(int T, const int *uf, const int *tstart, int *cl_size, int *info) {
    __shared__ int ncl, nmulti;
    if (threadIdx.x == 0) ncl = nmulti = 0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) cl_size[t] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x)
        if (tstart[t] >= 0) atomicAdd(&cl_size[uf[t]], 1);
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x)
        if (cl_size[t] > 0) {
            atomicAdd(&ncl, 1);
            if (cl_size[t] > 1) atomicAdd(&nmulti, 1);
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        info[7] = ncl;
        info[8] = nmulti;
    }
}

// ------------------------------------------------------------------------------------------------
// dual ascent: column passes
// ------------------------------------------------------------------------------------------------
// min over runs of equal tree id inside a warp (columns are sorted by tree); returns true on the
// first lane of each run, which then holds the run minimum.
__device__ __forceinline__ bool warp_run_min(int t, unsigned long long &key) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ok = __shfl_down_sync(0xffffffffu, key, o);
        const int ot = __shfl_down_sync(0xffffffffu, t, o);
        if (lane + o < 32 && ot == t && ok < key) key = ok;
    }
    const int pt = __shfl_up_sync(0xffffffffu, t, 1);
    return lane == 0 || pt != t;
}

template <bool FORCE>
__device__ __forceinline__ void dual_rc_body(ColView c, AssocWork w) {
    if (!FORCE && w.info[0]) return;
    const int n = *c.n_ptr;
    const int nround = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
        int t = -1;
        unsigned long long key = kKeyInf;
        if (i < n) {
            const int j = c.idx ? c.idx[i] : i;
            t = c.tree[j];
            if (FORCE || !w.tdone[t]) {
                double v = col_cost(c, j, t);
                for (int k = 0; k < c.width; ++k) {
                    const int r = c.rows[(long long)k * c.stride + j];
                    if (r >= 0) v += w.u[r];
                }
                w.rc[j] = v;
                key = f64_key(v);
            } else {
                t = -1;
            }
        }
        const bool head = warp_run_min(t, key);
        if (head && t >= 0) atomicMin(&w.tmin[t], key);
    }
}
template <bool FORCE>
__global__ void __launch_bounds__(256) dual_rc_kernel(ColView c, AssocWork w) { dual_rc_body<FORCE>(c, w); }


template <bool FORCE>
__device__ __forceinline__ void dual_arg_body(ColView c, AssocWork w) {
    if (!FORCE && w.info[0]) return;
    const int n = *c.n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = c.idx ? c.idx[i] : i;
        const int t = c.tree[j];
        if (!FORCE && w.tdone[t]) continue;
        if (f64_key(w.rc[j]) == w.tmin[t]) atomicMax(&w.targ[t], j);
    }
}
template <bool FORCE>
__global__ void __launch_bounds__(256) dual_arg_kernel(ColView c, AssocWork w) { dual_arg_body<FORCE>(c, w); }


// subgradient, per-cluster Polyak step and bookkeeping, as four small grid kernels
// (accumulators cl_m/cl_u/cl_cost/cl_nrm are zero on entry: assoc_init / du_apply leave them so)
__device__ __forceinline__ bool du_skip(const ColView &c, const AssocWork &w) {
    return w.info[0] || (c.idx && w.act_n[2]);
}

// warp-aggregated accumulation into per-cluster sums: when every active lane targets the same cluster
// (the normal case: one giant cluster, or consecutive trees/rows of one cluster) the warp reduces first
// and issues ONE atomic; otherwise lanes fall back to individual atomics.  Integer adds: order-free.
__device__ __forceinline__ void warp_add_ll(long long *base, int cl, long long v, bool active) {
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!mask) return;
    const int leader = __ffs(mask) - 1;
    const int cl0 = __shfl_sync(0xffffffffu, cl, leader);
    if (__all_sync(0xffffffffu, !active || cl == cl0)) {
        long long x = active ? v : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == leader) atomicAdd((unsigned long long *)&base[cl0], (unsigned long long)x);
    } else if (active) {
        atomicAdd((unsigned long long *)&base[cl], (unsigned long long)v);
    }
}
__device__ __forceinline__ void warp_add_i(int *base, int cl, int v, bool active) {
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!mask) return;
    const int leader = __ffs(mask) - 1;
    const int cl0 = __shfl_sync(0xffffffffu, cl, leader);
    if (__all_sync(0xffffffffu, !active || cl == cl0)) {
        int x = active ? v : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == leader && x) atomicAdd(&base[cl0], x);
    } else if (active && v) {
        atomicAdd(&base[cl], v);
    }
}

__device__ __forceinline__ void du_trees_body(ColView c, AssocWork w) {
    if (du_skip(c, w)) return;
    const int T = c.n_trees;
    const int Tround = (T + 31) & ~31;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Tround; t += gridDim.x * blockDim.x) {
        int cl = 0;
        bool active = t < T && w.tstart[t] >= 0;
        if (active) {
            cl = w.uf[t];
            active = !w.cl_done[cl];
        }
        long long m = 0, cost = 0;
        if (active) {
            const int j = w.targ[t];
            w.freq[j] += 1;  // ergodic primal estimate: how often this column is the tree's Lagrangian choice
            m = to_fix(key_f64(w.tmin[t]));
            cost = to_fix(col_cost(c, j, t));
            for (int k = 0; k < c.width; ++k) {
                const int r = c.rows[(long long)k * c.stride + j];
                if (r >= 0) atomicAdd(&w.usage[r], 1);
            }
        }
        warp_add_ll(w.cl_m, cl, m, active);
        warp_add_ll(w.cl_cost, cl, cost, active);
    }
}

__global__ void du_trees_kernel(ColView c, AssocWork w) { du_trees_body(c, w); }

__device__ __forceinline__ void du_rows_body(ColView c, AssocWork w) {
    if (du_skip(c, w)) return;
    const int nr = *w.row_n;
    const int nround = (nr + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
        int cl = 0, g = 0;
        long long uf = 0;
        bool active = i < nr;
        if (active) {
            const int r = w.row_list[i];
            cl = w.uf[w.row_owner[r]];
            active = !w.cl_done[cl];
            if (active) {
                g = w.usage[r] - 1;
                const double ur = w.u[r];
                if (ur <= 0.0 && g < 0) g = 0;
                w.usage[r] = g;
                if (ur > 0.0) uf = to_fix(ur);
            }
        }
        warp_add_i(w.cl_nrm, cl, g * g, active);
        warp_add_ll(w.cl_u, cl, uf, active);
    }
}

__global__ void du_rows_kernel(ColView c, AssocWork w) { du_rows_body(c, w); }

// the decision of ONE cluster (t = its root, a tree that has columns)
__device__ __forceinline__ void du_decide_one(const AssocWork &w, int t) {
    w.cl_flag[t] = 0;   // flags live from here to the end of the apply phase (only roots ever carry one)
    if (w.cl_done[t]) return;
    const double L = from_fix(w.cl_m[t] - w.cl_u[t]);
    const int nrm = w.cl_nrm[t];
    if (nrm == 0) {  // conflict-free and complementary: argmins are optimal
        w.cl_done[t] = 1;
        w.cl_best[t] = L;
        w.cl_ub[t] = from_fix(w.cl_cost[t]);
        w.cl_flag[t] = 3;
        w.cl_step[t] = 0.0;
        atomicOr(&w.stall_ctr[2], 1);
        return;
    }
    if (L > w.cl_best[t] + 1e-12) {
        // flag 1 = keep these multipliers; flag 8 = the gain is large enough to keep iterating
        const bool big = L > w.cl_best[t] + kBigGain * fmax(1.0, fabs(L));
        w.cl_flag[t] = big ? 9 : 1;
        if (big) atomicOr(&w.stall_ctr[2], 1);
        w.cl_best[t] = L;
        w.cl_stall[t] = 0;
    } else if (++w.cl_stall[t] >= kPatience) {
        w.cl_theta[t] *= kShrink;
        w.cl_stall[t] = 0;
    }
    if (w.cl_ub[t] - w.cl_best[t] < 1e-9) {
        w.cl_done[t] = 1;
        w.cl_step[t] = 0.0;
    } else {
        // no step before the first primal solution provides an upper bound
        w.cl_step[t] = w.cl_ub[t] < 1e299 ? w.cl_theta[t] * (w.cl_ub[t] - L) / (double)nrm : 0.0;
        atomicAdd(&w.stall_ctr[1], 1);
    }
}

// du_decide_one with every load issued before the first use (one dependent hop instead of a chain of five); same arithmetic,
// same stores.
__device__ __forceinline__ void du_decide_one_hoisted(const AssocWork &w, int t) {
    const int done = w.cl_done[t];
    const long long m = w.cl_m[t], us = w.cl_u[t], cs = w.cl_cost[t];
    const int nrm = w.cl_nrm[t];
    double best = w.cl_best[t], theta = w.cl_theta[t];
    const double ub0 = w.cl_ub[t];
    int stall = w.cl_stall[t];
    if (done) {
        w.cl_flag[t] = 0;
        return;
    }
    const double L = from_fix(m - us);
    if (nrm == 0) {  // conflict-free and complementary: argmins are optimal
        w.cl_done[t] = 1;
        w.cl_best[t] = L;
        w.cl_ub[t] = from_fix(cs);
        w.cl_flag[t] = 3;
        w.cl_step[t] = 0.0;
        atomicOr(&w.stall_ctr[2], 1);
        return;
    }
    int flag = 0;
    if (L > best + 1e-12) {
        const bool big = L > best + kBigGain * fmax(1.0, fabs(L));
        flag = big ? 9 : 1;
        if (big) atomicOr(&w.stall_ctr[2], 1);
        best = L;
        w.cl_best[t] = L;
        w.cl_stall[t] = 0;
    } else if (++stall >= kPatience) {
        theta *= kShrink;
        w.cl_theta[t] = theta;
        w.cl_stall[t] = 0;
    } else {
        w.cl_stall[t] = stall;
    }
    w.cl_flag[t] = flag;
    if (ub0 - best < 1e-9) {
        w.cl_done[t] = 1;
        w.cl_step[t] = 0.0;
    } else {
        w.cl_step[t] = ub0 < 1e299 ? theta * (ub0 - L) / (double)nrm : 0.0;
        atomicAdd(&w.stall_ctr[1], 1);
    }
}

__device__ __forceinline__ void du_decide_body(ColView c, AssocWork w) {
    if (du_skip(c, w)) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += gridDim.x * blockDim.x) {
        if (w.tstart[t] < 0 || w.uf[t] != t) continue;
        du_decide_one(w, t);
    }
}
__global__ void du_decide_kernel(ColView c, AssocWork w) { du_decide_body(c, w); }


__device__ __forceinline__ void du_apply_body(ColView c, AssocWork w) {
    if (du_skip(c, w)) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && c.idx && w.act_n[2]) w.info[0] = 1;
        return;
    }
    const int nr = *w.row_n;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
        const int r = w.row_list[i];
        const int cl = w.uf[w.row_owner[r]];
        const int fl = w.cl_flag[cl];
        if (fl & 1) w.best_u[r] = w.u[r];
        if (!w.cl_done[cl]) w.u[r] = fmax(0.0, w.u[r] + w.cl_step[cl] * (double)w.usage[r]);
        w.usage[r] = 0;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += stride) {
        if (w.tstart[t] < 0) continue;
        const int cl = w.uf[t];
        if (w.cl_flag[cl] & 2) w.sel[t] = w.targ[t];
        w.tdone[t] = w.cl_done[cl];
        w.tmin[t] = kKeyInf;
        w.targ[t] = -1;
    }
}
__global__ void du_apply_kernel(ColView c, AssocWork w) { du_apply_body(c, w); }


// runs after du_apply: clears the per-cluster accumulators/flags and advances the iteration state
__device__ __forceinline__ void du_finish_body(ColView c, AssocWork w) {
    if (w.info[0]) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += gridDim.x * blockDim.x) {
        w.cl_m[t] = 0;
        w.cl_u[t] = 0;
        w.cl_cost[t] = 0;
        w.cl_nrm[t] = 0;
        w.cl_flag[t] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        w.info[1] += 1;
        // stop when every cluster is settled, or no bound moved noticeably for kStallStop iterations
        w.stall_ctr[0] = w.stall_ctr[2] ? 0 : w.stall_ctr[0] + 1;
        if (w.stall_ctr[1] == 0 || w.stall_ctr[0] >= kStallStop) w.info[0] = 1;
        w.stall_ctr[1] = 0;
        w.stall_ctr[2] = 0;
    }
}
__global__ void du_finish_kernel(ColView c, AssocWork w) { du_finish_body(c, w); }


// du_apply + du_finish in ONE phase for the persistent loop (one grid barrier less per iteration).  The
// early-out word info[0] is not read here: it is known to be 0 inside an iteration, and only the bookkeeping
// thread below may set it, for the check at the top of the next iteration.  cl_flag is not cleared here
// (other threads of this phase still read it): du_decide resets it for every cluster root.
__device__ __forceinline__ void du_apply_finish_body(ColView c, AssocWork w) {
    if (c.idx && w.act_n[2]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) w.info[0] = 1;
        return;
    }
    const int nr = *w.row_n;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
        const int r = w.row_list[i];
        const int cl = w.uf[w.row_owner[r]];
        const int fl = w.cl_flag[cl];
        if (fl & 1) w.best_u[r] = w.u[r];
        if (!w.cl_done[cl]) w.u[r] = fmax(0.0, w.u[r] + w.cl_step[cl] * (double)w.usage[r]);
        w.usage[r] = 0;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += stride) {
        w.cl_m[t] = 0;
        w.cl_u[t] = 0;
        w.cl_cost[t] = 0;
        w.cl_nrm[t] = 0;
        if (w.tstart[t] < 0) continue;
        const int cl = w.uf[t];
        if (w.cl_flag[cl] & 2) w.sel[t] = w.targ[t];
        w.tdone[t] = w.cl_done[cl];
        w.tmin[t] = kKeyInf;
        w.targ[t] = -1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        w.info[1] += 1;
        // stop when every cluster is settled, or no bound moved noticeably for kStallStop iterations
        w.stall_ctr[0] = w.stall_ctr[2] ? 0 : w.stall_ctr[0] + 1;
        if (w.stall_ctr[1] == 0 || w.stall_ctr[0] >= kStallStop) w.info[0] = 1;
        w.stall_ctr[1] = 0;
        w.stall_ctr[2] = 0;
    }
}

static void dual_update(const ColView &c, AssocWork &w, cudaStream_t s) {
    const int tb = (c.n_trees + 127) / 128;
    const int rb = tb > 64 ? tb : 64;
    count_launch(), du_trees_kernel<<<tb, 128, 0, s>>>(c, w);
    count_launch(), du_rows_kernel<<<rb, 256, 0, s>>>(c, w);
    count_launch(), du_decide_kernel<<<tb, 128, 0, s>>>(c, w);
    count_launch(), du_apply_kernel<<<rb, 256, 0, s>>>(c, w);
    count_launch(), du_finish_kernel<<<tb, 128, 0, s>>>(c, w);
}

// ------------------------------------------------------------------------------------------------
// primal greedy in parallel rounds
// ------------------------------------------------------------------------------------------------
// block 0 resets the per-tree state; the row arrays (n_rows entries) are cleared by the WHOLE grid
__device__ __forceinline__ void greedy_init_body(ColView c, AssocWork w, const int *tstart) {
    if (w.info[0]) return;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < c.n_rows; r += gridDim.x * blockDim.x) {
        w.row_taken[r] = 0;
        w.row_bid[r] = kKeyInf;
    }
    if (blockIdx.x != 0) return;
    __shared__ int left;
    if (threadIdx.x == 0) left = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < c.n_trees; t += blockDim.x) {
        const bool part = tstart[t] >= 0 && !w.cl_done[w.uf[t]];
        w.committed[t] = part ? 0 : 1;
        w.prop_key[t] = kKeyInf;
        w.prop_col[t] = -1;
        w.sel_new[t] = -1;
        if (part) atomicAdd(&left, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) w.info[2] = left;
}
__global__ void __launch_bounds__(1024, 1) greedy_init_kernel(ColView c, AssocWork w, const int *tstart) { greedy_init_body(c, w, tstart); }


// bidding key (lower wins): mode 0 = reduced cost; mode 1 = most frequent Lagrangian choice first
// (rounding of the ergodic primal average), reduced cost as tie break
__device__ __forceinline__ unsigned long long greedy_key(const AssocWork &w, int j) {
    const unsigned long long k = f64_key(w.rc[j]);
    if (w.stall_ctr[3] == 0) return k;
    const unsigned long long f = (unsigned long long)(4095 - min(w.freq[j], 4095));
    return (f << 52) | (k >> 12);
}

__device__ __forceinline__ bool rows_free(const ColView &c, const int *taken, int j) {
    for (int k = 0; k < c.width; ++k) {
        const int r = c.rows[(long long)k * c.stride + j];
        if (r >= 0 && taken[r]) return false;
    }
    return true;
}

__device__ __forceinline__ void greedy_prop_body(ColView c, AssocWork w) {
    if (w.info[0] || w.info[2] == 0) return;
    const int n = *c.n_ptr;
    const int nround = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
        int t = -1;
        unsigned long long key = kKeyInf;
        if (i < n) {
            const int j = c.idx ? c.idx[i] : i;
            t = c.tree[j];
            if (!w.committed[t] && rows_free(c, w.row_taken, j))
                key = greedy_key(w, j);
            else
                t = -1;
        }
        const bool head = warp_run_min(t, key);
        if (head && t >= 0) atomicMin(&w.prop_key[t], key);
    }
}
__global__ void __launch_bounds__(256) greedy_prop_kernel(ColView c, AssocWork w) { greedy_prop_body(c, w); }


__device__ __forceinline__ void greedy_arg_body(ColView c, AssocWork w) {
    if (w.info[0] || w.info[2] == 0) return;
    const int n = *c.n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = c.idx ? c.idx[i] : i;
        const int t = c.tree[j];
        if (w.committed[t]) continue;
        if (greedy_key(w, j) == w.prop_key[t] && rows_free(c, w.row_taken, j)) atomicMax(&w.prop_col[t], j);
    }
}
__global__ void __launch_bounds__(256) greedy_arg_kernel(ColView c, AssocWork w) { greedy_arg_body(c, w); }


__device__ __forceinline__ void greedy_commit_body(ColView c, AssocWork w) {
    if (w.info[0] || w.info[2] == 0) return;
    const int T = c.n_trees, R = c.n_rows;
    __shared__ int left;
    if (threadIdx.x == 0) left = 0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (w.committed[t]) continue;
        const int j = w.prop_col[t];
        if (j < 0) continue;
        const unsigned long long bid = (w.prop_key[t] & ~0xFFFFFFull) | (unsigned long long)t;
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0) atomicMin(&w.row_bid[r], bid);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (w.committed[t]) continue;
        const int j = w.prop_col[t];
        bool win = j >= 0;
        if (win) {
            const unsigned long long bid = (w.prop_key[t] & ~0xFFFFFFull) | (unsigned long long)t;
            for (int k = 0; k < c.width; ++k) {
                const int r = c.rows[(long long)k * c.stride + j];
                if (r >= 0 && w.row_bid[r] != bid) win = false;
            }
        }
        if (win) {
            w.committed[t] = 1;
            w.sel_new[t] = j;
            for (int k = 0; k < c.width; ++k) {
                const int r = c.rows[(long long)k * c.stride + j];
                if (r >= 0) w.row_taken[r] = 1;
            }
        } else {
            atomicAdd(&left, 1);
            w.prop_key[t] = kKeyInf;
        }
    }
    __syncthreads();
    // only the rows that received a bid this round need their bid cleared (not all n_rows of them)
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int j = w.prop_col[t];
        if (j < 0) continue;
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0) w.row_bid[r] = kKeyInf;
        }
        w.prop_col[t] = -1;
    }
    if (threadIdx.x == 0) w.info[2] = left;
}
__global__ void __launch_bounds__(1024, 1) greedy_commit_kernel(ColView c, AssocWork w) { greedy_commit_body(c, w); }


// adopt the greedy solution for every cluster it improves
__device__ __forceinline__ void greedy_finish_body(ColView c, AssocWork w, const int *tstart) {
    if (w.info[0] || w.info[2] != 0) return;
    const int T = c.n_trees;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        w.cl_cost[t] = 0;
        w.cl_flag[t] = 0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (tstart[t] < 0) continue;
        const int cl = w.uf[t];
        if (w.cl_done[cl]) continue;
        atomicAdd((unsigned long long *)&w.cl_cost[cl], (unsigned long long)to_fix(col_cost(c, w.sel_new[t], t)));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (tstart[t] < 0 || w.uf[t] != t || w.cl_done[t]) continue;
        const double ub = from_fix(w.cl_cost[t]);
        if (ub < w.cl_ub[t]) {
            w.cl_ub[t] = ub;
            w.cl_flag[t] = 4;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (tstart[t] < 0) continue;
        if (w.cl_flag[w.uf[t]] & 4) w.sel[t] = w.sel_new[t];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {  // leave the dual-update accumulators clean
        w.cl_cost[t] = 0;
        w.cl_flag[t] = 0;
    }
}
__global__ void __launch_bounds__(1024, 1) greedy_finish_kernel(ColView c, AssocWork w, const int *tstart) { greedy_finish_body(c, w, tstart); }


// ------------------------------------------------------------------------------------------------
// The whole dual loop as ONE persistent cooperative kernel (grid = resident CTAs, grid.sync between
// phases): `iters` subgradient iterations, a greedy primal pass every kGreedyEvery iterations whose
// bidding rounds stop as soon as every tree is committed.  Replaces ~7 launches per iteration.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dual_loop_persistent_kernel(ColView c, AssocWork w, int iters, int greedy_every,
                                                                   const int *declined) {
    cg::grid_group grid = cg::this_grid();
    if (declined && *declined == 0) return;   // the cluster version ran this loop
    for (int it = 0; it < iters; ++it) {
        if (((volatile int *)w.info)[0]) break;  // uniform: written before the last grid.sync
        if (it % greedy_every == 0) {
            dual_rc_body<false>(c, w);
            grid.sync();
            for (int mode = 0; mode < (it ? 2 : 1); ++mode) {
                if (blockIdx.x == 0 && threadIdx.x == 0) w.stall_ctr[3] = mode;
                greedy_init_body(c, w, w.tstart);
                grid.sync();
                for (int r = 0; r < kGreedyRounds; ++r) {
                    if (((volatile int *)w.info)[2] == 0) break;
                    greedy_prop_body(c, w);
                    grid.sync();
                    greedy_arg_body(c, w);
                    grid.sync();
                    if (blockIdx.x == 0) greedy_commit_body(c, w);
                    grid.sync();
                }
                if (blockIdx.x == 0) greedy_finish_body(c, w, w.tstart);
                grid.sync();
            }
        }
        dual_rc_body<false>(c, w);
        grid.sync();
        dual_arg_body<false>(c, w);
        grid.sync();
        du_trees_body(c, w);
        grid.sync();
        du_rows_body(c, w);
        grid.sync();
        du_decide_body(c, w);
        grid.sync();
        du_apply_finish_body(c, w);
        grid.sync();
    }
}

// ------------------------------------------------------------------------------------------------
// The same loop inside ONE thread-block cluster (16 CTAs x 1024 threads on sm_100): the iteration is latency
// bound -- a few tens of thousands of columns, six dependent phases -- so what it needs is a cheap barrier and
// short load chains, not more threads.  Versus the grid version above:
//   * barrier.cluster (hardware, ~0.2 us) instead of the cooperative-groups grid barrier (~1.8 us each);
//   * every CTA keeps its slice of the iterated columns {cost, tree, rows} resident in shared memory for the whole
//     launch (the column set is fixed between pricing passes), so the reduced-cost pass is shared-memory reads +
//     ONE hop to the multipliers, and the argmin pass needs no column data at all;
//   * every thread keeps the (row, cluster) pairs it owns in registers, so the row phases are one hop as well.
// Same arithmetic, same fixed-point sums: the result is bit-identical to the grid version.  The kernel declines
// (info[kAssocInfo - 1] bit 30 stays clear -> the grid version runs) when the slice does not fit.
// ------------------------------------------------------------------------------------------------
__device__ unsigned long long g_loop_prof[16];   // ns per phase of the cluster loop (MHT_LOOP_PROF=1 prints them)
constexpr int kClusterCtas = 16;
constexpr int kClusterThreads = 1024;
constexpr int kClusterRowsPerThread = 6;    // rows with a multiplier: <= 16 * 1024 * 6
constexpr int kClusterTreesPerCta = 2048;   // span of tree ids one CTA's slice may cover

__host__ __device__ inline size_t cluster_slice_bytes(int nc, int W) {
    return (size_t)nc * (8 + 8 + 4 + 4 * (size_t)W) + (size_t)kClusterTreesPerCta * 16 + 64;
}

// first position >= i of the iterated list where a new tree starts (columns are sorted by tree)
__device__ __forceinline__ int tree_start_at_or_after(const ColView &c, int n, int i) {
    if (i <= 0) return 0;
    if (i >= n) return n;
    int prev = c.tree[c.idx ? c.idx[i - 1] : i - 1];
    while (i < n) {
        const int t = c.tree[c.idx ? c.idx[i] : i];
        if (t != prev) break;
        ++i;
    }
    return i;
}

__global__ void __launch_bounds__(kClusterThreads, 1)
dual_loop_cluster_kernel(ColView c, AssocWork w, int iters, int greedy_every, int nc_cap, int *declined) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char dl_smem[];
    __shared__ int s_lo, s_hi;
    const int n = *c.n_ptr;
    const int nr = *w.row_n;
    const int nctas = (int)gridDim.x, nth = nctas * (int)blockDim.x;
    const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int per = (n + nctas - 1) / nctas;
    // uniform over the cluster: every CTA sees the same n / nr
    if (per > nc_cap || nr > nth * kClusterRowsPerThread || c.n_trees > nth) {
        if (gtid == 0) *declined = 1;
        return;
    }
    // every CTA owns WHOLE trees: its slice starts at the first tree boundary at or after blockIdx * per, so the
    // per-tree minimum / argmin never leave the CTA
    if (threadIdx.x == 0) {
        s_lo = tree_start_at_or_after(c, n, (int)blockIdx.x * per);
        s_hi = blockIdx.x + 1 == gridDim.x ? n : tree_start_at_or_after(c, n, ((int)blockIdx.x + 1) * per);
    }
    if (gtid == 0) *declined = 0;
    __syncthreads();
    const int lo = s_lo, hi = s_hi, nc = hi - lo;
    const int t_first = nc > 0 ? c.tree[c.idx ? c.idx[lo] : lo] : 0;
    const int t_last = nc > 0 ? c.tree[c.idx ? c.idx[hi - 1] : hi - 1] : -1;
    const int ntl = t_last - t_first + 1;
    cluster.sync();                                   // *declined = 0 is visible before anybody raises it
    if (threadIdx.x == 0 && (nc > nc_cap || ntl > kClusterTreesPerCta)) atomicExch(declined, 1);
    cluster.sync();
    if (*(volatile int *)declined) return;            // uniform: the grid version takes over
    double *s_cost = (double *)dl_smem;
    double *s_rc = s_cost + nc_cap;
    unsigned long long *s_tmin = (unsigned long long *)(s_rc + nc_cap);   // [kClusterTreesPerCta]
    int *s_tree = (int *)(s_tmin + kClusterTreesPerCta);
    int *s_rows = s_tree + nc_cap;             // [W][nc_cap]
    int *s_targ = s_rows + (size_t)c.width * nc_cap;                     // [kClusterTreesPerCta] local column of the argmin
    int *s_cl = s_targ + kClusterTreesPerCta;                            // [kClusterTreesPerCta] cluster label of the tree
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        const int j = c.idx ? c.idx[lo + k] : lo + k;
        const int t = c.tree[j];
        s_tree[k] = t;
        s_cost[k] = col_cost(c, j, t);
        for (int q = 0; q < c.width; ++q) s_rows[q * nc_cap + k] = c.rows[(long long)q * c.stride + j];
    }
    // rows this thread owns for the whole launch: (row, cluster label); the label of a row never changes
    int my_r[kClusterRowsPerThread], my_cl[kClusterRowsPerThread];
    const int nq = (nr + nth - 1) / nth;
#pragma unroll
    for (int q = 0; q < kClusterRowsPerThread; ++q) {
        const int i = gtid + q * nth;
        my_r[q] = -1;
        my_cl[q] = 0;
        if (q < nq && i < nr) {
            my_r[q] = w.row_list[i];
            my_cl[q] = w.uf[w.row_owner[my_r[q]]];
        }
    }
    // the tree this thread serves in the per-tree phases (the kernel declines when n_trees > nth): its cluster label
    // and whether it has columns / is its cluster's root never change during the launch
    const int my_t = gtid < c.n_trees ? gtid : -1;
    int my_t_cl = 0;
    bool my_t_has = false, my_t_root = false;
    if (my_t >= 0) {
        my_t_has = w.tstart[my_t] >= 0;
        my_t_cl = w.uf[my_t];
        my_t_root = my_t_has && my_t_cl == my_t;
    }
    // cluster label of the trees of this CTA's slice (phase A)
    for (int tl = threadIdx.x; tl < ntl; tl += blockDim.x) s_cl[tl] = w.uf[t_first + tl];
    __syncthreads();
    const int nc_round = (nc + 31) & ~31;
    unsigned long long tp = 0;
#define LOOP_PROF(slot)                                               \
    if (gtid == 0) {                                                  \
        unsigned long long now_;                                      \
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now_));      \
        g_loop_prof[slot] += now_ - tp;                               \
        tp = now_;                                                    \
    }
    if (gtid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tp));
    // every phase below is ONE dependent hop to L2: its loads are issued together, before anything consumes them
    const bool overflow = c.idx && w.act_n[2];   // loop invariant: set by the active-list kernels before the launch
    for (int it = 0; it < iters; ++it) {
        // stop flag (uniform: written before the last cluster barrier); consumed once phase A's loads are in flight
        const int stop = ((volatile int *)w.info)[0];
        bool skip = overflow;     // du_skip(): the subgradient phases are skipped once the solve is finished
        if (it % greedy_every == 0) {
            if (stop) break;
            dual_rc_body<false>(c, w);
            cluster.sync();
            for (int mode = 0; mode < (it ? 2 : 1); ++mode) {
                if (gtid == 0) w.stall_ctr[3] = mode;
                greedy_init_body(c, w, w.tstart);
                cluster.sync();
                for (int r = 0; r < kGreedyRounds; ++r) {
                    if (((volatile int *)w.info)[2] == 0) break;
                    greedy_prop_body(c, w);
                    cluster.sync();
                    greedy_arg_body(c, w);
                    cluster.sync();
                    if (blockIdx.x == 0) greedy_commit_body(c, w);
                    cluster.sync();
                }
                if (blockIdx.x == 0) greedy_finish_body(c, w, w.tstart);
                cluster.sync();
            }
            skip = skip || ((volatile int *)w.info)[0] != 0;   // the primal phase may have closed the last gap
            LOOP_PROF(0)
        }
        // ---- phase A, CTA local: reduced costs, per-tree minimum and argmin (ties -> last column) in shared memory,
        //      then the tree's share of the subgradient: row usage, bound and cost sums ----
        for (int tl = threadIdx.x; tl < ntl; tl += blockDim.x) {
            // done flag of the tree (trees without columns in the slice never get looked at): one hop, all trees at once
            s_tmin[tl] = w.tdone[t_first + tl] ? 0ull : kKeyInf;
            s_targ[tl] = -1;
        }
        // pass 1: reduced costs -- nothing but loads and adds, so the gathers of a thread's columns overlap (with each
        // other and with the done flags above)
#pragma unroll 4
        for (int k = threadIdx.x; k < nc; k += blockDim.x) {
            double v = s_cost[k];
            for (int q = 0; q < c.width; ++q) {
                const int r = s_rows[q * nc_cap + k];
                if (r >= 0) v += w.u[r];
            }
            s_rc[k] = v;
        }
        __syncthreads();
        if (stop) break;                                  // uniform over the cluster; nothing global was written yet
        // pass 2: per-tree minimum (warp-aggregated over runs of equal tree); key 0 marks a finished tree
        for (int k = threadIdx.x; k < nc_round; k += blockDim.x) {
            int t = -1;
            unsigned long long key = kKeyInf;
            if (k < nc) {
                t = s_tree[k];
                if (s_tmin[t - t_first] != 0ull) key = f64_key(s_rc[k]);
                else t = -1;
            }
            const bool head = warp_run_min(t, key);
            if (head && t >= 0) atomicMin(&s_tmin[t - t_first], key);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < nc; k += blockDim.x) {
            const int tl = s_tree[k] - t_first;
            if (s_tmin[tl] != 0ull && f64_key(s_rc[k]) == s_tmin[tl]) atomicMax(&s_targ[tl], k);
        }
        __syncthreads();
        LOOP_PROF(1)
        if (!skip) {
            const int ntl_round = (ntl + 31) & ~31;
            for (int tl = threadIdx.x; tl < ntl_round; tl += blockDim.x) {
                const int k = tl < ntl ? s_targ[tl] : -1;
                const bool active = k >= 0;
                int cl = 0;
                long long m = 0, cost = 0;
                if (active) {
                    const int t = t_first + tl;
                    const int j = c.idx ? c.idx[lo + k] : lo + k;
                    cl = s_cl[tl];
                    w.freq[j] += 1;  // ergodic primal estimate: how often this column is the tree's Lagrangian choice
                    w.targ[t] = j;
                    w.tmin[t] = s_tmin[tl];
                    m = to_fix(key_f64(s_tmin[tl]));
                    cost = to_fix(s_cost[k]);
                    for (int q = 0; q < c.width; ++q) {
                        const int r = s_rows[q * nc_cap + k];
                        if (r >= 0) atomicAdd(&w.usage[r], 1);
                    }
                }
                warp_add_ll(w.cl_m, cl, m, active);
                warp_add_ll(w.cl_cost, cl, cost, active);
            }
        }
        cluster.sync();
        LOOP_PROF(3)
        // ---- rows: subgradient, norms (this thread's rows) ----
        if (!skip) {
            int r_done[kClusterRowsPerThread], r_use[kClusterRowsPerThread];
            double r_u[kClusterRowsPerThread];
#pragma unroll
            for (int q = 0; q < kClusterRowsPerThread; ++q) {
                r_done[q] = 1;
                r_use[q] = 0;
                r_u[q] = 0.0;
                if (q < nq && my_r[q] >= 0) {
                    r_done[q] = w.cl_done[my_cl[q]];
                    r_use[q] = w.usage[my_r[q]];
                    r_u[q] = w.u[my_r[q]];
                }
            }
#pragma unroll
            for (int q = 0; q < kClusterRowsPerThread; ++q) {
                if (q >= nq) break;                       // uniform
                const int r = my_r[q], cl = my_cl[q];
                int g = 0;
                long long uf = 0;
                const bool active = r >= 0 && !r_done[q];
                if (active) {
                    g = r_use[q] - 1;
                    const double ur = r_u[q];
                    if (ur <= 0.0 && g < 0) g = 0;
                    w.usage[r] = g;
                    if (ur > 0.0) uf = to_fix(ur);
                }
                warp_add_i(w.cl_nrm, cl, g * g, active);
                warp_add_ll(w.cl_u, cl, uf, active);
            }
        }
        cluster.sync();
        LOOP_PROF(4)
        if (my_t_root && !skip) du_decide_one_hoisted(w, my_t);
        cluster.sync();
        LOOP_PROF(5)
        // ---- apply the step (this thread's rows), per-tree reset, bookkeeping ----
        if (overflow) {
            if (gtid == 0) w.info[0] = 1;
        } else {
            int a_fl[kClusterRowsPerThread], a_done[kClusterRowsPerThread], a_use[kClusterRowsPerThread];
            double a_u[kClusterRowsPerThread], a_step[kClusterRowsPerThread];
#pragma unroll
            for (int q = 0; q < kClusterRowsPerThread; ++q) {
                a_fl[q] = 0;
                a_done[q] = 1;
                a_use[q] = 0;
                a_u[q] = 0.0;
                a_step[q] = 0.0;
                if (q < nq && my_r[q] >= 0) {
                    a_fl[q] = w.cl_flag[my_cl[q]];
                    a_done[q] = w.cl_done[my_cl[q]];
                    a_step[q] = w.cl_step[my_cl[q]];
                    a_u[q] = w.u[my_r[q]];
                    a_use[q] = w.usage[my_r[q]];
                }
            }
            int t_fl = 0, t_done = 0, t_arg = -1;
            if (my_t >= 0 && my_t_has) {
                t_fl = w.cl_flag[my_t_cl];
                t_done = w.cl_done[my_t_cl];
                t_arg = w.targ[my_t];
            }
            int b_iters = 0, b_s0 = 0, b_s1 = 0, b_s2 = 0;
            if (gtid == 0) {
                b_iters = w.info[1];
                b_s0 = w.stall_ctr[0];
                b_s1 = w.stall_ctr[1];
                b_s2 = w.stall_ctr[2];
            }
#pragma unroll
            for (int q = 0; q < kClusterRowsPerThread; ++q) {
                if (q >= nq) break;
                const int r = my_r[q];
                if (r < 0) continue;
                const double ur = a_u[q];
                if (a_fl[q] & 1) w.best_u[r] = ur;
                if (!a_done[q]) w.u[r] = fmax(0.0, ur + a_step[q] * (double)a_use[q]);
                w.usage[r] = 0;
            }
            if (my_t >= 0) {
                const int t = my_t;
                w.cl_m[t] = 0;
                w.cl_u[t] = 0;
                w.cl_cost[t] = 0;
                w.cl_nrm[t] = 0;
                if (my_t_has) {
                    if (t_fl & 2) w.sel[t] = t_arg;
                    w.tdone[t] = t_done;
                    w.tmin[t] = kKeyInf;
                    w.targ[t] = -1;
                }
            }
            if (gtid == 0) {
                w.info[1] = b_iters + 1;
                const int s0 = b_s2 ? 0 : b_s0 + 1;
                w.stall_ctr[0] = s0;
                if (b_s1 == 0 || s0 >= kStallStop) w.info[0] = 1;
                w.stall_ctr[1] = 0;
                w.stall_ctr[2] = 0;
            }
        }
        cluster.sync();
        LOOP_PROF(6)
        if (gtid == 0) g_loop_prof[7] += 1;
    }
#undef LOOP_PROF
}

// ------------------------------------------------------------------------------------------------
// sifting: the dual iterations run on an ACTIVE subset of the columns (reduced cost within a
// threshold of the tree minimum at the last full pricing pass, plus every conflict-free all-miss
// column and the incumbent); full passes re-price all columns, so the final bound is exact.
// ------------------------------------------------------------------------------------------------
__device__ __constant__ double kActDelta[4] = {3.0, 1.5, 0.75, 0.25};

__device__ __forceinline__ int active_mask(const ColView &c, const AssocWork &w, int j) {
    const int t = c.tree[j];
    if (w.tdone[t]) return 0;
    const double exc = w.rc[j] - key_f64(w.tmin[t]);
    // the first column of a tree is its all-miss leaf (the miss child is the first child at every level):
    // it uses no row, so it keeps the restricted problem feasible; the incumbent stays in as well
    const bool keep = j == w.sel[t] || j == w.tstart[t];
    int m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) m |= (keep || exc <= kActDelta[q]) ? (1 << q) : 0;
    return m;
}

__global__ void __launch_bounds__(256) active_count_kernel(ColView c, AssocWork w) {
    const int n = *c.n_ptr;
    const int ntiles = (n + 255) / 256;
    const long long plane = (long long)(ntiles + 1);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j = tile * 256 + threadIdx.x;
        const int m = j < n ? active_mask(c, w, j) : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cnt = __syncthreads_count(m & (1 << q));
            if (threadIdx.x == 0) w.act_tile[q * plane + tile] = cnt;
        }
    }
}

// single CTA: pick the widest threshold whose list fits, exclusive-scan its tile counts in place
__global__ void __launch_bounds__(1024, 1) active_scan_kernel(ColView c, AssocWork w) {
    const int n = *c.n_ptr;
    const int ntiles = (n + 255) / 256;
    const long long plane = (long long)(ntiles + 1);
    __shared__ long long tot[4];
    __shared__ int pick;
    __shared__ int carry;
    __shared__ int wsum[32];
    if (threadIdx.x < 4) tot[threadIdx.x] = 0;
    __syncthreads();
    for (int q = 0; q < 4; ++q) {
        long long s = 0;
        for (int i = threadIdx.x; i < ntiles; i += blockDim.x) s += w.act_tile[q * plane + i];
        atomicAdd((unsigned long long *)&tot[q], (unsigned long long)s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pick = 3;
        for (int q = 3; q >= 0; --q)
            if (tot[q] <= w.cap_act) pick = q;
        const bool fits = tot[pick] <= w.cap_act;
        w.act_n[0] = fits ? (int)tot[pick] : 0;
        w.act_n[1] = pick;
        w.act_n[2] = fits ? 0 : 1;   // 1 = even the tightest threshold does not fit: sifting disabled
        carry = 0;
    }
    __syncthreads();
    int *cnt = w.act_tile + pick * plane;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int kPer = 8;   // consecutive tile counts per thread (serial in registers): 8x fewer block barriers
    for (int base = 0; base < ntiles; base += blockDim.x * kPer) {
        const int i0 = base + threadIdx.x * kPer;
        int v[kPer], tsum = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            v[q] = (i0 + q < ntiles) ? cnt[i0 + q] : 0;
            tsum += v[q];
        }
        int incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tt;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ss = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tt = __shfl_up_sync(0xffffffffu, ss, o);
                if (lane >= o) ss += tt;
            }
            wsum[lane] = ss;
        }
        __syncthreads();
        int run = carry + (wid ? wsum[wid - 1] : 0) + incl - tsum;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if (i0 + q < ntiles) cnt[i0 + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += wsum[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) active_scatter_kernel(ColView c, AssocWork w) {
    if (w.act_n[2]) return;
    const int n = *c.n_ptr;
    const int ntiles = (n + 255) / 256;
    const long long plane = (long long)(ntiles + 1);
    const int q = w.act_n[1];
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j = tile * 256 + threadIdx.x;
        const int f = (j < n && (active_mask(c, w, j) & (1 << q))) ? 1 : 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[wid] = __popc(b);
        __syncthreads();
        int before = __popc(b & ((1u << lane) - 1));
        for (int k = 0; k < wid; ++k) before += wsum[k];
        if (f) {
            w.act_col[w.act_tile[q * plane + tile] + before] = j;
            w.freq[j] = 0;
        }
        __syncthreads();
    }
}

// between sifting rounds: continue from the best multipliers, re-arm the iteration
__global__ void __launch_bounds__(1024, 1) sift_rearm_kernel(ColView c, AssocWork w) {
    for (int r = threadIdx.x; r < c.n_rows; r += blockDim.x) w.u[r] = w.best_u[r];
    for (int t = threadIdx.x; t < c.n_trees; t += blockDim.x) {
        w.tmin[t] = kKeyInf;
        w.targ[t] = -1;
        if (w.uf[t] == t && !w.cl_done[t]) {
            w.cl_best[t] = -1e300;   // bounds from a restricted column set are not comparable across rounds
            w.cl_stall[t] = 0;
        }
    }
    if (threadIdx.x == 0) {
        w.info[0] = 0;
        w.stall_ctr[0] = 0;
    }
}

__global__ void reset_tree_min_kernel(int T, AssocWork w) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        w.tmin[t] = kKeyInf;
        w.targ[t] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// final bound, candidates, components, exact search
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) final_prepare_kernel(ColView c, AssocWork w) {
    for (int r = threadIdx.x; r < c.n_rows; r += blockDim.x) {
        w.u[r] = w.best_u[r];
        w.usage[r] = 0;
        w.row_mark[r] = -1;
        w.row_taken[r] = 0;
        w.row_holder[r] = -1;
    }
    for (int t = threadIdx.x; t < c.n_trees; t += blockDim.x) {
        w.tmin[t] = kKeyInf;
        w.targ[t] = -1;
    }
    if (threadIdx.x == 0) w.info[0] = 0;  // re-arm the early-out word for the column passes
}

// per-cluster bound at best_u and gap -> cl_step; totals -> objective[0..1]
__global__ void __launch_bounds__(1024, 1) final_bound_kernel(ColView c, AssocWork w, const int *tstart) {
    const int T = c.n_trees, R = c.n_rows;
    __shared__ long long lb, ob;
    if (threadIdx.x == 0) lb = ob = 0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        w.cl_m[t] = 0;
        w.cl_u[t] = 0;
        w.cl_cost[t] = 0;
        w.comp_uf[t] = t;
        w.cand_fill[t] = 0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (tstart[t] < 0) continue;
        const int cl = w.uf[t];
        if (w.sel[t] < 0) {
            // no primal solution reached this tree (the greedy ran out of rounds): fall back to the tree's first
            // column when it uses no row (forest columns: the all-miss leaf) -- always conflict free.  Only a tree
            // without such a column keeps its Lagrangian choice, which may collide: flagged, never certified.
            const int j0 = tstart[t];
            bool row_free = true;
            for (int k = 0; k < c.width; ++k) row_free = row_free && c.rows[(long long)k * c.stride + j0] < 0;
            if (row_free) {
                w.sel[t] = j0;
            } else {
                w.sel[t] = w.targ[t];
                atomicAdd(&w.info[11], 1);
            }
        }
        atomicAdd((unsigned long long *)&w.cl_m[cl], (unsigned long long)to_fix(key_f64(w.tmin[t])));
        atomicAdd((unsigned long long *)&w.cl_cost[cl], (unsigned long long)to_fix(col_cost(c, w.sel[t], t)));
    }
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int o = w.row_owner[r];
        if (o >= 0 && w.u[r] > 0.0)
            atomicAdd((unsigned long long *)&w.cl_u[w.uf[o]], (unsigned long long)to_fix(w.u[r]));
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (tstart[t] < 0 || w.uf[t] != t) continue;
        const double L = from_fix(w.cl_m[t] - w.cl_u[t]);
        const double ub = from_fix(w.cl_cost[t]);
        w.cl_best[t] = L;
        w.cl_ub[t] = ub;
        const double gap = ub - L;
        w.cl_step[t] = gap;                     // candidate threshold of this cluster
        w.cl_done[t] = gap <= 1e-9 ? 1 : 0;     // closed: incumbent proven optimal
        atomicAdd((unsigned long long *)&lb, (unsigned long long)(w.cl_m[t] - w.cl_u[t]));
        atomicAdd((unsigned long long *)&ob, (unsigned long long)w.cl_cost[t]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        w.objective[0] = from_fix(lb);
        w.objective[1] = from_fix(ob);
    }
}

__device__ __forceinline__ bool is_candidate(const ColView &c, const AssocWork &w, int j, int t) {
    const int cl = w.uf[t];
    if (w.cl_done[cl]) return false;
    const double exc = w.rc[j] - key_f64(w.tmin[t]);
    return exc <= w.cl_step[cl] + 1e-9 || j == w.sel[t];
}

__global__ void __launch_bounds__(256) cand_count_kernel(ColView c, AssocWork w) {
    const int n = *c.n_ptr;
    const int nround = (n + 31) & ~31;
    const int lane = threadIdx.x & 31;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += gridDim.x * blockDim.x) {
        const int t = j < n ? c.tree[j] : -1;
        int v = (j < n && is_candidate(c, w, j, t)) ? 1 : 0;
        // sum over runs of equal tree inside the warp, one atomic per run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ov = __shfl_down_sync(0xffffffffu, v, o);
            const int ot = __shfl_down_sync(0xffffffffu, t, o);
            if (lane + o < 32 && ot == t) v += ov;
        }
        const int pt = __shfl_up_sync(0xffffffffu, t, 1);
        if ((lane == 0 || pt != t) && t >= 0 && v) atomicAdd(&w.cand_cnt[t], v);
    }
}

__global__ void __launch_bounds__(1024, 1) cand_scan_kernel(ColView c, AssocWork w) {
    // T is small (<= ~10^4): serial chunks per thread + one block scan
    const int T = c.n_trees;
    __shared__ int part[1024];
    __shared__ int over;
    if (threadIdx.x == 0) over = 0;
    const int per = (T + blockDim.x - 1) / blockDim.x;
    const int lo = min(T, (int)threadIdx.x * per), hi = min(T, lo + per);
    int s = 0;
    for (int t = lo; t < hi; ++t) {
        s += w.cand_cnt[t];
        if (w.cand_cnt[t] > kMaxCandPerTree) over = 1;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) {
            const int v = part[i];
            part[i] = acc;
            acc += v;
        }
        w.info[3] = acc;
        if (acc > w.cap_cand || over) w.info[6] = 1;
    }
    __syncthreads();
    int acc = part[threadIdx.x];
    for (int t = lo; t < hi; ++t) {
        w.cand_off[t] = acc;
        acc += w.cand_cnt[t];
    }
    if (hi == T && lo <= T) w.cand_off[T] = acc;
}

__global__ void __launch_bounds__(256) cand_fill_kernel(ColView c, AssocWork w) {
    if (w.info[6]) return;
    const int n = *c.n_ptr;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int t = c.tree[j];
        if (is_candidate(c, w, j, t)) w.cand_col[w.cand_off[t] + atomicAdd(&w.cand_fill[t], 1)] = j;
    }
}

// per tree: sort candidates by (excess, column)
__global__ void cand_sort_kernel(ColView c, AssocWork w) {
    if (w.info[6]) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += gridDim.x * blockDim.x) {
        const int cnt = w.cand_cnt[t];
        if (cnt == 0 || cnt > kDomMax) continue;   // long lists stay unsorted: only the branch & bound sees them
        int *v = w.cand_col + w.cand_off[t];
        for (int i = 1; i < cnt; ++i) {
            const int key = v[i];
            const double kr = w.rc[key];
            int p = i - 1;
            while (p >= 0 && (w.rc[v[p]] > kr || (w.rc[v[p]] == kr && v[p] > key))) {
                v[p + 1] = v[p];
                --p;
            }
            v[p + 1] = key;
        }
    }
}

// ---- dominance reduction of the candidate lists ----------------------------------------------------
// A row is CONTESTED when candidates of two or more trees use it.  Within a tree, candidate a dominates
// candidate b when it is not more expensive and its contested rows are a subset of b's: any solution
// using b stays feasible and gets no worse with a (a's other rows are wanted by nobody else).  Removing
// dominated candidates keeps an optimal solution and collapses the "independent cheap alternatives" that
// make plain enumeration explode; rows stop being contested as lists shrink, so the pass is repeated.
__global__ void contest_reset_kernel(AssocWork w) {
    if (w.info[6]) return;
    const int nr = *w.row_n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
        const int r = w.row_list[i];
        w.row_mark[r] = -1;
        w.row_cont[r] = 0;
    }
}

// also used for the component union (uf != null): trees sharing a row among surviving candidates
__global__ void contest_mark_kernel(ColView c, AssocWork w, int *uf) {
    if (w.info[6]) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp; t < c.n_trees; t += nwarp) {
        const int cnt = w.cand_cnt[t];
        const int *v = w.cand_col + w.cand_off[t];
        for (int i = lane; i < cnt; i += 32) {
            const int j = v[i];
            for (int k = 0; k < c.width; ++k) {
                const int r = c.rows[(long long)k * c.stride + j];
                if (r < 0) continue;
                int o = w.row_mark[r];
                if (o < 0) {
                    o = atomicCAS(&w.row_mark[r], -1, t);
                    if (o < 0) o = t;
                }
                if (o != t) {
                    w.row_cont[r] = 1;
                    if (uf) uf_union(uf, t, o);
                }
            }
        }
    }
}

__device__ __forceinline__ bool contested_subset(const ColView &c, const int *row_cont, int a, int b) {
    for (int k = 0; k < c.width; ++k) {
        const int r = c.rows[(long long)k * c.stride + a];
        if (r < 0 || !row_cont[r]) continue;
        bool found = false;
        for (int q = 0; q < c.width; ++q) found = found || c.rows[(long long)q * c.stride + b] == r;
        if (!found) return false;
    }
    return true;
}

// one WARP per tree: lane = the candidate b under test (lists are at most kDomMax = 64 long: two per lane)
__global__ void __launch_bounds__(256) cand_dominance_kernel(ColView c, AssocWork w) {
    if (w.info[6]) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp; t < c.n_trees; t += nwarp) {
        const int cnt = w.cand_cnt[t];
        if (cnt < 2 || cnt > kDomMax) continue;
        int *v = w.cand_col + w.cand_off[t];
        unsigned long long drop = 0ull;
        for (int half = 0; half < 2; ++half) {
            const int ib = lane + 32 * half;
            bool dominated = false;
            if (ib < cnt) {
                const int b = v[ib];
                const double cb = col_cost(c, b, t);
                for (int ia = 0; ia < cnt && !dominated; ++ia) {
                    if (ia == ib) continue;
                    const int a = v[ia];
                    const double ca = col_cost(c, a, t);
                    if (!(ca < cb || (ca == cb && a > b))) continue;   // ties: the later leaf wins, like the reference
                    dominated = contested_subset(c, w.row_cont, a, b);
                }
            }
            drop |= (unsigned long long)__ballot_sync(0xffffffffu, dominated) << (32 * half);
        }
        // the incumbent must stay inside the lists (components are only independent over listed columns):
        // if it was dominated, a surviving dominator replaces it -- feasible and not more expensive
        if (lane == 0) {
            const int inc = w.sel[t];
            for (int ib = 0; ib < cnt; ++ib) {
                if (v[ib] != inc || !(drop >> ib & 1ull)) continue;
                const double cb = col_cost(c, inc, t);
                bool replaced = false;
                for (int ia = 0; ia < cnt && !replaced; ++ia) {
                    if (drop >> ia & 1ull) continue;
                    const int a = v[ia];
                    const double ca = col_cost(c, a, t);
                    if ((ca < cb || (ca == cb && a > inc)) && contested_subset(c, w.row_cont, a, inc)) {
                        w.sel[t] = a;
                        replaced = true;
                    }
                }
                if (!replaced) drop &= ~(1ull << ib);
            }
            int n = 0;
            for (int i = 0; i < cnt; ++i)
                if (!(drop >> i & 1ull)) v[n++] = v[i];
            w.cand_cnt[t] = n;
        }
        __syncwarp();
    }
}

// single CTA: group candidate trees by component; singletons are resolved here
__global__ void __launch_bounds__(1024, 1) comp_build_kernel(ColView c, AssocWork w) {
    if (w.info[6]) return;
    const int T = c.n_trees;
    __shared__ int ncomp, maxc;
    if (threadIdx.x == 0) ncomp = maxc = 0;
    __shared__ int nsurv;
    if (threadIdx.x == 0) nsurv = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        w.comp_cnt[t] = 0;
        if (w.cand_cnt[t] > 0) {
            w.comp_uf[t] = uf_find(w.comp_uf, t);
            atomicAdd(&nsurv, w.cand_cnt[t]);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x)
        if (w.cand_cnt[t] > 0) atomicAdd(&w.comp_cnt[w.comp_uf[t]], 1);
    __syncthreads();
    // serial offsets (T small); components with >= 2 trees get a slot in comp_off order
    if (threadIdx.x == 0) {
        int acc = 0, k = 0;
        for (int t = 0; t < T; ++t) {
            const int n = w.comp_cnt[t];
            w.cl_nrm[t] = -1;  // component slot of label t
            if (n >= 2) {
                w.comp_off[k] = acc;
                w.cl_nrm[t] = k++;
                acc += n;
                maxc = max(maxc, n);
            }
        }
        w.comp_off[k] = acc;
        ncomp = k;
        w.info[4] = k;
        w.info[9] = maxc;
        w.info[13] = w.info[3];   // candidates after reduced-cost fixing
        w.info[3] = nsurv;        // ... and after dominance reduction
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) w.cand_fill[t] = 0;  // reuse as per-component cursor
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (w.cand_cnt[t] == 0) continue;
        const int lab = w.comp_uf[t];
        const int slot = w.cl_nrm[lab];
        if (slot >= 0) {
            w.comp_trees[w.comp_off[slot] + atomicAdd(&w.cand_fill[lab], 1)] = t;
        } else {  // alone: cheapest candidate wins
            const int *v = w.cand_col + w.cand_off[t];
            int best = v[0];
            double bc = col_cost(c, best, t);
            for (int i = 1; i < w.cand_cnt[t]; ++i) {
                const double cc = col_cost(c, v[i], t);
                if (cc < bc || (cc == bc && v[i] > best)) {
                    bc = cc;
                    best = v[i];
                }
            }
            w.sel[t] = best;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Parallel local search on the incumbent, BEFORE the candidate lists are cut (so it also runs when those
// overflow: the giant clusters, where the greedy primal alone is tens of NLLR units above the optimum).
// Every tree has a shortlist of kLsK columns (per-lane minima of its reduced costs at the final multipliers +
// incumbent + all-miss leaf).  Round: (1) one warp per tree evaluates, lane = shortlisted column, the move
// "take this column; the at most ONE tree it collides with moves to its best shortlisted column that is free
// afterwards" and proposes the best improving one, bidding (cost change, tree) on the trees and rows it would
// touch; (2) proposals that hold the lowest bid everywhere are applied -- they are disjoint by construction;
// (3) bids are cleared.  Deterministic: bids are atomicMin keys.
// ------------------------------------------------------------------------------------------------
__global__ void ls_begin_kernel(ColView c, AssocWork w, const int *tstart) {
    const int T = c.n_trees;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        w.ls_tbid[t] = kKeyInf;
        w.ls_j[t] = -1;
        if (tstart[t] < 0) continue;
        const int j = w.sel[t];
        if (j < 0) {
            w.ls_ctr[2] = 1;   // no feasible incumbent: nothing to improve on
            continue;
        }
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0) w.row_holder[r] = t;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) w.ls_ctr[0] = w.ls_ctr[1] = 0;
}

// one CTA per tree: thread i scans columns tstart + i, + 256, ...; the 8 warps' minima of each lane class
// (column index mod 32) are combined in shared memory
__global__ void __launch_bounds__(256) ls_shortlist_kernel(ColView c, AssocWork w, const int *tstart, const int *tend) {
    __shared__ double s_v[8][32];
    __shared__ int s_j[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int t = blockIdx.x; t < c.n_trees; t += gridDim.x) {
        int best = -1;
        double bv = 1e300;
        if (tstart[t] >= 0 && !w.cl_done[w.uf[t]])
            for (int j = tstart[t] + threadIdx.x; j < tend[t]; j += 256) {
                const double v = w.rc[j];
                if (v <= bv) {   // ties: the later leaf, like everywhere else
                    bv = v;
                    best = j;
                }
            }
        s_v[wid][lane] = bv;
        s_j[wid][lane] = best;
        __syncthreads();
        if (wid == 0) {
            for (int k = 1; k < 8; ++k) {
                const double v = s_v[k][lane];
                const int j = s_j[k][lane];
                if (j >= 0 && (best < 0 || v < bv || (v == bv && j > best))) {
                    bv = v;
                    best = j;
                }
            }
            w.ls_short[t * kLsK + lane] = best;
            if (lane == 0) {
                w.ls_short[t * kLsK + 32] = tstart[t] >= 0 ? w.sel[t] : -1;
                w.ls_short[t * kLsK + 33] = tstart[t];
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ bool col_uses_row(const ColView &c, int j, int r) {
    for (int k = 0; k < c.width; ++k)
        if (c.rows[(long long)k * c.stride + j] == r) return true;
    return false;
}

__global__ void __launch_bounds__(256) ls_propose_kernel(ColView c, AssocWork w, const int *tstart) {
    if (w.ls_ctr[2]) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp; t < c.n_trees; t += nwarp) {
        if (tstart[t] < 0 || w.cl_done[w.uf[t]]) continue;
        const int cur = w.sel[t];
        const double cur_cost = col_cost(c, cur, t);
        double bd = 0.0;
        int bj = -1, bo = -1, bjo = -1;
        for (int q = lane; q < kLsK; q += 32) {
            const int j = w.ls_short[t * kLsK + q];
            if (j < 0 || j == cur) continue;
            double delta = col_cost(c, j, t) - cur_cost;
            int other = -1, nother = 0;
            for (int k = 0; k < c.width && nother < 2; ++k) {
                const int r = c.rows[(long long)k * c.stride + j];
                if (r < 0) continue;
                const int h = w.row_holder[r];
                if (h >= 0 && h != t && h != other) {
                    other = h;
                    ++nother;
                }
            }
            if (nother > 1) continue;
            int jo = -1;
            if (nother == 1) {
                const double cur2 = col_cost(c, w.sel[other], other);
                double best2 = 1e300;
                for (int q2 = 0; q2 < kLsK; ++q2) {
                    const int cand = w.ls_short[other * kLsK + q2];
                    if (cand < 0 || cand == w.sel[other]) continue;
                    const double cc = col_cost(c, cand, other);
                    if (cc > best2 || (cc == best2 && cand <= jo)) continue;
                    bool ok = true;
                    for (int k = 0; k < c.width && ok; ++k) {
                        const int r2 = c.rows[(long long)k * c.stride + cand];
                        if (r2 < 0) continue;
                        if (col_uses_row(c, j, r2)) ok = false;                      // j takes its rows
                        const int h = w.row_holder[r2];
                        if (h >= 0 && h != other && h != t) ok = false;              // rows t holds now are released
                    }
                    if (ok) {
                        best2 = cc;
                        jo = cand;
                    }
                }
                if (jo < 0) continue;
                delta += best2 - cur2;
            }
            if (delta < -1e-12 && (delta < bd || (delta == bd && j > bj))) {
                bd = delta;
                bj = j;
                bo = nother ? other : -1;
                bjo = jo;
            }
        }
        // best over the lanes (lowest cost change, ties -> larger column)
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
            const int ojo = __shfl_xor_sync(0xffffffffu, bjo, o);
            if (oj >= 0 && (bj < 0 || od < bd || (od == bd && oj > bj))) {
                bd = od;
                bj = oj;
                bo = oo;
                bjo = ojo;
            }
        }
        if (lane == 0) {
            w.ls_j[t] = bj;
            if (bj >= 0) {
                w.ls_delta[t] = bd;
                w.ls_o[t] = bo;
                w.ls_jo[t] = bjo;
                const unsigned long long bid = (f64_key(bd) & ~0xFFFFFFull) | (unsigned long long)t;
                atomicMin(&w.ls_tbid[t], bid);
                if (bo >= 0) atomicMin(&w.ls_tbid[bo], bid);
                for (int k = 0; k < c.width; ++k) {
                    const int r = c.rows[(long long)k * c.stride + bj];
                    if (r >= 0) atomicMin(&w.row_bid[r], bid);
                    const int r2 = bo >= 0 ? c.rows[(long long)k * c.stride + bjo] : -1;
                    if (r2 >= 0) atomicMin(&w.row_bid[r2], bid);
                }
            }
        }
    }
}

__global__ void ls_apply_kernel(ColView c, AssocWork w) {
    if (w.ls_ctr[2]) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += gridDim.x * blockDim.x) {
        const int j = w.ls_j[t];
        if (j < 0) continue;
        const int o = w.ls_o[t], jo = w.ls_jo[t];
        const unsigned long long bid = (f64_key(w.ls_delta[t]) & ~0xFFFFFFull) | (unsigned long long)t;
        bool win = w.ls_tbid[t] == bid && (o < 0 || w.ls_tbid[o] == bid);
        for (int k = 0; k < c.width && win; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0 && w.row_bid[r] != bid) win = false;
            const int r2 = o >= 0 ? c.rows[(long long)k * c.stride + jo] : -1;
            if (r2 >= 0 && w.row_bid[r2] != bid) win = false;
        }
        if (!win) continue;
        // release the old rows of the moving trees, then take the new ones
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + w.sel[t]];
            if (r >= 0) w.row_holder[r] = -1;
            const int r2 = o >= 0 ? c.rows[(long long)k * c.stride + w.sel[o]] : -1;
            if (r2 >= 0) w.row_holder[r2] = -1;
        }
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0) w.row_holder[r] = t;
            const int r2 = o >= 0 ? c.rows[(long long)k * c.stride + jo] : -1;
            if (r2 >= 0) w.row_holder[r2] = o;
        }
        w.sel[t] = j;
        if (o >= 0) w.sel[o] = jo;
        atomicAdd(&w.ls_ctr[0], 1);
    }
}

// clear this round's bids; stop when a round applied nothing
__global__ void ls_clear_kernel(ColView c, AssocWork w) {
    if (w.ls_ctr[2]) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < c.n_trees; t += gridDim.x * blockDim.x) {
        const int j = w.ls_j[t];
        if (j < 0) continue;
        const int o = w.ls_o[t], jo = w.ls_jo[t];
        w.ls_tbid[t] = kKeyInf;
        if (o >= 0) w.ls_tbid[o] = kKeyInf;
        for (int k = 0; k < c.width; ++k) {
            const int r = c.rows[(long long)k * c.stride + j];
            if (r >= 0) w.row_bid[r] = kKeyInf;
            const int r2 = o >= 0 ? c.rows[(long long)k * c.stride + jo] : -1;
            if (r2 >= 0) w.row_bid[r2] = kKeyInf;
        }
    }
}
__global__ void ls_round_end_kernel(AssocWork w) {
    if (w.ls_ctr[2]) return;
    if (w.ls_ctr[0] == 0) w.ls_ctr[2] = 1;
    w.ls_ctr[1] += w.ls_ctr[0];
    w.ls_ctr[0] = 0;
}

// one CTA per component, thread 0 searches depth first with the Lagrangian bound
// (components of up to kDfsMaxTrees trees: enumeration is cheap there; everything else, and whatever this search
// gives up on, goes to the branch & bound below)
__global__ void __launch_bounds__(32) branch_bound_kernel(ColView c, AssocWork w, int budget, double *fscratch) {
    if (w.info[6]) return;
    const int ncomp = w.info[4];
    for (int comp = blockIdx.x; comp < ncomp; comp += gridDim.x) {
        if (threadIdx.x != 0) continue;
        const int off = w.comp_off[comp], k = w.comp_off[comp + 1] - off;
        w.bbw.comp_state[comp] = 0;
        if (k > kDfsMaxTrees) continue;
        int *trees = w.comp_trees + off;
        bool lists_short = true;
        for (int i = 0; i < k; ++i) lists_short = lists_short && w.cand_cnt[trees[i]] <= kDomMax;
        if (!lists_short) continue;
        // deterministic order: fewest candidates first, then tree index
        for (int i = 1; i < k; ++i) {
            const int key = trees[i];
            int p = i - 1;
            while (p >= 0 && (w.cand_cnt[trees[p]] > w.cand_cnt[key] ||
                              (w.cand_cnt[trees[p]] == w.cand_cnt[key] && trees[p] > key))) {
                trees[p + 1] = trees[p];
                --p;
            }
            trees[p + 1] = key;
        }
        int *pos = w.cand_stack + 3 * (long long)off, *chosen = pos + k, *bestsel = chosen + k;
        double *exc_acc = fscratch + 2 * ((long long)off + comp), *cost_acc = exc_acc + (k + 1);
        // component bound: sum of tree minima minus the multipliers of every row the candidates touch
        double sum_m = 0.0, sum_u = 0.0, best = 0.0;
        for (int i = 0; i < k; ++i) {
            const int t = trees[i];
            sum_m += key_f64(w.tmin[t]);
            best += col_cost(c, w.sel[t], t);
            bestsel[i] = w.sel[t];
            const int *v = w.cand_col + w.cand_off[t];
            for (int q = 0; q < w.cand_cnt[t]; ++q)
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + v[q]];
                    if (r >= 0 && w.usage[r] == 0) {
                        w.usage[r] = 1;
                        sum_u += w.u[r];
                    }
                }
        }
        const double Lcomp = sum_m - sum_u;
        const unsigned long long node_budget = (unsigned long long)budget;
        unsigned long long nodes = 0;
        int depth = 0;
        pos[0] = 0;
        exc_acc[0] = 0.0;
        cost_acc[0] = 0.0;
        bool exhausted = false;
        while (depth >= 0) {
            const int t = trees[depth];
            const int cnt = w.cand_cnt[t];
            const int *v = w.cand_col + w.cand_off[t];
            const double mt = key_f64(w.tmin[t]);
            bool found = false;
            while (pos[depth] < cnt) {
                const int j = v[pos[depth]++];
                const double e = fmax(0.0, w.rc[j] - mt);
                if (Lcomp + exc_acc[depth] + e >= best - 1e-12) {
                    pos[depth] = cnt;
                    break;
                }
                if (!rows_free(c, w.row_taken, j)) continue;
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + j];
                    if (r >= 0) w.row_taken[r] = 1;
                }
                chosen[depth] = j;
                exc_acc[depth + 1] = exc_acc[depth] + e;
                cost_acc[depth + 1] = cost_acc[depth] + col_cost(c, j, t);
                found = true;
                break;
            }
            if (!found) {
                --depth;
                if (depth >= 0)
                    for (int kk = 0; kk < c.width; ++kk) {
                        const int r = c.rows[(long long)kk * c.stride + chosen[depth]];
                        if (r >= 0) w.row_taken[r] = 0;
                    }
                continue;
            }
            if (++nodes > node_budget) {
                exhausted = true;
                break;
            }
            if (depth + 1 == k) {
                if (cost_acc[k] < best - 1e-12) {
                    best = cost_acc[k];
                    for (int i = 0; i < k; ++i) bestsel[i] = chosen[i];
                }
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + chosen[depth]];
                    if (r >= 0) w.row_taken[r] = 0;
                }
            } else {
                ++depth;
                pos[depth] = 0;
            }
        }
        if (exhausted)   // out of budget in the middle of a branch: give the rows of the partial assignment back
            for (int d = 0; d <= depth; ++d)
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + chosen[d]];
                    if (r >= 0) w.row_taken[r] = 0;
                }
        for (int i = 0; i < k; ++i) w.sel[trees[i]] = bestsel[i];
        atomicAdd(w.bb_nodes, nodes);
        if (!exhausted) w.bbw.comp_state[comp] = 1;
        // the search marked the rows its candidates touch (bound bookkeeping): clear them for the next stage
        for (int i = 0; i < k; ++i) {
            const int t = trees[i];
            const int *v = w.cand_col + w.cand_off[t];
            for (int q = 0; q < w.cand_cnt[t]; ++q)
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + v[q]];
                    if (r >= 0) w.usage[r] = 0;
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exact repair of the open components: best-first Lagrangian branch & bound (bb_core.h), one CTA per
// node evaluation, every CTA of the grid serving one shared node pool
// ------------------------------------------------------------------------------------------------
// single CTA: which components still need the search, offsets of their compacted cores
__global__ void __launch_bounds__(1024, 1) bb_plan_kernel(ColView c, AssocWork w, int max_cols_now, double max_gap_now) {
    BBWork &b = w.bbw;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) b.hdr[i] = 0;
        b.p_ctr[0] = b.p_ctr[1] = b.p_ctr[2] = b.p_ctr[3] = 0;
    }
    __syncthreads();
    if (w.info[6]) return;
    const int ncomp = w.info[4];
    __shared__ int n_open;
    if (threadIdx.x == 0) n_open = 0;
    __syncthreads();
    // columns per component (cl_nrm is free here: reuse as scratch)
    for (int k = threadIdx.x; k < ncomp; k += blockDim.x) {
        int cols = 0;
        for (int i = w.comp_off[k]; i < w.comp_off[k + 1]; ++i) cols += w.cand_cnt[w.comp_trees[i]];
        w.cl_stall[k] = cols;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0, col_off = 0, tree_off = 0, max_c = 0, max_t = 0;
        for (int k = 0; k < ncomp; ++k) {
            if (b.comp_state[k]) continue;
            const int nT = w.comp_off[k + 1] - w.comp_off[k], nC = w.cl_stall[k];
            if (nC > b.max_cols || nC > max_cols_now || nT > b.max_trees || n >= kBBMaxNodes / 4) continue;   // stays open: uncertified
            // the effort of closing a component grows exponentially with its cluster's gap: what the time box cannot
            // hope to close is not started (measured on the bench: nothing above a gap of 2 closes in 8 ms)
            if (nT > kDfsMaxTrees && w.cl_step[w.uf[w.comp_trees[w.comp_off[k]]]] > max_gap_now) continue;
            bb::Comp &p = b.comps[n];
            p.nC = nC;
            p.nT = nT;
            p.nR = 0;
            p.W = c.width;
            p.row_stride = b.cap;
            p.cost = b.c_cost + col_off;
            p.tree = b.c_tree + col_off;
            p.rows = b.c_rows + col_off;
            p.tstart = b.t_start + tree_off + n;
            p.nwords = (nC + 31) / 32;
            p.ub_key = b.ub_key + n;
            p.best_sel = b.best_sel + tree_off;
            p.lock = b.lock + n;
            b.lock[n] = 0;
            b.row_cnt[n] = 0;
            b.comp_unproven[n] = 0;
            b.comp_nodes[n] = 0;
            b.comp_slot[n] = k;
            col_off += nC;
            tree_off += nT;
            max_c = max(max_c, nC);
            max_t = max(max_t, nT);
            ++n;
        }
        b.hdr[0] = n;
        b.hdr[1] = max_c;
        b.hdr[3] = max_t;
        b.hdr[5] = col_off;
    }
}

// one CTA per search component: compact its candidate columns, number its rows, cost of the incumbent
__global__ void __launch_bounds__(256) bb_compact_kernel(ColView c, AssocWork w) {
    BBWork &b = w.bbw;
    const int n = b.hdr[0];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        bb::Comp &p = b.comps[i];
        const int k = b.comp_slot[i];
        const int *trees = w.comp_trees + w.comp_off[k];
        int *tstart = (int *)p.tstart;
        const long long col_off = p.cost - b.c_cost;
        const long long tree_off = p.best_sel - b.best_sel;
        {   // tstart = exclusive prefix of the candidate counts, ub = cost of the incumbent (fixed summation order)
            __shared__ int s_cnt[256];
            __shared__ double s_part[256];
            const int per = (p.nT + 255) / 256;
            const int lo = min(p.nT, (int)threadIdx.x * per), hi = min(p.nT, lo + per);
            int acc = 0;
            double ub = 0.0;
            for (int t = lo; t < hi; ++t) {
                acc += w.cand_cnt[trees[t]];
                ub += col_cost(c, w.sel[trees[t]], trees[t]);
                b.t_gtree[tree_off + t] = trees[t];
            }
            s_cnt[threadIdx.x] = acc;
            s_part[threadIdx.x] = ub;
            __syncthreads();
            if (threadIdx.x == 0) {
                int run = 0;
                double tot = 0.0;
                for (int q = 0; q < 256; ++q) {
                    const int v = s_cnt[q];
                    s_cnt[q] = run;
                    run += v;
                    tot += s_part[q];
                }
                tstart[p.nT] = run;
                b.ub_key[i] = bb::key_of(tot);
            }
            __syncthreads();
            acc = s_cnt[threadIdx.x];
            for (int t = lo; t < hi; ++t) {
                tstart[t] = acc;
                acc += w.cand_cnt[trees[t]];
            }
        }
        __syncthreads();
        // columns: warp per tree, lanes over its candidates; rows marked for numbering
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int t = wid; t < p.nT; t += nw) {
            const int gt = trees[t];
            const int *v = w.cand_col + w.cand_off[gt];
            const int cnt = w.cand_cnt[gt], base = tstart[t];
            const int inc = w.sel[gt];
            for (int q = lane; q < cnt; q += 32) {
                const int j = v[q];
                const long long lj = col_off + base + q;
                b.c_cost[lj] = col_cost(c, j, gt);
                b.c_tree[lj] = t;
                b.c_gcol[lj] = j;
                if (j == inc) p.best_sel[t] = base + q;
                for (int kk = 0; kk < c.width; ++kk) {
                    const int r = c.rows[(long long)kk * c.stride + j];
                    if (r >= 0) b.row_local[r] = -2;
                }
            }
        }
        __syncthreads();
        int *grow = b.r_grow + (long long)c.width * col_off;
        for (int lj = threadIdx.x; lj < p.nC; lj += blockDim.x) {
            const int j = b.c_gcol[col_off + lj];
            for (int kk = 0; kk < c.width; ++kk) {
                const int r = c.rows[(long long)kk * c.stride + j];
                if (r >= 0 && atomicCAS(&b.row_local[r], -2, -3) == -2) grow[atomicAdd(&b.row_cnt[i], 1)] = r;
            }
        }
        __syncthreads();
        const int nR = b.row_cnt[i];
        for (int q = threadIdx.x; q < nR; q += blockDim.x) b.row_local[grow[q]] = q;
        __syncthreads();
        for (int lj = threadIdx.x; lj < p.nC; lj += blockDim.x) {
            const int j = b.c_gcol[col_off + lj];
            for (int kk = 0; kk < c.width; ++kk) {
                const int r = c.rows[(long long)kk * c.stride + j];
                b.c_rows[(long long)kk * b.cap + col_off + lj] = r >= 0 ? b.row_local[r] : -1;
            }
        }
        if (threadIdx.x == 0) {
            p.nR = nR;
            atomicMax(&b.hdr[2], nR);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// pool geometry + one root node per component (all columns alive, the dual loop's multipliers)
__global__ void __launch_bounds__(256) bb_root_kernel(ColView c, AssocWork w, double budget_ms) {
    BBWork &b = w.bbw;
    const int n = b.hdr[0];
    if (n == 0) return;
    const int node_words = (b.hdr[1] + 31) / 32, node_rows = max(b.hdr[2], 1);
    const long long node_bytes = 4ll * node_words + 4ll * node_rows;
    long long cap = b.pool_bytes / node_bytes;
    if (cap > kBBMaxNodes) cap = kBBMaxNodes;
    if (cap < 2ll * n + 8) {   // the pool cannot even hold the roots: the components stay open (uncertified)
        if (blockIdx.x == 0 && threadIdx.x == 0) b.hdr[6] = 1;
        return;
    }
    bb::Pool &pl = *b.pool;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pl.cap = (int)cap;
        pl.node_words = node_words;
        pl.node_rows = node_rows;
        pl.state = b.p_state;
        pl.key = b.p_key;
        pl.bound = b.p_bound;
        pl.comp = b.p_comp;
        pl.bt = b.p_bt;
        pl.br = b.p_br;
        pl.alive = (unsigned *)b.pool_mem;
        pl.u = (float *)(b.pool_mem + 4ll * node_words * cap);
        pl.outstanding = b.p_ctr;
        pl.stop = b.p_ctr + 1;
        pl.nodes = b.p_ctr + 2;
        pl.iters = b.p_ctr + 3;
        pl.comp_unproven = b.comp_unproven;
        pl.comp_nodes = b.comp_nodes;
        b.hdr[4] = (int)cap;
        b.p_ctr[0] = n;
        const unsigned long long now = global_timer_ns();
        b.deadline[1] = now;
        b.deadline[0] = now + (unsigned long long)(budget_ms * 1.0e6);
    }
    unsigned *alive = (unsigned *)b.pool_mem;
    float *u = (float *)(b.pool_mem + 4ll * node_words * cap);
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x)
        b.p_state[i] = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const bb::Comp &p = b.comps[i];
        const long long col_off = p.cost - b.c_cost;
        const int *grow = b.r_grow + (long long)c.width * col_off;
        for (int wd = threadIdx.x; wd < p.nwords; wd += blockDim.x) {
            const int left = p.nC - 32 * wd;
            alive[(long long)i * node_words + wd] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
        }
        for (int r = threadIdx.x; r < p.nR; r += blockDim.x) u[(long long)i * node_rows + r] = (float)w.u[grow[r]];
        if (threadIdx.x == 0) {
            b.p_comp[i] = i;
            b.p_bt[i] = -1;
            b.p_br[i] = -1;
            b.p_bound[i] = -1e300;
            b.p_key[i] = -1e300;
        }
    }
}
// (separate launch: the root states become visible after every slot was cleared)
__global__ void bb_open_roots_kernel(AssocWork w) {
    BBWork &b = w.bbw;
    if (b.hdr[6]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) b.hdr[0] = 0;
        return;
    }
    const int n = b.hdr[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b.p_state[i] = 1;
}

struct DeviceCtx {
    double *red_d;                 // [32] shared
    long long *red_l;              // [32] shared
    unsigned long long *word;      // [1] shared
    unsigned char *smem;           // dynamic shared memory for the component state
    size_t smem_bytes;
    char *gscratch;                // this worker's global scratch
    const BBWork *b;
    __device__ int tid() const { return (int)threadIdx.x; }
    __device__ int nthr() const { return (int)blockDim.x; }
    __device__ void sync() { __syncthreads(); }
    __device__ void amin64(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
    __device__ void amax(int *p, int v) { atomicMax(p, v); }
    __device__ void aadd(int *p, int v) { atomicAdd(p, v); }
    __device__ int acas(int *p, int cmp, int val) { return atomicCAS(p, cmp, val); }
    __device__ void fence() { __threadfence(); }
    __device__ void backoff() { __nanosleep(2000); }
    __device__ bool expired() { return global_timer_ns() > b->deadline[0]; }
    // deterministic block sums: warp shuffles, then every thread adds the warp partials in warp order
    __device__ double sum(double v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red_d[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red_d[i];
        __syncthreads();
        return s;
    }
    __device__ long long maxll(long long v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const long long x = __shfl_xor_sync(0xffffffffu, v, o);
            v = x > v ? x : v;
        }
        if ((threadIdx.x & 31) == 0) red_l[threadIdx.x >> 5] = v;
        __syncthreads();
        long long s = red_l[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) s = red_l[i] > s ? red_l[i] : s;
        __syncthreads();
        return s;
    }
    // sums of a and b, maximum of m, and thread 0's x to everybody: one pair of barriers
    __device__ void reduce(double &a, double &b, long long &m, unsigned long long &x) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            const long long y = __shfl_xor_sync(0xffffffffu, m, o);
            m = y > m ? y : m;
        }
        const int nw = (int)(blockDim.x >> 5);
        if ((threadIdx.x & 31) == 0) {
            red_d[threadIdx.x >> 5] = a;
            red_d[32 + (threadIdx.x >> 5)] = b;
            red_l[threadIdx.x >> 5] = m;
        }
        if (threadIdx.x == 0) *word = x;
        __syncthreads();
        double sa = 0.0, sb = 0.0;
        long long sm = red_l[0];
        for (int i = 0; i < nw; ++i) {
            sa += red_d[i];
            sb += red_d[32 + i];
            sm = red_l[i] > sm ? red_l[i] : sm;
        }
        x = *word;
        __syncthreads();
        a = sa;
        b = sb;
        m = sm;
    }
    __device__ unsigned long long bcast(unsigned long long v) {
        if (threadIdx.x == 0) *word = v;
        __syncthreads();
        const unsigned long long r = *word;
        __syncthreads();
        return r;
    }
    // component state in shared memory when it fits, else in this worker's global scratch
    __device__ bool bind(const bb::Comp &p, bb::Scratch &s) {
        if (p.nC > b->max_cols || p.nR > b->max_rows || p.nT > b->max_trees) return false;
        char *g = gscratch;
        auto take = [&](size_t bytes) { char *r = g; g += (bytes + 15) / 16 * 16; return r; };
        s.rc = (double *)take(8ull * b->max_cols);
        s.freq = (int *)take(4ull * b->max_cols);
        s.ubest = (double *)take(8ull * b->max_rows);
        s.best_targ = (int *)take(4ull * b->max_trees);
        s.cand_d = (double *)take(8ull * b->max_trees);
        s.cand_r = (int *)take(4ull * b->max_trees);
        s.alive2 = (unsigned *)take(4ull * ((b->max_cols + 31) / 32));
        const size_t need = (8ull * p.nR + 15) / 16 * 16 + (4ull * p.nR + 15) / 16 * 16 + (8ull * p.nT + 15) / 16 * 16 +
                            (4ull * p.nT + 15) / 16 * 16 + (4ull * p.nwords + 15) / 16 * 16;
        if (need <= smem_bytes) {
            unsigned char *q = smem;
            auto stake = [&](size_t bytes) { unsigned char *r = q; q += (bytes + 15) / 16 * 16; return r; };
            s.u = (double *)stake(8ull * p.nR);
            s.tmin = (unsigned long long *)stake(8ull * p.nT);
            s.usage = (int *)stake(4ull * p.nR);
            s.targ = (int *)stake(4ull * p.nT);
            s.alive = (unsigned *)stake(4ull * p.nwords);
        } else {
            s.u = (double *)take(8ull * b->max_rows);
            s.tmin = (unsigned long long *)take(8ull * b->max_trees);
            s.usage = (int *)take(4ull * b->max_rows);
            s.targ = (int *)take(4ull * b->max_trees);
            s.alive = (unsigned *)take(4ull * ((b->max_cols + 31) / 32));
        }
        return true;
    }
};

__host__ __device__ inline long long bb_scratch_bytes(long long max_cols, long long max_rows, long long max_trees) {
    auto al16 = [](long long v) { return (v + 15) / 16 * 16; };
    return al16(8 * max_cols) + al16(4 * max_cols) + al16(8 * max_rows) + al16(4 * max_trees) + al16(8 * max_trees) +
           al16(4 * max_trees) + al16(8 * max_rows) + al16(8 * max_trees) + al16(4 * max_rows) + al16(4 * max_trees) +
           2 * al16(4 * ((max_cols + 31) / 32)) + 256;
}

__global__ void __launch_bounds__(kBBThreads, 1) bb_search_kernel(AssocWork w, int K_root, int K_node, int max_nodes,
                                                                  int sb_cands, int sb_iters) {
    extern __shared__ __align__(16) unsigned char bb_smem[];
    __shared__ double red_d[64];
    __shared__ long long red_l[32];
    __shared__ unsigned long long word;
    const BBWork &b = w.bbw;
    if (b.hdr[0] == 0) return;
    DeviceCtx ctx;
    ctx.red_d = red_d;
    ctx.red_l = red_l;
    ctx.word = &word;
    ctx.smem = bb_smem;
    unsigned dyn;
    asm volatile("mov.u32 %0, %dynamic_smem_size;" : "=r"(dyn));
    ctx.smem_bytes = dyn;
    ctx.gscratch = b.scratch_mem + (long long)blockIdx.x * b.scratch_per_worker;
    ctx.b = &b;
    bb::Scratch s;
    bb::Pool pl = *b.pool;
    bb::worker(ctx, b.comps, pl, s, K_root, K_node, max_nodes, sb_cands, sb_iters);
}

// results of the search: selection of every searched component, proven flags
__global__ void __launch_bounds__(256) bb_writeback_kernel(ColView c, AssocWork w) {
    BBWork &b = w.bbw;
    const int n = b.hdr[0];
    if (n == 0) return;
    // nodes still in the pool belong to unfinished components
    const int cap = b.hdr[4];
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < cap; i += blockDim.x)
            if (b.p_state[i] != 0) b.comp_unproven[b.p_comp[i]] = 1;
        if (threadIdx.x == 0) {
            atomicAdd(w.bb_nodes, (unsigned long long)b.p_ctr[2]);
            w.info[14] = b.p_ctr[3];   // subgradient iterations spent inside the search
        }
    }
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const bb::Comp &p = b.comps[i];
        const long long col_off = p.cost - b.c_cost;
        const long long tree_off = p.best_sel - b.best_sel;
        for (int t = threadIdx.x; t < p.nT; t += blockDim.x)
            w.sel[b.t_gtree[tree_off + t]] = b.c_gcol[col_off + p.best_sel[t]];
        // rows of this component leave the numbering table clean for the next solve
        const int *grow = b.r_grow + (long long)c.width * col_off;
        for (int r = threadIdx.x; r < p.nR; r += blockDim.x) b.row_local[grow[r]] = -1;
    }
}
// (after the grid finished marking unproven components)
__global__ void bb_state_kernel(AssocWork w) {
    BBWork &b = w.bbw;
    const int n = b.hdr[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!b.comp_unproven[i]) b.comp_state[b.comp_slot[i]] = 1;
}

// single CTA: final objective and certificate
// ---- last line of defence: the selection that leaves the solver is conflict free, whatever the parallel primal heuristics
//      and the time-boxed exact search did.  Every selected column claims its rows (smallest tree index wins a row); a tree
//      that lost a row falls back to its all-miss column (forest columns: the first column of the tree uses no row) and is
//      counted in info[11], which also withdraws the certificate.  Three tiny launches per solve.
__global__ void feas_reset_kernel(int R, int *claim) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) claim[r] = 0x7fffffff;
}
__global__ void feas_claim_kernel(ColView c, AssocWork w, const int *tstart, int *claim) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c.n_trees || tstart[t] < 0) return;
    const int j = w.sel[t];
    if (j < 0) return;
    for (int k = 0; k < c.width; ++k) {
        const int r = c.rows[(long long)k * c.stride + j];
        if (r >= 0) atomicMin(&claim[r], t);
    }
}
__global__ void feas_repair_kernel(ColView c, AssocWork w, const int *tstart, const int *claim) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c.n_trees || tstart[t] < 0) return;
    const int j = w.sel[t];
    if (j < 0) return;
    bool lost = false;
    for (int k = 0; k < c.width; ++k) {
        const int r = c.rows[(long long)k * c.stride + j];
        if (r >= 0 && claim[r] != t) lost = true;
    }
    if (!lost) return;
    const int j0 = tstart[t];
    bool row_free = true;
    for (int k = 0; k < c.width; ++k) row_free = row_free && c.rows[(long long)k * c.stride + j0] < 0;
    if (row_free) w.sel[t] = j0;
    atomicAdd(&w.info[11], 1);
}

__global__ void __launch_bounds__(1024, 1) final_objective_kernel(ColView c, AssocWork w, const int *tstart) {
    __shared__ long long ob;
    if (threadIdx.x == 0) ob = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < c.n_trees; t += blockDim.x)
        if (tstart[t] >= 0)
            atomicAdd((unsigned long long *)&ob, (unsigned long long)to_fix(col_cost(c, w.sel[t], t)));
    __syncthreads();
    if (threadIdx.x == 0) {
        w.objective[1] = from_fix(ob);
        w.info[12] = w.act_n[0];
        int open = 0;
        if (!w.info[6])
            for (int k = 0; k < w.info[4]; ++k) open += w.bbw.comp_state[k] ? 0 : 1;
        w.info[5] = open;   // components whose exact search did not finish
        w.info[10] = (open == 0 && w.info[6] == 0 && w.info[11] == 0) ? 1 : 0;  // certified
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int64_t al(int64_t b) { return (b + 255) / 256 * 256; }

struct Carver {
    char *p;
    int64_t used = 0;
    template <class T> T *take(int64_t n) {
        T *r = p ? (T *)(p + used) : nullptr;
        used += al(n * (int64_t)sizeof(T));
        return r;
    }
};

static int64_t carve_all(Carver &cv, int64_t cap_cols, int64_t T, int64_t R, int64_t cap_cand, AssocWork *w,
                         int **tstart, int **tend, double **fscratch) {
    AssocWork d;
    d.rc = cv.take<double>(cap_cols);
    d.freq = cv.take<int>(cap_cols);
    d.tmin = cv.take<unsigned long long>(T);
    d.targ = cv.take<int>(T);
    d.uf = cv.take<int>(T);
    d.sel = cv.take<int>(T);
    d.sel_new = cv.take<int>(T);
    d.committed = cv.take<int>(T);
    d.prop_key = cv.take<unsigned long long>(T);
    d.prop_col = cv.take<int>(T);
    d.cl_m = cv.take<long long>(T);
    d.cl_u = cv.take<long long>(T);
    d.cl_cost = cv.take<long long>(T);
    d.cl_nrm = cv.take<int>(T);
    d.cl_best = cv.take<double>(T);
    d.cl_ub = cv.take<double>(T);
    d.cl_theta = cv.take<double>(T);
    d.cl_step = cv.take<double>(T);
    d.cl_stall = cv.take<int>(T);
    d.cl_done = cv.take<int>(T);
    d.cl_flag = cv.take<int>(T);
    d.tdone = cv.take<int>(T);
    d.stall_ctr = cv.take<int>(4);
    d.cand_cnt = cv.take<int>(T + 1);
    d.cand_off = cv.take<int>(T + 1);
    d.cand_fill = cv.take<int>(T + 1);
    d.comp_uf = cv.take<int>(T);
    d.comp_trees = cv.take<int>(T);
    d.comp_off = cv.take<int>(T + 1);
    d.comp_cnt = cv.take<int>(T);
    d.row_owner = cv.take<int>(R);
    d.u = cv.take<double>(R);
    d.best_u = cv.take<double>(R);
    d.usage = cv.take<int>(R);
    d.row_bid = cv.take<unsigned long long>(R);
    d.row_taken = cv.take<int>(R);
    d.row_mark = cv.take<int>(R);
    d.row_cont = cv.take<int>(R);
    d.row_holder = cv.take<int>(R);
    d.ls_short = cv.take<int>(T * kLsK);
    d.ls_delta = cv.take<double>(T);
    d.ls_j = cv.take<int>(T);
    d.ls_o = cv.take<int>(T);
    d.ls_jo = cv.take<int>(T);
    d.ls_tbid = cv.take<unsigned long long>(T);
    d.ls_ctr = cv.take<int>(4);
    d.row_list = cv.take<int>(R);
    d.row_n = cv.take<int>(4);
    d.cap_act = cap_cols < (1 << 20) ? cap_cols : (cap_cols / 8 > (1 << 20) ? cap_cols / 8 : (1 << 20));
    d.act_col = cv.take<int>(d.cap_act);
    d.act_tile = cv.take<int>(4 * (cap_cols / 256 + 2));
    d.act_n = cv.take<int>(4);
    d.cand_col = cv.take<int>(cap_cand);
    d.cand_stack = cv.take<int>(3 * T + 8);
    d.cap_cand = cap_cand;
    d.info = cv.take<int>(kAssocInfo);
    d.bb_nodes = cv.take<unsigned long long>(1);
    d.objective = cv.take<double>(2);
    int *ts = cv.take<int>(T), *te = cv.take<int>(T);
    double *fs = cv.take<double>(4 * T + 8);
    {   // exact repair (bb_core.h)
        BBWork &b = d.bbw;
        const int64_t cap = cap_cand;
        b.cap = cap;
        b.comps = cv.take<bb::Comp>(T);
        b.comp_state = cv.take<int>(T);
        b.comp_slot = cv.take<int>(T);
        b.hdr = cv.take<int>(16);
        b.deadline = cv.take<unsigned long long>(2);
        b.c_cost = cv.take<double>(cap);
        b.c_tree = cv.take<int>(cap);
        b.c_gcol = cv.take<int>(cap);
        b.c_rows = cv.take<int>(cap * MHT_MAX_WINDOW);
        b.r_grow = cv.take<int>(cap * MHT_MAX_WINDOW);
        b.t_start = cv.take<int>(2 * T + 2);
        b.t_gtree = cv.take<int>(T);
        b.best_sel = cv.take<int>(T);
        b.ub_key = cv.take<unsigned long long>(T);
        b.lock = cv.take<int>(T);
        b.row_local = cv.take<int>(R);
        b.row_cnt = cv.take<int>(T);
        b.comp_unproven = cv.take<int>(T);
        b.comp_nodes = cv.take<int>(T);
        b.pool = cv.take<bb::Pool>(1);
        b.p_state = cv.take<int>(kBBMaxNodes);
        b.p_comp = cv.take<int>(kBBMaxNodes);
        b.p_bt = cv.take<int>(kBBMaxNodes);
        b.p_br = cv.take<int>(kBBMaxNodes);
        b.p_key = cv.take<double>(kBBMaxNodes);
        b.p_bound = cv.take<double>(kBBMaxNodes);
        b.p_ctr = cv.take<int>(8);
        b.max_cols = (int)(cap < kBBMaxCols ? cap : kBBMaxCols);
        const int64_t rows_bound = (int64_t)b.max_cols * MHT_MAX_WINDOW;
        b.max_rows = (int)(R < rows_bound ? R : rows_bound);
        b.max_trees = (int)T;
        b.scratch_per_worker = bb_scratch_bytes(b.max_cols, b.max_rows, b.max_trees);
        b.scratch_mem = cv.take<char>(b.scratch_per_worker * kBBWorkers);
        int64_t pool = cap * 2048;
        if (pool < (16ll << 20)) pool = 16ll << 20;
        if (pool > (1ll << 30)) pool = 1ll << 30;
        b.pool_bytes = pool;
        b.pool_mem = cv.take<char>(pool);
    }
    d.tstart = ts;
    if (w) *w = d;
    if (tstart) *tstart = ts;
    if (tend) *tend = te;
    if (fscratch) *fscratch = fs;
    return cv.used;
}

int64_t assoc_workspace_bytes(int64_t cap_cols, int64_t T, int64_t R, int64_t cap_cand) {
    Carver cv{nullptr};
    return carve_all(cv, cap_cols, T, R, cap_cand, nullptr, nullptr, nullptr, nullptr) + 256;
}

static thread_local int *g_tstart, *g_tend;
static thread_local double *g_fscratch;

void assoc_carve(void *d_work, int64_t cap_cols, int64_t T, int64_t R, int64_t cap_cand, AssocWork *w) {
    Carver cv{(char *)d_work};
    carve_all(cv, cap_cols, T, R, cap_cand, w, &g_tstart, &g_tend, &g_fscratch);
}

int assoc_begin(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s, bool warm) {
    (void)grid_dim;
    int n = (c.n_trees + 1 > c.n_rows ? c.n_trees + 1 : c.n_rows);
    if (n < kAssocInfo) n = kAssocInfo;  // the status words are cleared by the same kernel
    count_launch(), assoc_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(c, w, w.tstart, g_tend, warm ? 1 : 0);
    MHT_CUDA(cudaGetLastError());
    return MHT_OK;
}

static int cluster_phase(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s, bool warm = false,
                         bool pre_unioned = false) {
    const int T = c.n_trees;
    if (!pre_unioned)
        if (int rc = assoc_begin(c, w, grid_dim, s, warm)) return rc;
    count_launch(), assoc_init_cols_kernel<<<grid_dim, 256, 0, s>>>(c, w, g_tstart, g_tend);
    if (!pre_unioned) count_launch(), uf_union_cols_kernel<<<grid_dim, 256, 0, s>>>(c, w.uf, w.row_owner, w.row_mark);
    if (warm) count_launch(), warm_fix_kernel<<<(c.n_rows + 255) / 256, 256, 0, s>>>(c.n_rows, w.row_mark, w.u, w.best_u);
    count_launch(), uf_flatten_kernel<<<(T + 255) / 256, 256, 0, s>>>(T, w.uf);
    count_launch(), row_list_kernel<<<(c.n_rows + 255) / 256, 256, 0, s>>>(c.n_rows, w.row_owner, w.row_list, w.row_n);
    count_launch(), cluster_stats_kernel<<<1, 1024, 0, s>>>(T, w.uf, g_tstart, w.cl_nrm, w.info);
    MHT_CUDA(cudaGetLastError());
    return MHT_OK;
}

int assoc_cluster(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s) {
    return cluster_phase(c, w, grid_dim, s);
}

static void greedy_pass(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s) {
    count_launch(), greedy_init_kernel<<<1, 1024, 0, s>>>(c, w, g_tstart);
    for (int r = 0; r < kGreedyRounds; ++r) {
        count_launch(), greedy_prop_kernel<<<grid_dim, 256, 0, s>>>(c, w);
        count_launch(), greedy_arg_kernel<<<grid_dim, 256, 0, s>>>(c, w);
        count_launch(), greedy_commit_kernel<<<1, 1024, 0, s>>>(c, w);
    }
    count_launch(), greedy_finish_kernel<<<1, 1024, 0, s>>>(c, w, g_tstart);
}

static int persistent_grid() {
    static int blocks = 0;
    if (!blocks) {
        int per_sm = 0, dev = 0, sms = kSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_loop_persistent_kernel, 256, 0);
        if (per_sm > 4) per_sm = 4;
        blocks = per_sm > 0 ? per_sm * sms : sms;
    }
    return blocks;
}

// cluster configuration the device can run: CTAs per cluster (16 non-portable, else 8, else 0 = no cluster
// version) and the shared-memory slice capacity in columns
struct ClusterPlan {
    int ctas = 0, nc_cap = 0;
    size_t smem = 0;
};
static ClusterPlan cluster_plan(int W) {
    static ClusterPlan plans[MHT_MAX_WINDOW + 1];
    static bool done[MHT_MAX_WINDOW + 1] = {false};
    if (done[W]) return plans[W];
    done[W] = true;
    ClusterPlan best;
    if (getenv("MHT_NO_CLUSTER_LOOP")) return plans[W] = best;
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t budget = (size_t)optin - 2048;   // static shared memory of the greedy bodies + reserve
    int nc_cap = (int)((budget - 64 - (size_t)kClusterTreesPerCta * 16) / (8 + 8 + 4 + 4 * (size_t)W));
    nc_cap &= ~31;
    if (nc_cap < 256) return plans[W] = best;
    const size_t smem = cluster_slice_bytes(nc_cap, W);
    if (cudaFuncSetAttribute(dual_loop_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess ||
        cudaFuncSetAttribute(dual_loop_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return plans[W] = best;
    }
    for (int ctas : {kClusterCtas, 8}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(kClusterThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ctas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, dual_loop_cluster_kernel, &cfg) == cudaSuccess && n_clusters >= 1) {
            best.ctas = ctas;
            best.nc_cap = nc_cap;
            best.smem = smem;
            break;
        }
        cudaGetLastError();
    }
    return plans[W] = best;
}

static void print_loop_prof() {
    unsigned long long h[16];
    if (cudaMemcpyFromSymbol(h, g_loop_prof, sizeof(h)) != cudaSuccess) return;
    const double it = h[7] ? (double)h[7] : 1.0;
    fprintf(stderr, "[mht] cluster loop, %.0f iterations: us/iteration greedy %.2f min+argmin %.2f trees %.2f rows %.2f "
            "decide %.2f apply %.2f\n", it, h[0] / it * 1e-3, h[1] / it * 1e-3, h[3] / it * 1e-3, h[4] / it * 1e-3,
            h[5] / it * 1e-3, h[6] / it * 1e-3);
}

static int dual_loop(const ColView &c, AssocWork &w, int iters, int grid_dim, cudaStream_t s) {
    static bool prof_hook = false;
    if (!prof_hook) {
        prof_hook = true;
        if (getenv("MHT_LOOP_PROF")) atexit(print_loop_prof);
    }
    ColView cc = c;
    AssocWork ww = w;
    static int greedy_every = getenv("MHT_GREEDY_EVERY") ? atoi(getenv("MHT_GREEDY_EVERY")) : kGreedyEvery;
    int *declined = w.act_n + 3;      // 0 = the cluster version ran the loop
    const ClusterPlan plan = cluster_plan(c.width);
    if (plan.ctas) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(plan.ctas);
        cfg.blockDim = dim3(kClusterThreads);
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = plan.ctas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nc_cap = plan.nc_cap;
        count_launch();
        MHT_CUDA(cudaLaunchKernelEx(&cfg, dual_loop_cluster_kernel, cc, ww, iters, greedy_every, nc_cap, declined));
    } else {
        static const int one = 1;
        MHT_CUDA(cudaMemcpyAsync(declined, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    }
    void *args[] = {&cc, &ww, &iters, &greedy_every, &declined};
    // grid.sync cost grows with the grid: the active list (~1e5 columns) gets one CTA per SM
    const int grid = grid_dim < persistent_grid() ? grid_dim : persistent_grid();
    count_launch();
    MHT_CUDA(cudaLaunchCooperativeKernel((void *)dual_loop_persistent_kernel, dim3(grid), dim3(256), args,
                                         0, s));
    return MHT_OK;
}

// incidences (rows >= 0) of the columns the dual loop iterates on -> info[15]
__global__ void __launch_bounds__(256) nnz_kernel(ColView c, AssocWork w) {
    const int n = *c.n_ptr;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = c.idx ? c.idx[i] : i;
        for (int k = 0; k < c.width; ++k) cnt += c.rows[(long long)k * c.stride + j] >= 0 ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&w.info[15], cnt);
}

int assoc_solve(const ColView &c, AssocWork &w, int max_iters, int bb_budget, int grid_dim, cudaStream_t s,
                AssocEvents *ev, bool warm_start, bool sift, bool pre_unioned, double exact_ms,
                int exact_nodes) {
    if (int rc = cluster_phase(c, w, grid_dim, s, warm_start, pre_unioned)) return rc;
    if (ev && ev->after_cluster) MHT_CUDA(cudaEventRecord(ev->after_cluster, s));
    if (ev) ev->n_dual = 0;
    auto timed_loop = [&](const ColView &v, int iters, int g) -> int {
        const bool rec = ev && ev->n_dual < 4 && ev->dual[2 * ev->n_dual];
        if (rec) MHT_CUDA(cudaEventRecord(ev->dual[2 * ev->n_dual], s));
        if (int rc = dual_loop(v, w, iters, g, s)) return rc;
        if (rec) {
            MHT_CUDA(cudaEventRecord(ev->dual[2 * ev->n_dual + 1], s));
            ev->n_dual += 1;
        }
        return MHT_OK;
    };
    // iteration 0 settles every conflict-free cluster (all singletons) exactly
    count_launch(), dual_rc_kernel<false><<<grid_dim, 256, 0, s>>>(c, w);
    count_launch(), dual_arg_kernel<false><<<grid_dim, 256, 0, s>>>(c, w);
    dual_update(c, w, s);
    if (!sift) {
        // small problem: iterations cost ~25 us each, so give the bound 4x the budget (the loop leaves as soon
        // as every cluster is settled or nothing has moved for kStallStop iterations)
        count_launch(), nnz_kernel<<<grid_dim, 256, 0, s>>>(c, w);
        if (int rc = timed_loop(c, 4 * max_iters, grid_dim)) return rc;
    } else {
        // sifting: price all columns, iterate on the active list, re-price; kSiftRounds times
        ColView a = c;
        a.idx = w.act_col;
        a.n_ptr = w.act_n;
        static const int act_grid_env = getenv("MHT_ACT_GRID") ? atoi(getenv("MHT_ACT_GRID")) : 2 * kSMs;
        const int act_grid = act_grid_env;
        static const int sift_rounds = getenv("MHT_SIFT_ROUNDS") ? atoi(getenv("MHT_SIFT_ROUNDS")) : kSiftRounds;
        for (int round = 0; round < sift_rounds; ++round) {
            if (round) count_launch(), sift_rearm_kernel<<<1, 1024, 0, s>>>(c, w);
            count_launch(), dual_rc_kernel<true><<<grid_dim, 256, 0, s>>>(c, w);
            count_launch(), active_count_kernel<<<grid_dim, 256, 0, s>>>(c, w);
            count_launch(), active_scan_kernel<<<1, 1024, 0, s>>>(c, w);
            count_launch(), active_scatter_kernel<<<grid_dim, 256, 0, s>>>(c, w);
            count_launch(), reset_tree_min_kernel<<<(c.n_trees + 255) / 256, 256, 0, s>>>(c.n_trees, w);
            if (round == sift_rounds - 1) count_launch(), nnz_kernel<<<act_grid, 256, 0, s>>>(a, w);
            if (int rc = timed_loop(a, max_iters, act_grid)) return rc;
        }
    }
    MHT_CUDA(cudaGetLastError());
    // final multipliers -> reduced costs, bound, candidates, exact repair
    count_launch(), final_prepare_kernel<<<1, 1024, 0, s>>>(c, w);
    count_launch(), dual_rc_kernel<true><<<grid_dim, 256, 0, s>>>(c, w);
    count_launch(), dual_arg_kernel<true><<<grid_dim, 256, 0, s>>>(c, w);
    {   // parallel local search on the incumbent (see ls_propose_kernel)
        static const int ls_rounds = getenv("MHT_LS_ROUNDS") ? atoi(getenv("MHT_LS_ROUNDS")) : 10;
        const int tb = (c.n_trees + 127) / 128, wb = (c.n_trees + 7) / 8;
        if (ls_rounds > 0) {
            MHT_CUDA(cudaMemsetAsync(w.ls_ctr, 0, 4 * sizeof(int), s));
            MHT_CUDA(cudaMemsetAsync(w.row_bid, 0xff, sizeof(unsigned long long) * (size_t)c.n_rows, s));
            count_launch(), ls_begin_kernel<<<tb, 128, 0, s>>>(c, w, g_tstart);
            count_launch(), ls_shortlist_kernel<<<c.n_trees < 8 * kSMs ? c.n_trees : 8 * kSMs, 256, 0, s>>>(c, w, g_tstart, g_tend);
            for (int round = 0; round < ls_rounds; ++round) {
                count_launch(), ls_propose_kernel<<<wb, 256, 0, s>>>(c, w, g_tstart);
                count_launch(), ls_apply_kernel<<<tb, 128, 0, s>>>(c, w);
                count_launch(), ls_clear_kernel<<<tb, 128, 0, s>>>(c, w);
                count_launch(), ls_round_end_kernel<<<1, 1, 0, s>>>(w);
            }
            // the candidate-based search below keeps its own holder table
            MHT_CUDA(cudaMemsetAsync(w.row_holder, 0xff, sizeof(int) * (size_t)c.n_rows, s));
        }
    }
    count_launch(), final_bound_kernel<<<1, 1024, 0, s>>>(c, w, g_tstart);
    count_launch(), cand_count_kernel<<<grid_dim, 256, 0, s>>>(c, w);
    count_launch(), cand_scan_kernel<<<1, 1024, 0, s>>>(c, w);
    count_launch(), cand_fill_kernel<<<grid_dim, 256, 0, s>>>(c, w);
    {
        const int tb = (c.n_trees + 127) / 128, rb = (c.n_rows + 255) / 256 < 1024 ? (c.n_rows + 255) / 256 : 1024;
        const int wb = (c.n_trees + 7) / 8;
        count_launch(), cand_sort_kernel<<<tb, 128, 0, s>>>(c, w);
        for (int round = 0; round < 2; ++round) {
            count_launch(), contest_reset_kernel<<<rb, 256, 0, s>>>(w);
            count_launch(), contest_mark_kernel<<<wb, 256, 0, s>>>(c, w, nullptr);
            count_launch(), cand_dominance_kernel<<<wb, 256, 0, s>>>(c, w);
        }
        count_launch(), contest_reset_kernel<<<rb, 256, 0, s>>>(w);
        count_launch(), contest_mark_kernel<<<wb, 256, 0, s>>>(c, w, w.comp_uf);
    }
    count_launch(), comp_build_kernel<<<1, 1024, 0, s>>>(c, w);
    count_launch(), branch_bound_kernel<<<kSMs * 4, 32, 0, s>>>(c, w, bb_budget, g_fscratch);
    {   // exact repair of what is still open (bb_core.h)
        static const int k_root = getenv("MHT_BB_KROOT") ? atoi(getenv("MHT_BB_KROOT")) : 100;
        static const int k_node = getenv("MHT_BB_KNODE") ? atoi(getenv("MHT_BB_KNODE")) : 30;
        static const double env_ms = getenv("MHT_BB_MS") ? atof(getenv("MHT_BB_MS")) : -1.0;
        static const int sb_cands = getenv("MHT_BB_SB") ? atoi(getenv("MHT_BB_SB")) : 0;      // strong-branching probes
        static const int sb_iters = getenv("MHT_BB_SB_ITERS") ? atoi(getenv("MHT_BB_SB_ITERS")) : 15;
        const double ms = env_ms >= 0.0 ? env_ms : exact_ms;
        if (ev && ev->exact_begin) MHT_CUDA(cudaEventRecord(ev->exact_begin, s));
        // a dual iteration on n columns costs ~n / 4e9 s in one CTA: components the time box cannot even evaluate
        // a few hundred times are left alone (they stay open, the scan is reported uncertified)
        const double cols_now = ms * 4000.0;
        static const double gap_per_ms = getenv("MHT_BB_GAP_PER_MS") ? atof(getenv("MHT_BB_GAP_PER_MS")) : 0.125;
        count_launch(), bb_plan_kernel<<<1, 1024, 0, s>>>(c, w, cols_now > 1e9 ? 1000000000 : (int)cols_now, 1.0 + gap_per_ms * ms);
        if (ms > 0.0) {
            count_launch(), bb_compact_kernel<<<kSMs, 256, 0, s>>>(c, w);
            count_launch(), bb_root_kernel<<<kSMs, 256, 0, s>>>(c, w, ms);
            count_launch(), bb_open_roots_kernel<<<4, 256, 0, s>>>(w);
            static int smem_bytes = 0;
            if (!smem_bytes) {
                int dev = 0, optin = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
                smem_bytes = optin - 4096;   // static shared memory of the kernel + reserve
                if (smem_bytes < 32768) smem_bytes = 32768;
                MHT_CUDA(cudaFuncSetAttribute(bb_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            }
            count_launch(), bb_search_kernel<<<kBBWorkers, kBBThreads, smem_bytes, s>>>(w, k_root, k_node, exact_nodes, sb_cands,
                                                                                        sb_iters);
            count_launch(), bb_writeback_kernel<<<kSMs, 256, 0, s>>>(c, w);
            count_launch(), bb_state_kernel<<<4, 256, 0, s>>>(w);
        }
        if (ev && ev->exact_end) MHT_CUDA(cudaEventRecord(ev->exact_end, s));
    }
    count_launch(), feas_reset_kernel<<<(c.n_rows + 255) / 256, 256, 0, s>>>(c.n_rows, w.row_holder);
    count_launch(), feas_claim_kernel<<<(c.n_trees + 255) / 256, 256, 0, s>>>(c, w, g_tstart, w.row_holder);
    count_launch(), feas_repair_kernel<<<(c.n_trees + 255) / 256, 256, 0, s>>>(c, w, g_tstart, w.row_holder);
    count_launch(), final_objective_kernel<<<1, 1024, 0, s>>>(c, w, g_tstart);
    MHT_CUDA(cudaGetLastError());
    return MHT_OK;
}

}  // namespace mht

using namespace mht;

extern "C" int64_t mht_assoc_workspace(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width) {
    (void)width;
    return assoc_workspace_bytes(n_cols, n_trees, n_rows, n_cols) + 256;
}

static int make_view(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width, const double *d_cost,
                     const int32_t *d_tree, const int32_t *d_rows, void *d_work, ColView *c, AssocWork *w,
                     cudaStream_t s, int64_t cap_cols = -1) {
    if (cap_cols < 0) cap_cols = n_cols;
    if (cap_cols < n_cols) {
        set_error("assoc: n_cols=%lld exceeds the column capacity %lld", (long long)n_cols, (long long)cap_cols);
        return MHT_E_INVALID;
    }
    if (n_cols < 0 || n_cols > 0x7ffffff0ll || n_trees <= 0 || n_trees >= (1 << 24) || n_rows < 0 ||
        n_rows > 0x7ffffff0ll || width < 0 || width > MHT_MAX_WINDOW || !d_work) {
        set_error("assoc: invalid argument (n_cols=%lld n_trees=%lld n_rows=%lld width=%d)", (long long)n_cols,
                  (long long)n_trees, (long long)n_rows, width);
        return MHT_E_INVALID;
    }
    int *n_dev = (int *)d_work;
    const int n32 = (int)n_cols;
    MHT_CUDA(cudaMemcpyAsync(n_dev, &n32, sizeof(int), cudaMemcpyHostToDevice, s));
    assoc_carve((char *)d_work + 256, cap_cols, n_trees, n_rows, cap_cols, w);
    c->n_ptr = n_dev;
    c->idx = nullptr;
    c->meas = nullptr;
    c->plane_new = -1;
    c->cost = d_cost;
    c->tree_base = nullptr;
    c->tree = d_tree;
    c->rows = d_rows;
    c->stride = cap_cols;
    c->width = width;
    c->n_trees = (int)n_trees;
    c->n_rows = (int)(n_rows ? n_rows : 1);
    return MHT_OK;
}

extern "C" int mht_cluster(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width, const int32_t *d_col_tree,
                           const int32_t *d_col_rows, int32_t *d_cluster_of_tree, void *d_work, void *stream) {
    if (int rc = check_device()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    ColView c;
    AssocWork w;
    if (int rc = make_view(n_cols, n_trees, n_rows, width, nullptr, d_col_tree, d_col_rows, d_work, &c, &w, s))
        return rc;
    if (int rc = assoc_cluster(c, w, kSMs * 4, s)) return rc;
    MHT_CUDA(cudaMemcpyAsync(d_cluster_of_tree, w.uf, n_trees * sizeof(int), cudaMemcpyDeviceToDevice, s));
    return MHT_OK;
}

static int assoc_solve_entry(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width, const double *d_col_cost,
                             const int32_t *d_col_tree, const int32_t *d_col_rows, int32_t *d_selected_col, double *h_info,
                             void *d_work, void *stream, bool warm, int64_t clear_lo, int64_t clear_hi,
                             int64_t cap_cols = -1, double exact_ms = 10000.0, int max_iters = 200) {
    if (int rc = check_device()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    ColView c;
    AssocWork w;
    if (int rc = make_view(n_cols, n_trees, n_rows, width, d_col_cost, d_col_tree, d_col_rows, d_work, &c, &w, s, cap_cols))
        return rc;
    if (warm && clear_hi > clear_lo && clear_lo >= 0 && clear_hi <= n_rows)
        MHT_CUDA(cudaMemsetAsync(w.u + clear_lo, 0, sizeof(double) * (size_t)(clear_hi - clear_lo), s));
    // sifting (iterate on an active column list) from 10^6 columns on, like the forest
    if (int rc = assoc_solve(c, w, max_iters, 4096, kSMs * 8, s, nullptr, warm, n_cols > 1000000, false, exact_ms)) return rc;
    MHT_CUDA(cudaMemcpyAsync(d_selected_col, w.sel, n_trees * sizeof(int), cudaMemcpyDeviceToDevice, s));
    int info[kAssocInfo];
    double obj[2];
    unsigned long long nodes;
    MHT_CUDA(cudaMemcpyAsync(info, w.info, sizeof(info), cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(obj, w.objective, sizeof(obj), cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(&nodes, w.bb_nodes, sizeof(nodes), cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    if (h_info) {
        h_info[0] = obj[0];
        h_info[1] = obj[1];
        h_info[2] = info[3];
        h_info[3] = info[4];
        h_info[4] = (double)nodes;
        h_info[5] = info[1];
        h_info[6] = info[10];
        h_info[7] = info[9];
    }
    if (!info[10]) {
        set_error("mht_assoc_solve: optimality not certified (search budget or candidate capacity exhausted)");
        return MHT_E_NOTOPTIMAL;
    }
    return MHT_OK;
}

extern "C" int mht_assoc_solve(int64_t n_cols, int64_t n_trees, int64_t n_rows, int32_t width,
                               const double *d_col_cost, const int32_t *d_col_tree, const int32_t *d_col_rows,
                               int32_t *d_selected_col, double *h_info, void *d_work, void *stream) {
    return assoc_solve_entry(n_cols, n_trees, n_rows, width, d_col_cost, d_col_tree, d_col_rows, d_selected_col, h_info,
                             d_work, stream, false, 0, 0);
}

extern "C" int mht_assoc_solve_warm(int64_t n_cols, int64_t cap_cols, int64_t n_trees, int64_t n_rows, int32_t width,
                                    const double *d_col_cost, const int32_t *d_col_tree, const int32_t *d_col_rows,
                                    int32_t *d_selected_col, double *h_info, void *d_work, void *stream, int32_t warm,
                                    int64_t clear_row_lo, int64_t clear_row_hi, double exact_ms, int32_t max_dual_iters) {
    return assoc_solve_entry(n_cols, n_trees, n_rows, width, d_col_cost, d_col_tree, d_col_rows, d_selected_col, h_info,
                             d_work, stream, warm != 0, clear_row_lo, clear_row_hi, cap_cols,
                             exact_ms > 0.0 ? exact_ms : 10000.0, max_dual_iters > 0 ? max_dual_iters : 200);
}
