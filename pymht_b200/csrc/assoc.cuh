// Association stage (cluster + global-hypothesis 0/1 program) shared between the stateless
// operators and the forest.  See assoc.cu for the algorithm.
#pragma once
#include "common.cuh"
#include "bb_core.h"

namespace mht {

constexpr int kAssocInfo = 16;  // ints in AssocWork::info
constexpr int kLsK = 34;        // shortlist: 32 per-lane minima + incumbent + all-miss leaf

// Column view: one column = one leaf hypothesis.  Columns of a tree are contiguous.
struct ColView {
    const int *n_ptr;          // device scalar: number of columns (entries of idx when idx != null)
    const int *idx;            // null, or ascending list of ACTIVE column indices (sifting)
    const double *cost;        // [n]   (cumulative NLLR for the forest)
    const double *tree_base;   // [T] or null: cost_j := cost[j] - tree_base[tree[j]]
    const int *tree;           // [n] non-decreasing
    const int *rows;           // [width][stride]  row id or <0
    const int *meas;           // null, or measurementNumber per column: 0 marks the first of a sibling run
                               // whose columns share every row plane except plane_new
    int plane_new;
    long long stride;
    int width;
    int n_trees;
    int n_rows;
};

// Exact repair (bb_core.h): compacted cores of the open components, node pool, worker scratch.
constexpr int kBBMaxNodes = 1 << 16;     // node slots (the pool memory may allow fewer)
constexpr int kBBMaxCols = 1 << 18;      // columns of one component the search accepts
constexpr int kBBWorkers = kSMs;         // one CTA per SM
constexpr int kBBThreads = 512;
constexpr int kDfsMaxTrees = 16;         // components up to this size go to the depth-first enumeration first

struct BBWork {
    bb::Comp *comps;            // [T]
    int *comp_state;            // [T] per component slot: 1 = solved exactly
    int *comp_slot;             // [T] search component -> component slot
    int *hdr;                   // [16]: 0 n search comps, 1 max nC, 2 max nR, 3 max nT, 4 pool cap, 5 total cols
    unsigned long long *deadline;   // [2] globaltimer deadline (ns), start
    double *c_cost;             // [cap]
    int *c_tree, *c_gcol;       // [cap]
    int *c_rows;                // [W][cap]
    int *r_grow;                // [W * cap] global row of each local row, per component at W * col_off
    int *t_start;               // [2T + 2]
    int *t_gtree;               // [T]
    int *best_sel;              // [T]
    unsigned long long *ub_key; // [T]
    int *lock;                  // [T]
    int *row_local;             // [R] global row -> local row of its component
    int *row_cnt;               // [T]
    int *comp_unproven, *comp_nodes;   // [T]
    bb::Pool *pool;             // device copy of the pool descriptor
    int *p_state, *p_comp, *p_bt, *p_br;   // [kBBMaxNodes]
    double *p_key, *p_bound;    // [kBBMaxNodes]
    int *p_ctr;                 // [8]: outstanding, stop, nodes, iters
    char *pool_mem;
    long long pool_bytes;
    char *scratch_mem;
    long long scratch_per_worker;
    long long cap;              // = cap_cand
    int max_cols, max_rows, max_trees;   // dimensions the worker scratch is sized for
};

struct AssocWork {
    // per column
    double *rc;                // [cap_cols]
    int *freq;                 // [cap_cols] how often the column was its tree's argmin (primal rounding)
    // per tree
    unsigned long long *tmin;  // ordered key of min reduced cost
    int *targ;                 // argmin column (ties -> last)
    int *uf;                   // union-find parent -> cluster label (smallest tree index)
    int *sel;                  // incumbent column per tree
    int *sel_new;              // greedy scratch
    int *committed;
    unsigned long long *prop_key;
    int *prop_col;
    long long *cl_m, *cl_u, *cl_cost;   // fixed-point per-cluster sums (indexed by label)
    int *cl_nrm;
    double *cl_best, *cl_ub, *cl_theta, *cl_step;
    int *cl_stall, *cl_done, *cl_flag;
    int *tdone;                // [T] cl_done of the tree's cluster (one load, not a uf -> cl_done chain)
    int *stall_ctr;            // [4] iterations without any bound improvement
    int *cand_cnt, *cand_off, *cand_fill;   // [T+1]
    int *comp_uf, *comp_trees, *comp_off, *comp_cnt;  // candidate components
    // per row
    int *row_owner;
    double *u, *best_u;
    int *usage;
    unsigned long long *row_bid;
    int *row_taken;
    int *row_mark;
    int *row_cont;             // [R] 1 = candidates of two or more trees use the row
    int *row_holder;           // [R] local search: tree currently using the row (-1 free)
    int *ls_short;             // [T][kLsK] shortlist of columns per tree (parallel local search)
    double *ls_delta;          // [T] proposed move: cost change ...
    int *ls_j, *ls_o, *ls_jo;  // [T] ... own new column, displaced tree (-1 none) and its new column
    unsigned long long *ls_tbid;   // [T] bids on trees
    int *ls_ctr;               // [4] moves applied this round / total, stop flag
    int *row_list;             // [R] rows touched by at least one column (order irrelevant)
    int *row_n;                // device: entries of row_list
    int *tstart;               // [T] first column of each tree (-1: the tree has no columns)
    // candidates
    int *act_col;              // [cap_act] active column list (sifting)
    int *act_tile;             // [4][cap_cols/256+2] per-tile counts for 4 thresholds
    int *act_n;                // [4] device: list length, chosen threshold index
    long long cap_act;
    int *cand_col;             // [cap_cand]
    int *cand_stack;           // [T] DFS cursors, [T] order etc. carved by the kernel
    long long cap_cand;
    // device status words
    int *info;                 // [kAssocInfo]: 0 all_done, 1 iters, 2 greedy_left, 3 n_cand, 4 n_comp,
                               // 5 open components, 6 cand_overflow, 7 n_clusters, 8 n_multi, 9 max_comp, 10 certified,
                               // 11 trees moved to their miss column (no primal solution, or lost a row in the final feasibility check), 12 n_active, 13 candidates before dominance,
                               // 14 iterations inside the exact search, 15 nnz of the columns the dual loop iterates on
    unsigned long long *bb_nodes;
    double *objective;         // [2]: lower bound, objective
    BBWork bbw;
};

// ---- lock-free union-find shared with the forest's emit kernel (label = smallest tree index) ----
__device__ __forceinline__ int uf_find(int *uf, int x) {
    int p = ((volatile int *)uf)[x];
    while (p != x) {
        const int gp = ((volatile int *)uf)[p];
        if (gp != p) uf[x] = gp;  // path halving (benign race)
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int *uf, int a, int b) {
    while (true) {
        a = uf_find(uf, a);
        b = uf_find(uf, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        if (atomicCAS(&uf[b], b, a) == b) return;
    }
}
// tree t uses measurement row r: first toucher owns the row, later trees are united with the owner.
// Fast path: equal (possibly stale, L1-cached) parent pointers prove t and o are already in one set --
// parents only ever move to other members of the same set -- so the volatile find chains are skipped.
__device__ __forceinline__ void uf_touch_row(int *uf, int *row_owner, int *row_multi, int r, int t) {
    int o = row_owner[r];
    if (o < 0) {
        o = atomicCAS(&row_owner[r], -1, t);
        if (o < 0) o = t;
    }
    if (o != t) {
        if (row_multi[r] == 0) row_multi[r] = 1;
        if (uf[t] != uf[o]) uf_union(uf, t, o);
    }
}

int64_t assoc_workspace_bytes(int64_t cap_cols, int64_t n_trees, int64_t n_rows, int64_t cap_cand);
void assoc_carve(void *d_work, int64_t cap_cols, int64_t n_trees, int64_t n_rows, int64_t cap_cand, AssocWork *w);
// per-solve reset of the tree / row state (everything that does not need the column count); a caller that
// unites trees itself (forest emit kernel) calls this first and passes pre_unioned = true to assoc_solve
int assoc_begin(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s, bool warm_start);
// cluster labels only (uf[t] = smallest tree index of the component); async
int assoc_cluster(const ColView &c, AssocWork &w, int grid_dim, cudaStream_t s);
// full solve; async; results in w.sel / w.info / w.objective
// optional CUDA events for the stage timers of mht_scan_info
struct AssocEvents {
    cudaEvent_t after_cluster = nullptr;
    cudaEvent_t dual[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // begin/end per loop launch
    int n_dual = 0;                       // out: event pairs recorded
    cudaEvent_t exact_begin = nullptr, exact_end = nullptr;
};
// exact_ms / exact_nodes: wall-clock and node budget of the exact repair (branch & bound) of this solve
int assoc_solve(const ColView &c, AssocWork &w, int max_iters, int bb_budget, int grid_dim, cudaStream_t s,
                AssocEvents *ev = nullptr, bool warm_start = false, bool sift = false,
                bool pre_unioned = false, double exact_ms = 50.0, int exact_nodes = 1 << 30);

}  // namespace mht
