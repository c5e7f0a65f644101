// Host build of pymht_b200/csrc/bb_core.h (the exact-repair branch & bound of libmht_b200) with a one-thread
// execution context: test infrastructure for tests/test_bb_core_host.py, which checks it against HiGHS on the
// cluster problems of the reference fixtures.  g++ -O2 -shared -fPIC
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../pymht_b200/csrc/bb_core.h"

struct HostCtx {
    int tid() const { return 0; }
    int nthr() const { return 1; }
    void sync() {}
    void amin64(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
    void amax(int *p, int v) { if (v > *p) *p = v; }
    void aadd(int *p, int v) { *p += v; }
    double sum(double v) { return v; }
    long long maxll(long long v) { return v; }
    unsigned long long bcast(unsigned long long v) { return v; }
    void reduce(double &, double &, long long &, unsigned long long &) {}
    int acas(int *p, int cmp, int val) { const int old = *p; if (old == cmp) *p = val; return old; }
    void fence() {}
    void backoff() {}
    bool expired() { return false; }
    bool bind(const bb::Comp &, bb::Scratch &) { return true; }
};

// columns sorted by tree; rows[W][nC]; sel0 = incumbent (local column per tree) or null (= all-miss columns unknown:
// then the first column of every tree must be row-free).  Returns 1 when the optimum is proven.
extern "C" int bb_solve_host(int nC, int nT, int nR, int W, const double *cost, const int *tree, const int *rows,
                             const double *u0, const int *sel0, int K_root, int K_node, int max_nodes, int pool_cap,
                             int *best_sel, double *best, int *nodes, int *iters, int sb_cands, int sb_iters,
                             double *open_bound) {
    std::vector<int> tstart(nT + 1, 0);
    for (int j = 0; j < nC; ++j) tstart[tree[j] + 1] = j + 1;
    for (int t = 0; t < nT; ++t) if (tstart[t + 1] < tstart[t]) tstart[t + 1] = tstart[t];
    unsigned long long ub_key;
    int lock = 0;
    std::vector<int> bsel(nT);
    double ub = 0.0;
    for (int t = 0; t < nT; ++t) {
        bsel[t] = sel0 ? sel0[t] : tstart[t];
        ub += cost[bsel[t]];
    }
    ub_key = bb::key_of(ub);
    bb::Comp p;
    p.nC = nC; p.nT = nT; p.nR = nR; p.W = W; p.row_stride = nC;
    p.cost = cost; p.tree = tree; p.rows = rows; p.tstart = tstart.data();
    p.nwords = (nC + 31) / 32;
    p.ub_key = &ub_key; p.best_sel = bsel.data(); p.lock = &lock;
    std::vector<double> su(nR), rc(nC), ubest(nR), cand_d(nT);
    std::vector<int> usage(nR), targ(nT), freq(nC), best_targ(nT), cand_r(nT);
    std::vector<unsigned long long> tmin(nT);
    std::vector<unsigned> alive(p.nwords), alive2(p.nwords);
    bb::Scratch s;
    s.u = su.data(); s.usage = usage.data(); s.tmin = tmin.data(); s.targ = targ.data(); s.alive = alive.data();
    s.rc = rc.data(); s.freq = freq.data(); s.ubest = ubest.data(); s.best_targ = best_targ.data();
    s.cand_d = cand_d.data(); s.cand_r = cand_r.data(); s.alive2 = alive2.data();
    bb::Pool pl;
    pl.cap = pool_cap; pl.node_words = p.nwords; pl.node_rows = nR > 0 ? nR : 1;
    std::vector<int> state(pool_cap, 0), comp(pool_cap), bt(pool_cap), br(pool_cap);
    std::vector<double> key(pool_cap), bound(pool_cap);
    std::vector<unsigned> palive((size_t)pool_cap * pl.node_words);
    std::vector<float> pu((size_t)pool_cap * pl.node_rows);
    int outstanding = 1, stop = 0, n_nodes = 0, n_iters = 0, unproven = 0, cnodes = 0;
    pl.state = state.data(); pl.key = key.data(); pl.bound = bound.data(); pl.comp = comp.data();
    pl.bt = bt.data(); pl.br = br.data(); pl.alive = palive.data(); pl.u = pu.data();
    pl.outstanding = &outstanding; pl.stop = &stop; pl.nodes = &n_nodes; pl.iters = &n_iters;
    pl.comp_unproven = &unproven; pl.comp_nodes = &cnodes;
    // root node: every column alive, caller's multipliers
    state[0] = 1; comp[0] = 0; bt[0] = -1; br[0] = -1; bound[0] = -1e300; key[0] = -1e300;
    for (int w = 0; w < p.nwords; ++w) {
        const int left = nC - 32 * w;
        palive[w] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
    }
    for (int r = 0; r < nR; ++r) pu[r] = u0 ? (float)u0[r] : 0.0f;
    HostCtx c;
    bb::worker(c, &p, pl, s, K_root, K_node, max_nodes, sb_cands, sb_iters);
    if (open_bound) {
        double lo = 1e300;
        for (int i = 0; i < pool_cap; ++i)
            if (state[i] != 0 && bound[i] < lo) lo = bound[i];
        *open_bound = lo;
    }
    for (int t = 0; t < nT; ++t) best_sel[t] = bsel[t];
    *best = bb::of_key(ub_key);
    *nodes = n_nodes;
    *iters = n_iters;
    return (!unproven && !stop && outstanding == 0) ? 1 : 0;
}
