// Host build of pymht_b200/csrc/experimental/lbb_core.h (single thread) for scripts/proto/lbb_host_check.py.
//   g++ -O2 -shared -fPIC -o /tmp/liblbb_host.so scripts/proto/lbb_host.cpp
#include <vector>
#include "../../pymht_b200/csrc/experimental/lbb_core.h"

struct HostCtx {
    int tid() const { return 0; }
    int nthr() const { return 1; }
    void sync() {}
    void amin(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
    void amax(int *p, int v) { if (v > *p) *p = v; }
    void aadd(int *p, int v) { *p += v; }
};

extern "C" int lbb_solve_host(int nC, int nT, int nR, int W, const double *cost, const int *tree, const int *rows,
                              const double *u0, double ub0, const int *sel0, int iters_root, int iters_node,
                              int max_nodes, int max_depth, int *best_sel, double *best, int *nodes) {
    lbb::Problem p{nC, nT, nR, W, cost, tree, rows, u0, ub0, sel0, iters_root, iters_node, max_nodes, max_depth};
    std::vector<double> u(nR), ust((size_t)(max_depth + 1) * nR), red(16);
    std::vector<int> usage(nR), rowtaken(nR), fixed(nT), targ(nT), fj(max_depth + 1), ft(max_depth + 1), fs(max_depth + 1);
    std::vector<unsigned> forbid((nC + 31) / 32);
    std::vector<unsigned long long> tmin(nT);
    std::vector<long long> redi(16);
    lbb::Work w{u.data(), ust.data(), usage.data(), rowtaken.data(), fixed.data(), forbid.data(), tmin.data(), targ.data(),
                best_sel, fj.data(), ft.data(), fs.data(), red.data(), redi.data(), 0.0, 0, 0};
    HostCtx c;
    lbb::Solver<HostCtx> s(p, w, c);
    s.run();
    *best = w.best;
    *nodes = w.nodes;
    return w.proven;
}
