// Host instantiation of pymht_b200/csrc/gnn_core.h with a one-thread execution context (tests/test_gnn_core_host.py).
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#include "../../pymht_b200/csrc/gnn_core.h"

using namespace mht::gnn;

struct HostCtx {
    int tid() const { return 0; }
    int nthr() const { return 1; }
    void sync() {}
    int ld(const int *p) const { return *p; }
    unsigned long long ld64(const unsigned long long *p) const { return *p; }
    unsigned long long amin64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
    int amin32(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
    int aadd(int *p, int v) { int o = *p; *p = o + v; return o; }
    int aexch(int *p, int v) { int o = *p; *p = v; return o; }
};

struct HostW {
    int lane() const { return 0; }
    int nlanes() const { return 1; }
    void wsync() {}
    unsigned long long wmin64(unsigned long long v) { return v; }
    int wmin32(int v) { return v; }
    int wall(int v) { return v; }
    int aadd(int *p, int v) { int o = *p; *p = o + v; return o; }
    int acas(int *p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
    unsigned long long amin64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
    unsigned long long ld64(const unsigned long long *p) const { return *p; }
};

static int find(std::vector<int> &uf, int x) {
    while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; }
    return x;
}

// CSR graph in, match_col[n_rows] out (-1 = unassigned).  stats: [searches, rounds, components, largest component rows]
// n_warps > 0: the speculative parallel phase first (simulated: the searches of a batch run one after the other against the
// same frozen state, exactly what concurrent warps see), the block-wide searches only for what it leaves.
// stats: [searches, rounds, components, largest component rows, batches, speculative commits, rows left to search()]
extern "C" int gnn_solve_host(int n_rows, int n_cols, const int *row_ptr, const int *col, const double *cost,
                              int *match_out, long long *stats, int n_warps) {
    Graph g{n_rows, n_cols, row_ptr, col, cost};
    std::vector<double> u(n_rows, 0.0), v(n_cols, 0.0), drow(n_rows, from_bits(kInfBits));
    std::vector<int> mc(n_rows, -1), mr(n_cols, -1), pred(n_cols, kNoPred), mark(n_cols, 0);
    std::vector<unsigned long long> dcol(n_cols, kInfBits);
    std::vector<int> la(n_rows + 1), lb(n_rows + 1), tr(n_rows + 1), tc(n_cols + 1), ch(n_cols + 1);
    State st{u.data(), v.data(), mc.data(), mr.data(), dcol.data(), drow.data(), pred.data(), mark.data(),
             la.data(), lb.data(), tr.data(), tc.data(), ch.data()};
    std::vector<int> uf(n_rows + n_cols);
    for (int i = 0; i < n_rows + n_cols; ++i) uf[i] = i;
    double cmax = 0.0;
    for (int i = 0; i < n_rows; ++i)
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
            int a = find(uf, i), b = find(uf, n_rows + col[e]);
            if (a != b) uf[std::max(a, b)] = std::min(a, b);
            cmax = std::max(cmax, cost[e]);
        }
    const double BIG = (double)(std::min(n_rows, n_cols) + 1) * (cmax + 1.0);
    std::vector<std::vector<int>> comp_rows(n_rows);
    std::vector<int> comp_cols(n_rows, 0);
    for (int i = 0; i < n_rows; ++i) comp_rows[find(uf, i)].push_back(i);
    for (int j = 0; j < n_cols; ++j) { int r = find(uf, n_rows + j); if (r < n_rows) comp_cols[r] += 1; }
    HostCtx c;
    Shared sh{};
    int off_r = 0, off_c = 0;
    for (int k = 0; k < 7; ++k) stats[k] = 0;
    if (n_warps > 0) {
        start_pass(c, 0, g, st, 0, 1, BIG);
        start_pass(c, 1, g, st, 0, 1, BIG);
        start_pass(c, 2, g, st, 0, 1, BIG);
        std::vector<unsigned long long> tr_(n_rows, ~0ull), mr_(n_rows, ~0ull), tc_(n_cols, ~0ull), mc_(n_cols, ~0ull);
        Claims cl{tr_.data(), mr_.data(), tc_.data(), mc_.data()};
        std::vector<Spec> sp(n_warps);
        std::vector<int> cursor(n_warps), cur_row(n_warps, -1);
        std::vector<char> hard(n_rows, 0);
        for (int w = 0; w < n_warps; ++w) cursor[w] = w;
        HostW hw;
        for (unsigned epoch = 1;; ++epoch) {
            int active = 0;
            for (int w = 0; w < n_warps; ++w) {
                cur_row[w] = -1;
                while (cursor[w] < n_rows) {
                    const int i = cursor[w];
                    if (mc[i] == -1 && row_ptr[i + 1] > row_ptr[i] && !hard[i]) break;
                    cursor[w] += n_warps;
                }
                if (cursor[w] >= n_rows) continue;
                cur_row[w] = cursor[w];
                ++active;
                spec_search(hw, g, st, &sp[w], cur_row[w], BIG, getenv("GNN_ROWCAP") ? atoi(getenv("GNN_ROWCAP")) : kSpecRows);
                if (sp[w].overflow) {
                    hard[cur_row[w]] = 1;
                    cur_row[w] = -1;
                    stats[6] += 1;
                    continue;
                }
                spec_claim(hw, &sp[w], cl, epoch);
            }
            if (!active) break;
            stats[4] += 1;
            { int mx = 0; for (int w = 0; w < n_warps; ++w) if (cur_row[w] >= 0) mx = std::max(mx, sp[w].nR); if (getenv("GNN_TRACE")) fprintf(stderr, "batch %lld active %d maxrows %d\n", stats[4], active, mx); }
            std::vector<int> win;
            for (int w = 0; w < n_warps; ++w)
                if (cur_row[w] >= 0 && spec_check(hw, &sp[w], cl, epoch)) win.push_back(w);
            for (int w : win) spec_commit(hw, g, st, &sp[w]);
            stats[5] += (long long)win.size();
        }
    }
    for (int l = 0; l < n_rows; ++l) {
        if (comp_rows[l].empty()) continue;
        solve_component(c, g, st, &sh, comp_rows[l].data(), (int)comp_rows[l].size(), off_r, off_c, BIG, n_warps > 0);
        off_r += (int)comp_rows[l].size();
        off_c += comp_cols[l];
        stats[0] += sh.searches; stats[1] += sh.rounds; stats[2] += 1;
        stats[3] = std::max<long long>(stats[3], (long long)comp_rows[l].size());
    }
    for (int i = 0; i < n_rows; ++i) match_out[i] = mc[i] >= 0 ? mc[i] : -1;
    return 0;
}
