// Global nearest neighbour assignment of the M-of-N initiator (reference pymht/initiators/m_of_n.py:24-104) as a SPARSE
// problem: the reference pads the gated distance matrix to a square one (invalid pairs cost bigM, padding costs 10 x the
// largest valid distance) and hands it to a dense O(n^3) Munkres solver.  With bigM above the sum of all valid costs the
// optimum of that padded problem is: a matching of MAXIMUM CARDINALITY on the gated pairs and, among those, one of
// minimum total distance.  This core computes exactly that on the gated graph alone (CSR by row), one connected component
// per thread block:
//   * every row gets a private "stay unassigned" column of cost BIG (> any achievable sum of real costs), which turns the
//     lexicographic objective into one minimum-cost assignment of all rows;
//   * successive shortest augmenting paths with node potentials (u rows, v columns; reduced cost c - u - v >= 0, = 0 on
//     matched pairs, v = 0 on free columns).  Each search is a label-correcting shortest-path sweep run by ALL threads of
//     the block over the frontier (synchronous rounds), pruned by the best end point found so far; rows are searched in
//     ascending order, so the result does not depend on thread scheduling.
// Written against an execution context Ctx (tid / nthr / sync / atomics) so that the same code runs as a CUDA block
// (initiator.cu) and as one host thread (tests/host/gnn_host.cpp, checked against scipy's linear_sum_assignment on the
// reference's padded matrix).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GNN_HD __host__ __device__ __forceinline__
#else
#define GNN_HD inline
#endif

namespace mht {
namespace gnn {

constexpr unsigned long long kInfBits = 0x7ff0000000000000ull;   // +inf as an ordered key (labels are >= 0)
constexpr int kNoPred = 0x7fffffff;

GNN_HD unsigned long long to_bits(double x) {
    union { double d; unsigned long long u; } c;
    c.d = x;
    return c.u;
}
GNN_HD double from_bits(unsigned long long u) {
    union { double d; unsigned long long u; } c;
    c.u = u;
    return c.d;
}

struct Graph {
    int n_rows, n_cols;
    const int *row_ptr;      // [n_rows + 1]
    const int *col;          // [E] ascending inside a row
    const double *cost;      // [E] >= 0
};

struct State {
    double *u;                    // [n_rows]
    double *v;                    // [n_cols] <= 0; 0 on free columns
    int *match_col;               // [n_rows] -1 free, -2 assigned to its own "unassigned" column, else the column
    int *match_row;               // [n_cols] -1 free
    unsigned long long *dcol;     // [n_cols] label of a column inside a search; +inf outside
    double *drow;                 // [n_rows] label of a row inside a search; +inf outside
    int *pred;                    // [n_cols] row the label came from; kNoPred outside
    int *mark;                    // [n_cols] round stamp (dedupes the changed list); never reset
    int *list_a, *list_b;         // [n_rows] frontier rows, this round / next round
    int *touched_rows;            // [n_rows]
    int *touched_cols;            // [n_cols]
    int *changed;                 // [n_cols]
};

// scalars shared by the threads of one component (shared memory on the device)
struct Shared {
    int n_active, n_next, n_trows, n_tcols, n_changed, end_id, stamp, free_rows;
    unsigned long long best;
    long long searches, rounds;
};

// one shortest-augmenting-path search from the free row s, followed by the potential update and the augmentation.
// off_r / off_c: start of this component's segment in the row- / column-sized scratch lists.
template <class Ctx>
GNN_HD void search(Ctx &c, const Graph &g, State &st, Shared *sh, int s, int off_r, int off_c, double BIG) {
    int *act = st.list_a + off_r, *nxt = st.list_b + off_r;
    int *trows = st.touched_rows + off_r, *tcols = st.touched_cols + off_c, *chg = st.changed + off_c;
    if (c.tid() == 0) {
        sh->n_active = 1;
        sh->n_next = 0;
        sh->n_trows = 1;
        sh->n_tcols = 0;
        sh->n_changed = 0;
        act[0] = s;
        trows[0] = s;
        st.drow[s] = 0.0;
        sh->best = to_bits(BIG - st.u[s]);      // s stays unassigned
        sh->searches += 1;
    }
    c.sync();
    while (true) {
        const int na = c.ld(&sh->n_active);
        if (na == 0) break;
        const int stamp = c.ld(&sh->stamp);
        // relax the edges of the frontier rows
        for (int a = c.tid(); a < na; a += c.nthr()) {
            const int i = act[a];
            const double di = st.drow[i], ui = st.u[i];
            if (!(di < from_bits(c.ld64(&sh->best)))) continue;
            const int mi = st.match_col[i];
            for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e) {
                const int j = g.col[e];
                if (j == mi) continue;
                double nd = di + ((g.cost[e] - ui) - st.v[j]);
                if (nd < di) nd = di;
                if (!(nd < from_bits(c.ld64(&sh->best)))) continue;
                const unsigned long long nb = to_bits(nd);
                const unsigned long long old = c.amin64(&st.dcol[j], nb);
                if (nb < old) {
                    if (old == kInfBits) tcols[c.aadd(&sh->n_tcols, 1)] = j;
                    if (c.aexch(&st.mark[j], stamp) != stamp) chg[c.aadd(&sh->n_changed, 1)] = j;
                }
            }
        }
        c.sync();
        // predecessor of every column whose label dropped this round (strict improvements only: the pointers form a tree)
        for (int a = c.tid(); a < na; a += c.nthr()) {
            const int i = act[a];
            const double di = st.drow[i], ui = st.u[i];
            const int mi = st.match_col[i];
            for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e) {
                const int j = g.col[e];
                if (j == mi || c.ld(&st.mark[j]) != stamp) continue;
                double nd = di + ((g.cost[e] - ui) - st.v[j]);
                if (nd < di) nd = di;
                if (to_bits(nd) == c.ld64(&st.dcol[j])) c.amin32(&st.pred[j], -1 - i);   // temporary negative code: smallest wins = largest row
            }
        }
        c.sync();
        const int nc = c.ld(&sh->n_changed);
        for (int k = c.tid(); k < nc; k += c.nthr()) {
            const int j = chg[k];
            const int pj = c.ld(&st.pred[j]);
            if (pj < 0) st.pred[j] = -1 - pj;
            const double d = from_bits(c.ld64(&st.dcol[j]));
            if (!(d < from_bits(c.ld64(&sh->best)))) continue;
            const int i2 = st.match_row[j];
            if (i2 < 0) {
                c.amin64(&sh->best, to_bits(d));
            } else if (d < st.drow[i2]) {
                if (to_bits(st.drow[i2]) == kInfBits) trows[c.aadd(&sh->n_trows, 1)] = i2;
                st.drow[i2] = d;
                nxt[c.aadd(&sh->n_next, 1)] = i2;
                double cand = d + (BIG - st.u[i2]);
                if (cand < d) cand = d;
                c.amin64(&sh->best, to_bits(cand));
            }
        }
        c.sync();
        if (c.tid() == 0) {
            sh->n_active = sh->n_next;
            sh->n_next = 0;
            sh->n_changed = 0;
            sh->stamp = stamp + 1;
            sh->end_id = kNoPred;
            sh->rounds += 1;
        }
        int *t = act;
        act = nxt;
        nxt = t;
        c.sync();
    }
    // end point: the free column (or the row that gives up its column and stays unassigned) whose label is delta
    const unsigned long long dbits = c.ld64(&sh->best);
    const double delta = from_bits(dbits);
    const int ntr = c.ld(&sh->n_trows), ntc = c.ld(&sh->n_tcols);
    if (c.tid() == 0 && ntc == 0) sh->end_id = kNoPred;
    c.sync();
    for (int k = c.tid(); k < ntc; k += c.nthr()) {
        const int j = tcols[k];
        if (st.match_row[j] < 0 && c.ld64(&st.dcol[j]) == dbits) c.amin32(&sh->end_id, j);
    }
    for (int k = c.tid(); k < ntr; k += c.nthr()) {
        const int i = trows[k];
        const double d = st.drow[i];
        double cand = d + (BIG - st.u[i]);
        if (cand < d) cand = d;
        if (to_bits(cand) == dbits) c.amin32(&sh->end_id, g.n_cols + i);
    }
    c.sync();
    // augment (one thread: the path is a chain)
    if (c.tid() == 0) {
        int end = sh->end_id, cur = -1, guard = ntr + 2;
        if (end >= g.n_cols) {
            const int i = end - g.n_cols;
            cur = st.match_col[i];
            st.match_col[i] = -2;
        } else {
            cur = end;
        }
        while (cur >= 0 && guard-- > 0) {
            const int i = c.ld(&st.pred[cur]);
            const int next = st.match_col[i];
            st.match_col[i] = cur;
            st.match_row[cur] = i;
            cur = (i == s) ? -1 : next;
        }
    }
    c.sync();
    // potentials: labels below delta are exact shortest distances
    for (int k = c.tid(); k < ntr; k += c.nthr()) {
        const int i = trows[k];
        const double d = st.drow[i];
        if (d < delta) st.u[i] += delta - d;
        st.drow[i] = from_bits(kInfBits);
    }
    for (int k = c.tid(); k < ntc; k += c.nthr()) {
        const int j = tcols[k];
        const double d = from_bits(c.ld64(&st.dcol[j]));
        if (d < delta) st.v[j] -= delta - d;
        st.dcol[j] = kInfBits;
        st.pred[j] = kNoPred;
    }
    c.sync();
}

// ---- speculative parallel searches ----------------------------------------------------------------------------------
// The searches of one component are short and local (a handful of rows around the free row), so most of them can run AT
// THE SAME TIME: in a batch every warp runs the search of one free row read-only against the current state, with its labels
// in a small private table (Dijkstra: one row scanned per step, lanes over its edges).  Then the batch is filtered to a set
// of searches that do not interfere -- search b is kept iff no search with a smaller priority value modifies a node b touched
// and none touched a node b modifies ("smaller" = higher priority, claim_prio; claims by atomicMin on epoch-stamped keys, so the outcome depends on the data only,
// never on timing) -- and the kept ones commit (potentials, augmentation) concurrently: applied one after the other they
// would have read exactly the same values, so the result is a valid sequence of shortest-augmenting-path steps.  Rejected
// searches are repeated in the next batch; searches that outgrow the private table are left to the block-wide search().
constexpr int kSpecCols = 512;
constexpr int kSpecRows = 384;
constexpr int kSpecHash = 1024;       // open-addressing column -> table index map (power of two, >= 2 x kSpecCols)

struct Spec {
    int nC, nR, end, overflow, root, pad;
    double best;
    double c_d[kSpecCols];
    double r_d[kSpecRows];
    int c_id[kSpecCols], c_pred[kSpecCols], c_mrow[kSpecCols];   // c_mrow: the column's matched row, loaded with its potential
    int r_id[kSpecRows];
    int h_key[kSpecHash];
    short h_idx[kSpecHash];
    unsigned char c_done[kSpecCols];
};

struct Claims {                       // [n_rows] / [n_cols] keys: (inverted epoch << 40) | best priority among the claimants
    unsigned long long *touch_r, *mod_r, *touch_c, *mod_c;
};

// priority of a search inside its batch (smaller wins): the searches that scanned MORE rows first -- they are the ones that are
// expensive to repeat -- then the smaller row index.  Depends on the data only.
GNN_HD unsigned long long claim_prio(int rows_scanned, int row, bool size_first = true) {
    const int cls = size_first ? 1023 - (rows_scanned < 1023 ? rows_scanned : 1023) : 0;
    return ((unsigned long long)cls << 30) | (unsigned)row;
}
GNN_HD unsigned long long claim_key(unsigned epoch, unsigned long long prio) {
    return ((unsigned long long)(0xffffffu - epoch) << 40) | prio;
}
GNN_HD unsigned long long claim_owner(unsigned long long key, unsigned epoch) {
    return (key >> 40) == (unsigned long long)(0xffffffu - epoch) ? (key & ((1ull << 40) - 1)) : ~0ull;
}

// W: lane() / nlanes() / wsync() / wmin64() / wmin32() / wall() over the lanes of one search, aadd() / acas() on the private table,
// amin64() on the global claim keys
template <class W>
GNN_HD void spec_search(W &w, const Graph &g, const State &st, Spec *sp, int s, double BIG, int row_cap = kSpecRows) {
    if (w.lane() == 0) {
        sp->nC = 0;
        sp->nR = 1;
        sp->r_id[0] = s;
        sp->r_d[0] = 0.0;
        sp->best = BIG - st.u[s];
        sp->end = g.n_cols + s;
        sp->overflow = 0;
        sp->root = s;
    }
    for (int k = w.lane(); k < kSpecHash; k += w.nlanes()) sp->h_key[k] = -1;
    w.wsync();
    // One scanned row per step, three dependent round trips to L2: {u, row_ptr} of the row -> {col, cost} of its edges ->
    // {v, match_row} of their columns; everything else lives in the private table.
    int i = s, mi = -1;
    double di = 0.0;
    while (true) {
        const double ui = st.u[i];
        const int e0 = g.row_ptr[i], e1 = g.row_ptr[i + 1];
        double best = sp->best;
        if (i != s) {                                  // the row may give up its column and stay unassigned
            double cand = di + (BIG - ui);
            if (cand < di) cand = di;
            if (cand < best) {
                best = cand;
                if (w.lane() == 0) {
                    sp->best = cand;
                    sp->end = g.n_cols + i;
                }
            }
        }
        for (int e = e0 + w.lane(); e < e1; e += w.nlanes()) {
            const int j = g.col[e];
            if (j == mi) continue;
            const int mrow = st.match_row[j];
            double nd = di + ((g.cost[e] - ui) - st.v[j]);
            if (nd < di) nd = di;
            if (!(nd < best)) continue;
            // the columns of one row are distinct: no other lane looks up or inserts j during this scan
            unsigned slot = ((unsigned)j * 2654435761u) >> 12 & (kSpecHash - 1);
            while (true) {
                int key = sp->h_key[slot];
                if (key == -1) key = w.acas(&sp->h_key[slot], -1, j);
                if (key == -1) {                       // inserted
                    const int k = w.aadd(&sp->nC, 1);
                    if (k < kSpecCols) {
                        sp->h_idx[slot] = (short)k;
                        sp->c_id[k] = j;
                        sp->c_d[k] = nd;
                        sp->c_pred[k] = i;
                        sp->c_mrow[k] = mrow;
                        sp->c_done[k] = 0;
                    } else {
                        sp->overflow = 1;
                    }
                    break;
                }
                if (key == j) {
                    const int idx = sp->h_idx[slot];
                    if (!sp->c_done[idx] && nd < sp->c_d[idx]) {
                        sp->c_d[idx] = nd;
                        sp->c_pred[idx] = i;
                    }
                    break;
                }
                slot = (slot + 1) & (kSpecHash - 1);
            }
        }
        w.wsync();
        if (sp->overflow) break;
        // nearest column not scanned yet (ties: smallest column index)
        const int nC = sp->nC;
        unsigned long long bd = kInfBits;
        int bj = kNoPred, bk = -1;
        for (int k = w.lane(); k < nC; k += w.nlanes())
            if (!sp->c_done[k]) {
                const unsigned long long b = to_bits(sp->c_d[k]);
                if (b < bd || (b == bd && sp->c_id[k] < bj)) {
                    bd = b;
                    bj = sp->c_id[k];
                    bk = k;
                }
            }
        const unsigned long long m = w.wmin64(bd);
        if (m == kInfBits || !(from_bits(m) < sp->best)) break;
        const int js = w.wmin32(bd == m ? bj : kNoPred);
        const int ks = w.wmin32((bd == m && bj == js) ? bk : kNoPred);
        const double dmin = from_bits(m);
        const int i2 = sp->c_mrow[ks];
        const int nR = sp->nR;
        w.wsync();
        if (w.lane() == 0) {
            sp->c_done[ks] = 1;
            if (i2 < 0) {
                sp->best = dmin;
                sp->end = js;
            } else if (nR >= row_cap) {
                sp->overflow = 1;
            } else {
                sp->r_id[nR] = i2;
                sp->r_d[nR] = dmin;
                sp->nR = nR + 1;
            }
        }
        w.wsync();
        if (i2 < 0 || sp->overflow) break;
        i = i2;
        mi = js;
        di = dmin;
    }
    if (w.lane() == 0 && sp->nC > kSpecCols) sp->nC = kSpecCols;
    w.wsync();
}

template <class W>
GNN_HD void spec_claim(W &w, const Spec *sp, Claims cl, unsigned epoch, bool size_first = true) {
    const unsigned long long key = claim_key(epoch, claim_prio(sp->nR, sp->root, size_first));
    const double delta = sp->best;
    for (int k = w.lane(); k < sp->nR; k += w.nlanes()) {
        w.amin64(&cl.touch_r[sp->r_id[k]], key);
        if (sp->r_d[k] <= delta) w.amin64(&cl.mod_r[sp->r_id[k]], key);
    }
    for (int k = w.lane(); k < sp->nC; k += w.nlanes()) {
        w.amin64(&cl.touch_c[sp->c_id[k]], key);
        if (sp->c_d[k] <= delta) w.amin64(&cl.mod_c[sp->c_id[k]], key);
    }
}

// true iff no search of this batch with a higher priority (smaller claim_prio) interferes
template <class W>
GNN_HD bool spec_check(W &w, const Spec *sp, Claims cl, unsigned epoch, bool size_first = true) {
    const unsigned long long s = claim_prio(sp->nR, sp->root, size_first);
    const double delta = sp->best;
    int ok = 1;
    for (int k = w.lane(); k < sp->nR; k += w.nlanes()) {
        const int id = sp->r_id[k];
        if (claim_owner(w.ld64(&cl.mod_r[id]), epoch) < s) ok = 0;
        if (sp->r_d[k] <= delta && claim_owner(w.ld64(&cl.touch_r[id]), epoch) < s) ok = 0;
    }
    for (int k = w.lane(); k < sp->nC; k += w.nlanes()) {
        const int id = sp->c_id[k];
        if (claim_owner(w.ld64(&cl.mod_c[id]), epoch) < s) ok = 0;
        if (sp->c_d[k] <= delta && claim_owner(w.ld64(&cl.touch_c[id]), epoch) < s) ok = 0;
    }
    return w.wall(ok) != 0;
}

template <class W>
GNN_HD void spec_commit(W &w, const Graph &g, State &st, Spec *sp) {
    const int s = sp->root;
    const double delta = sp->best;
    for (int k = w.lane(); k < sp->nR; k += w.nlanes()) {
        const double d = sp->r_d[k];
        if (d < delta) st.u[sp->r_id[k]] += delta - d;
    }
    for (int k = w.lane(); k < sp->nC; k += w.nlanes()) {
        const double d = sp->c_d[k];
        if (d < delta) st.v[sp->c_id[k]] -= delta - d;
    }
    w.wsync();
    if (w.lane() == 0) {
        int cur, guard = sp->nR + 2;
        if (sp->end >= g.n_cols) {
            const int i = sp->end - g.n_cols;
            cur = st.match_col[i];
            st.match_col[i] = -2;
        } else {
            cur = sp->end;
        }
        while (cur >= 0 && guard-- > 0) {
            unsigned slot = ((unsigned)cur * 2654435761u) >> 12 & (kSpecHash - 1);
            while (sp->h_key[slot] != cur) slot = (slot + 1) & (kSpecHash - 1);
            const int i = sp->c_pred[sp->h_idx[slot]];
            const int next = st.match_col[i];
            st.match_col[i] = cur;
            st.match_row[cur] = i;
            cur = (i == s) ? -1 : next;
        }
    }
    w.wsync();
}

// start values of a whole problem (all components at once, any thread layout): u = smallest cost of the row, and a row whose
// nearest column is wanted by no smaller row takes it.  Three passes separated by barriers of the caller:
template <class Ctx>
GNN_HD void start_pass(Ctx &c, int pass, const Graph &g, State &st, int first, int stride, double BIG) {
    for (int i = first; i < g.n_rows; i += stride) {
        if (pass == 0) {
            double m = BIG;
            int jm = -1;
            for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e)
                if (g.cost[e] < m) {
                    m = g.cost[e];
                    jm = g.col[e];
                }
            st.u[i] = m;
            st.list_a[i] = jm;
            if (jm >= 0) c.amin32(&st.pred[jm], i);
        } else if (pass == 1) {
            const int jm = st.list_a[i];
            if (jm >= 0 && c.ld(&st.pred[jm]) == i) {
                st.match_col[i] = jm;
                st.match_row[jm] = i;
            }
        } else {
            const int jm = st.list_a[i];
            if (jm >= 0) st.pred[jm] = kNoPred;
        }
    }
}

// Solve one connected component: rows[0..nr) ascending (global row indices).  The state arrays of its rows and columns
// must hold: match_col = -1, match_row = -1, v = 0, drow = +inf, dcol = +inf, pred = kNoPred, mark = 0.
template <class Ctx>
GNN_HD void solve_component(Ctx &c, const Graph &g, State &st, Shared *sh, const int *rows, int nr, int off_r, int off_c,
                            double BIG, bool started = false) {
    if (c.tid() == 0) {
        sh->stamp = 1;
        sh->searches = 0;
        sh->rounds = 0;
    }
    if (started) c.sync();
    if (!started) {
    // start: u = the row's smallest cost; a row whose nearest column is free takes it (smallest row index wins the column)
    for (int k = c.tid(); k < nr; k += c.nthr()) {
        const int i = rows[k];
        double m = BIG;
        int jm = -1;
        for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e)
            if (g.cost[e] < m) {
                m = g.cost[e];
                jm = g.col[e];
            }
        st.u[i] = m;
        st.list_a[off_r + k] = jm;
        if (jm >= 0) c.amin32(&st.pred[jm], i);
    }
    c.sync();
    for (int k = c.tid(); k < nr; k += c.nthr()) {
        const int i = rows[k];
        const int jm = st.list_a[off_r + k];
        if (jm >= 0 && c.ld(&st.pred[jm]) == i) {
            st.match_col[i] = jm;
            st.match_row[jm] = i;
        }
    }
    c.sync();
    for (int k = c.tid(); k < nr; k += c.nthr()) {
        const int jm = st.list_a[off_r + k];
        if (jm >= 0) st.pred[jm] = kNoPred;
    }
    c.sync();
    }
    for (int k = 0; k < nr; ++k) {
        const int i = rows[k];
        if (st.match_col[i] != -1) continue;       // uniform: written before the last barrier
        if (g.row_ptr[i + 1] == g.row_ptr[i]) continue;
        search(c, g, st, sh, i, off_r, off_c, BIG);
    }
}

}  // namespace gnn
}  // namespace mht
