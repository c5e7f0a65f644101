// EXPERIMENTAL -- not part of libmht_b200.so (pymht_b200/build.py does not compile this directory).
//
// Lagrangian branch & bound with dual re-optimisation at every node: the exact repair DESIGN.md section 9 plans for
// clusters whose LP relaxation has a gap (cfg3 scan 2: 279 trees).  Same algorithm as scripts/proto/lbb_proto.py,
// written once for a generic execution context so that
//   * scripts/proto/lbb_host.cpp runs it single-threaded on the CPU (checked against HiGHS by
//     scripts/proto/lbb_host_check.py), and
//   * lbb_repair.cu instantiates it for one CTA per component (compiles for sm_100a; NOT yet run on a GPU).
//
// Problem of one component: columns 0..nC-1 sorted by tree, each with a cost and <= W local rows; pick one column
// per tree, every row at most once, minimum cost  (reference pymht/tracker.py:1155-1217 restricted to the core).
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define LBB_HD __host__ __device__
#else
#define LBB_HD
#endif

namespace lbb {

struct Problem {
    int nC, nT, nR, W;
    const double *cost;    // [nC]
    const int *tree;       // [nC] non-decreasing, 0..nT-1
    const int *rows;       // [W][nC] local row id or -1
    const double *u0;      // [nR] multipliers to start from (>= 0)
    double ub0;            // cost of the incumbent
    const int *sel0;       // [nT] incumbent column per tree
    int iters_root, iters_node, max_nodes, max_depth;
};

struct Work {              // scratch of one component; sizes in brackets
    double *u;             // [nR]
    double *u_stack;       // [max_depth + 1][nR] multipliers at the end of each open node
    int *usage;            // [nR]
    int *rowtaken;         // [nR] tree whose FIXED column holds the row, -1 free
    int *fixed;            // [nT] column the tree is fixed to, -1 free
    unsigned *forbid;      // [(nC + 31) / 32] forbidden columns
    unsigned long long *tmin;   // [nT] ordered key of the tree's minimum reduced cost
    int *targ;             // [nT] argmin column (ties -> larger index, like the reference's '<=')
    int *best_sel;         // [nT] out: best selection found
    int *frame_j, *frame_t, *frame_state;   // [max_depth + 1] decision stack
    double *red;           // [threads + 8] reduction scratch
    long long *redi;       // [threads + 8]
    double best;           // out: its cost
    int proven;            // out: 1 = search exhausted (best is optimal)
    int nodes;             // out
};

LBB_HD inline unsigned long long key_of(double v) {
    union { double d; unsigned long long u; } c;
    c.d = v;
    return (c.u & 0x8000000000000000ull) ? ~c.u : (c.u | 0x8000000000000000ull);
}
LBB_HD inline double of_key(unsigned long long k) {
    union { double d; unsigned long long u; } c;
    c.u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return c.d;
}
constexpr unsigned long long kInfKey = ~0ull;

// Ctx: int tid(), int nthr(), void sync(), void amin(unsigned long long*, unsigned long long),
//      void amax(int*, int), void aadd(int*, int)
template <class Ctx>
struct Solver {
    const Problem &p;
    Work &w;
    Ctx &c;
    LBB_HD Solver(const Problem &p_, Work &w_, Ctx &c_) : p(p_), w(w_), c(c_) {}

    LBB_HD bool alive(int j) const {
        if (w.forbid[j >> 5] >> (j & 31) & 1u) return false;
        const int t = p.tree[j], f = w.fixed[t];
        if (f >= 0 && f != j) return false;
        for (int k = 0; k < p.W; ++k) {
            const int r = p.rows[(long long)k * p.nC + j];
            if (r >= 0 && w.rowtaken[r] >= 0 && w.rowtaken[r] != t) return false;
        }
        return true;
    }
    LBB_HD double rc(int j) const {
        double v = p.cost[j];
        for (int k = 0; k < p.W; ++k) {
            const int r = p.rows[(long long)k * p.nC + j];
            if (r >= 0) v += w.u[r];
        }
        return v;
    }
    // block reductions through w.red / w.redi (every thread returns the same value)
    LBB_HD double sum(double v) {
        w.red[c.tid()] = v;
        c.sync();
        if (c.tid() == 0) {
            double s = 0.0;
            for (int i = 0; i < c.nthr(); ++i) s += w.red[i];
            w.red[c.nthr()] = s;
        }
        c.sync();
        const double s = w.red[c.nthr()];
        c.sync();
        return s;
    }
    LBB_HD long long maxi(long long v) {
        w.redi[c.tid()] = v;
        c.sync();
        if (c.tid() == 0) {
            long long s = w.redi[0];
            for (int i = 1; i < c.nthr(); ++i) s = w.redi[i] > s ? w.redi[i] : s;
            w.redi[c.nthr()] = s;
        }
        c.sync();
        const long long s = w.redi[c.nthr()];
        c.sync();
        return s;
    }

    // Subgradient iterations under the current fixings.  Returns the best bound seen (1e300 = infeasible node);
    // leaves targ / usage of the LAST iterate for branching; updates the incumbent from conflict-free iterates.
    LBB_HD double evaluate(int iters) {
        double bestL = -1e300, theta = 0.5;
        int stall = 0;
        for (int it = 0; it < iters; ++it) {
            for (int t = c.tid(); t < p.nT; t += c.nthr()) {
                w.tmin[t] = kInfKey;
                w.targ[t] = -1;
            }
            for (int r = c.tid(); r < p.nR; r += c.nthr()) w.usage[r] = 0;
            c.sync();
            for (int j = c.tid(); j < p.nC; j += c.nthr())
                if (alive(j)) c.amin(&w.tmin[p.tree[j]], key_of(rc(j)));
            c.sync();
            for (int j = c.tid(); j < p.nC; j += c.nthr())
                if (alive(j) && key_of(rc(j)) == w.tmin[p.tree[j]]) c.amax(&w.targ[p.tree[j]], j);
            c.sync();
            double lsum = 0.0, csum = 0.0;
            long long dead = 0;
            for (int t = c.tid(); t < p.nT; t += c.nthr()) {
                const int j = w.targ[t];
                if (j < 0) {
                    dead = 1;
                    continue;
                }
                lsum += of_key(w.tmin[t]);
                csum += p.cost[j];
                for (int k = 0; k < p.W; ++k) {
                    const int r = p.rows[(long long)k * p.nC + j];
                    if (r >= 0) c.aadd(&w.usage[r], 1);
                }
            }
            if (maxi(dead)) return 1e300;
            lsum = sum(lsum);
            csum = sum(csum);
            double usum = 0.0, nrm = 0.0;
            long long worst = 0;
            for (int r = c.tid(); r < p.nR; r += c.nthr()) {
                int g = w.usage[r] - 1;
                if (w.u[r] <= 0.0 && g < 0) g = 0;
                usum += w.u[r];
                nrm += (double)g * (double)g;
                if (w.usage[r] > worst) worst = w.usage[r];
            }
            usum = sum(usum);
            nrm = sum(nrm);
            worst = maxi(worst);
            const double L = lsum - usum;
            if (worst <= 1 && csum < w.best - 1e-12) {   // conflict-free argmins: a feasible selection
                c.sync();
                if (c.tid() == 0) w.best = csum;
                for (int t = c.tid(); t < p.nT; t += c.nthr()) w.best_sel[t] = w.targ[t];
                c.sync();
            }
            if (L > bestL + 1e-12) {
                bestL = L;
                stall = 0;
            } else if (++stall >= 5) {
                theta *= 0.7;
                stall = 0;
            }
            if (bestL >= w.best - 1e-9 || nrm == 0.0) break;
            const double step = theta * (w.best - L) / nrm;
            for (int r = c.tid(); r < p.nR; r += c.nthr()) {
                int g = w.usage[r] - 1;
                if (w.u[r] <= 0.0 && g < 0) g = 0;
                const double v = w.u[r] + step * (double)g;
                w.u[r] = v > 0.0 ? v : 0.0;
            }
            c.sync();
        }
        return bestL;
    }

    // Branching column: an argmin user of the most contested row; if the argmins are conflict free (gap without
    // conflict), the costliest argmin of a free tree.  Returns false when every tree is fixed.
    LBB_HD bool pick(int &t_out, int &j_out) {
        long long best = -1;   // (usage << 32) | (nR - r): most used row, lowest index
        for (int r = c.tid(); r < p.nR; r += c.nthr())
            if (w.usage[r] > 1) {
                const long long k = ((long long)w.usage[r] << 32) | (long long)(p.nR - r);
                if (k > best) best = k;
            }
        best = maxi(best);
        long long pickk = -1;  // (priority << 32) | (nT - t)
        if (best >= 0) {
            const int r = p.nR - (int)(best & 0xffffffffll);
            for (int t = c.tid(); t < p.nT; t += c.nthr()) {
                const int j = w.targ[t];
                bool uses = false;
                for (int k = 0; k < p.W && j >= 0; ++k) uses = uses || p.rows[(long long)k * p.nC + j] == r;
                if (uses) {
                    const long long k = (1ll << 32) | (long long)(p.nT - t);
                    if (k > pickk) pickk = k;
                }
            }
        } else {
            for (int t = c.tid(); t < p.nT; t += c.nthr())
                if (w.fixed[t] < 0 && w.targ[t] >= 0) {
                    // rank by cost through the ordered key's upper bits (deterministic, thread-count independent)
                    const long long k = (long long)(key_of(p.cost[w.targ[t]]) >> 33) << 32 | (long long)(p.nT - t);
                    if (k > pickk) pickk = k;
                }
        }
        pickk = maxi(pickk);
        if (pickk < 0) return false;
        t_out = p.nT - (int)(pickk & 0xffffffffll);
        j_out = w.targ[t_out];
        return j_out >= 0;
    }

    LBB_HD void set_fix(int t, int j, bool on) {
        if (c.tid() == 0) {
            w.fixed[t] = on ? j : -1;
            for (int k = 0; k < p.W; ++k) {
                const int r = p.rows[(long long)k * p.nC + j];
                if (r >= 0) w.rowtaken[r] = on ? t : -1;
            }
        }
        c.sync();
    }
    LBB_HD void set_forbid(int j, bool on) {
        if (c.tid() == 0) {
            if (on) w.forbid[j >> 5] |= 1u << (j & 31);
            else w.forbid[j >> 5] &= ~(1u << (j & 31));
        }
        c.sync();
    }

    LBB_HD void run() {
        for (int t = c.tid(); t < p.nT; t += c.nthr()) {
            w.fixed[t] = -1;
            w.best_sel[t] = p.sel0[t];
        }
        for (int r = c.tid(); r < p.nR; r += c.nthr()) w.rowtaken[r] = -1;
        for (int i = c.tid(); i < (p.nC + 31) / 32; i += c.nthr()) w.forbid[i] = 0u;
        if (c.tid() == 0) {
            w.best = p.ub0;
            w.nodes = 0;
            w.proven = 0;
        }
        c.sync();
        int depth = 0, nodes = 0;
        bool descend = true, complete = true;
        while (true) {
            if (descend) {
                if (++nodes > p.max_nodes) {
                    complete = false;
                    break;
                }
                const double *src = depth == 0 ? p.u0 : w.u_stack + (long long)(depth - 1) * p.nR;
                for (int r = c.tid(); r < p.nR; r += c.nthr()) w.u[r] = src[r];
                c.sync();
                const double L = evaluate(depth == 0 ? p.iters_root : p.iters_node);
                int t = -1, j = -1;
                if (L >= w.best - 1e-9) {
                    descend = false;
                } else if (depth >= p.max_depth) {
                    complete = false;       // cannot go deeper: the subtree stays unexplored
                    descend = false;
                } else if (!pick(t, j)) {
                    descend = false;        // every tree fixed: the node is a leaf, evaluate() recorded it
                } else {
                    for (int r = c.tid(); r < p.nR; r += c.nthr()) w.u_stack[(long long)depth * p.nR + r] = w.u[r];
                    if (c.tid() == 0) {
                        w.frame_j[depth] = j;
                        w.frame_t[depth] = t;
                        w.frame_state[depth] = 0;
                    }
                    c.sync();
                    set_fix(t, j, true);    // child A: the tree takes the column
                    ++depth;
                }
            } else {
                if (depth == 0) break;
                const int d = depth - 1;
                const int j = w.frame_j[d], t = w.frame_t[d], st = w.frame_state[d];
                c.sync();
                if (st == 0) {              // child B: the column is forbidden
                    set_fix(t, j, false);
                    set_forbid(j, true);
                    if (c.tid() == 0) w.frame_state[d] = 1;
                    c.sync();
                    descend = true;
                } else {
                    set_forbid(j, false);
                    depth = d;
                }
            }
        }
        if (c.tid() == 0) {
            w.nodes = nodes;
            w.proven = complete ? 1 : 0;
        }
        c.sync();
    }
};

}  // namespace lbb
