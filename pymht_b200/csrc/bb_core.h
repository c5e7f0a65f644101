// Exact repair of the global-hypothesis 0/1 program (reference pymht/tracker.py:1155-1217, the CBC solve of
// _solveBLP_OR_TOOLS) for the components the dual loop leaves open: a best-first LAGRANGIAN BRANCH & BOUND whose
// nodes are evaluated by whole thread blocks.
//
//   problem of one component: columns sorted by tree, each with a cost and <= W rows; one column per tree, every
//   row at most once, minimum cost.
//   node       = set of alive columns (bit mask) + the multipliers its parent ended with
//   bound      = max over K projected-subgradient iterations of  sum_t min_j rc_j - sum_r u_r  (alive columns)
//   fixing     = columns whose reduced cost exceeds their tree minimum by more than (incumbent - bound) die
//   branching  = a (tree t, row r) pair whose ergodic usage y[t,r] (how often t's Lagrangian choice used r over
//                the second half of the iterations -- the subgradient method's estimate of the LP primal) is
//                closest to 1/2:   child 0: t does not use r      child 1: t uses r and nobody else does
//   incumbents = conflict-free Lagrangian choices
// Measured on the reference fixtures (host build, one thread): the 279-tree cluster of cfg3 scan 2 (LP gap 0.74,
// 6 430 columns) is proven in 140 nodes / 1.7e4 iterations; the previous depth-first column fix/forbid search
// needed 8 585 nodes / 3.4e5 iterations.
//
// The code is written once against an execution context (Ctx) so that tests/host/bb_host.cpp runs the SAME
// functions single-threaded on the CPU against HiGHS (tests/test_bb_core_host.py), and assoc.cu instantiates them
// with one CTA per node (many CTAs share one node pool).
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define BB_HD __host__ __device__
#else
#define BB_HD
#endif

namespace bb {

constexpr unsigned long long kInfKey = ~0ull;
constexpr double kPruneEps = 1e-9;

BB_HD inline unsigned long long key_of(double v) {
    union { double d; unsigned long long u; } c;
    c.d = v;
    return (c.u & 0x8000000000000000ull) ? ~c.u : (c.u | 0x8000000000000000ull);
}
BB_HD inline double of_key(unsigned long long k) {
    union { double d; unsigned long long u; } c;
    c.u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return c.d;
}

// One component (all arrays are views into the compacted core, global memory).
struct Comp {
    int nC, nT, nR, W;
    long long row_stride;        // rows[k * row_stride + j]
    const double *cost;          // [nC]
    const int *tree;             // [nC] local tree, non-decreasing
    const int *rows;             // [W][row_stride] local row or -1
    const int *tstart;           // [nT + 1]
    int nwords;                  // (nC + 31) / 32
    // shared state of the search (written with atomics)
    unsigned long long *ub_key;  // ordered key of the incumbent's cost
    int *best_sel;               // [nT] incumbent: local column per tree
    int *lock;                   // incumbent lock
};

// Scratch of ONE worker (a CTA).  Shared memory on the device when it fits, global memory otherwise.
struct Scratch {
    double *u;                   // [nR]
    int *usage;                  // [nR]
    unsigned long long *tmin;    // [nT]
    int *targ;                   // [nT]
    unsigned *alive;             // [nwords]
    // global memory
    double *rc;                  // [nC]
    int *freq;                   // [nC]
    double *ubest;               // [nR]
    int *best_targ;              // [nT]
    double *cand_d;              // [nT] branching candidate per tree: distance of y from 1/2
    int *cand_r;                 // [nT] ... and its row
    unsigned *alive2;            // [nwords] alive mask of a probed child (strong branching)
};

struct EvalResult {
    double bound;                // best bound seen (1e300: infeasible)
    int bt, br;                  // branching pair (-1: none)
    int solved;                  // 1 = conflict-free choices met the bound (node closed)
    int aborted;                 // 1 = the deadline passed during the iterations (bound still valid)
    int iters;
};

// Ctx: int tid(); int nthr(); void sync();
//      void amin64(unsigned long long*, unsigned long long); void amax(int*, int); void aadd(int*, int);
//      double sum(double); long long maxll(long long);   (block-wide, every thread gets the result)
//      unsigned long long bcast(unsigned long long)      (thread 0's value to every thread)
//      void reduce(double &a, double &b, long long &m, unsigned long long &x)   a, b summed, m maximised, x = thread 0's
//      int acas(int*, int cmp, int val); void fence();
//      bool bind(const Comp&, Scratch&)  point the scratch at memory that holds the component (false: too big)
//      void backoff(); bool expired()    (thread 0's view; callers broadcast it)
// Values another worker may change (incumbent, stop flag, counters) are read by thread 0 and broadcast, so that
// every thread of the block takes the same branch.

// block-wide broadcast of thread 0's value
template <class Ctx>
BB_HD inline long long bcast0(Ctx &c, long long v) { return (long long)c.bcast((unsigned long long)v); }
template <class Ctx>
BB_HD inline double read_ub(Ctx &c, const Comp &p) {
    unsigned long long k = 0;
    if (c.tid() == 0) k = *(volatile unsigned long long *)p.ub_key;
    return of_key(c.bcast(k));
}

template <class Ctx>
BB_HD inline bool is_alive(const Scratch &s, int j) { return (s.alive[j >> 5] >> (j & 31)) & 1u; }

template <class Ctx>
BB_HD inline bool col_uses(const Comp &p, int j, int r) {
    for (int k = 0; k < p.W; ++k)
        if (p.rows[(long long)k * p.row_stride + j] == r) return true;
    return false;
}

// alive &= decision(t, r, type).  type 0: t does not use r.  type 1: t uses r, nobody else does.  type < 0: none.
template <class Ctx>
BB_HD inline void apply_decision(Ctx &c, const Comp &p, Scratch &s, int t, int r, int type) {
    if (type < 0) return;
    for (int w = c.tid(); w < p.nwords; w += c.nthr()) {
        unsigned m = s.alive[w];
        if (!m) continue;
        unsigned out = m;
        for (int b = 0; b < 32; ++b) {
            if (!((m >> b) & 1u)) continue;
            const int j = w * 32 + b;
            const bool mine = p.tree[j] == t;
            if (type == 0 && !mine) continue;
            const bool uses = col_uses<Ctx>(p, j, r);
            const bool dead = type == 0 ? uses : (mine ? !uses : uses);
            if (dead) out &= ~(1u << b);
        }
        s.alive[w] = out;
    }
    c.sync();
}

// Incumbent update from the conflict-free choices in s.targ (cost csum).
template <class Ctx>
BB_HD inline void offer_incumbent(Ctx &c, const Comp &p, const Scratch &s, double csum) {
    // every thread sees the same csum; thread 0 takes the lock
    if (c.tid() == 0) {
        while (c.acas(p.lock, 0, 1) != 0) {}
        c.fence();
    }
    c.sync();
    const bool better = csum < read_ub(c, p);   // under the lock
    if (better) {
        for (int t = c.tid(); t < p.nT; t += c.nthr()) p.best_sel[t] = s.targ[t];
        c.sync();
        if (c.tid() == 0) {
            c.fence();
            *(volatile unsigned long long *)p.ub_key = key_of(csum);
        }
    }
    c.sync();
    if (c.tid() == 0) {
        c.fence();
        *(volatile int *)p.lock = 0;
    }
    c.sync();
}

// Strong-branching probe: bound of the child (t, r, type) of the node whose alive mask is s.alive and whose best
// multipliers are s.ubest, after K iterations.  Uses s.alive2 / s.u / s.rc / s.tmin / s.targ / s.usage as scratch;
// s.alive, s.ubest, s.best_targ stay untouched.  Returns 1e300 for an infeasible child.
template <class Ctx>
BB_HD inline double probe_child(Ctx &c, const Comp &p, Scratch &s, int t, int r, int type, int K, double ub) {
    for (int w = c.tid(); w < p.nwords; w += c.nthr()) {
        unsigned m = s.alive[w];
        unsigned out = m;
        for (int b = 0; b < 32 && m; ++b) {
            if (!((m >> b) & 1u)) continue;
            const int j = w * 32 + b;
            const bool mine = p.tree[j] == t;
            if (type == 0 && !mine) continue;
            const bool uses = col_uses<Ctx>(p, j, r);
            const bool dead = type == 0 ? uses : (mine ? !uses : uses);
            if (dead) out &= ~(1u << b);
        }
        s.alive2[w] = out;
    }
    for (int q = c.tid(); q < p.nR; q += c.nthr()) s.u[q] = s.ubest[q];
    c.sync();
    double bestL = -1e300, theta = 1.0;
    int stall = 0;
    for (int it = 0; it < K; ++it) {
        for (int q = c.tid(); q < p.nT; q += c.nthr()) s.tmin[q] = kInfKey;
        for (int q = c.tid(); q < p.nR; q += c.nthr()) s.usage[q] = 0;
        c.sync();
        for (int j = c.tid(); j < p.nC; j += c.nthr()) {
            if (!((s.alive2[j >> 5] >> (j & 31)) & 1u)) continue;
            double v = p.cost[j];
            for (int k = 0; k < p.W; ++k) {
                const int q = p.rows[(long long)k * p.row_stride + j];
                if (q >= 0) v += s.u[q];
            }
            s.rc[j] = v;
            const unsigned long long key = key_of(v);
            const int tt = p.tree[j];
            if (key < *(volatile unsigned long long *)&s.tmin[tt]) c.amin64(&s.tmin[tt], key);
        }
        c.sync();
        // usage of the argmins: every column that attains its tree's minimum counts once per tree (ties are rare and
        // only perturb the direction, never the bound)
        for (int q = c.tid(); q < p.nT; q += c.nthr()) s.targ[q] = -1;
        c.sync();
        for (int j = c.tid(); j < p.nC; j += c.nthr()) {
            if (!((s.alive2[j >> 5] >> (j & 31)) & 1u)) continue;
            if (key_of(s.rc[j]) == s.tmin[p.tree[j]]) c.amax(&s.targ[p.tree[j]], j);
        }
        c.sync();
        double lsum = 0.0, dummy = 0.0;
        long long dead = 0;
        for (int q = c.tid(); q < p.nT; q += c.nthr()) {
            const int j = s.targ[q];
            if (j < 0) {
                dead = 1;
                continue;
            }
            lsum += of_key(s.tmin[q]);
            for (int k = 0; k < p.W; ++k) {
                const int rr = p.rows[(long long)k * p.row_stride + j];
                if (rr >= 0) c.aadd(&s.usage[rr], 1);
            }
        }
        unsigned long long x = 0;
        c.reduce(lsum, dummy, dead, x);
        if (dead) return 1e300;
        double usum = 0.0, nrm = 0.0;
        long long worst = 0;
        for (int q = c.tid(); q < p.nR; q += c.nthr()) {
            int g = s.usage[q] - 1;
            const double ur = s.u[q];
            if (ur <= 0.0 && g < 0) g = 0;
            usum += ur;
            nrm += (double)(g * g);
        }
        c.reduce(usum, nrm, worst, x);
        const double L = lsum - usum;
        if (L > bestL + 1e-12) {
            bestL = L;
            stall = 0;
        } else if (++stall >= 3) {
            theta *= 0.7;
            stall = 0;
        }
        if (nrm == 0.0 || bestL >= ub - kPruneEps) break;
        const double step = theta * (ub - L) / nrm;
        for (int q = c.tid(); q < p.nR; q += c.nthr()) {
            int g = s.usage[q] - 1;
            const double ur = s.u[q];
            if (ur <= 0.0 && g < 0) g = 0;
            const double v = ur + step * (double)g;
            s.u[q] = v > 0.0 ? v : 0.0;
        }
        c.sync();
    }
    c.sync();
    return bestL;
}

// K subgradient iterations on the alive columns, starting from s.u.  On return: s.ubest / s.best_targ hold the
// best iterate, s.alive has lost the columns fixed out at the best iterate, res has bound / branching pair.
template <class Ctx>
BB_HD inline void evaluate(Ctx &c, const Comp &p, Scratch &s, int K, EvalResult &res, int sb_cands = 0, int sb_iters = 15) {
    double bestL = -1e300, theta = 1.0;
    int stall = 0, it = 0;
    bool solved = false, infeasible = false, aborted = false;
    const int half = K / 2;
    for (int j = c.tid(); j < p.nC; j += c.nthr()) s.freq[j] = 0;
    for (int r = c.tid(); r < p.nR; r += c.nthr()) s.ubest[r] = s.u[r];
    for (int t = c.tid(); t < p.nT; t += c.nthr()) s.best_targ[t] = -1;
    c.sync();
    for (it = 0; it < K; ++it) {
        if ((it & 15) == 15 && bcast0(c, c.expired() ? 1 : 0)) {
            aborted = true;
            break;
        }
        for (int t = c.tid(); t < p.nT; t += c.nthr()) {
            s.tmin[t] = kInfKey;
            s.targ[t] = -1;
        }
        for (int r = c.tid(); r < p.nR; r += c.nthr()) s.usage[r] = 0;
        c.sync();
        for (int j = c.tid(); j < p.nC; j += c.nthr()) {
            if (!is_alive<Ctx>(s, j)) continue;
            double v = p.cost[j];
            for (int k = 0; k < p.W; ++k) {
                const int r = p.rows[(long long)k * p.row_stride + j];
                if (r >= 0) v += s.u[r];
            }
            s.rc[j] = v;
            const unsigned long long key = key_of(v);
            const int t = p.tree[j];
            if (key < *(volatile unsigned long long *)&s.tmin[t]) c.amin64(&s.tmin[t], key);
        }
        c.sync();
        for (int j = c.tid(); j < p.nC; j += c.nthr()) {
            if (!is_alive<Ctx>(s, j)) continue;
            const int t = p.tree[j];
            if (key_of(s.rc[j]) == s.tmin[t]) c.amax(&s.targ[t], j);   // ties -> later leaf, like the reference's '<='
        }
        c.sync();
        double lsum = 0.0, csum = 0.0;
        long long dead = 0;
        for (int t = c.tid(); t < p.nT; t += c.nthr()) {
            const int j = s.targ[t];
            if (j < 0) {
                dead = 1;
                continue;
            }
            lsum += of_key(s.tmin[t]);
            csum += p.cost[j];
            for (int k = 0; k < p.W; ++k) {
                const int r = p.rows[(long long)k * p.row_stride + j];
                if (r >= 0) c.aadd(&s.usage[r], 1);
            }
            if (it >= half) s.freq[j] += 1;
        }
        unsigned long long xk = 0;
        c.reduce(lsum, csum, dead, xk);
        if (dead) {
            infeasible = true;
            break;
        }
        double usum = 0.0, nrm = 0.0;
        long long worst = 0;
        for (int r = c.tid(); r < p.nR; r += c.nthr()) {
            int g = s.usage[r] - 1;
            const double ur = s.u[r];
            if (ur <= 0.0 && g < 0) g = 0;
            usum += ur;
            nrm += (double)(g * g);
            if (s.usage[r] > worst) worst = s.usage[r];
        }
        // the incumbent may have been improved by another worker: thread 0 reads it, everybody gets the same value
        if (c.tid() == 0) xk = *(volatile unsigned long long *)p.ub_key;
        c.reduce(usum, nrm, worst, xk);
        const double L = lsum - usum;
        double ub = of_key(xk);
        if (worst <= 1 && csum < ub - 1e-12) {
            offer_incumbent(c, p, s, csum);
            ub = read_ub(c, p);
        }
        if (L > bestL + 1e-12) {
            bestL = L;
            stall = 0;
            for (int r = c.tid(); r < p.nR; r += c.nthr()) s.ubest[r] = s.u[r];
            for (int t = c.tid(); t < p.nT; t += c.nthr()) s.best_targ[t] = s.targ[t];
        } else if (++stall >= 5) {
            theta *= 0.7;
            stall = 0;
        }
        if (nrm == 0.0) {   // conflict free and complementary: the choices are optimal for this node
            solved = true;
            break;
        }
        if (bestL >= ub - kPruneEps) break;
        const double step = theta * (ub - L) / nrm;
        for (int r = c.tid(); r < p.nR; r += c.nthr()) {
            int g = s.usage[r] - 1;
            const double ur = s.u[r];
            if (ur <= 0.0 && g < 0) g = 0;
            const double v = ur + step * (double)g;
            s.u[r] = v > 0.0 ? v : 0.0;
        }
        c.sync();
    }
    c.sync();
    res.iters = it;
    res.bt = res.br = -1;
    res.solved = solved ? 1 : 0;
    res.aborted = aborted ? 1 : 0;
    if (infeasible) {
        res.bound = 1e300;
        return;
    }
    res.bound = bestL;
    const double ub = read_ub(c, p);
    if (solved || aborted || bestL >= ub - kPruneEps) return;

    // ---- reduced costs at the best iterate: fixing + tree minima ----
    for (int r = c.tid(); r < p.nR; r += c.nthr()) s.u[r] = s.ubest[r];
    for (int t = c.tid(); t < p.nT; t += c.nthr()) s.tmin[t] = kInfKey;
    c.sync();
    for (int j = c.tid(); j < p.nC; j += c.nthr()) {
        if (!is_alive<Ctx>(s, j)) continue;
        double v = p.cost[j];
        for (int k = 0; k < p.W; ++k) {
            const int r = p.rows[(long long)k * p.row_stride + j];
            if (r >= 0) v += s.u[r];
        }
        s.rc[j] = v;
        const unsigned long long key = key_of(v);
        const int t = p.tree[j];
        if (key < *(volatile unsigned long long *)&s.tmin[t]) c.amin64(&s.tmin[t], key);
    }
    c.sync();
    const double gap = ub - bestL;
    for (int w = c.tid(); w < p.nwords; w += c.nthr()) {
        unsigned m = s.alive[w];
        if (!m) continue;
        unsigned out = m;
        for (int b = 0; b < 32; ++b) {
            if (!((m >> b) & 1u)) continue;
            const int j = w * 32 + b;
            if (s.rc[j] - of_key(s.tmin[p.tree[j]]) > gap + 1e-9) out &= ~(1u << b);
        }
        s.alive[w] = out;
    }
    c.sync();

    // ---- branching pair: ergodic usage y[t,r] closest to 1/2 ----
    const int denom = it - half > 0 ? it - half : 1;
    constexpr int kTab = 48;
    for (int t = c.tid(); t < p.nT; t += c.nthr()) {
        int tab_r[kTab], tab_n[kTab], nt = 0;
        for (int j = p.tstart[t]; j < p.tstart[t + 1]; ++j) {
            const int f = s.freq[j];
            if (!f) continue;
            for (int k = 0; k < p.W; ++k) {
                const int r = p.rows[(long long)k * p.row_stride + j];
                if (r < 0) continue;
                int q = 0;
                while (q < nt && tab_r[q] != r) ++q;
                if (q == nt) {
                    if (nt == kTab) continue;
                    tab_r[nt] = r;
                    tab_n[nt] = 0;
                    ++nt;
                }
                tab_n[q] += f;
            }
        }
        double bd = 2.0;
        int brow = -1;
        for (int q = 0; q < nt; ++q) {
            double y = (double)tab_n[q] / (double)denom;
            if (y > 1.0) y = 1.0;
            const double d = fabs(y - 0.5);
            if (d < bd || (d == bd && tab_r[q] < brow)) {
                bd = d;
                brow = tab_r[q];
            }
        }
        s.cand_d[t] = bd;
        s.cand_r[t] = brow;
    }
    c.sync();
    {
        // block argmin over trees of (distance, tree): quantise the distance to 2^-20 so the key fits 64 bits
        long long best = -1;
        for (int t = c.tid(); t < p.nT; t += c.nthr()) {
            if (s.cand_r[t] < 0 || s.cand_d[t] >= 0.5 - 1e-9) continue;
            const long long q = (long long)((0.5 - s.cand_d[t]) * 1048576.0);    // larger = closer to 1/2
            const long long key = (q << 32) | (long long)(0x7fffffff - t);
            if (key > best) best = key;
        }
        best = c.maxll(best);
        if (best >= 0 && sb_cands <= 1) {
            res.bt = 0x7fffffff - (int)(best & 0xffffffffll);
            res.br = s.cand_r[res.bt];
            return;
        }
        if (best >= 0) {
            // ---- strong branching: probe the sb_cands most fractional pairs, keep the one whose weaker child gains
            //      most (a pruned / infeasible child counts as the whole gap) ----
            double best_score = -1.0;
            int best_t = -1, best_r = -1;
            for (int cand = 0; cand < sb_cands; ++cand) {
                long long pick = -1;
                for (int t = c.tid(); t < p.nT; t += c.nthr()) {
                    if (s.cand_r[t] < 0 || s.cand_d[t] >= 0.5 - 1e-9) continue;
                    const long long q = (long long)((0.5 - s.cand_d[t]) * 1048576.0);
                    const long long key = (q << 32) | (long long)(0x7fffffff - t);
                    if (key > pick) pick = key;
                }
                pick = c.maxll(pick);
                if (pick < 0) break;
                const int t = 0x7fffffff - (int)(pick & 0xffffffffll);
                const int r = s.cand_r[t];
                c.sync();
                if (c.tid() == 0) s.cand_d[t] = 1.0;      // taken
                c.sync();
                const double gap_now = ub - bestL;
                double g0 = probe_child(c, p, s, t, r, 0, sb_iters, ub) - bestL;
                double g1 = probe_child(c, p, s, t, r, 1, sb_iters, ub) - bestL;
                if (g0 > gap_now) g0 = gap_now;
                if (g1 > gap_now) g1 = gap_now;
                if (g0 < 1e-6) g0 = 1e-6;
                if (g1 < 1e-6) g1 = 1e-6;
                const double score = g0 * g1;
                if (score > best_score) {
                    best_score = score;
                    best_t = t;
                    best_r = r;
                }
                if (bcast0(c, c.expired() ? 1 : 0)) break;
            }
            // the probes used s.u as scratch: the node's multipliers are s.ubest (what store_node keeps)
            for (int q = c.tid(); q < p.nR; q += c.nthr()) s.u[q] = s.ubest[q];
            c.sync();
            res.bt = best_t;
            res.br = best_r;
            if (best_t >= 0) return;
        }
    }
    // ---- fallbacks (the choices were stable over the second half) ----
    for (int r = c.tid(); r < p.nR; r += c.nthr()) s.usage[r] = 0;
    c.sync();
    for (int t = c.tid(); t < p.nT; t += c.nthr()) {
        const int j = s.best_targ[t];
        for (int k = 0; k < p.W && j >= 0; ++k) {
            const int r = p.rows[(long long)k * p.row_stride + j];
            if (r >= 0) c.aadd(&s.usage[r], 1);
        }
    }
    c.sync();
    {   // most contested row at the best iterate, lowest tree among its users
        long long best = -1;
        for (int r = c.tid(); r < p.nR; r += c.nthr())
            if (s.usage[r] > 1) {
                const long long key = ((long long)s.usage[r] << 32) | (long long)(0x7fffffff - r);
                if (key > best) best = key;
            }
        best = c.maxll(best);
        if (best >= 0) {
            const int r = 0x7fffffff - (int)(best & 0xffffffffll);
            long long bt = -1;
            for (int t = c.tid(); t < p.nT; t += c.nthr()) {
                const int j = s.best_targ[t];
                if (j >= 0 && col_uses<Ctx>(p, j, r)) {
                    const long long key = (long long)(0x7fffffff - t);
                    if (key > bt) bt = key;
                }
            }
            bt = c.maxll(bt);
            if (bt >= 0) {
                res.bt = 0x7fffffff - (int)bt;
                res.br = r;
                return;
            }
        }
    }
    {   // conflict free but a row with a positive multiplier is unused: the tree with the cheapest alive column on it
        long long best = -1;
        for (int r = c.tid(); r < p.nR; r += c.nthr())
            if (s.usage[r] == 0 && s.u[r] > 0.0) {
                long long q = (long long)(s.u[r] * 1048576.0);
                if (q > 0x3fffffff) q = 0x3fffffff;
                const long long key = (q << 32) | (long long)(0x7fffffff - r);
                if (key > best) best = key;
            }
        best = c.maxll(best);
        if (best >= 0) {
            const int r = 0x7fffffff - (int)(best & 0xffffffffll);
            long long bt = -1;
            for (int j = c.tid(); j < p.nC; j += c.nthr()) {
                if (!is_alive<Ctx>(s, j) || !col_uses<Ctx>(p, j, r)) continue;
                const double exc = s.rc[j] - of_key(s.tmin[p.tree[j]]);
                long long q = (long long)((exc < 0 ? 0 : exc) * 1048576.0);
                if (q > 0x3fffffff) q = 0x3fffffff;
                const long long key = ((0x3fffffffll - q) << 32) | (long long)(0x7fffffff - p.tree[j]);
                if (key > bt) bt = key;
            }
            bt = c.maxll(bt);
            if (bt >= 0) {
                res.bt = 0x7fffffff - (int)(bt & 0xffffffffll);
                res.br = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// node pool shared by all workers + the worker loop
// ------------------------------------------------------------------------------------------------
struct Pool {
    int cap;                 // node slots
    int node_words;          // alive words per slot
    int node_rows;           // multipliers per slot
    int *state;              // [cap] 0 free, 1 open, 2 taken
    double *key;             // [cap] priority: bound - incumbent at publication (most negative first)
    double *bound;           // [cap]
    int *comp;               // [cap]
    int *bt, *br;            // [cap] pair to branch on; bt = -1: unevaluated root
    unsigned *alive;         // [cap][node_words]
    float *u;                // [cap][node_rows] multipliers the node ended with (any u >= 0 is a valid start)
    int *outstanding;        // open + in-flight nodes
    int *stop;               // budget / deadline hit
    int *nodes;              // nodes evaluated
    int *iters;              // subgradient iterations
    int *comp_unproven;      // [n_comp] 1 = part of the component's tree was dropped
    int *comp_nodes;         // [n_comp]
};

// claim the open node with the smallest key (ties: lowest slot); -1 when nothing is open
template <class Ctx>
BB_HD inline int claim_best(Ctx &c, Pool &pl) {
    for (int attempt = 0; attempt < 8; ++attempt) {
        long long best = -1;
        for (int i = c.tid(); i < pl.cap; i += c.nthr())
            if (*(volatile int *)&pl.state[i] == 1) {
                const long long k = (long long)(~key_of(*(volatile double *)&pl.key[i]) >> 1);
                if (k > best) best = k;
            }
        best = c.maxll(best);
        if (best < 0) return -1;
        long long slot = -1;
        for (int i = c.tid(); i < pl.cap; i += c.nthr())
            if (*(volatile int *)&pl.state[i] == 1 &&
                (long long)(~key_of(*(volatile double *)&pl.key[i]) >> 1) == best) {
                const long long k = 0x7fffffff - i;
                if (k > slot) slot = k;
            }
        slot = c.maxll(slot);
        if (slot < 0) continue;
        const int i = 0x7fffffff - (int)slot;
        long long ok = 0;
        if (c.tid() == 0) {
            ok = c.acas(&pl.state[i], 1, 2) == 1 ? 1 : 0;
            c.fence();
        }
        ok = bcast0(c, ok);
        if (ok) return i;
    }
    return -1;
}

// reserve a free slot (state 0 -> 2); -1 when the pool is full
template <class Ctx>
BB_HD inline int alloc_slot(Ctx &c, Pool &pl, int hint) {
    for (int attempt = 0; attempt < 8; ++attempt) {
        long long slot = -1;
        for (int q = c.tid(); q < pl.cap; q += c.nthr()) {
            const int i = (q + hint) % pl.cap;
            if (*(volatile int *)&pl.state[i] == 0) {
                const long long k = 0x7fffffff - q;   // first free slot after the hint
                if (k > slot) slot = k;
                break;
            }
        }
        slot = c.maxll(slot);
        if (slot < 0) return -1;
        const int i = ((0x7fffffff - (int)slot) + hint) % pl.cap;
        long long ok = 0;
        if (c.tid() == 0) ok = c.acas(&pl.state[i], 0, 2) == 0 ? 1 : 0;
        ok = bcast0(c, ok);
        if (ok) return i;
        hint = i + 1;
    }
    return -1;
}

template <class Ctx>
BB_HD inline void store_node(Ctx &c, const Comp &p, Pool &pl, int slot, int comp, const Scratch &s, double bound,
                             int bt, int br) {
    unsigned *a = pl.alive + (long long)slot * pl.node_words;
    float *u = pl.u + (long long)slot * pl.node_rows;
    for (int w = c.tid(); w < p.nwords; w += c.nthr()) a[w] = s.alive[w];
    for (int r = c.tid(); r < p.nR; r += c.nthr()) u[r] = (float)s.ubest[r];
    if (c.tid() == 0) {
        pl.comp[slot] = comp;
        pl.bound[slot] = bound;
        pl.key[slot] = bound - of_key(*(volatile unsigned long long *)p.ub_key);
        pl.bt[slot] = bt;
        pl.br[slot] = br;
    }
    c.sync();
}

// Expand the (taken) node `cur`: evaluate its children, publish what survives.  Returns the slot of the child
// this worker continues with (taken), or -1.
template <class Ctx>
BB_HD inline int expand(Ctx &c, const Comp *comps, Pool &pl, int cur, Scratch &s, int K_root, int K_node,
                        int max_nodes, int sb_cands = 0, int sb_iters = 15) {
    const int ci = ((volatile int *)pl.comp)[cur];
    const Comp &p = comps[ci];
    const double pbound = ((volatile double *)pl.bound)[cur];
    const int pbt = ((volatile int *)pl.bt)[cur], pbr = ((volatile int *)pl.br)[cur];
    const bool root = pbt < 0;
    const double ub0 = read_ub(c, p);
    if (!root && pbound >= ub0 - kPruneEps) {   // the incumbent improved since the node was published
        c.sync();
        if (c.tid() == 0) {
            c.fence();
            *(volatile int *)&pl.state[cur] = 0;
            c.aadd(pl.outstanding, -1);
        }
        c.sync();
        return -1;
    }
    if (bcast0(c, (*(volatile int *)pl.nodes >= max_nodes || c.expired()) ? 1 : 0)) {
        if (c.tid() == 0) {
            *(volatile int *)pl.stop = 1;
            *(volatile int *)&pl.comp_unproven[ci] = 1;
        }
        c.sync();
        return -1;
    }
    const unsigned *pa = pl.alive + (long long)cur * pl.node_words;
    const float *pu = pl.u + (long long)cur * pl.node_rows;
    int surv[2] = {-1, -1};
    double sbound[2] = {0.0, 0.0};
    int nsurv = 0;
    const int nchild = root ? 1 : 2;
    for (int ch = 0; ch < nchild; ++ch) {
        // another SM wrote these: volatile loads bypass this SM's (non-coherent) L1
        for (int w = c.tid(); w < p.nwords; w += c.nthr()) s.alive[w] = ((const volatile unsigned *)pa)[w];
        for (int r = c.tid(); r < p.nR; r += c.nthr()) s.u[r] = (double)((const volatile float *)pu)[r];
        c.sync();
        apply_decision(c, p, s, pbt, pbr, root ? -1 : ch);
        EvalResult res;
        evaluate(c, p, s, root ? K_root : K_node, res, sb_cands, sb_iters);
        if (c.tid() == 0) {
            c.aadd(pl.nodes, 1);
            c.aadd(pl.iters, res.iters);
            c.aadd(&pl.comp_nodes[ci], 1);
        }
        const double ub = read_ub(c, p);
        double b = res.bound;
        if (!root && b < pbound) b = pbound;      // a child is never weaker than its parent
        if (res.solved || b >= ub - kPruneEps) continue;
        if (res.aborted) {                         // out of time inside the node: its subtree stays unexplored
            if (c.tid() == 0) {
                *(volatile int *)pl.stop = 1;
                *(volatile int *)&pl.comp_unproven[ci] = 1;
            }
            continue;
        }
        if (res.bt < 0) {                          // open but nothing to branch on (numerical corner): give up on it
            if (c.tid() == 0) *(volatile int *)&pl.comp_unproven[ci] = 1;
            continue;
        }
        int slot;
        const bool last = ch == nchild - 1;
        if (last) {
            slot = cur;                            // the parent's data is no longer needed
        } else {
            slot = alloc_slot(c, pl, cur + 1);
            if (slot < 0) {                        // pool full: the subtree is lost
                if (c.tid() == 0) *(volatile int *)&pl.comp_unproven[ci] = 1;
                continue;
            }
        }
        store_node(c, p, pl, slot, ci, s, b, res.bt, res.br);
        surv[nsurv] = slot;
        sbound[nsurv] = b;
        ++nsurv;
    }
    c.sync();
    const bool cur_reused = nsurv > 0 && surv[nsurv - 1] == cur;
    int keep = -1, publish = -1;
    if (nsurv == 1) keep = surv[0];
    if (nsurv == 2) {
        const bool first_better = sbound[0] <= sbound[1];
        keep = first_better ? surv[0] : surv[1];
        publish = first_better ? surv[1] : surv[0];
    }
    if (c.tid() == 0) {
        c.fence();
        if (nsurv != 1) c.aadd(pl.outstanding, nsurv - 1);
        if (publish >= 0) *(volatile int *)&pl.state[publish] = 1;
        if (!cur_reused) *(volatile int *)&pl.state[cur] = 0;
        c.fence();
    }
    c.sync();
    return keep;
}

// One worker (a CTA): claim the most promising open node, dive from it, repeat until the pool drains.
template <class Ctx>
BB_HD inline void worker(Ctx &c, const Comp *comps, Pool &pl, Scratch &s, int K_root, int K_node, int max_nodes,
                         int sb_cands = 0, int sb_iters = 15) {
    while (true) {
        if (bcast0(c, *(volatile int *)pl.stop)) break;
        int cur = claim_best(c, pl);
        if (cur < 0) {
            if (bcast0(c, *(volatile int *)pl.outstanding) <= 0) break;
            c.backoff();
            continue;
        }
        while (cur >= 0) {
            const int ci = ((volatile int *)pl.comp)[cur];
            if (!c.bind(comps[ci], s)) {   // scratch of this worker cannot hold the component
                if (c.tid() == 0) {
                    *(volatile int *)&pl.comp_unproven[ci] = 1;
                    *(volatile int *)&pl.state[cur] = 0;
                    c.aadd(pl.outstanding, -1);
                }
                c.sync();
                break;
            }
            cur = expand(c, comps, pl, cur, s, K_root, K_node, max_nodes, sb_cands, sb_iters);
            if (bcast0(c, *(volatile int *)pl.stop)) break;
        }
    }
}

}  // namespace bb
