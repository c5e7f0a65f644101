"""CPU prototype: Lagrangian bound + reduced-cost fixing + MRV/forward-checking DFS with the dynamic bound on the
biggest cluster of a fixture scan.  Counts search nodes."""
import sys, os, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.sparse import csr_matrix
from oracle import mht_oracle as mo
from conftest import golden
sys.setrecursionlimit(10000)

name, upto = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 500
g = golden(name)
T, lam_phi, lam_nu, N, Pd, eta2, R = g["params"]
trk = mo.OracleTracker(T, lam_phi, lam_nu, eta2=eta2, N=int(N), P_d=Pd)
for x in g["init_x"]:
    trk.initiate(x, float(g["init_time"]))
for k in range(upto):
    pre = "s%d_" % k
    trk.n_scans += 1
    trk._grow(g[pre + "z"], float(g[pre + "time"]), trk.n_scans)
    cls = trk._cluster()
    if k < upto - 1:
        trk._select(cls); trk._terminate(); trk._prune()
cl = max(cls, key=len)
cost, ct, ptr, idx, nr, nodes = trk._columns(cl)
cost = cost * trk.N
n, nT = len(cost), len(cl)
cols = np.repeat(np.arange(n), np.diff(ptr))
A = csr_matrix((np.ones(len(idx)), (cols, idx)), shape=(n, nr))      # col x row
tstart = np.searchsorted(ct, np.arange(nT)); tend = np.append(tstart[1:], n)
sel_opt, opt = mo.solve_blp(cost, ct, ptr, idx, nT, nr)
print("cluster trees %d cols %d rows %d  optimum %.6f" % (nT, n, nr, opt))
rows_of = [idx[ptr[j]:ptr[j + 1]] for j in range(n)]

def greedy(rc):
    order = np.argsort(rc, kind="stable")
    taken = np.zeros(nr, bool); sel = -np.ones(nT, int)
    for j in order:
        t = ct[j]
        if sel[t] >= 0 or taken[rows_of[j]].any():
            continue
        sel[t] = j; taken[rows_of[j]] = True
    return sel, cost[sel].sum()

# subgradient with Polyak step
u = np.zeros(nr); best_L, best_u, ub, best_sel = -1e300, u.copy(), 1e300, None
theta, stall = 1.0, 0
for it in range(iters):
    rc = cost + A @ u
    mins = np.minimum.reduceat(rc, tstart)
    arg = np.array([tstart[t] + np.argmin(rc[tstart[t]:tend[t]]) for t in range(nT)])
    L = mins.sum() - u.sum()
    if it % 40 == 0:
        s, c = greedy(rc)
        if c < ub: ub, best_sel = c, s
    if L > best_L + 1e-12: best_L, best_u, stall = L, u.copy(), 0
    else:
        stall += 1
        if stall >= 10: theta *= 0.7; stall = 0
    gsub = np.asarray(A[arg].sum(axis=0)).ravel() - 1.0
    gsub[(u <= 0) & (gsub < 0)] = 0
    nrm = (gsub ** 2).sum()
    if nrm == 0: break
    u = np.maximum(0, u + theta * (ub - L) / nrm * gsub)
u = best_u
rc = cost + A @ u
mins = np.minimum.reduceat(rc, tstart)
L = mins.sum() - u.sum()
s, c = greedy(rc)
if c < ub: ub, best_sel = c, s
print("after %d iters: L %.6f  UB %.6f  (optimum %.6f) gap %.4f" % (iters, L, ub, opt, ub - L))

def search(ub0, sel0, budget):
    gap = ub0 - L
    exc = rc - mins[ct]
    cand = [[j for j in range(tstart[t], tend[t]) if exc[j] <= gap + 1e-9 or j == sel0[t]] for t in range(nT)]
    for t in range(nT): cand[t].sort(key=lambda j: (exc[j], j))
    ncand = sum(len(c) for c in cand)
    taken = np.zeros(nr, np.int32)
    assigned = -np.ones(nT, int)
    best = [ub0, sel0.copy()]
    nodes = [0]
    free = [t for t in range(nT)]
    def feas(j): return not taken[rows_of[j]].any()
    def rec(exc_acc, cost_acc, unassigned):
        if nodes[0] > budget: return
        # evaluate unassigned trees
        bt, bn, summin, bfe = -1, 1 << 30, 0.0, None
        for t in unassigned:
            fe = [j for j in cand[t] if feas(j)]
            if not fe: return
            summin += max(0.0, exc[fe[0]])
            if len(fe) < bn: bt, bn, bfe = t, len(fe), fe
        if L + exc_acc + summin >= best[0] - 1e-12: return
        if bt < 0:
            return
        rest = [t for t in unassigned if t != bt]
        for j in bfe:
            e = max(0.0, exc[j])
            if L + exc_acc + e + (summin - max(0.0, exc[bfe[0]])) >= best[0] - 1e-12: break
            nodes[0] += 1
            taken[rows_of[j]] += 1; assigned[bt] = j
            if not rest:
                if cost_acc + cost[j] < best[0] - 1e-12:
                    best[0] = cost_acc + cost[j]; best[1] = assigned.copy()
            else:
                rec(exc_acc + e, cost_acc + cost[j], rest)
            taken[rows_of[j]] -= 1; assigned[bt] = -1
    t0 = time.time()
    rec(0.0, 0.0, free)
    return best[0], nodes[0], ncand, time.time() - t0

for budget in (200000,):
    b, nn, nc, dt = search(ub, best_sel, budget)
    print("MRV/FC search: candidates %d  nodes %d  best %.6f (optimum %.6f)  proven=%s  %.1fs" % (nc, nn, b, opt, nn <= budget, dt))
