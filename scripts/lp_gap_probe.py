"""CPU probe: LP-relaxation vs integer optimum of the cluster problems of a fixture scan (oracle columns, HiGHS).
usage: lp_gap_probe.py cfg3_head 2"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.optimize import linprog, milp, LinearConstraint, Bounds
from scipy.sparse import csr_matrix
from oracle import mht_oracle as mo
from conftest import golden

name, upto = sys.argv[1], int(sys.argv[2])
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
    if k == upto - 1:
        tot_lp = tot_ip = singles = 0.0
        for cl in cls:
            if len(cl) < 2:
                t = cl[0]
                singles += min(l.cnllr for l in trk.leaves[t]) - trk.roots[t].cnllr
                continue
            cost, ct, ptr, idx, nr, nodes = trk._columns(cl)
            cost = cost * trk.N
            n = len(cost)
            cols = np.repeat(np.arange(n), np.diff(ptr))
            A1 = csr_matrix((np.ones(len(idx)), (idx, cols)), shape=(nr, n))
            A2 = csr_matrix((np.ones(n), (ct, np.arange(n))), shape=(len(cl), n))
            t0 = time.time()
            lp = linprog(cost, A_ub=A1, b_ub=np.ones(nr), A_eq=A2, b_eq=np.ones(len(cl)), bounds=(0, 1), method="highs")
            t1 = time.time()
            sel, obj = mo.solve_blp(cost, ct, ptr, idx, len(cl), nr)
            t2 = time.time()
            frac = int(np.sum((lp.x > 1e-6) & (lp.x < 1 - 1e-6)))
            tot_lp += lp.fun; tot_ip += obj
            if len(cl) > 20 or abs(lp.fun - obj) > 1e-7:
                print("cluster trees=%d cols=%d rows=%d  LP=%.6f IP=%.6f gap=%.2e fractional=%d  (lp %.2fs, mip %.2fs)" % (
                    len(cl), n, nr, lp.fun, obj, obj - lp.fun, frac, t1 - t0, t2 - t1))
        print("sum over multi clusters: LP %.6f IP %.6f" % (tot_lp, tot_ip))
        print("whole scan (singletons %.6f): LP bound %.6f, optimum %.6f" % (singles, singles + tot_lp, singles + tot_ip))
    trk._select(cls)
    trk._terminate()
    trk._prune()
