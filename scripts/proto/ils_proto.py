import sys, time, os
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:3]
exec(open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read().split("def search(")[0])
rng = np.random.RandomState(1)
K = 40
cols_of_tree = []
sel0, c0 = greedy(rc)
for t in range(nT):
    idxs = np.arange(tstart[t], tend[t]); o = np.argsort(rc[idxs], kind='stable')[:K]
    cols_of_tree.append(np.array(sorted(set(idxs[o].tolist()) | {int(sel0[t]), int(tstart[t])})))
rowsets = [set(r.tolist()) for r in rows_of]

def ls(sel, max_sweeps=20):
    sel = sel.copy(); holder = -np.ones(nr, int)
    for t in range(nT): holder[rows_of[sel[t]]] = t
    for sw in range(max_sweeps):
        improved = False
        for t in range(nT):
            cur = sel[t]; best = None
            for j in cols_of_tree[t]:
                if j == cur: continue
                d = cost[j] - cost[cur]
                hs = set(int(h) for h in holder[rows_of[j]] if h >= 0 and h != t)
                if len(hs) > 1: continue
                if not hs:
                    if d < -1e-12 and (best is None or d < best[0]): best = (d, j, -1, -1)
                    continue
                o = hs.pop(); bo = None
                for jo in cols_of_tree[o]:
                    if jo == sel[o] or rowsets[jo] & rowsets[j]: continue
                    if any(h >= 0 and h != o and h != t for h in holder[rows_of[jo]]): continue
                    dd = cost[jo] - cost[sel[o]]
                    if bo is None or dd < bo[0]: bo = (dd, jo)
                if bo is None: continue
                if d + bo[0] < -1e-12 and (best is None or d + bo[0] < best[0]): best = (d + bo[0], j, o, bo[1])
            if best is not None:
                d, j, o, jo = best
                holder[rows_of[sel[t]]] = -1
                if o >= 0: holder[rows_of[sel[o]]] = -1; sel[o] = jo; holder[rows_of[jo]] = o
                sel[t] = j; holder[rows_of[j]] = t
                improved = True
        if not improved: break
    return sel, cost[sel].sum()
t0 = time.time()
s, c = ls(sel0)
print("greedy %.3f -> LS %.3f (optimum %.3f) %.0fs" % (c0, c, opt, time.time() - t0))
best_s, best_c = s, c
for it in range(int(args[3]) if len(args) > 3 else 30):
    s2 = best_s.copy()
    # kick: a random tree and the trees sharing rows with its neighbourhood go to all-miss
    t = rng.randint(nT); kick = {t}
    for j in cols_of_tree[t][:10]:
        for r in rows_of[j]:
            for tt in range(nT):
                if r in rowsets[best_s[tt]]: kick.add(tt)
    for tt in kick: s2[tt] = tstart[tt]
    s2, c2 = ls(s2, 6)
    if c2 < best_c - 1e-12: best_s, best_c = s2, c2
    if it % 10 == 9: print("  ILS %d: best %.3f (%.0fs)" % (it + 1, best_c, time.time() - t0))
print("ILS best %.3f, optimum %.3f, above by %.3f" % (best_c, opt, best_c - opt))
