import sys, time
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:3]
exec(open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read().split("def search(")[0])
sel0, c0 = greedy(rc)
print("greedy: %.4f  optimum %.4f  L %.4f" % (c0, opt, L))
cols_of_tree = [np.arange(tstart[t], tend[t]) for t in range(nT)]

def local_search(sel, depth_max=3, sweeps=30):
    sel = sel.copy()
    holder = -np.ones(nr, int)
    for t in range(nT): holder[rows_of[sel[t]]] = t
    def release(t):
        holder[rows_of[sel[t]]] = -1
    def take(t, j):
        sel[t] = j; holder[rows_of[j]] = t
    # recursive ejection: tree t wants column j; conflicting trees must move (depth limited); returns best total delta and moves
    def try_move(t, j, depth, banned):
        delta = cost[j] - cost[sel[t]]
        confl = set(int(holder[r]) for r in rows_of[j] if holder[r] >= 0 and holder[r] != t)
        if not confl: return delta, [(t, j)]
        if depth == 0 or len(confl) > 2: return None
        moves = [(t, j)]
        blocked = set(rows_of[j].tolist())
        for o in confl:
            if o in banned: return None
            best = None
            for jo in cols_of_tree[o]:
                if jo == sel[o] or blocked & set(rows_of[jo].tolist()): continue
                # rows of jo must be free except those held by o itself or by trees already moving
                hs = set(int(holder[r]) for r in rows_of[jo] if holder[r] >= 0)
                hs -= {o, t} | confl
                if hs:
                    if depth <= 1: continue
                    sub = None   # keep it simple: no deeper chains for displaced trees
                    continue
                d = cost[jo] - cost[sel[o]]
                if best is None or d < best[0]: best = (d, jo)
            if best is None: return None
            delta += best[0]; moves.append((o, best[1])); blocked |= set(rows_of[best[1]].tolist())
        return delta, moves
    total = cost[sel].sum()
    for sw in range(sweeps):
        improved = False
        for t in range(nT):
            cur = sel[t]
            for j in cols_of_tree[t]:
                if cost[j] >= cost[cur] + 3.0: continue
                r = try_move(t, j, depth_max, {t})
                if r is None or r[0] >= -1e-12: continue
                for (tt, jj) in r[1]: release(tt)
                for (tt, jj) in r[1]: take(tt, jj)
                total += r[0]; improved = True
                break
        if not improved: break
    return sel, cost[sel].sum(), sw
t0 = time.time()
s1, c1, sw = local_search(sel0)
print("greedy + ejection(<=2 displaced) local search: %.4f after %d sweeps (%.1fs)" % (c1, sw, time.time() - t0))
