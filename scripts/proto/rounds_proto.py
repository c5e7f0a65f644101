import sys
sys.argv = [sys.argv[0], "cfg3_lowclutter", "2", "500"]
exec(open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read().split("def search(")[0])
# parallel-rounds greedy as on the GPU: every uncommitted tree proposes its best feasible column; a proposal wins
# if it holds the lowest bid on all its rows
def rounds_greedy(rc, max_rounds=10**6):
    taken = np.zeros(nr, bool); sel = -np.ones(nT, int); r = 0
    while (sel < 0).any() and r < max_rounds:
        r += 1
        prop = {}
        for t in np.flatnonzero(sel < 0):
            best = None
            for j in range(tstart[t], tend[t]):
                if not taken[rows_of[j]].any() and (best is None or rc[j] < rc[best] or (rc[j] == rc[best] and j > best)):
                    best = j
            prop[t] = best
        bid = {}
        for t, j in prop.items():
            for rr in rows_of[j]:
                if rr not in bid or (rc[j], t) < bid[rr]: bid[rr] = (rc[j], t)
        for t, j in prop.items():
            if all(bid[rr] == (rc[j], t) for rr in rows_of[j]):
                sel[t] = j; taken[rows_of[j]] = True
    return sel, r
sel, r = rounds_greedy(rc)
print("parallel-rounds greedy: rounds needed %d, cost %.6f (sequential greedy %.6f, optimum %.6f)" % (r, cost[sel].sum(), greedy(rc)[1], opt))
