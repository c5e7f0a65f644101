import sys, time
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:3]
exec(open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read().split("def search(")[0])
print("greedy at best duals: UB %.4f ; optimum %.4f ; L %.4f" % (greedy(rc)[1], opt, L))

def dive(u0, iters_per_round=30, rounds=60, frac=1.0):
    u = u0.copy(); BIG = 1e6
    fixed = -np.ones(nT, int); taken = np.zeros(nr, bool)
    theta = 0.5
    ub_local = ub
    for rd in range(rounds):
        free = np.flatnonzero(fixed < 0)
        if len(free) == 0: break
        # a few subgradient iterations on the free trees (taken rows blocked)
        for it in range(iters_per_round if rd else 1):
            ue = np.where(taken, BIG, u)
            rcx = cost + A @ ue
            arg = np.array([tstart[t] + np.argmin(rcx[tstart[t]:tend[t]]) for t in free])
            use = np.asarray(A[arg].sum(axis=0)).ravel()
            gsub = use - 1.0
            gsub[taken] = 0
            gsub[(u <= 0) & (gsub < 0)] = 0
            nrm = (gsub ** 2).sum()
            if nrm == 0: break
            Lf = rcx[arg].sum() - u[~taken].sum()
            step = theta * max(1e-3, abs(ub_local - (Lf + cost[fixed[fixed >= 0]].sum()))) / nrm
            u = np.maximum(0, u + step * gsub)
        ue = np.where(taken, BIG, u)
        rcx = cost + A @ ue
        arg = np.array([tstart[t] + np.argmin(rcx[tstart[t]:tend[t]]) for t in free])
        use = np.asarray(A[arg].sum(axis=0)).ravel()
        # commit conflict-free argmins (all rows used exactly once), best margin first
        ok = [(t, j) for t, j in zip(free, arg) if (use[rows_of[j]] <= 1).all()]
        if not ok:   # nothing conflict-free: commit the single lowest reduced cost column
            k = int(np.argmin(rcx[arg])); ok = [(free[k], arg[k])]
        else:
            ok = ok[:max(1, int(frac * len(ok)))]
        for t, j in ok:
            fixed[t] = j; taken[rows_of[j]] = True
        theta = max(0.05, theta * 0.9)
    # leftovers: greedy
    free = np.flatnonzero(fixed < 0)
    if len(free):
        ue = np.where(taken, BIG, u); rcx = cost + A @ ue
        for j in np.argsort(rcx, kind="stable"):
            t = ct[j]
            if fixed[t] >= 0 or taken[rows_of[j]].any(): continue
            fixed[t] = j; taken[rows_of[j]] = True
    assert (fixed >= 0).all()
    return cost[fixed].sum(), rd

for ipr, rounds in ((10, 40), (30, 60)):
    t0 = time.time()
    c, rd = dive(u, ipr, rounds)
    print("dive iters/round %d rounds %d: UB %.4f (optimum %.4f) in %d rounds, %.1fs" % (ipr, rounds, c, opt, rd, time.time() - t0))
