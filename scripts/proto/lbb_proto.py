"""Lagrangian branch & bound with dual re-optimisation at every node (numpy prototype)."""
import sys, time
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:3]
src = open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read()
exec(src.split("def search(")[0])
K_NODE = int(args[3]) if len(args) > 3 else 40
INF = 1e9
At = A.T.tocsr()   # row x col
col_rows = rows_of

def solve_node(alive, u0, ub, iters):
    """alive: bool mask of usable columns.  Returns best L, its u, the argmin selection at the best u, conflict-free flag."""
    u = u0.copy(); best = (-1e300, None, None, None); theta, stall = 0.5, 0
    cmask = np.where(alive, cost, INF)
    for it in range(iters):
        rcx = cmask + A @ u
        mins = np.minimum.reduceat(rcx, tstart)
        if (mins >= INF / 2).any(): return (1e300, u, None, False)      # a tree has no column left: infeasible
        arg = np.array([tstart[t] + int(np.argmin(rcx[tstart[t]:tend[t]])) for t in range(nT)])
        Ln = mins.sum() - u.sum()
        use = np.asarray(A[arg].sum(axis=0)).ravel()
        if Ln > best[0] + 1e-12:
            best = (Ln, u.copy(), arg, bool((use <= 1).all())); stall = 0
        else:
            stall += 1
            if stall >= 5: theta *= 0.7; stall = 0
        if Ln >= ub - 1e-9: break
        g = use - 1.0; g[(u <= 0) & (g < 0)] = 0
        nrm = (g ** 2).sum()
        if nrm == 0: break
        u = np.maximum(0, u + theta * (ub - Ln) / nrm * g)
    return best

ub = opt + 0.76          # the GPU's incumbent quality
best_sel = None
nodes = 0
t0 = time.time()
stack = [(np.ones(n, bool), best_u.copy(), 0)]
while stack:
    alive, u0, depth = stack.pop()
    nodes += 1
    Ln, un, arg, free = solve_node(alive, u0, ub, K_NODE if depth else 200)
    if Ln >= ub - 1e-9: continue
    cst = cost[arg].sum()
    use = np.asarray(A[arg].sum(axis=0)).ravel()
    if (use <= 1).all():
        if cst < ub - 1e-12: ub, best_sel = cst, arg.copy(); print("  node %d depth %d: new incumbent %.6f (L %.6f)" % (nodes, depth, ub, Ln))
        if Ln >= ub - 1e-9: continue
    # branch on the most contested row: its argmin users
    r = int(np.argmax(use))
    if use[r] <= 1:
        # conflict-free but L < cost: complementary slackness fails; branch on the tree with the largest rc - min spread... fix it
        t = int(np.argmax([cost[a] for a in arg])); j = arg[t]
    else:
        users = [t for t in range(nT) if r in col_rows[arg[t]]]
        t = users[0]; j = arg[t]
    # child B: forbid column j ; child A: fix tree t to j (remove every other column of t and every column sharing a row with j)
    b = alive.copy(); b[j] = False
    a = alive.copy(); a[tstart[t]:tend[t]] = False; a[j] = True
    for rr in col_rows[j]:
        cols_r = At[rr].indices
        a[cols_r] = False
    a[j] = True
    stack.append((b, un, depth + 1))
    stack.append((a, un, depth + 1))
    if nodes % 1000 == 0: print("  nodes %d, stack %d, ub %.6f, %.0fs" % (nodes, len(stack), ub, time.time() - t0))
    if nodes >= 1500: break
print("Lagrangian B&B: nodes %d, ub %.6f (optimum %.6f), proven=%s, %.0fs" % (nodes, ub, opt, not stack, time.time() - t0))
