import sys, time
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:3]
src = open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "search_proto.py")).read()
exec(src.split("for budget in (200000,):")[0])
# incumbent like the GPU's: 0.76 above the optimum (use the optimal selection as the incumbent column set, but the
# GPU's upper bound for the fixing threshold)
sel_inc = -np.ones(nT, int)
for j in sel_opt: sel_inc[ct[j]] = j
for slack in (0.76, 0.0):
    b, nn, nc, dt = search(opt + slack, sel_inc, 400000)
    print("UB = optimum + %.2f, L = LP - %.3f: candidates %d, MRV/FC nodes %d, best %.6f, proven=%s (%.0fs)" % (
        slack, -248.569584 - L if args[0] == "cfg3_head" else 0.0, nc, nn, b, nn <= 400000, dt))
