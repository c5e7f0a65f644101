import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_gpu_parity as tg
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_lowclutter"
for k, g, pre, trk, nodes, hist, info in tg._replay_tracker(name, maxTargets=1024, maxNodes=1 << 20):
    ids = [n.ID for n in nodes]; want = list(g[pre + "ids"])
    print("scan", k + 1, {kk: info[kk] for kk in ("n_trees", "n_parents", "n_children", "n_clusters", "n_multi_clusters", "certified", "dual_iters", "n_candidates", "bb_nodes", "lower_bound", "objective", "n_dead", "max_component", "n_components", "ms_assoc")})
    miss = sorted(set(want) - set(ids)); extra = sorted(set(ids) - set(want))
    print("  missing", miss, "extra", extra)
    for t in trk.__terminatedTargets__:
        if t.ID in miss:
            print("  terminated", t.ID, t.status, "cnllr", t.cumulativeNLLR, "x", t.x_0, "meas", t.measurementNumber, "scan", t.scanNumber)
    H = g[pre + "hist"]
    common = [i for i in ids if i in want]
    same = sum(hist[ids.index(i)] == list(H[want.index(i), :len(hist[ids.index(i)])]) for i in common)
    print("  identical histories %d / %d" % (same, len(common)))
    if miss or extra:
        break
