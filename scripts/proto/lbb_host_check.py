"""Check pymht_b200/csrc/experimental/lbb_core.h (host build, single thread) against HiGHS on the biggest cluster of a
fixture scan.   python scripts/proto/lbb_host_check.py cfg3_head 2 [iters_node]"""
import ctypes as C, os, subprocess, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
args = sys.argv[1:]
sys.argv = [sys.argv[0]] + args[:2] + ["600"]
exec(open(os.path.join(HERE, "search_proto.py")).read().split("def search(")[0])
lib_path = "/tmp/liblbb_host.so"
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", lib_path, os.path.join(HERE, "lbb_host.cpp")])
lib = C.CDLL(lib_path)
W = int(max(np.diff(ptr)))
RM = -np.ones((W, n), dtype=np.int32)
for j in range(n):
    r = idx[ptr[j]:ptr[j + 1]]
    RM[:len(r), j] = r
sel_inc = np.zeros(nT, dtype=np.int32)
for j in sel_opt: sel_inc[ct[j]] = j
iters_node = int(args[2]) if len(args) > 2 else 40
slacks = [float(v) for v in args[3].split(",")] if len(args) > 3 else [0.76, 5.0]
max_nodes = int(args[4]) if len(args) > 4 else 200000
max_depth = int(args[5]) if len(args) > 5 else 60
for slack, label in [(v, "incumbent %.2f above the optimum" % v) for v in slacks]:
    best_sel = np.zeros(nT, dtype=np.int32); best = C.c_double(); nodes = C.c_int()
    cst = np.ascontiguousarray(cost, dtype=np.float64); tr = np.ascontiguousarray(ct, dtype=np.int32)
    u0 = np.ascontiguousarray(best_u, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    t0 = time.time()
    lib.lbb_solve_host.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_double, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    proven = lib.lbb_solve_host(n, nT, nr, W, p(cst), p(tr), p(np.ascontiguousarray(RM)), p(u0), float(opt + slack), p(sel_inc),
                                200, iters_node, max_nodes, max_depth, p(best_sel), C.byref(best), C.byref(nodes))
    ok = abs(best.value - opt) < 1e-9
    taken = RM[:, best_sel][RM[:, best_sel] >= 0]
    feas = len(taken) == len(set(taken.tolist())) and list(ct[best_sel]) == list(range(nT))
    print("%s: proven=%d nodes=%d best %.6f (HiGHS %.6f) match=%s feasible=%s %.1fs" % (
        label, proven, nodes.value, best.value, opt, ok, feas or slack > 0 and best.value >= opt + slack - 1e-9, time.time() - t0))
