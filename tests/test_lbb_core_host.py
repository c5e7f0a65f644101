"""CPU check of the EXPERIMENTAL exact-repair core (pymht_b200/csrc/experimental/lbb_core.h, the Lagrangian branch &
bound planned for the next round): its host build must find and PROVE the HiGHS optimum of every multi-tree cluster
of the small reference fixtures, starting from the all-miss incumbent and zero multipliers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import mht_oracle as mo


@pytest.fixture(scope="module")
def lbb_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lbb") / "liblbb_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "scripts", "proto", "lbb_host.cpp")])
    lib = C.CDLL(out)
    lib.lbb_solve_host.argtypes = ([C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_double, C.c_void_p] + [C.c_int] * 4 +
                                   [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)])
    return lib


def _clusters(name, upto):
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
        for cl in cls:
            if len(cl) > 1:
                yield k, trk, cl
        trk._select(cls)
        trk._terminate()
        trk._prune()


@pytest.mark.parametrize("name,upto", [("cfg1_crossing", 10), ("cfg5_small", 8), ("cfg2_small", 8)])
def test_host_core_proves_the_highs_optimum(lbb_lib, name, upto):
    n_checked = 0
    for k, trk, cl in _clusters(name, upto):
        cost, ct, ptr, idx, nr, nodes = trk._columns(cl)
        cost = np.ascontiguousarray(cost * trk.N, dtype=np.float64)
        n, nT = len(cost), len(cl)
        if n > 4000:
            continue
        sel_opt, opt = mo.solve_blp(cost, ct, ptr, idx, nT, nr)
        W = int(max(np.diff(ptr).max(), 1))
        RM = -np.ones((W, n), dtype=np.int32)
        for j in range(n):
            r = idx[ptr[j]:ptr[j + 1]]
            RM[:len(r), j] = r
        tstart = np.searchsorted(ct, np.arange(nT))
        assert all(ptr[j + 1] == ptr[j] for j in tstart), "first column of a tree must be its all-miss leaf"
        sel0 = np.ascontiguousarray(tstart, dtype=np.int32)
        u0 = np.zeros(max(nr, 1))
        best_sel = np.zeros(nT, dtype=np.int32)
        best, nn = C.c_double(), C.c_int()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        proven = lbb_lib.lbb_solve_host(n, nT, max(nr, 1), W, p(cost), p(np.ascontiguousarray(ct, dtype=np.int32)),
                                        p(np.ascontiguousarray(RM)), p(u0), float(cost[sel0].sum()), p(sel0), 200, 40,
                                        200000, 64, p(best_sel), C.byref(best), C.byref(nn))
        assert proven == 1, (name, k, nn.value)
        assert abs(best.value - opt) <= 1e-9 * max(1.0, abs(opt)), (name, k, best.value, opt)
        rows_used = RM[:, best_sel][RM[:, best_sel] >= 0]
        assert len(rows_used) == len(set(rows_used.tolist())) and list(np.asarray(ct)[best_sel]) == list(range(nT))
        assert abs(cost[best_sel].sum() - opt) <= 1e-9 * max(1.0, abs(opt))
        n_checked += 1
    assert n_checked > 0
