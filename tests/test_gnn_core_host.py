"""CPU check of the initiator's assignment core (pymht_b200/csrc/gnn_core.h: sparse maximum-cardinality / minimum-distance
matching by successive shortest paths, one thread block per connected component on the GPU).  tests/host/gnn_host.cpp
instantiates the SAME header with a one-thread execution context; its result must equal the reference's
_solve_global_nearest_neighbour (m_of_n.py:24-104: dense padded matrix + Munkres; restated in oracle/initiator_oracle.py
with scipy's linear_sum_assignment) on random gated point sets from sparse (small components) to dense (one giant
component), square and rectangular."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import initiator_oracle as io


@pytest.fixture(scope="module")
def gnn_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gnn") / "libgnn_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "host", "gnn_host.cpp")])
    lib = C.CDLL(out)
    lib.gnn_solve_host.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def solve_sparse(lib, dist, gate, n_warps=0):
    n1, n2 = dist.shape
    valid = dist <= gate
    row_ptr = np.zeros(n1 + 1, dtype=np.int32)
    row_ptr[1:] = np.cumsum(valid.sum(axis=1))
    rr, cc = np.nonzero(valid)
    col = np.ascontiguousarray(cc, dtype=np.int32)
    cost = np.ascontiguousarray(dist[rr, cc], dtype=np.float64)
    match = np.empty(n1, dtype=np.int32)
    stats = np.zeros(7, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.gnn_solve_host(n1, n2, p(row_ptr), p(col), p(cost), p(match), p(stats), n_warps) == 0
    return [(i, int(match[i])) for i in range(n1) if match[i] >= 0], stats


@pytest.mark.parametrize("n1,n2,side,gate,seed", [
    (40, 40, 400.0, 50.0, 1), (60, 35, 300.0, 50.0, 2), (35, 60, 300.0, 50.0, 3), (150, 150, 400.0, 50.0, 4),
    (300, 280, 530.0, 50.0, 5), (280, 300, 300.0, 50.0, 6), (200, 200, 150.0, 50.0, 7), (1, 5, 50.0, 50.0, 8),
    (5, 1, 50.0, 50.0, 9), (50, 50, 5000.0, 50.0, 10), (400, 120, 250.0, 50.0, 11), (120, 400, 250.0, 50.0, 12)])
def test_sparse_assignment_equals_reference_formulation(gnn_lib, n1, n2, side, gate, seed):
    rng = np.random.RandomState(seed)
    a = rng.uniform(0, side, (n1, 2)).astype(np.float32)
    b = rng.uniform(0, side, (n2, 2)).astype(np.float32)
    d = np.linalg.norm((b[None, :, :] - a[:, None, :]).astype(np.float64), axis=2)
    want = io.solve_gnn(d, gate)
    for n_warps in (0, 7, 64):          # block-wide searches only / speculative parallel batches first
        got, stats = solve_sparse(gnn_lib, d, gate, n_warps)
        assert len(got) == len(want), (n_warps, len(got), len(want), stats)
        cw, cg = sum(d[i, j] for i, j in want), sum(d[i, j] for i, j in got)
        assert abs(cw - cg) <= 1e-9 * max(1.0, cw), (n_warps, cw, cg)
        assert got == sorted(want), (n_warps, stats)


def test_clustered_points_with_many_unassigned_rows(gnn_lib):
    """Rows outnumber the columns they can reach: most searches end with a row left unassigned, which one is decided by cost."""
    rng = np.random.RandomState(77)
    cent = rng.uniform(0, 1000, (12, 2))
    a = (cent[rng.randint(0, 12, 240)] + rng.normal(scale=15, size=(240, 2))).astype(np.float32)
    b = (cent[rng.randint(0, 12, 90)] + rng.normal(scale=15, size=(90, 2))).astype(np.float32)
    d = np.linalg.norm((b[None, :, :] - a[:, None, :]).astype(np.float64), axis=2)
    want = io.solve_gnn(d, 40.0)
    for n_warps in (0, 16):
        got, stats = solve_sparse(gnn_lib, d, 40.0, n_warps)
        assert got == sorted(want), (n_warps, stats)


@pytest.mark.parametrize("seed", range(8))
def test_ties_give_the_optimal_cardinality_and_cost(gnn_lib, seed):
    """Degenerate costs (a handful of distinct integer values, many zero-reduced-cost alternating cycles): the assignment is
    not unique, but its cardinality and total cost must be the reference formulation's optimum -- with and without the
    speculative batches, which must also agree with each other run to run (the outcome depends on the data only)."""
    rng = np.random.RandomState(100 + seed)
    n1, n2 = rng.randint(20, 120), rng.randint(20, 120)
    d = rng.randint(1, 5, size=(n1, n2)).astype(np.float64)
    d[rng.uniform(size=(n1, n2)) > 0.08] = 1e9          # sparse gate
    want = io.solve_gnn(d, 10.0)
    cw = sum(d[i, j] for i, j in want)
    outs = []
    for n_warps in (0, 5, 64, 64):
        got, stats = solve_sparse(gnn_lib, d, 10.0, n_warps)
        assert len(set(j for _, j in got)) == len(got)                       # a matching
        assert all(d[i, j] <= 10.0 for i, j in got)
        assert len(got) == len(want), (n_warps, len(got), len(want), stats)
        assert sum(d[i, j] for i, j in got) == cw, (n_warps, stats)
        outs.append(got)
    assert outs[2] == outs[3]


def _brute_force(d, gate):
    """All matchings of the gated pairs of a tiny problem: (maximum cardinality, then minimum total cost)."""
    n1, n2 = d.shape
    best = (0, 0.0)

    def rec(i, used, card, cost):
        nonlocal best
        if i == n1:
            if card > best[0] or (card == best[0] and cost < best[1] - 1e-12):
                best = (card, cost)
            return
        rec(i + 1, used, card, cost)
        for j in range(n2):
            if not (used >> j) & 1 and d[i, j] <= gate:
                rec(i + 1, used | (1 << j), card + 1, cost + d[i, j])
    rec(0, 0, 0, 0.0)
    return best


@pytest.mark.parametrize("seed", range(12))
def test_tiny_problems_against_brute_force(gnn_lib, seed):
    """Independent of SciPy and of the reference's padding construction: on problems small enough to enumerate, the core's
    assignment has the maximum cardinality and, among those, the minimum cost -- what m_of_n.py:24-104 computes."""
    rng = np.random.RandomState(900 + seed)
    n1, n2 = rng.randint(1, 7), rng.randint(1, 7)
    d = rng.uniform(0, 10, (n1, n2))
    gate = 6.0
    card, cost = _brute_force(d, gate)
    for n_warps in (0, 3):
        got, stats = solve_sparse(gnn_lib, d, gate, n_warps)
        assert len(got) == card, (n_warps, got, card)
        assert abs(sum(d[i, j] for i, j in got) - cost) < 1e-9, (n_warps, got, cost)
    want = io.solve_gnn(d, gate)                      # and the oracle's restatement agrees with the enumeration too
    assert len(want) == card and abs(sum(d[i, j] for i, j in want) - cost) < 1e-9
