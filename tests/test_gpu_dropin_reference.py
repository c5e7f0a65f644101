"""Drop-in proof: the UNMODIFIED reference `pymht.tracker.Tracker` with
exactly three method bodies replaced by ctypes stubs into libmht_b200.so --

    Tracker._processLeafNodes      (pymht/tracker.py:383-398)   -> mht_gate_batch_host
    Tracker._findClustersFromSets  (pymht/tracker.py:961-974)   -> mht_cluster
    Tracker._solveBLP_OR_TOOLS     (pymht/tracker.py:1155-1217) -> mht_assoc_solve

-- everything else (Target.spawnNewNodes, _createA1/_createA2/_createC, pruning, termination ...) stays the
reference's own Python.  Replaying the reference-generated fixtures must reproduce the reference's tracks.
The reference is taken from /root/reference when present, else from the installed copy baseline/_ref (which
travels to the GPU box); without either the test is skipped."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu

_REF = "/root/reference" if os.path.isdir("/root/reference/pymht") else os.path.join(ROOT, "baseline", "_ref")


def _reference_tracker(*a, **kw):
    os.environ["PYMHT_REFERENCE_ROOT"] = _REF
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = _REF
    if not ref_shim.available():
        pytest.skip("no reference tree (/root/reference or baseline/_ref)")
    return ref_shim.make_reference_tracker(*a, **kw)


def _install_stubs(trk, calls):
    """The three stubs of INTEGRATION.md section 2, bound to ONE reference Tracker instance."""
    import types
    import torch
    import pymht.models.pv as pv
    from pymht_b200 import _lib
    lib = _lib.load()
    p = _lib.ptr
    dev = torch.device("cuda:0")

    def _processLeafNodes(self, targetNodes, scanList, aisList):                       # tracker.py:383-398
        L, z = len(targetNodes), np.ascontiguousarray(scanList.measurements, dtype=np.float64)
        M, cap = len(z), max(64 * len(targetNodes), 64)
        m = _lib.Model.from_arrays(self.A, self.Q, pv.C_RADAR, self.R_RADAR, self.eta2, self.lambda_ex)
        x0 = np.ascontiguousarray([n.x_0 for n in targetNodes], dtype=np.float64).reshape(L, 4)
        P0 = np.ascontiguousarray([n.P_0 for n in targetNodes], dtype=np.float32).reshape(L, 4, 4)
        Pd = np.ascontiguousarray([n.P_d for n in targetNodes], dtype=np.float64)
        cn = np.ascontiguousarray([n.cumulativeNLLR for n in targetNodes], dtype=np.float64)
        while True:
            x_bar, P_bar, P_hat = np.empty((L, 4)), np.empty((L, 4, 4), np.float32), np.empty((L, 4, 4), np.float32)
            miss, off = np.empty(L), np.empty(L + 1, np.int32)
            meas, pcn, xhat = np.empty(cap, np.int32), np.empty(cap), np.empty((cap, 4))
            used = np.empty(max(M, 1), np.uint8)
            rc = lib.mht_gate_batch_host(C.byref(m), L, M, p(x0), p(P0), p(Pd), p(cn), p(z), p(x_bar), p(P_bar),
                                         p(P_hat), p(miss), p(off), p(meas), p(pcn), p(xhat), cap, p(used))
            if rc == _lib.MHT_E_CAPACITY:
                cap = int(off[L]) + 16
                continue
            _lib.check(rc)
            break
        calls["gate"] += 1
        idx = [meas[off[i]:off[i + 1]].astype(np.int64) for i in range(L)]          # gated_indices_list (:832)
        xh = [xhat[off[i]:off[i + 1]] for i in range(L)]                             # gated_x_hat_list   (:841)
        nllr = [pcn[off[i]:off[i + 1]] - cn[i] for i in range(L)]                    # nllr_list          (:846)
        empty = [np.array([]) for _ in range(L)]
        return (x_bar, P_bar), (xh, P_hat, idx, nllr), (empty, empty, empty, empty, empty)

    def _findClustersFromSets(self):                                                  # tracker.py:961-974
        sets = self.__associatedMeasurements__
        nT = len(sets)
        rowid, tree, rows = {}, [], []
        for t, st in enumerate(sets):
            for key in st:
                tree.append(t)
                rows.append(rowid.setdefault(key, len(rowid)))
            tree.append(t)                 # every tree needs at least one column
            rows.append(-1)
        order = np.argsort(np.asarray(tree), kind="stable")
        tree = np.ascontiguousarray(np.asarray(tree, dtype=np.int32)[order])
        RM = np.ascontiguousarray(np.asarray(rows, dtype=np.int32)[order]).reshape(1, -1)
        n, nR = len(tree), max(len(rowid), 1)
        d_tree, d_rows = torch.from_numpy(tree).to(dev), torch.from_numpy(RM).to(dev)
        d_lab = torch.empty(nT, dtype=torch.int32, device=dev)
        d_work = torch.empty(lib.mht_assoc_workspace(n, nT, nR, 1), dtype=torch.uint8, device=dev)
        _lib.check(lib.mht_cluster(n, nT, nR, 1, d_tree.data_ptr(), d_rows.data_ptr(), d_lab.data_ptr(),
                                   d_work.data_ptr(), None))
        torch.cuda.synchronize()
        lab = d_lab.cpu().numpy()
        calls["cluster"] += 1
        return [np.where(lab == c)[0] for c in np.unique(lab)]

    def _solveBLP_OR_TOOLS(self, A1, A2, f):                                          # tracker.py:1155-1217
        A1, A2 = np.asarray(A1, dtype=bool), np.asarray(A2, dtype=bool)
        nR, n = A1.shape
        nT = A2.shape[0]
        tree = np.ascontiguousarray(np.argmax(A2, axis=0), dtype=np.int32)
        width = max(int(A1.sum(axis=0).max()), 1) if nR else 1
        RM = -np.ones((width, n), dtype=np.int32)
        rr, cc = np.nonzero(A1.T)                      # (column, row) pairs, columns ascending
        slot = np.zeros(n, dtype=np.int64)
        for j, r in zip(rr, cc):
            RM[slot[j], j] = r
            slot[j] += 1
        d_cost = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float64)).to(dev)
        d_tree, d_rows = torch.from_numpy(tree).to(dev), torch.from_numpy(np.ascontiguousarray(RM)).to(dev)
        d_sel = torch.empty(nT, dtype=torch.int32, device=dev)
        d_work = torch.empty(lib.mht_assoc_workspace(n, nT, max(nR, 1), width), dtype=torch.uint8, device=dev)
        info = np.zeros(8)
        _lib.check(lib.mht_assoc_solve(n, nT, max(nR, 1), width, d_cost.data_ptr(), d_tree.data_ptr(),
                                       d_rows.data_ptr(), d_sel.data_ptr(), p(info), d_work.data_ptr(), None))
        torch.cuda.synchronize()
        assert info[6] == 1
        calls["blp"] += 1
        return sorted(int(j) for j in d_sel.cpu().numpy())

    trk._processLeafNodes = types.MethodType(_processLeafNodes, trk)
    trk._findClustersFromSets = types.MethodType(_findClustersFromSets, trk)
    trk._solveBLP_OR_TOOLS = types.MethodType(_solveBLP_OR_TOOLS, trk)


@pytest.mark.parametrize("name", ["cfg1_crossing", "cfg2_small", "cfg5_small", "cfg2"])
def test_patched_reference_reproduces_reference_tracks(name):
    g = golden(name)
    T, lam_phi, lam_nu, N, Pd, eta2, R = [float(v) for v in g["params"]]
    trk = _reference_tracker(T, lam_phi, lam_nu, N=int(N), P_d=Pd, eta2=eta2)
    import pymht.models.pv as pv
    import pymht.utils.helpFunctions as hpf
    from pymht.pyTarget import Target
    from pymht.utils.classDefinitions import MeasurementList
    calls = {"gate": 0, "cluster": 0, "blp": 0}
    _install_stubs(trk, calls)
    for x in g["init_x"]:
        trk.initiateTarget(Target(float(g["init_time"]), None, np.asarray(x, dtype=np.float64), pv.P0,
                                  status="preinitialized"))
    for k in range(int(g["n_scans"])):
        pre = "s%d_" % k
        trk.addMeasurementList(MeasurementList(float(g[pre + "time"]), g[pre + "z"]))
        nodes = list(trk.getTrackNodes())
        assert [n.ID for n in nodes] == list(g[pre + "ids"]), (name, k)
        hist = hpf.backtrackMeasurementNumbers(nodes)
        H = g[pre + "hist"]
        for i, h in enumerate(hist):
            assert list(h) == list(H[i, :len(h)]), (name, k, i)
        np.testing.assert_allclose(np.array([n.x_0 for n in nodes]).reshape(-1, 4), g[pre + "x"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose([n.cumulativeNLLR for n in nodes], g[pre + "cnllr"], rtol=1e-5, atol=1e-6)
        assert len(trk.__clusterList__) == int(g[pre + "nclusters"])
    assert calls["gate"] > 0 and calls["cluster"] == int(g["n_scans"])
    if any(int(g["s%d_n_ilp" % k]) for k in range(int(g["n_scans"]))):
        assert calls["blp"] > 0
