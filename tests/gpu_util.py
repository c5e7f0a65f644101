"""Helpers for the GPU parity tests: device buffers via torch, oracle problem extraction."""
import ctypes as C

import numpy as np

from pymht_b200 import _lib


def model_from_oracle(mo, T, eta2, lambda_ex):
    A, Q, Cm, R, P0 = mo.cv_model(T)
    return _lib.Model.from_arrays(A, Q, Cm, R, eta2, lambda_ex), (A, Q, Cm, R, P0)


def gate_batch_host(model, x0, P0, Pd, cnllr, z, cap=None):
    lib = _lib.load()
    L, M = x0.shape[0], z.shape[0]
    cap = cap if cap is not None else max(16, 64 * L)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    P0 = np.ascontiguousarray(P0, dtype=np.float32)
    Pd = np.ascontiguousarray(np.broadcast_to(Pd, (L,)), dtype=np.float64)
    cnllr = np.ascontiguousarray(cnllr, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = dict(x_bar=np.zeros((L, 4)), P_bar=np.zeros((L, 4, 4), np.float32), P_hat=np.zeros((L, 4, 4), np.float32),
               miss=np.zeros(L), off=np.zeros(L + 1, np.int32), meas=np.zeros(cap, np.int32), cnllr=np.zeros(cap),
               xhat=np.zeros((cap, 4)), used=np.zeros(max(M, 1), np.uint8))
    rc = lib.mht_gate_batch_host(C.byref(model), L, M, _lib.ptr(x0), _lib.ptr(P0), _lib.ptr(Pd), _lib.ptr(cnllr),
                                 _lib.ptr(z), _lib.ptr(out["x_bar"]), _lib.ptr(out["P_bar"]), _lib.ptr(out["P_hat"]),
                                 _lib.ptr(out["miss"]), _lib.ptr(out["off"]), _lib.ptr(out["meas"]),
                                 _lib.ptr(out["cnllr"]), _lib.ptr(out["xhat"]), cap, _lib.ptr(out["used"]))
    out["rc"] = rc
    return out


def oracle_columns(trk):
    """Global association problem (all trees) of an OracleTracker after _grow: cost, tree, rows[w][n]."""
    cost, tree, rows, rowid = [], [], [], {}
    for t, (root, leaves) in enumerate(zip(trk.roots, trk.leaves)):
        for leaf in leaves:
            r, n = [], leaf
            while n is not root:
                if n.meas:
                    r.append(rowid.setdefault((n.scan, n.meas), len(rowid)))
                n = n.parent
            rows.append(r)
            cost.append(leaf.cnllr - root.cnllr)
            tree.append(t)
    n = len(cost)
    width = max([len(r) for r in rows] + [1])
    RM = -np.ones((width, n), dtype=np.int32)
    for j, r in enumerate(rows):
        RM[:len(r), j] = r
    return np.array(cost), np.array(tree, dtype=np.int32), RM, max(len(rowid), 1)


def assoc_solve_device(cost, tree, RM, n_trees, n_rows):
    import torch
    lib = _lib.load()
    n, width = len(cost), RM.shape[0]
    dev = torch.device("cuda:0")
    d_cost = torch.from_numpy(np.ascontiguousarray(cost, dtype=np.float64)).to(dev)
    d_tree = torch.from_numpy(np.ascontiguousarray(tree, dtype=np.int32)).to(dev)
    d_rows = torch.from_numpy(np.ascontiguousarray(RM, dtype=np.int32)).to(dev)
    d_sel = torch.full((n_trees,), -7, dtype=torch.int32, device=dev)
    wbytes = lib.mht_assoc_workspace(n, n_trees, n_rows, width)
    d_work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    info = np.zeros(8)
    torch.cuda.synchronize()
    rc = lib.mht_assoc_solve(n, n_trees, n_rows, width, d_cost.data_ptr(), d_tree.data_ptr(), d_rows.data_ptr(),
                             d_sel.data_ptr(), _lib.ptr(info), d_work.data_ptr(), None)
    torch.cuda.synchronize()
    return rc, d_sel.cpu().numpy(), info


def cluster_device(tree, RM, n_trees, n_rows):
    import torch
    lib = _lib.load()
    n, width = len(tree), RM.shape[0]
    dev = torch.device("cuda:0")
    d_tree = torch.from_numpy(np.ascontiguousarray(tree, dtype=np.int32)).to(dev)
    d_rows = torch.from_numpy(np.ascontiguousarray(RM, dtype=np.int32)).to(dev)
    d_lab = torch.full((n_trees,), -7, dtype=torch.int32, device=dev)
    d_work = torch.empty(lib.mht_assoc_workspace(n, n_trees, n_rows, width), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    _lib.check(lib.mht_cluster(n, n_trees, n_rows, width, d_tree.data_ptr(), d_rows.data_ptr(), d_lab.data_ptr(),
                               d_work.data_ptr(), None))
    torch.cuda.synchronize()
    return d_lab.cpu().numpy()
