"""ctypes binding of libmht_b200.so (include/mht_b200.h).  There is NO CPU fallback: a missing
library or a missing sm_100 device raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmht_b200.so")

MHT_OK, MHT_E_INVALID, MHT_E_CUDA, MHT_E_CAPACITY, MHT_E_NODEVICE, MHT_E_NOTOPTIMAL = 0, -1, -2, -3, -4, -5
MAX_WINDOW = 16


class MhtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmht_b200 error %d: %s" % (code, msg))
        self.code = code


class Model(C.Structure):
    _fields_ = [("A", C.c_float * 16), ("Q", C.c_float * 16), ("C", C.c_float * 8), ("R", C.c_float * 4),
                ("eta2", C.c_double), ("lambda_ex", C.c_double)]

    @classmethod
    def from_arrays(cls, A, Q, Cm, R, eta2, lambda_ex):
        m = cls()
        for name, arr, n in (("A", A, 16), ("Q", Q, 16), ("C", Cm, 8), ("R", R, 4)):
            flat = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
            if flat.size != n:
                raise ValueError("model matrix %s has %d elements, expected %d" % (name, flat.size, n))
            getattr(m, name)[:] = flat.tolist()
        m.eta2, m.lambda_ex = float(eta2), float(lambda_ex)
        return m


class ForestConfig(C.Structure):
    _fields_ = [("model", Model), ("n_scan_window", C.c_int32), ("max_trees", C.c_int32), ("max_meas", C.c_int32),
                ("max_nodes", C.c_int64), ("max_parents", C.c_int64), ("default_Pd", C.c_double),
                ("score_upper", C.c_double), ("cnllr_upper", C.c_double), ("radar_range", C.c_double),
                ("position", C.c_double * 2), ("max_dual_iters", C.c_int32), ("exact_ms", C.c_int32)]


class ScanInfo(C.Structure):
    _fields_ = [("n_parents", C.c_int64), ("n_children", C.c_int64), ("n_pairs", C.c_int64),
                ("n_trees", C.c_int32), ("n_clusters", C.c_int32), ("n_multi_clusters", C.c_int32),
                ("n_dead", C.c_int32), ("dual_iters", C.c_int32), ("certified", C.c_int32),
                ("n_candidates", C.c_int64), ("bb_nodes", C.c_int64), ("lower_bound", C.c_double),
                ("objective", C.c_double), ("ms_gate", C.c_float), ("ms_cluster", C.c_float),
                ("ms_assoc", C.c_float), ("ms_prune", C.c_float), ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("n_active", C.c_int64), ("max_component", C.c_int32), ("n_components", C.c_int32),
                ("ms_dual", C.c_float), ("ms_exact", C.c_float), ("nnz_active", C.c_int64),
                ("rows_active", C.c_int32), ("bb_iters", C.c_int32), ("open_components", C.c_int32),
                ("repaired_trees", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GnnInfo(C.Structure):
    _fields_ = [("n_edges", C.c_int32), ("n_components", C.c_int32), ("largest_component", C.c_int32),
                ("n_assigned", C.c_int32), ("searches", C.c_int32), ("rounds", C.c_int32),
                ("batches", C.c_int32), ("spec_commits", C.c_int32), ("spec_overflow", C.c_int32),
                ("reserved", C.c_int32), ("ms_gate", C.c_float), ("ms_solve", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
_SIGNATURES = {
    "mht_version": (C.c_int, []),
    "mht_last_error": (C.c_char_p, []),
    "mht_device_count": (C.c_int, []),
    "mht_launch_count": (_i64, []),
    "mht_gate_batch_workspace": (_i64, [_i64, _i64]),
    "mht_gate_batch": (C.c_int, [C.POINTER(Model), _i64, _i64] + [_vp] * 13 + [_i64, _vp, _vp, _vp]),
    "mht_gate_batch_host": (C.c_int, [C.POINTER(Model), _i64, _i64] + [_vp] * 13 + [_i64, _vp]),
    "mht_assoc_workspace": (_i64, [_i64, _i64, _i64, _i32]),
    "mht_cluster": (C.c_int, [_i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mht_assoc_solve": (C.c_int, [_i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mht_assoc_solve_warm": (C.c_int, [_i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _dbl, _i32]),
    "mht_record_bytes": (_i32, [_i32]),
    "mht_forest_export_records": (C.c_int, [_vp, _i32, _vp, _i64]),
    "mht_unpack_records": (C.c_int, [_i32, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mht_forest_create": (C.c_int, [C.POINTER(ForestConfig), C.POINTER(_vp)]),
    "mht_forest_destroy": (None, [_vp]),
    "mht_forest_bytes": (_i64, [_vp]),
    "mht_forest_initiate": (C.c_int, [_vp, _vp, _vp, _dbl, C.POINTER(_i32)]),
    "mht_forest_release": (C.c_int, [_vp, _i32]),
    "mht_forest_scan": (C.c_int, [_vp, _i64, _vp, _dbl, C.POINTER(ScanInfo), _vp]),
    "mht_forest_scan_device": (C.c_int, [_vp, _i64, _vp, _dbl, C.POINTER(ScanInfo)]),
    "mht_forest_grow": (C.c_int, [_vp, _i64, _vp, _i32, _dbl, C.POINTER(ScanInfo), _vp]),
    "mht_forest_export_columns": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _vp]),
    "mht_forest_select": (C.c_int, [_vp, _vp, _vp, C.POINTER(ScanInfo)]),
    "mht_forest_tracks": (C.c_int, [_vp, _i32, C.POINTER(_i32), _vp, _vp, _vp, _vp, _vp, _vp]),
    "mht_forest_set_dynamic_window": (C.c_int, [_vp, _i32, _i32, _i32]),
    "mht_forest_windows": (C.c_int, [_vp, _i32, C.POINTER(_i32), _vp, _vp]),
    "mht_forest_history": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i32), _vp, _vp, _vp, _vp]),
    "mht_forest_histories": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i32), _vp, _vp, _vp, _vp, _vp, _vp]),
    "mht_forest_histories_of": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mht_forest_measurement_set": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i32), _vp, _vp]),
    "mht_forest_min_leaf_distance": (C.c_int, [_vp, _dbl, _dbl, C.POINTER(_dbl)]),
    "mht_forest_leaves": (C.c_int, [_vp, _i32, _i64, C.POINTER(_i64), _vp, _vp, _vp]),
    "mht_gnn_create": (C.c_int, [_i64, _i64, _i64, C.POINTER(_vp)]),
    "mht_gnn_destroy": (None, [_vp]),
    "mht_gnn_assign": (C.c_int, [_vp, C.c_int, _i64, _vp, _vp, _i64, _vp, _dbl, _vp, C.POINTER(GnnInfo)]),
    "mht_gnn_similar": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _dbl, _vp, _i64, C.POINTER(_i64)]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the in-tree CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("%s is missing: run `python -m pymht_b200.build` (needs nvcc). "
                              "pymht_b200 has no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, allow=()):
    if rc != MHT_OK and rc not in allow:
        raise MhtError(rc, load().mht_last_error().decode())
    return rc


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
