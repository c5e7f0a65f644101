"""Import shim that makes the UNMODIFIED reference (/root/reference) importable read-only.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it.

The reference cannot be imported as-is in this image:
  * matplotlib, termcolor, munkres, ortools, pykalman are absent;
  * NumPy 2.x removed np.int / np.Inf / np.NaN (reference tracker.py:1004,1111;
    m_of_n.py:24; pyTarget.py:588);
  * reference tracker.py:29-30 asserts the NumPy *minor* version >= 12.
This module installs throw-away stand-ins in sys.modules and never edits or copies the reference.
The `pywraplp` facade solves the reference's own BLP (tracker.py:1155-1217) exactly with
scipy.optimize.milp (HiGHS, relative MIP gap 0) because OR-Tools/CBC cannot be installed offline.

/root/reference does not exist on the GPU box; `available()` says whether the live reference
can be used.  Golden vectors produced with it are committed under tests/golden/.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PYMHT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pymht"))


class _Anything:
    """Permissive dummy: any attribute / call / index returns another dummy."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    def _fallback(attr):  # PEP 562 module-level fallback; dunders (__file__, __path__ ...) stay absent so that
        if attr.startswith("__") and attr.endswith("__"):   # inspect / importlib treat the stub like a builtin
            raise AttributeError(attr)
        return _Anything()
    mod.__getattr__ = _fallback
    sys.modules[name] = mod
    return mod


# ---------------------------------------------------------------------------------------
# pywraplp facade: exactly the surface reference tracker.py:1167-1210 touches.
# ---------------------------------------------------------------------------------------
class _Var:
    __slots__ = ("index", "value")

    def __init__(self, index):
        self.index = index
        self.value = 0.0

    def solution_value(self):
        return self.value

    def __mul__(self, coef):
        return _Term(float(coef), self.index)

    __rmul__ = __mul__


class _Term:
    __slots__ = ("coef", "index")

    def __init__(self, coef, index):
        self.coef = coef
        self.index = index


class _LinExpr:
    def __init__(self, terms):
        self.terms = terms

    def __le__(self, rhs):
        return _Constraint(self.terms, -np.inf, float(rhs))

    def __eq__(self, rhs):  # noqa: used as constraint builder, like OR-Tools
        return _Constraint(self.terms, float(rhs), float(rhs))

    __hash__ = None


class _Constraint:
    def __init__(self, terms, lo, hi):
        self.terms, self.lo, self.hi = terms, lo, hi


class _Solver:
    CBC_MIXED_INTEGER_PROGRAMMING = 1
    OPTIMAL = 0
    FEASIBLE = 1
    INFEASIBLE = 2

    def __init__(self, name, backend):
        self._vars = []
        self._cons = []
        self._obj = None
        self._wall_ms = 0.0

    def BoolVar(self, name):
        v = _Var(len(self._vars))
        self._vars.append(v)
        return v

    def Sum(self, items):
        terms = []
        for it in items:
            if isinstance(it, _Var):
                terms.append(_Term(1.0, it.index))
            else:
                terms.append(it)
        return _LinExpr(terms)

    def Minimize(self, expr):
        self._obj = expr

    def Add(self, constraint):
        self._cons.append(constraint)

    def WallTime(self):
        return self._wall_ms

    def Solve(self):
        import time
        from scipy.optimize import milp, LinearConstraint, Bounds
        from scipy.sparse import csr_matrix

        t0 = time.time()
        n = len(self._vars)
        c = np.zeros(n)
        for t in self._obj.terms:
            c[t.index] += t.coef
        rows, cols, vals, lo, hi = [], [], [], [], []
        for r, con in enumerate(self._cons):
            for t in con.terms:
                rows.append(r)
                cols.append(t.index)
                vals.append(t.coef)
            lo.append(con.lo)
            hi.append(con.hi)
        A = csr_matrix((vals, (rows, cols)), shape=(len(self._cons), n))
        res = milp(c, constraints=LinearConstraint(A, lo, hi), integrality=np.ones(n),
                   bounds=Bounds(0, 1), options={"mip_rel_gap": 0.0})
        self._wall_ms = (time.time() - t0) * 1e3
        if res.x is None:
            return self.INFEASIBLE
        for v, x in zip(self._vars, np.round(res.x)):
            v.value = float(x)
        return self.OPTIMAL if res.status == 0 else self.FEASIBLE


def _munkres(cost):
    """Contract of reference m_of_n.py:63-67: boolean assignment matrix of a cost matrix."""
    from scipy.optimize import linear_sum_assignment
    cost = np.asarray(cost, dtype=float)
    r, c = linear_sum_assignment(cost)
    out = np.zeros(cost.shape, dtype=bool)
    out[r, c] = True
    return out


_INSTALLED = False


def install():
    """Idempotently install stubs + aliases and put the reference on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for alias, target in (("int", int), ("Inf", np.inf), ("NaN", np.nan), ("float", float)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    mpl = _stub_module("matplotlib", get_backend=lambda: "Agg", use=lambda *a, **k: None)
    mpl.pyplot = _stub_module("matplotlib.pyplot")
    mpl.patches = _stub_module("matplotlib.patches")
    mpl.cm = _stub_module("matplotlib.cm")
    _stub_module("termcolor", cprint=lambda *a, **k: None)
    _stub_module("munkres", munkres=_munkres)
    ort = _stub_module("ortools")
    ort.linear_solver = _stub_module("ortools.linear_solver")
    ort.linear_solver.pywraplp = _stub_module("ortools.linear_solver.pywraplp", Solver=_Solver)
    import scipy.sparse.csgraph  # noqa: F401  (must be imported before the version spoof)
    import scipy.stats  # noqa: F401
    import scipy.optimize  # noqa: F401
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    real_version = np.__version__
    np.__version__ = "1.26.0"  # reference tracker.py:29-30 checks the minor version only
    try:
        import pymht.tracker  # noqa: F401
        import pymht.utils.simulator  # noqa: F401
    finally:
        np.__version__ = real_version
    _INSTALLED = True


class NullInitiator:
    """Benchmark hygiene: M-of-N initiation is out of scope."""

    def processMeasurements(self, *a, **k):
        return []


def make_reference_tracker(radarPeriod, lambda_phi, lambda_nu, **kw):
    """Reference Tracker with the initiator nulled and mergeThreshold=0 (exactly T trees)."""
    install()
    import pymht.tracker as rtracker
    import pymht.models.pv as rpv
    trk = rtracker.Tracker(rpv, radarPeriod, lambda_phi, lambda_nu, **kw)
    trk.initiator = NullInitiator()
    trk.mergeThreshold = 0.0
    return trk
