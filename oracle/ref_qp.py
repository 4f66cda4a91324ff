"""The reference's OWN ``QP`` class (``solver/qp_jit.py:22-628``), constructed and driven here with backend ``"osqp"``:
the module is imported unmodified with (i) a numeric stand-in for the CasADi calls of its assembly code (``DM``, ``vertcat``,
``kron``; ``SX.sym`` / ``Function`` only build the unused CasADi-backend helpers) and (ii) a stand-in ``osqp`` module whose
``OSQP.setup / update / solve`` record exactly what the reference hands over (P, q, A, l, u, settings) and solve it with the
restated algorithm of ``oracle/osqp_restated.py``.  So the reference's statements do the assembly (``_assemble_structures``,
``_reassemble_numeric_same_sparsity``), the bound bookkeeping (``offset_constraints``, ``update_ubg``, the x0 rows and the
+-1e20 sanitising inside ``solve``) and the unpacking (``_pack_solution``); only the ADMM iteration is ours.

Pins row a3 of SURVEY.md section 8 -- "is the QP we solve the reference's QP?" -- against the reference itself: the recorded
(P, A, l, u) must equal ``oracle/qp.py``'s dense assembly entry for entry.  Works only where ``/root/reference`` exists;
``oracle/gen_golden_ref_qp.py`` commits the recorded problems and returned dictionaries as ``tests/golden/refqp_*.npz``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import importlib
import sys
import types
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

from . import osqp_restated
from .ref_models import REFERENCE_ROOT, _Anything, available  # noqa: F401


class DM(np.ndarray):
    """Dense numeric stand-in for casadi.DM (always 2-D)."""

    def __new__(cls, a=0.0, *rest):
        if sp.issparse(a):
            a = a.toarray()
        a = np.array(a, dtype=float)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        return a.view(cls)

    def full(self):
        return np.asarray(self)

    def sparsity(self):
        return np.asarray(self) != 0

    @staticmethod
    def zeros(n, m=1):
        return DM(np.zeros(n if isinstance(n, tuple) else (n, m)))

    @staticmethod
    def ones(n, m=1):
        return DM(np.ones(n if isinstance(n, tuple) else (n, m)))

    @staticmethod
    def inf(n, m=1):
        return DM(np.full(n if isinstance(n, tuple) else (n, m), np.inf))


def _casadi():
    ca = types.ModuleType("casadi")
    ca.DM = DM
    ca.vertcat = lambda *args: DM(np.vstack([DM(a) for a in args])) if args else DM.zeros(0, 1)
    ca.kron = lambda a, b: DM(np.kron(DM(a), DM(b)))
    ca.SX = SimpleNamespace(sym=lambda name, n=1, m=1: DM.zeros(n, m))
    ca.Function = lambda *a, **k: (lambda *x: None)
    ca.conic = lambda *a, **k: None
    return ca


class _Recorder:
    """Everything the reference handed to ``osqp.OSQP`` objects, in order."""

    def __init__(self):
        self.setups, self.solves = [], []


def _osqp(rec, solver_settings):
    mod = types.ModuleType("osqp")

    class OSQP:
        def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
            self.P, self.A = sp.csc_matrix(P).copy(), sp.csc_matrix(A).copy()
            self.q, self.l, self.u = np.array(q, float), np.array(l, float), np.array(u, float)
            self.settings = dict(settings)
            rec.setups.append(dict(settings=dict(settings), P=self.P.toarray(), A=self.A.toarray()))

        def update(self, Ax=None, Px=None, q=None, l=None, u=None):
            if Ax is not None:
                self.A.data[:] = Ax
            if Px is not None:
                self.P.data[:] = Px
            for k, v in (("q", q), ("l", l), ("u", u)):
                if v is not None:
                    setattr(self, k, np.array(v, float))

        def solve(self):
            s = dict(solver_settings)
            P, A = self.P.toarray(), self.A.toarray()
            x, y, info = osqp_restated.solve(P, self.q, A, self.l, self.u, **s)
            rec.solves.append(dict(P=P, A=A, q=self.q.copy(), l=self.l.copy(), u=self.u.copy(), settings=dict(self.settings),
                                   x=x.copy(), y=y.copy(), status=info["status"], polished=info["polished"]))
            code = {"solved": 1, "solved inaccurate": 2, "primal infeasible": -3}.get(info["status"], -2)
            return SimpleNamespace(x=x, y=y, info=SimpleNamespace(status_val=code, status=info["status"], obj_val=info["obj_val"]))
    mod.OSQP = OSQP
    return mod


def load(rec, solver_settings=None):
    """Imports the reference's ``solver.qp_jit`` over the stand-ins; returns (module, LTV class)."""
    if not available():
        raise RuntimeError("reference tree not present; tests/golden/refqp_*.npz are the pinned outputs")
    names = ("casadi", "osqp", "matplotlib", "matplotlib.pyplot")
    saved = {k: sys.modules.get(k) for k in names}
    ours = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("dyn", "solver", "util")}
    for k in ours:
        del sys.modules[k]
    sys.modules["casadi"] = _casadi()
    sys.modules["osqp"] = _osqp(rec, solver_settings or {})
    for k in ("matplotlib", "matplotlib.pyplot"):
        if saved[k] is None:
            sys.modules[k] = _Anything(k)
    sys.path.insert(0, REFERENCE_ROOT)
    prev = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        mod = importlib.import_module("solver.qp_jit")
        ltv = importlib.import_module("dyn.LTV").LTV
    finally:
        sys.dont_write_bytecode = prev
        for k in [k for k in sys.modules if k.split(".")[0] in ("dyn", "solver", "util")]:
            del sys.modules[k]
        sys.modules.update(ours)
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, ltv


def reference_qp(N, m, Q, R, Qf, rec, solver_settings=None):
    """``QP(N, Q, R, LTV(m, N), Qf, backend='osqp')`` of the reference (as ``fast_SLS.__init__`` builds it, fast_SLS_jit.py:228-241)."""
    mod, LTV = load(rec, solver_settings)
    model = SimpleNamespace(nx=m.nx, nu=m.nu, nw=m.nw, ni=m.ni, ni_f=m.ni_f, G=np.array(m.G, float), Gf=np.array(m.Gf, float),
                            gf=np.array(m.gf, float), g=np.array(m.g, float))
    return mod.QP(N, np.array(Q, float), np.array(R, float), LTV(model, N), np.array(Qf, float), backend="osqp", verbose=False)
