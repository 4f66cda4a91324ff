"""The reference's OWN model sources, executed here: ``dyn/pendulum.py``, ``dyn/quadrotor.py``, ``dyn/rocket.py`` and
``dyn/model.py`` are imported unmodified from ``/root/reference`` with a NUMERIC stand-in for the dozen CasADi
functions they call (``vertcat, hcat, vertsplit, mtimes, cross, inv, diag, sin, cos, atan, sqrt, SX.sym``: CasADi
3.7.1 is absent; every one of these is plain element-wise or small-matrix arithmetic, restated on NumPy below).
``ode(X, u)`` and ``Model.ddyn(x, u, h)`` (``dyn/model.py:15-34``) then run the reference's statements on numbers.

Purpose: pin the TRANSCRIPTION of the ODEs (``oracle/models.py``, ``csrc/rnmpc_models.cuh``) and the constant tables
(bounds, E, parameters) against the reference itself rather than against a reading of it: ``oracle/gen_golden_ref_models.py``
writes ``tests/golden/refdyn_*.npz`` (inputs, ``ode``, ``ddyn``, ``g``, ``gf``, ``E``); the CPU tests compare the
oracle with them, the GPU tests the CUDA kernels.  Like ``oracle/ref_kernels.py`` this works only where the reference
tree exists (the build container); the fixtures travel.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RNMPC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dyn"))


def _mat(a):
    """CasADi semantics: everything is a 2-D matrix; scalars are 1 x 1, flat sequences are columns."""
    a = np.asarray(a)
    if not np.iscomplexobj(a):
        a = a.astype(float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)
    return a


def _numeric_casadi():
    ca = types.ModuleType("casadi")
    ca.vertcat = lambda *args: np.vstack([_mat(a) for a in args])
    ca.hcat = lambda args: np.hstack([_mat(a) for a in args])
    ca.vertsplit = lambda a: [r.reshape(1, 1) for r in _mat(a)]
    ca.mtimes = lambda a, b: _mat(a) @ _mat(b)
    ca.cross = lambda a, b: np.cross(_mat(a).ravel(), _mat(b).ravel()).reshape(-1, 1)
    ca.inv = lambda a: np.linalg.inv(_mat(a))
    ca.diag = lambda a: np.diag(_mat(a).ravel())
    ca.sin, ca.cos, ca.atan, ca.sqrt = np.sin, np.cos, np.arctan, np.sqrt

    class SX:                                   # only used to build name lists / look up bounds by name in the constructors
        def __init__(self, name):
            self._name = name

        def name(self):
            return self._name

        @staticmethod
        def sym(name, *shape):
            return SX(name)
    ca.SX = SX
    return ca


class _Anything(types.ModuleType):
    """Stand-in for the plotting imports of the model files (never called on this path)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(name)

    def __call__(self, *a, **k):
        return _Anything("call")


def load():
    """Returns {'pendulum': Pendulum, 'quadrotor': Quadrotor, 'rocket': Rocket}: the reference's classes."""
    if not available():
        raise RuntimeError("reference tree not present; tests/golden/refdyn_*.npz are the pinned outputs")
    saved = {k: sys.modules.get(k) for k in ("casadi", "matplotlib", "matplotlib.pyplot")}
    sys.modules["casadi"] = _numeric_casadi()
    for k in ("matplotlib", "matplotlib.pyplot"):
        if saved[k] is None:
            sys.modules[k] = _Anything(k)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    prev_dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True              # /root/reference is read-only
    mirror = {k: v for k, v in sys.modules.items() if k == "dyn" or k.startswith("dyn.")}
    try:
        for k in mirror:
            del sys.modules[k]                  # the product's mirror package has the same module names
        out = dict(pendulum=importlib.import_module("dyn.pendulum").Pendulum,
                   quadrotor=importlib.import_module("dyn.quadrotor").Quadrotor,
                   rocket=importlib.import_module("dyn.rocket").Rocket)
    finally:
        sys.dont_write_bytecode = prev_dont_write
        for k in [k for k in sys.modules if k == "dyn" or k.startswith("dyn.")]:
            del sys.modules[k]
        sys.modules.update(mirror)
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return out


def ode(model, x, u):
    """``model.ode`` of a reference instance on numbers: x (nx,), u (nu,) -> (nx,)."""
    return np.asarray(model.ode(_mat(x), _mat(u)), float).ravel()


def ddyn(model, x, u, h=0.05):
    """``Model.ddyn`` (dyn/model.py:15-34; RK4 unless the instance says 'euler')."""
    return np.asarray(model.ddyn(_mat(x), _mat(u), h), float).ravel()


def jacobians(model, x, u, h=0.05, step=1e-30):
    """A = d ddyn / dx, B = d ddyn / du of the reference's own discretised map by the complex-step derivative
    (Im f(x + i step e_j) / step: no subtraction, exact to rounding for the analytic functions the models use) -- what the
    reference obtains from ``ca.jacobian`` (solver/SCP_SLS_jit.py:199-218)."""
    x, u = np.asarray(x, float), np.asarray(u, float)
    f = lambda xx, uu: np.asarray(model.ddyn(_mat(xx), _mat(uu), h)).ravel()
    A = np.empty((len(x), len(x)))
    B = np.empty((len(x), len(u)))
    for j in range(len(x)):
        xc = x.astype(complex)
        xc[j] += 1j * step
        A[:, j] = f(xc, u.astype(complex)).imag / step
    for j in range(len(u)):
        uc = u.astype(complex)
        uc[j] += 1j * step
        B[:, j] = f(x.astype(complex), uc).imag / step
    return A, B
