"""Import the reference's OWN Numba kernels (solver/fast_SLS_jit.py:43-188) under import stubs.

Only usable where ``/root/reference`` exists (the build container); used by ``oracle/gen_golden.py`` to pin the
NumPy restatement in ``oracle/sls.py`` and to produce ``tests/golden/sls_*.npz``.  ``casadi``, ``prettytable`` and
``osqp_generated`` are absent from the image and are stubbed with empty modules *before* the import; the four
kernels themselves are executed unmodified.  ``NUMBA_CACHE_DIR`` is pointed outside the (read-only) reference.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "solver"))


def load():
    """Returns the module ``solver.fast_SLS_jit`` of the reference (kernels: ``_backward_solve_numba``,
    ``_propagate``, ``_backoff_from_phi``, ``riccati_step_njit``)."""
    if not available():
        raise RuntimeError("reference tree not present; golden vectors under tests/golden/ are the pinned outputs")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/rnmpc_numba_cache")
    for name in ("casadi", "osqp_generated", "osqp"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "prettytable" not in sys.modules:
        pt = types.ModuleType("prettytable")
        pt.PrettyTable, pt.NONE, pt.HEADER = object, 0, 1
        sys.modules["prettytable"] = pt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("solver.fast_SLS_jit")
