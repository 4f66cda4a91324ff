"""The reference's OWN warm-start shift, executed here: ``SCP_SLS.reset_warm_start`` (``solver/SCP_SLS_jit.py:500-551``) and
``OCP.pack_primal_nominal`` / ``unpack_primal_nominal`` (``solver/ocp.py:43-74``) are called as unbound methods of the unmodified
reference classes on an attribute bag whose model is the reference's own ``dyn/*.py`` class (numeric CasADi stand-in,
``oracle/ref_models.py``).  ``SCP_SLS`` itself cannot be constructed without CasADi's symbolic differentiation; this method needs
none.  Pins rows a11 and a12 of SURVEY.md section 8.  Works only where ``/root/reference`` exists; the outputs are committed in
``tests/golden/refshift_*.npz`` (recipe: ``python -m oracle.ref_scp``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

from . import models as om
from . import ref_kernels, ref_models
from .ref_models import REFERENCE_ROOT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load():
    """The reference's ``solver.SCP_SLS_jit`` module (import stubs as in oracle/ref_kernels.py, plus IPOPT's front end)."""
    ref_kernels.load()                                    # casadi / osqp / prettytable stubs, solver.fast_SLS_jit imported
    for attr in ("DM", "SX", "Function"):                 # annotations / defaults evaluated at class-definition time
        if not hasattr(sys.modules["casadi"], attr):
            setattr(sys.modules["casadi"], attr, object)
    ours = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("dyn",) and REFERENCE_ROOT not in (getattr(v, "__file__", "") or "")}
    for k in ours:
        del sys.modules[k]
    if "solver.nlp" not in sys.modules:
        sys.modules["solver.nlp"] = types.ModuleType("solver.nlp")
        sys.modules["solver.nlp"].NLP = object
    added = REFERENCE_ROOT not in sys.path
    if added:
        sys.path.insert(0, REFERENCE_ROOT)
    prev = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        return importlib.import_module("solver.SCP_SLS_jit")
    finally:
        sys.dont_write_bytecode = prev
        for k in [k for k in sys.modules if k.split(".")[0] == "dyn"]:
            del sys.modules[k]
        sys.modules.update(ours)
        if added:
            sys.path.remove(REFERENCE_ROOT)


def reference_shift(name, X, U, ubg):
    """X (nx, N+1), U (nu, N), ubg of the QP -> what reset_warm_start leaves in current_iteration_scp."""
    mod = load()
    ref_model = ref_models.load()[name]()
    N = U.shape[1]
    model = SimpleNamespace(nx=ref_model.nx, nu=ref_model.nu, ni=ref_model.ni, ni_f=ref_model.ni_f,
                            ddyn=lambda x, u: ref_models.ddyn(ref_model, x, u))      # CasADi's (n,) + (n,1) coercion made explicit
    calls = []
    bag = SimpleNamespace(N=N, m=model, current_iteration_scp=dict(primal_x=X.copy(), primal_u=U.copy()),
                          fast_SLS_solver=SimpleNamespace(solver_forward=SimpleNamespace(ubg=ubg.copy()),
                                                          reset_solver_to_zeros=lambda: calls.append("reset_solver_to_zeros")))
    bag.pack_primal_nominal = lambda Xn, Un: mod.OCP.pack_primal_nominal(bag, Xn, Un)
    mod.SCP_SLS.reset_warm_start(bag)
    cur = bag.current_iteration_scp
    Xb, Ub = mod.OCP.unpack_primal_nominal(bag, cur["primal_vec"])
    return dict(X=np.array(cur["primal_x"]), U=np.array(cur["primal_u"]), primal_vec=np.array(cur["primal_vec"], float).ravel(),
                pending=np.array(cur["pending_ubg_shift_ineq"]), unpacked_X=np.array(Xb, float), unpacked_U=np.array(Ub, float),
                calls=np.array(calls))


def main():
    for name in ("pendulum", "quadrotor", "rocket"):
        m, cfg = om.CONFIGS[name]()
        N = cfg["N"]
        rng = np.random.default_rng(3)
        X = rng.uniform(m.x_lb[:, None] / 3, m.x_ub[:, None] / 3, size=(m.nx, N + 1))
        U = rng.uniform(m.u_lb[:, None], m.u_ub[:, None], size=(m.nu, N))
        ubg = rng.normal(size=N * (m.nx + m.ni) + m.ni_f)
        out = reference_shift(name, X, U, ubg)
        np.savez(os.path.join(OUT, f"refshift_{name}.npz"), X_in=X, U_in=U, ubg_in=ubg,
                 source="reference SCP_SLS.reset_warm_start + OCP.pack/unpack_primal_nominal executed through oracle/ref_scp.py", **out)
        print(name, "written", out["X"].shape, out["primal_vec"].shape, out["calls"])


if __name__ == "__main__":
    main()
