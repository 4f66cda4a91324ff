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




# ---------------------------------------------------------------------------------------------------------------------------------
# SCP control flow (row a10): the reference's own SCP_SLS.solve / socp_step / check_convergence_scp / post_processing_solution /
# reset_warm_start (solver/SCP_SLS_jit.py:65-152, 404-483, 500-551, 610-622) run unbound on an attribute bag.  What they call that
# needs CasADi's symbolic differentiation or the generated OSQP module -- update_jacobian and the fast_SLS object -- is supplied by
# the oracle (whose pieces are pinned separately: refdyn_*, refglue_*, refqp_*, sls_*); everything else -- the RTI loop, the nominal
# update, cost, the signed-max primal_infeasibility, success / iterations bookkeeping, the result dictionary -- is the reference's.
class _Col(np.ndarray):
    """A CasADi DM column: combined with a flat NumPy vector it stays a column (CasADi coerces the (n,) operand to (n, 1)),
    where NumPy would broadcast to (n, n).  SCP_SLS_jit.py:452-454 relies on this."""
    __array_priority__ = 100.0

    def __sub__(self, other):
        other = np.asarray(other, float)
        return (np.asarray(self) - (other.reshape(-1, 1) if other.ndim == 1 else other)).view(_Col)

    def __add__(self, other):
        other = np.asarray(other, float)
        return (np.asarray(self) + (other.reshape(-1, 1) if other.ndim == 1 else other)).view(_Col)


def reference_closed_loop_steps(name, x0s, X, U, rti=1, sls_steps=1, fast="oracle", eps=1e-6, plant=None):
    """Runs solve(x0s[0]), reset_warm_start(), solve(x0s[1]), ... and returns the list of result dictionaries.
    fast = "reference": the fast_SLS object is the reference's own too (oracle/ref_fast_sls.py: its solve iteration, QP class and
    Numba kernels; ADMM restated), and only the VALUES of the linearisation come from the oracle.  plant: optional callable
    (step index, x, u0) -> next measured state; then x0s[0] starts a closed loop of len(x0s) steps (the other entries are ignored)."""
    import scipy.sparse as sp

    from . import ref_qp
    from . import scp as oscp
    mod = load()
    mod.ca.DM = ref_qp.DM
    ref_model = ref_models.load()[name]()
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    orc = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=rti, fast_sls_rti_steps=sls_steps)
    model = SimpleNamespace(nx=m.nx, nu=m.nu, ni=m.ni, ni_f=m.ni_f,
                            ddyn=lambda x, u: ref_models.ddyn(ref_model, x, u).reshape(-1, 1).view(_Col))
    bag = SimpleNamespace(N=N, m=model, Q=np.array(cfg["Q"], float), R=np.array(cfg["R"], float), Qf=np.array(cfg["Qf"], float),
                          verbose=False, rti=rti, MAX_ITER_SCP=100, epsilon_convergence=1e-6, save_it_data=False, it_data={},
                          convergence_data_scp={}, current_iteration_scp={}, H_mat=sp.csc_matrix(orc.fast.qp.H))
    cls = mod.SCP_SLS
    bag.pack_primal_nominal = lambda Xn, Un: mod.OCP.pack_primal_nominal(bag, Xn, Un)
    bag.unpack_primal_nominal = lambda y: mod.OCP.unpack_primal_nominal(bag, y)
    bag.scipy_csc_to_casadi_dm = lambda A: A.toarray()
    bag.eval_deviation_mismatch = lambda e, d: {}                      # diagnostic, result unused (SCP_SLS_jit.py:422)
    bag.socp_step = lambda x0: cls.socp_step(bag, x0)
    bag.check_convergence_scp = lambda: cls.check_convergence_scp(bag)
    bag.post_processing_solution = lambda: cls.post_processing_solution(bag)

    def update_jacobian():                                             # SCP_SLS_jit.py:251-366, by the oracle
        c = bag.current_iteration_scp
        orc.set_nominal(np.array(c["primal_x"], float), np.array(c["primal_u"], float))
        orc.update_jacobian()
    bag.update_jacobian = update_jacobian
    ref_fast = None
    if fast == "reference":
        from .ref_fast_sls import ReferenceFastSLS
        ref_fast = ReferenceFastSLS(name, sls_steps, eps)
        if isinstance(plant, dict):                                    # diagnostics: {'plant': callable, 'rec': list} collects the QP records
            plant["rec"].append(ref_fast.rec)
            plant = plant["plant"]

        def update_jacobian():                                         # values by the oracle, pushed as SCP_SLS_jit.py:343-366 does
            c = bag.current_iteration_scp
            Xc, Uc = np.array(c["primal_x"], float), np.array(c["primal_u"], float)
            A, B, cl, g_list = orc.linearize(Xc, Uc)
            ref_fast.update_dynamics_list(A, B, [m.E] * (N + 1), g_list, cl)
            ref_fast.update_linear_cost(2 * (orc.H @ np.array(c["primal_vec"], float).ravel()))
        bag.update_jacobian = update_jacobian
    bag.fast_SLS_solver = ref_fast if ref_fast is not None else SimpleNamespace(solve=lambda dx0: orc.fast.solve(np.asarray(dx0, float).reshape(-1)),
                                          reset_solver_to_zeros=orc.fast.reset_solver_to_zeros,
                                          solver_forward=SimpleNamespace(ubg=np.zeros(N * (m.nx + m.ni) + m.ni_f)))
    c = bag.current_iteration_scp
    c["primal_x"], c["primal_u"] = np.array(X, float), np.array(U, float)
    c["primal_vec"] = np.array(bag.pack_primal_nominal(c["primal_x"], c["primal_u"]), float)
    results = []
    x = np.asarray(x0s[0], float)
    for i, x0 in enumerate(x0s):
        if i > 0:
            if ref_fast is None:
                bag.fast_SLS_solver.solver_forward.ubg = np.array(orc.fast.ubg, float)
            cls.reset_warm_start(bag)
        x = x if plant is not None else np.asarray(x0, float)
        sol = cls.solve(bag, x)
        sol["x_measured"] = x.copy()
        results.append(sol)
        if plant is not None and sol.get("success"):
            x = plant(i, x, np.asarray(sol["primal_u"], float)[:, 0])
    return results


KEYS = ("primal_x", "primal_u", "primal_vec", "delta_vec", "dual_mu", "dual_mu_f", "cost", "cost_QP", "primal_infeasibility",
        "backoff", "backoff_x", "backoff_u", "dual_eta_f", "SOCP_steps", "iterations", "success")


def main_solve():
    from .gen_golden_ref_glue import scenario
    for name, rti, steps in (("pendulum", 3, 2), ("quadrotor", 3, 2), ("rocket", 1, 1)):         # the drivers' settings
        m, cfg, _, x0, X, U = scenario(name)
        rng = np.random.default_rng(9)
        # measured states: the start state, then the shifted nominal's first state plus a disturbance E w, |w| <= 1
        x1 = X[:, 1] + m.E @ rng.uniform(-1, 1, m.nx)
        res = reference_closed_loop_steps(name, [x0, x1], X, U, rti=rti, sls_steps=steps)
        out = dict(X_in=X, U_in=U, x0s=np.array([x0, x1]), rti=rti, sls_steps=steps)
        for i, r in enumerate(res):
            assert r["success"], (name, i)
            out.update({f"s{i}_{k}": np.asarray(r[k], float) for k in KEYS})
            out[f"s{i}_keys"] = np.array(sorted(r.keys()))
        np.savez_compressed(os.path.join(OUT, f"refsolve_{name}.npz"),
                            source="reference SCP_SLS.solve / socp_step / reset_warm_start control flow executed through oracle/ref_scp.py "
                                   "(update_jacobian and the fast_SLS object supplied by the oracle)", **out)
        print(name, "written; cost", [float(np.asarray(r["cost"]).ravel()[0]) for r in res], "infeas", [float(np.asarray(r["primal_infeasibility"]).ravel()[0]) for r in res],
              "iterations", [r["iterations"] for r in res])


if __name__ == "__main__":
    main()
    main_solve()
    main_loop()


def main_loop():
    """tests/golden/refloop_*.npz: closed loops run by the reference's own MPC stack -- SCP_SLS.solve / reset_warm_start over its own
    fast_SLS.solve, QP class and Numba kernels, plant = its own Model.ddyn + E w -- with only the VALUES of the linearisation (oracle,
    pinned by refdyn_*) and OSQP's ADMM iteration (restated, polished) supplied."""
    from .gen_golden_ref_glue import scenario
    for name, rti, sls_steps, steps, eps in (("pendulum", 3, 2, 6, 1e-6), ("quadrotor", 3, 2, 3, 1e-6), ("rocket", 1, 1, 3, 1e-5)):
        m, cfg, _, x0, X, U = scenario(name)
        refm = ref_models.load()[name]()
        W = np.random.default_rng(123).uniform(-1, 1, (steps, m.nx))
        states = [np.array(x0, float)]

        def plant(i, x, u):
            states.append(ref_models.ddyn(refm, x, u) + m.E @ W[i])
            return states[-1]
        res = reference_closed_loop_steps(name, [x0] * steps, X, U, rti=rti, sls_steps=sls_steps, fast="reference", eps=eps, plant=plant)
        assert all(bool(r["success"]) for r in res), name
        out = dict(X_in=X, U_in=U, x0=x0, W=W, rti=rti, sls_steps=sls_steps, states=np.array(states))
        for key in ("primal_x", "primal_u", "backoff_x", "backoff_u", "cost", "primal_infeasibility", "dual_mu"):
            out[key] = np.array([np.asarray(r[key], float) for r in res])
        np.savez_compressed(os.path.join(OUT, f"refloop_{name}.npz"),
                            source="reference SCP_SLS.solve / reset_warm_start over the reference's fast_SLS.solve, QP class and Numba kernels "
                                   f"(oracle/ref_scp.py, fast='reference'); linearisation values by the oracle, ADMM restated at eps {eps:g}", **out)
        print(name, "written;", steps, "steps, cost", out["cost"].ravel())
