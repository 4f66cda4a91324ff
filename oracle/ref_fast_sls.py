"""The reference's OWN fast-SLS iteration, executed end to end: ``fast_SLS.solve`` / ``_step`` / ``forward_solve`` /
``evaluate_dual_eta`` / ``check_convergence_socp`` / ``backward_solve`` / ``update_tightening`` / ``initialize_backoff`` /
``post_processing_solution`` / ``_finish_success`` (``solver/fast_SLS_jit.py:278-646``) run as unbound methods of the unmodified
class on an attribute bag whose ``solver_forward`` is the reference's own ``QP`` object (``oracle/ref_qp.py``: assembly, bounds and
unpacking by the reference, ADMM by the restated OSQP) and whose Riccati / propagation / back-off kernels are the reference's Numba
kernels.  Inputs: a linearisation (A, B, c, g lists) and the linear cost.  So between "linearised model in" and "result dictionary
out" the only statements that are not the reference's are OSQP's ADMM iteration.  Both modes: RTI (``rti_steps`` > 0) and iterate
until the Q3 test fires.  Works only where ``/root/reference`` exists; outputs committed as ``tests/golden/reffast_*.npz``
(recipe: ``python -m oracle.ref_fast_sls``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

from . import models as om
from . import ref_kernels, ref_qp
from .ref_glue import _dm

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
METHODS = ("_step", "forward_solve", "evaluate_dual_eta", "check_convergence_socp", "backward_solve", "update_tightening",
           "initialize_backoff", "post_processing_solution", "_finish_success", "_finish_failure")


class ReferenceFastSLS:
    """A long-lived ``fast_SLS`` of the reference (attribute bag + unbound methods + the reference's own QP object), with the
    interface ``SCP_SLS`` uses: update_dynamics_list, update_linear_cost, solve, reset_solver_to_zeros, solver_forward."""

    def __init__(self, name, rti_steps, eps):
        self.m, cfg = om.CONFIGS[name]()
        self.bag, self.cls, self.solver_forward, self.rec = _make_bag(name, rti_steps, eps)

    def update_dynamics_list(self, A, B, E, g_list, c_list):
        self.cls.update_dynamics_list(self.bag, [np.array(a, float) for a in A], [np.array(b, float) for b in B],
                                      [np.array(e, float) for e in E], [np.array(g, float).reshape(-1, 1) for g in g_list],
                                      [np.array(c, float).reshape(-1, 1) for c in c_list])

    def update_linear_cost(self, q):
        self.cls.update_linear_cost(self.bag, np.array(q, float).reshape(-1, 1))

    def solve(self, dx0):
        return self.cls.solve(self.bag, np.asarray(dx0, float).reshape(-1, 1))

    def reset_solver_to_zeros(self):
        """fast_SLS_jit.py:424-442 without the list re-initialisation of the OCP base class: initialize_solver + initialize_backoff
        (both the reference's own) and the QP's reset_ubg / reset_lbg / reset_q_cost_lin; ``_prev_primal_vec`` survives (Q3)."""
        b = self.bag
        b.A_list, b.B_list, b.E_list, b.c_offset_list, b.g_list = [], [], [], [], []
        self.cls.initialize_solver(b)
        b.current_iteration.update(success=False)
        b.initialize_backoff()
        self.solver_forward.reset_ubg()
        self.solver_forward.reset_lbg()
        self.solver_forward.reset_q_cost_lin()


def _make_bag(name, rti_steps, eps):
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    rec = ref_qp._Recorder()
    qp = ref_qp.reference_qp(N, m, cfg["Q"], cfg["R"], cfg["Qf"], rec, dict(eps_abs=eps, eps_rel=eps, max_iter=20000))
    cls = ref_kernels.load().fast_SLS
    model = SimpleNamespace(nx=m.nx, nu=m.nu, nw=m.nw, ni=m.ni, ni_f=m.ni_f, G=_dm(m.G), Gf=_dm(m.Gf), gf=np.array(m.gf, float))
    bag = SimpleNamespace(N=N, m=model, epsilon_backoff=1e-10, MAX_ITER=30, verbose=False, _tab="", solver_forward=qp,
                          rti_steps=(int(rti_steps) if rti_steps else None),
                          Q_reg=np.array(cfg["Q_reg"], float), R_reg=np.array(cfg["R_reg"], float), Q_reg_f=np.array(cfg["Q_reg_f"], float),
                          current_iteration=dict(eta=np.zeros((N, N, m.ni)), eta_f=np.zeros((N + 1, m.ni_f)), iteration_number=0,
                                                 success=False, t_backward_ms=0.0),
                          reset_solver_to_zeros=lambda: None)
    for name_ in METHODS:
        setattr(bag, name_, (lambda fn: (lambda *a: fn(bag, *a)))(getattr(cls, name_)))
    return bag, cls, qp, rec


def reference_fast_sls(name, A, B, E, g_list, c_list, q, dx0s, rti_steps, eps):
    """One ``fast_SLS`` object of the reference; ``solve(dx0)`` for every dx0 in turn (the Q3 history carries over, as in the
    reference).  Returns the result dictionaries."""
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    bag, cls, qp, rec = _make_bag(name, rti_steps, eps)
    # fast_SLS.update_dynamics_list (:250-273), fast_SLS.update_linear_cost (:575)
    bag.A_list, bag.B_list = [np.array(a, float) for a in A], [np.array(b, float) for b in B]
    bag.E_list = [np.array(e, float) for e in E]
    bag.g_list = [np.array(g, float).reshape(-1, 1) for g in g_list]
    bag.c_offset_list = [np.array(c, float).reshape(-1, 1) for c in c_list]
    qp.update_dynamics(bag.A_list, bag.B_list, bag.E_list, bag.g_list)
    qp.offset_constraints(np.hstack(bag.c_offset_list))
    qp.update_q_cost_lin(np.array(q, float).reshape(-1, 1))
    out = []
    for dx0 in dx0s:
        bag.current_iteration["iteration_number"] = 0
        sol = cls.solve(bag, np.asarray(dx0, float).reshape(-1, 1))
        out.append({k: (np.array(v, float) if v is not None else None) for k, v in sol.items()})
    return out, rec


KEYS = ("iteration_number", "success", "cost_nominal", "primal_x", "primal_u", "primal_vec", "dual_mu", "dual_mu_f", "eta", "eta_f", "K",
        "K_mat", "beta", "beta_f", "backoff", "backoff_f", "backoff_x", "backoff_u")


def main():
    from .gen_golden_ref_glue import scenario
    for name, rti_steps, eps in (("pendulum", None, 1e-6), ("quadrotor", 2, 1e-6), ("rocket", 1, 1e-5)):
        m, cfg, s, x0, X, U = scenario(name)
        x0 = x0 + 0.02 * (m.x_ub - m.x_lb) / 2 * np.random.default_rng(5).uniform(-1, 1, m.nx)
        s.set_nominal(X, U)
        s.update_jacobian()
        f = s.fast
        dx0 = X[:, 0] - x0
        sols, rec = reference_fast_sls(name, f.A, f.B, f.E, f.g_list, f.c_list, f.q, [dx0, dx0 * 0.9998], rti_steps, eps)
        out = dict(X=X, U=U, x0=x0, dx0s=np.array([dx0, dx0 * 0.9998]), rti_steps=(rti_steps or 0), n_qp_solves=len(rec.solves),
                   all_polished=all(sv["polished"] for sv in rec.solves))
        for i, sol in enumerate(sols):
            out.update({f"s{i}_{k}": sol[k] for k in KEYS})
            out[f"s{i}_keys"] = np.array(sorted(sol.keys()))
        np.savez_compressed(os.path.join(OUT, f"reffast_{name}.npz"), source="reference fast_SLS.solve executed through oracle/ref_fast_sls.py "
                            f"(reference QP class + Numba kernels; ADMM by oracle/osqp_restated.py at eps {eps:g})", **out)
        print(name, "written; iterations", [int(s_["iteration_number"]) for s_ in sols], "success", [bool(s_["success"]) for s_ in sols],
              "QP solves", len(rec.solves), "all polished", out["all_polished"])


if __name__ == "__main__":
    main()
