"""The reference's OWN fast-SLS glue, executed here: ``fast_SLS.evaluate_dual_eta`` (``solver/fast_SLS_jit.py:475-487``),
``fast_SLS.backward_solve`` (``:489-512``), ``fast_SLS.update_tightening`` (``:517-571``, incl. quirks Q4 / Q7) and
``fast_SLS.check_convergence_socp`` (``:581-600``, Q3) are called as the unbound methods of the reference class on a plain
attribute bag -- the class itself cannot be constructed without CasADi and the generated OSQP module, these four methods
need neither.  ``oracle/ref_kernels.py`` provides the module (import stubs), so this works only where ``/root/reference``
exists; ``oracle/gen_golden_ref_glue.py`` commits the outputs as ``tests/golden/refglue_*.npz``.

Pins rows a4, a5-a7 (as a chain, from multipliers to tightened bounds) and a8 / a9 of SURVEY.md section 8 against the
reference's statements rather than a reading of them.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from . import ref_kernels


class _DM(np.ndarray):
    """m.G / m.Gf are CasADi DM in the reference: sliceable, convertible with np.asarray, and with .full()."""

    def full(self):
        return np.asarray(self)


def _dm(a):
    return np.asarray(a, float).view(_DM)


class _Recorder:
    ubg = None

    def update_ubg(self, new_ubg):          # QP.update_ubg (solver/qp_jit.py:578-582)
        self.ubg = np.array(new_ubg, float)


def reference_pass(N, m, Q_reg, R_reg, Q_reg_f, A, B, E, g_list, c_list, mu, mu_f, beta, beta_f, primal_vecs=()):
    """One pass of the reference between two QP solves.  mu (ni, N) and mu_f (ni_f,) as the QP returns them
    (``solver/qp_jit.py:503-513``), beta (N, N, ni), beta_f (N+1, ni_f) the state before the pass.  ``primal_vecs``: a
    sequence of primal vectors fed one after the other to check_convergence_socp (its decisions are returned)."""
    mod = ref_kernels.load()
    cls = mod.fast_SLS
    model = SimpleNamespace(nx=m.nx, nu=m.nu, nw=m.nw, ni=m.ni, ni_f=m.ni_f, G=_dm(m.G), Gf=_dm(m.Gf), gf=np.array(m.gf, float))
    rec = _Recorder()
    cur = dict(beta=np.array(beta, float), beta_f=np.array(beta_f, float), dual_mu=np.array(mu, float), dual_mu_f=np.array(mu_f, float),
               eta=np.zeros((N, N, m.ni)), eta_f=np.zeros((N + 1, m.ni_f)))
    bag = SimpleNamespace(N=N, m=model, epsilon_backoff=1e-10, current_iteration=cur, Q_reg=np.array(Q_reg, float),
                          R_reg=np.array(R_reg, float), Q_reg_f=np.array(Q_reg_f, float),
                          A_list=[np.array(a, float) for a in A], B_list=[np.array(b, float) for b in B],
                          E_list=[np.array(e, float) for e in E], g_list=[np.array(g, float).reshape(-1, 1) for g in g_list],
                          c_offset_list=[np.array(c, float).reshape(-1, 1) for c in c_list], solver_forward=rec)
    cls.evaluate_dual_eta(bag)
    eta, eta_f = cur["eta"].copy(), cur["eta_f"].copy()
    cls.backward_solve(bag)
    cls.update_tightening(bag)
    decisions = []
    for p in primal_vecs:
        cur["primal_vec"] = np.array(p, float)
        decisions.append(bool(cls.check_convergence_socp(bag)))
    return dict(eta=eta, eta_f=eta_f, K=cur["K"], beta=cur["beta"], beta_f=cur["beta_f"], backoff=cur["backoff"],
                backoff_f=cur["backoff_f"], backoff_x=cur["backoff_x"], backoff_u=cur["backoff_u"], ubg=rec.ubg,
                cost_tube=float(cur["cost_tube"]), q3=np.array(decisions, bool))
