"""``QP``: the reference's nominal-QP front-end (solver/qp_jit.py:22-628) on the CUDA library.

Same constructor, update methods and result keys.  The reference assembles a CSC matrix and calls OSQP; here the state
(``A_list``, ``B_list``, ``ubg``, ``lbg``, ``q_cost_lin``) is kept in the reference's vector layout and ``solve(x0)`` hands
it to ``rnmpc_qp_general`` (Riccati interior point).  Only the ``"osqp"`` backend semantics exist (the one the hot path
uses, fast_SLS_jit.py:229-233); the CasADi / code-generation back-ends and the exporters are out of scope.
"""
from __future__ import annotations

import time

import numpy as np

from ..engine import BatchSolver
from .ocp import OCP

EPSILON = 1e-10   # solver/qp_jit.py:19


class QP(OCP):
    def __init__(self, N, Q, R, m, Qf, *, backend="osqp", codegen_module=None, verbose=True, export_standard_QP=False,
                 export_dir=None, device=0):
        super().__init__(N, Q, R, m, Qf)
        if backend not in {"casadi_osqp", "osqp", "osqp_codegen"}:
            raise AssertionError("unknown backend")
        self.backend = "osqp"          # every request is served by the CUDA solver
        self.verbose = verbose
        self.export_standard_qp = bool(export_standard_QP)
        nx, nu, ni, ni_f = m.nx, m.nu, m.ni, m.ni_f
        nz = nx + nu
        G, Gf = np.asarray(m.G, float), np.asarray(m.Gf, float)
        if not (np.array_equal(G, np.vstack([np.eye(nz), -np.eye(nz)])) and np.array_equal(Gf, np.vstack([np.eye(nx), -np.eye(nx)]))):
            raise NotImplementedError("box constraints G = [I; -I], Gf = [I; -I] only")
        # LTV placeholders (dyn/LTV.py:28-32) or the model's own lists
        self.A_list = list(getattr(m, "A_list", [np.ones((nx, nx))] * N))
        self.B_list = list(getattr(m, "B_list", [np.ones((nx, nu))] * N))
        self.E_list = list(getattr(m, "E_list", [np.ones((nx, m.nw))] * (N + 1)))
        self.g_list = list(getattr(m, "g_list", [np.ones(ni)] * N + [np.ones(ni_f)]))
        big = 1e3 * np.ones(nx)
        self._engine = BatchSolver(m.model_name, N, self.Q, self.R, self.Qf, np.eye(nx), np.eye(nu), np.eye(nx), np.eye(nx),
                                   big, -big, 1e3 * np.ones(nu), -1e3 * np.ones(nu), capacity=1, device=device)
        self._rebuild_bounds()
        self.nominal_ubg = self.ubg.copy()
        self.nominal_lbg = self.lbg.copy()
        self.reset_q_cost_lin()

    # ---------------------------------------------------------------------------------------------- bounds bookkeeping
    def _rebuild_bounds(self):
        """ubg = [0; g_k]_k ++ g_N, lbg = [0; -inf]_k ++ -inf  (qp_jit.py:141-156, 268-273)."""
        nx, ni, ni_f, N = self.m.nx, self.m.ni, self.m.ni_f, self.N
        self.ubg = np.concatenate([np.concatenate([np.zeros(nx), np.asarray(self.g_list[k], float).reshape(-1)]) for k in range(N)]
                                  + [np.asarray(self.g_list[N], float).reshape(-1)])
        self.lbg_ZEROS_INF_NOMINAL = np.concatenate([np.concatenate([np.zeros(nx), -np.inf * np.ones(ni)]) for _ in range(N)]
                                                    + [-np.inf * np.ones(ni_f)])
        self.lbg = self.lbg_ZEROS_INF_NOMINAL.copy()

    def update_dynamics(self, new_A_list, new_B_list, new_E_list=None, new_g_list=None):
        """qp_jit.py:518-548."""
        assert len(new_A_list) == self.N and len(new_B_list) == self.N
        self.A_list, self.B_list = list(new_A_list), list(new_B_list)
        if new_E_list is not None:
            assert len(new_E_list) == self.N + 1
            self.E_list = list(new_E_list)
        if new_g_list is not None:
            assert len(new_g_list) == self.N + 1
            self.g_list = list(new_g_list)
        self._rebuild_bounds()

    def update_ubg(self, new_ubg):
        self.ubg = np.asarray(new_ubg, float).reshape(-1).copy()

    def reset_ubg(self):
        self.ubg = self.nominal_ubg.copy()

    def reset_lbg(self):
        self.lbg = self.lbg_ZEROS_INF_NOMINAL.copy()

    def offset_constraints(self, offset_equality_constraints):
        """qp_jit.py:595-610: note the +EPSILON lands on every row of ubg."""
        off_in = np.asarray(offset_equality_constraints, float)
        assert off_in.shape == (self.m.nx, self.N)
        nx, ni = self.m.nx, self.m.ni
        off = np.concatenate([np.concatenate([off_in[:, k], np.zeros(ni)]) for k in range(self.N)] + [np.zeros(self.m.ni_f)])
        self.ubg = self.ubg - off + EPSILON
        self.lbg = self.lbg_ZEROS_INF_NOMINAL - off - EPSILON

    def update_q_cost_lin(self, q_cost_lin):
        q = np.asarray(q_cost_lin, float).reshape(-1)
        assert q.shape[0] == (self.m.nx + self.m.nu) * self.N + self.m.nx
        self.q_cost_lin = q.copy()

    def add_q_cost_lin(self, q_cost_lin):
        q = np.asarray(q_cost_lin, float).reshape(-1)
        assert q.shape[0] == (self.m.nx + self.m.nu) * self.N + self.m.nx
        self.q_cost_lin = self.q_cost_lin + q

    def reset_q_cost_lin(self):
        self.q_cost_lin = np.zeros((self.m.nx + self.m.nu) * self.N + self.m.nx)

    # ------------------------------------------------------------------------------------------------------------ solve
    def _entry_arrays(self):
        """(c, hi, lo, q) in the library's layouts from ubg / lbg / q_cost_lin."""
        nx, nu, ni, ni_f, N = self.m.nx, self.m.nu, self.m.ni, self.m.ni_f, self.N
        nz = nx + nu
        st_u = self.ubg[:N * (nx + ni)].reshape(N, nx + ni)
        st_l = self.lbg[:N * (nx + ni)].reshape(N, nx + ni)
        c = -0.5 * (st_u[:, :nx] + st_l[:, :nx])                     # dynamics rows: [A B -I] y in [-c - eps, -c (+ eps)]
        hi = np.full((N + 1, nz), np.inf)
        lo = np.full((N + 1, nz), -np.inf)
        hi[:N] = st_u[:, nx:nx + nz]
        lo[:N] = -st_u[:, nx + nz:]
        hi[N, :nx] = self.ubg[N * (nx + ni):][:nx]
        lo[N, :nx] = -self.ubg[N * (nx + ni):][nx:]
        hi[N, nx:], lo[N, nx:] = 0.0, 0.0
        big = 1e20
        hi, lo = np.where(np.isposinf(hi), big, hi), np.where(np.isneginf(lo), -big, lo)
        q = np.zeros((N + 1, nz))
        q[:N] = self.q_cost_lin[:N * nz].reshape(N, nz)
        q[N, :nx] = self.q_cost_lin[N * nz:]
        return c, hi, lo, q

    def solve(self, x0):
        """qp_jit.py:343-402 + _pack_solution :487-513.  The first state is pinned to -x0 (x(0) + x0 = 0)."""
        nx, nu, ni, ni_f, N = self.m.nx, self.m.nu, self.m.ni, self.m.ni_f, self.N
        c, hi, lo, q = self._entry_arrays()
        A = np.stack([np.asarray(a, float) for a in self.A_list])[None]
        B = np.stack([np.asarray(b, float) for b in self.B_list])[None]
        pin = -np.asarray(x0, float).reshape(1, nx)
        t0 = time.perf_counter()
        out = self._engine.qp_general(A, B, c[None], hi[None], lo[None], q[None], pin)
        status = int(out["status"].cpu().numpy()[0])
        t_ms = (time.perf_counter() - t0) * 1e3
        if status != 0:
            if self.verbose:
                print(f"QP(rnmpc_b200): status {status}")
            return {'success': False}
        dX, dU = out["dX"].cpu().numpy()[0], out["dU"].cpu().numpy()[0]
        mu, mu_f = out["mu"].cpu().numpy()[0], out["mu_f"].cpu().numpy()[0]
        primal_x, primal_u = dX.T.copy(), dU.T.copy()
        lam = np.full((N, nx + ni), np.nan)           # multipliers of the dynamics rows are not formed by the Riccati solver
        lam[:, nx:] = mu
        return {
            'success': True,
            'primal_vec': self.pack_primal_nominal(primal_x, primal_u),
            'primal_x': primal_x, 'primal_u': primal_u,
            'dual_vec': np.concatenate([lam.reshape(-1), mu_f]),
            'dual_mu': mu.T.copy(), 'dual_mu_f': mu_f.copy(),
            'cost': float(out["cost"].cpu().numpy()[0]),
            'time_ms': t_ms,
        }
