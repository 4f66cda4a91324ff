"""``fast_SLS``: the reference's fast-SLS iteration class (solver/fast_SLS_jit.py:195-646) on the CUDA library.

Same constructor, methods and result keys; the iteration control is the reference's Python control flow, the numerics
are two library calls per step: ``QP.solve`` (``rnmpc_qp_general``) and ``rnmpc_sls_duals`` (dual weights, Riccati sweep,
Phi propagation and back-offs in one launch).  ``SCP_SLS`` does not go through this class (it keeps the whole solve on the
device); this is the drop-in for callers of ``fast_SLS`` itself.
"""
from __future__ import annotations

import numpy as np

from .ocp import OCP
from .qp_jit import QP

PRIMAL_INFEASIBILITY = 0.0


def convert_tensor_to_matrix(tensor):
    """util/SLS.py:49-63."""
    N, M, n, m = tensor.shape
    return tensor.transpose(0, 2, 1, 3).reshape(N * n, M * m)


class fast_SLS(OCP):
    def __init__(self, N, Q, R, m, Qf, Q_reg=None, R_reg=None, Q_reg_f=None):
        super().__init__(N, Q, R, m, Qf, Q_reg, R_reg, Q_reg_f)
        self.epsilon_backoff = 1e-10
        self.MAX_ITER = 30
        self.verbose = True
        self.rti_steps = None
        self.solver_forward = QP(self.N, self.Q, self.R, self.m, self.Qf, backend="osqp")
        self.nominal_lbg = self.solver_forward.lbg
        self.A_list, self.B_list = self.solver_forward.A_list, self.solver_forward.B_list
        self.E_list, self.g_list = self.solver_forward.E_list, self.solver_forward.g_list
        self.c_offset_list = []
        self.current_iteration = {}
        self.initialize_solver()
        self.initialize_backoff()
        self._sls_engine = None

    def _engine(self):
        """Handle carrying the tube regularisers (they may be re-assigned after construction, SCP_SLS_jit.py:386-388)."""
        from ..engine import BatchSolver
        key = (self.Q_reg.tobytes(), self.R_reg.tobytes(), self.Q_reg_f.tobytes(), np.asarray(self.E_list[0], float).tobytes())
        if self._sls_engine is None or self._sls_engine[0] != key:
            m, nx, nu = self.m, self.m.nx, self.m.nu
            E = np.asarray(self.E_list[0], float)
            if any(not np.array_equal(np.asarray(e, float), E) for e in self.E_list):
                raise NotImplementedError("one disturbance map E for every column (the reference passes [m.E]*(N+1), quirk Q8)")
            big = 1e3
            eng = BatchSolver(m.model_name, self.N, self.Q, self.R, self.Qf, self.Q_reg, self.R_reg, self.Q_reg_f, E, big * np.ones(nx),
                              -big * np.ones(nx), big * np.ones(nu), -big * np.ones(nu), capacity=1)
            self._sls_engine = (key, eng)
        return self._sls_engine[1]

    def set_rti_steps(self, steps):
        self.rti_steps = None if (steps is None or steps <= 0) else int(steps)

    def update_dynamics_list(self, new_list_A, new_list_B, new_list_E=None, new_list_g=None, c_offset_list=None):
        """fast_SLS_jit.py:250-273."""
        self.A_list, self.B_list = new_list_A, new_list_B
        if new_list_E is not None:
            self.E_list = new_list_E
        if new_list_g is not None:
            self.g_list = new_list_g
        self.solver_forward.update_dynamics(self.A_list, self.B_list, self.E_list, self.g_list)
        if c_offset_list is not None:
            self.c_offset_list = c_offset_list
            self.solver_forward.offset_constraints(np.hstack([np.asarray(c, float).reshape(-1, 1) for c in c_offset_list]))

    def update_linear_cost(self, q_cost_lin):
        self.solver_forward.update_q_cost_lin(q_cost_lin)

    def add_linear_cost(self, q_cost_lin):
        self.solver_forward.add_q_cost_lin(q_cost_lin)

    # ------------------------------------------------------------------------------------------------------- state
    def initialize_solver(self):
        N, ni, ni_f = self.N, self.m.ni, self.m.ni_f
        self.current_iteration = {'primal_vec': np.zeros(1), 'dual_vec': np.zeros(1), 'eta': np.zeros((N, N, ni)),
                                  'eta_f': np.zeros((N + 1, ni_f)), 'iteration_number': 0, 't_qp_ms': np.nan, 't_backward_ms': np.nan}

    def reset_solver_to_zeros(self):
        """fast_SLS_jit.py:424-442 (the Q3 history ``_prev_primal_vec`` survives)."""
        self.A_list, self.B_list, self.E_list, self.c_offset_list, self.g_list = [], [], [], [], []
        self.initialize_solver()
        self.initialize_backoff()
        self.solver_forward.reset_ubg()
        self.solver_forward.reset_lbg()
        self.solver_forward.reset_q_cost_lin()

    def initialize_backoff(self):
        N, nx, nu = self.N, self.m.nx, self.m.nu
        c = self.current_iteration
        c['beta'] = np.full((N, N, self.m.ni), self.epsilon_backoff)
        c['beta_f'] = np.full((N + 1, self.m.ni_f), self.epsilon_backoff)
        c['backoff'] = np.sqrt(c['beta']).sum(axis=1)
        c['backoff_f'] = np.sqrt(c['beta_f']).sum(axis=0)
        c['backoff_x'] = np.zeros((N + 1, nx))
        c['backoff_u'] = np.zeros((N, nu))
        self._beta_valid = False

    # --------------------------------------------------------------------------------------------------- iteration
    def forward_solve(self, x0):
        sol = self.solver_forward.solve(x0)
        if sol['success'] is False:
            return False
        c = self.current_iteration
        c['t_qp_ms'] = sol.get('time_ms', np.nan)
        for k in ('primal_vec', 'primal_x', 'primal_u', 'dual_vec', 'dual_mu', 'dual_mu_f'):
            c[k] = sol[k]
        c['cost_nominal'] = sol['cost']
        return True

    def evaluate_dual_eta(self):
        """fast_SLS_jit.py:475-487 (only needed for the result dict: the library forms the weights on the fly)."""
        N, c = self.N, self.current_iteration
        beta, beta_f = np.maximum(c['beta'], self.epsilon_backoff), np.maximum(c['beta_f'], self.epsilon_backoff)
        for k in range(N):
            c['eta'][k, :k + 1] = c['dual_mu'][:, k][None, :] / (2.0 * np.sqrt(beta[k, :k + 1]))
        c['eta_f'][:] = c['dual_mu_f'][None, :] / (2.0 * np.sqrt(beta_f))

    def check_convergence_socp(self):
        """fast_SLS_jit.py:581-600 (quirk Q3)."""
        cur_p = self.current_iteration['primal_vec']
        prev_p = getattr(self, '_prev_primal_vec', None)
        self._prev_primal_vec = cur_p.copy()
        if prev_p is None:
            return False
        return bool(np.max(np.abs(cur_p - prev_p)) <= 1e-3)

    def backward_solve_and_tighten(self):
        """backward_solve + update_tightening (fast_SLS_jit.py:489-571) in one library call."""
        import time
        N, nx, nu, c = self.N, self.m.nx, self.m.nu, self.current_iteration
        A = np.stack([np.asarray(a, float) for a in self.A_list])[None]
        B = np.stack([np.asarray(b, float) for b in self.B_list])[None]
        t0 = time.perf_counter()
        out = self._engine().sls_duals(A, B, c['dual_mu'].T[None], c['dual_mu_f'][None],
                                       c['beta'][None] if self._beta_valid else None, c['beta_f'][None] if self._beta_valid else None)
        out = {k: v.cpu().numpy()[0] for k, v in out.items()}
        c['t_backward_ms'] = (time.perf_counter() - t0) * 1e3
        for k in ('K', 'S', 'Phi_x', 'Phi_u', 'beta', 'beta_f', 'backoff', 'backoff_f'):
            c[k] = out[k]
        self._beta_valid = True
        backoff, backoff_f = c['backoff'], c['backoff_f']
        c['backoff_x'] = np.vstack((backoff[:, :nx], backoff_f[:nx]))
        c['backoff_u'] = backoff[:, nx:nx + nu]
        tight = np.stack([np.asarray(g, float).reshape(-1) for g in self.g_list[:-1]]) - backoff
        cmat = np.hstack([np.asarray(ci, float).reshape(-1, 1) for ci in self.c_offset_list]) - PRIMAL_INFEASIBILITY
        stage = np.vstack([-cmat, tight.T]).reshape(-1, order='F')
        self.solver_forward.update_ubg(np.concatenate([stage, np.asarray(self.m.gf, float).reshape(-1) - backoff_f]))   # quirk Q4

    def _step(self, x0):
        if not self.forward_solve(x0):
            return False
        self.evaluate_dual_eta()
        if self.check_convergence_socp():
            self.current_iteration['success'] = True
            return True
        self.backward_solve_and_tighten()
        self.current_iteration['iteration_number'] += 1
        return None

    def solve(self, x0):
        """fast_SLS_jit.py:278-312."""
        self.current_iteration['success'] = False
        self.initialize_backoff()
        if self.rti_steps is not None and self.rti_steps > 0:
            infeasible = False
            for _ in range(self.rti_steps):
                if self._step(x0) is False:
                    infeasible = True
                    break
            if not infeasible:
                self.forward_solve(x0)
            self.current_iteration['success'] = (not infeasible) or bool(self.current_iteration.get('success', False))
            return self.post_processing_solution()
        for i in range(self.MAX_ITER):
            state = self._step(x0)
            if state is False:
                return self._finish_failure()
            if state is True:
                return self.post_processing_solution()
        self.forward_solve(x0)
        return self._finish_failure()

    def _finish_failure(self):
        self.current_iteration['success'] = False
        sol = self.post_processing_solution()
        self.reset_solver_to_zeros()
        return sol

    def post_processing_solution(self):
        """fast_SLS_jit.py:602-646."""
        c = self.current_iteration
        if 'K' in c:
            c['K_mat'] = np.nan_to_num(convert_tensor_to_matrix(c['K']))
        if 'Phi_u' in c and 'Phi_x' in c:
            c['Phi_u_mat'] = convert_tensor_to_matrix(np.nan_to_num(c['Phi_u']))
            c['Phi_x_mat'] = convert_tensor_to_matrix(np.nan_to_num(c['Phi_x']))
        keys = ('primal_x', 'primal_u', 'primal_vec', 'dual_vec', 'dual_mu', 'dual_mu_f', 'eta', 'eta_f', 'K', 'K_mat', 'Phi_x',
                'Phi_u', 'Phi_x_mat', 'Phi_u_mat', 'beta', 'beta_f', 'backoff', 'backoff_f', 'backoff_x', 'backoff_u', 't_qp_ms',
                't_backward_ms')
        out = {'iteration_number': c['iteration_number'], 'success': c['success'], 'cost_nominal': c.get('cost_nominal', np.nan),
               'cost_tube': np.nan, 'cost': np.nan}
        out.update({k: c.get(k, None) for k in keys})
        return out
