"""NumPy restatement of the reference's SCP / fast-SLS control flow (one MPC instance).

Follows (reference file:line):
  * ``solver/SCP_SLS_jit.py:65-152``   SCP_SLS.solve (RTI and run-until-converged policies)
  * ``solver/SCP_SLS_jit.py:251-366``  update_jacobian (A_k, B_k, c_k, shifted g_k, linear cost q = 2 H y_nom)
  * ``solver/SCP_SLS_jit.py:404-473``  socp_step (nominal update, cost, primal infeasibility)
  * ``solver/SCP_SLS_jit.py:500-551``  reset_warm_start (shift + terminal roll-out)
  * ``solver/fast_SLS_jit.py:250-327`` update_dynamics_list / solve / _step
  * ``solver/fast_SLS_jit.py:444-600`` initialize_backoff / forward_solve / evaluate_dual_eta / backward_solve /
                                       update_tightening / check_convergence_socp
Quirks Q1-Q8 of SURVEY.md section 7 are reproduced and marked inline.  The first-step nominal trajectory, which the
reference obtains from IPOPT (``solver/nlp.py``; absent here and not reproducible), is an explicit input
(``set_nominal``) or is produced by ``nominal_sqp`` -- the same tube-free SQP the CUDA path uses.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np

from . import sls
from .qp import EPSILON, QPData, solve_qp

PRIMAL_INFEASIBILITY = 0.0  # solver/fast_SLS_jit.py:37


class FastSLS:
    """solver/fast_SLS_jit.py:195-646, numerics only."""

    def __init__(self, N, Q, R, m, Qf, Q_reg, R_reg, Q_reg_f, kernels=None):
        self.N, self.m = N, m
        self.Q_reg, self.R_reg, self.Q_reg_f = Q_reg, R_reg, Q_reg_f
        self.epsilon_backoff = 1e-10      # :205
        self.MAX_ITER = 30                # :206
        self.rti_steps = None
        self.qp = QPData(N, m.nx, m.nu, Q, R, Qf, m.G, m.Gf)
        self.gf_model = np.array(m.gf, float).copy()   # LTV copy of gf (dyn/LTV.py:26) used by Q4
        self.kernels = kernels            # optional: the reference's own Numba kernels (oracle/ref_kernels.py)
        self.qp_log = []                  # (q, l, u, x, y) of every QP solved, for fixtures
        self.keep_qp_log = False
        self._prev_primal_vec = None      # Q3: never reset
        self.reset_solver_to_zeros()

    # -- state ------------------------------------------------------------------------------------------
    def reset_solver_to_zeros(self):
        """:424-442 (+ initialize_solver :408-422).  Note: ``_prev_primal_vec`` survives (Q3)."""
        N, m = self.N, self.m
        self.cur = dict(eta=np.zeros((N, N, m.ni)), eta_f=np.zeros((N + 1, m.ni_f)), iteration_number=0)
        self.initialize_backoff()
        self.q = np.zeros(self.qp.nv)
        self.A = self.B = self.E = self.g_list = self.c_list = None
        self.ubg = self.lbg = None

    def initialize_backoff(self):
        """:444-454."""
        N, m = self.N, self.m
        c = self.cur
        c["beta"] = np.full((N, N, m.ni), self.epsilon_backoff)
        c["beta_f"] = np.full((N + 1, m.ni_f), self.epsilon_backoff)
        c["backoff"] = np.sqrt(c["beta"]).sum(axis=1)
        c["backoff_f"] = np.sqrt(c["beta_f"]).sum(axis=0)
        c["backoff_x"] = np.zeros((N + 1, m.nx))
        c["backoff_u"] = np.zeros((N, m.nu))

    def set_rti_steps(self, steps):
        self.rti_steps = None if (steps is None or steps <= 0) else int(steps)

    def update_dynamics_list(self, A, B, E, g_list, c_list):
        """:250-273 -> qp_jit.py:518-548 (+ :268-273 bounds rebuild) and :595-610 offset_constraints."""
        N, m = self.N, self.m
        self.A, self.B, self.E, self.g_list, self.c_list = A, B, E, g_list, c_list
        self.qp.set_dynamics(A, B)
        ub = np.concatenate([np.concatenate([np.zeros(m.nx), g_list[k]]) for k in range(N)] + [g_list[N]])
        lb0 = np.concatenate([np.concatenate([np.zeros(m.nx), -np.inf * np.ones(m.ni)]) for _ in range(N)]
                             + [-np.inf * np.ones(m.ni_f)])
        off = np.concatenate([np.concatenate([c_list[k], np.zeros(m.ni)]) for k in range(N)] + [np.zeros(m.ni_f)])
        self.ubg = ub - off + EPSILON     # note: +1e-10 lands on every row (Q2)
        self.lbg = lb0 - off - EPSILON

    def update_linear_cost(self, q):
        self.q = np.asarray(q, float).reshape(-1)

    # -- iteration ----------------------------------------------------------------------------------------
    def forward_solve(self, x0):
        """:459-473 -> qp_jit.py:362-402."""
        l, u = self.qp.bounds(self.lbg, self.ubg, x0)
        x, y, info = solve_qp(self.qp.H, self.q, self.qp.Acon, l, u)
        if not np.all(np.isfinite(x)) or info["mu"] > 1e-6:
            return False
        sol = self.qp.pack_solution(x, y, x @ self.qp.H @ x + self.q @ x)
        if self.keep_qp_log:
            self.qp_log.append(dict(q=self.q.copy(), l=l, u=u, x=x.copy(), y=y.copy(),
                                    A=np.array(self.A), B=np.array(self.B)))
        c = self.cur
        for k in ("primal_vec", "primal_x", "primal_u", "dual_vec", "dual_mu", "dual_mu_f"):
            c[k] = sol[k]
        c["cost_nominal"] = sol["cost"]
        return True

    def evaluate_dual_eta(self):
        c = self.cur
        sls.evaluate_dual_eta(self.N, c["dual_mu"], c["dual_mu_f"], c["beta"], c["beta_f"], self.epsilon_backoff,
                              c["eta"], c["eta_f"])

    def check_convergence_socp(self):
        """:581-600 (Q3: compares with the previous *first* QP solve, across MPC steps; only ok_p is used)."""
        cur_p = self.cur["primal_vec"]
        if self._prev_primal_vec is None:
            self._prev_primal_vec = cur_p.copy()
            return False
        ok_p = bool(np.max(np.abs(cur_p - self._prev_primal_vec)) <= 1e-3)
        self._prev_primal_vec = cur_p.copy()
        return ok_p

    def backward_solve(self):
        """:489-512."""
        m, N, c = self.m, self.N, self.cur
        A, B = np.ascontiguousarray(self.A), np.ascontiguousarray(self.B)
        args = (N, m.nx, m.nu, A, B, np.ascontiguousarray(m.G), np.ascontiguousarray(m.Gf),
                np.ascontiguousarray(c["eta"]), np.ascontiguousarray(c["eta_f"]),
                np.ascontiguousarray(self.Q_reg), np.ascontiguousarray(self.R_reg), np.ascontiguousarray(self.Q_reg_f))
        fn = self.kernels._backward_solve_numba if self.kernels is not None else sls.backward_solve
        c["S"], c["K"] = fn(*args)

    def update_tightening(self):
        """:517-571."""
        m, N, c = self.m, self.N, self.cur
        nx, nu = m.nx, m.nu
        A, B, E = np.ascontiguousarray(self.A), np.ascontiguousarray(self.B), np.ascontiguousarray(self.E)
        K = np.ascontiguousarray(c["K"])
        Gx, Gu = np.ascontiguousarray(m.G[:, :nx]), np.ascontiguousarray(m.G[:, nx:])
        Gf = np.ascontiguousarray(m.Gf)
        if self.kernels is not None:
            Phi_x, Phi_u = self.kernels._propagate(A, B, E, K)
            out = self.kernels._backoff_from_phi(Phi_x, Phi_u, Gx, Gu, Gf, float(self.epsilon_backoff))
        else:
            Phi_x, Phi_u = sls.propagate(A, B, E, K)
            out = sls.backoff_from_phi(Phi_x, Phi_u, Gx, Gu, Gf, float(self.epsilon_backoff))
        c["Phi_x"], c["Phi_u"] = Phi_x, Phi_u
        c["beta"], c["beta_f"], c["backoff"], c["backoff_f"] = out
        backoff, backoff_f = c["backoff"], c["backoff_f"]
        c["backoff_x"] = np.vstack((backoff[:, :nx], backoff_f[:nx]))      # Q7
        c["backoff_u"] = backoff[:, nx:nx + nu]
        tight = np.stack(self.g_list[:-1]) - backoff                        # (N, ni)
        cmat = np.stack(self.c_list, axis=1) - PRIMAL_INFEASIBILITY         # (nx, N)
        stage = np.vstack([-cmat, tight.T]).reshape(-1, order="F")
        self.ubg = np.concatenate([stage, self.gf_model - backoff_f])       # Q4: un-shifted gf

    def _step(self, x0):
        """:314-327."""
        if not self.forward_solve(x0):
            return False
        self.evaluate_dual_eta()
        if self.check_convergence_socp():
            self.cur["success"] = True
            return True
        self.backward_solve()
        self.update_tightening()
        self.cur["iteration_number"] += 1
        return None

    def solve(self, x0):
        """:278-312."""
        self.cur["success"] = False
        if self.rti_steps is not None and self.rti_steps > 0:
            self.initialize_backoff()
            infeasible = False
            for _ in range(self.rti_steps):
                if self._step(x0) is False:
                    infeasible = True
                    break
            if not infeasible:
                self.forward_solve(x0)
            self.cur["success"] = (not infeasible) or bool(self.cur.get("success", False))
            return dict(self.cur)
        self.initialize_backoff()
        for i in range(self.MAX_ITER):
            state = self._step(x0)
            if state is False:
                self.cur["success"] = False
                out = dict(self.cur)
                self.reset_solver_to_zeros()
                return out
            if state is True:
                return dict(self.cur)
        self.forward_solve(x0)
        self.cur["success"] = False
        out = dict(self.cur)
        self.reset_solver_to_zeros()
        return out


class SCPSLS:
    """solver/SCP_SLS_jit.py:15-551, numerics only."""

    def __init__(self, N, Q, R, m, Qf, Q_reg=None, R_reg=None, Q_reg_f=None, rti=-1, fast_sls_rti_steps=None,
                 kernels=None):
        self.N, self.m, self.Q, self.R, self.Qf = N, m, np.asarray(Q, float), np.asarray(R, float), np.asarray(Qf, float)
        self.Q_reg = np.eye(m.nx) if Q_reg is None else np.asarray(Q_reg, float)      # solver/ocp.py:15-26
        self.R_reg = np.eye(m.nu) if R_reg is None else np.asarray(R_reg, float)
        self.Q_reg_f = np.eye(m.nx) if Q_reg_f is None else np.asarray(Q_reg_f, float)
        self.rti = rti
        self.MAX_ITER_SCP = 100
        self.epsilon_convergence = 1e-10
        self.fast = FastSLS(N, self.Q, self.R, m, self.Qf, self.Q_reg, self.R_reg, self.Q_reg_f, kernels=kernels)
        self.fast.set_rti_steps(fast_sls_rti_steps)
        self.H = self.fast.qp.H
        self.cur = {}

    # -- nominal -----------------------------------------------------------------------------------------
    def set_nominal(self, X, U):
        self.cur["primal_x"] = np.array(X, float)
        self.cur["primal_u"] = np.array(U, float)
        self.cur["primal_vec"] = self.fast.qp.pack(self.cur["primal_x"], self.cur["primal_u"])

    def linearize(self, X, U):
        """SCP_SLS_jit.py:274-277, 348-356."""
        m, N = self.m, self.N
        A, B, c = [], [], []
        for k in range(N):
            a, b = m.jacobians(X[:, k], U[:, k])
            A.append(a)
            B.append(b)
            c.append(m.ddyn(X[:, k], U[:, k]) - X[:, k + 1])
        g_list = [m.g - m.G @ np.concatenate([X[:, k], U[:, k]]) for k in range(N)]
        g_list.append(m.gf - m.Gf @ X[:, N])
        return A, B, c, g_list

    def update_jacobian(self):
        """SCP_SLS_jit.py:251-366."""
        X, U = self.cur["primal_x"], self.cur["primal_u"]
        A, B, c, g_list = self.linearize(X, U)
        E = [self.m.E for _ in range(self.N + 1)]                         # Q8
        self.cur.update(A_list=A, B_list=B, c_offset_list=c, g_list=g_list, E_list=E)
        self.fast.update_dynamics_list(A, B, E, g_list, c)
        self.fast.update_linear_cost(2 * (self.H @ self.cur["primal_vec"]))

    def socp_step(self, x0):
        """SCP_SLS_jit.py:404-473."""
        m, N, c = self.m, self.N, self.cur
        x0 = np.asarray(x0, float).reshape(-1)
        sol = self.fast.solve(c["primal_x"][:, 0] - x0)
        if not sol["success"]:
            return False
        dX, dU = sol["primal_x"], sol["primal_u"]
        c["primal_x"] = c["primal_x"] + dX
        c["primal_u"] = c["primal_u"] + dU
        c["delta_vec"] = self.fast.qp.pack(dX, dU)
        c["primal_vec"] = c["primal_vec"] + c["delta_vec"]
        for key in ("dual_vec", "dual_mu", "dual_mu_f"):
            c[key] = sol[key]
        c["cost_QP"] = sol["cost_nominal"]
        c["SOCP_steps"] = sol["iteration_number"]
        X, U = c["primal_x"], c["primal_u"]
        cost_nlp = sum(X[:, i] @ self.Q @ X[:, i] + U[:, i] @ self.R @ U[:, i] for i in range(N)) + X[:, N] @ self.Qf @ X[:, N]
        c["cost"] = sol["cost_nominal"] + cost_nlp
        c["primal_infeasibility"] = max(np.max(m.ddyn(X[:, i], U[:, i]) - X[:, i + 1]) for i in range(N))   # signed max, :452-456
        c["dual_eta"], c["dual_eta_f"] = sol["eta"].copy(), sol["eta_f"].copy()
        for key in ("backoff", "backoff_f", "backoff_x", "backoff_u", "beta", "beta_f"):
            c[key] = np.array(sol[key]).copy()
        for key in ("K", "Phi_x", "Phi_u", "S"):
            c[key] = sol.get(key)
        return True

    def solve(self, x0):
        """SCP_SLS_jit.py:65-152."""
        if "primal_x" not in self.cur:
            X, U, ok = nominal_sqp(self.m, self.N, self.Q, self.R, self.Qf, x0)
            if not ok:
                return {"success": False}
            self.set_nominal(X, U)
        self.update_jacobian()
        until_conv = not (self.rti is not None and self.rti > 0)
        max_iters = int(self.MAX_ITER_SCP) if until_conv else int(self.rti)
        last_success, last_iter = False, 0
        for ii in range(max_iters):
            last_iter = ii
            last_success = bool(self.socp_step(x0))
            if not last_success:
                break
            if until_conv and np.max(np.abs(self.cur["delta_vec"])) < self.epsilon_convergence:
                self.cur["success"], self.cur["iterations"] = True, ii
                return dict(self.cur)
            self.update_jacobian()          # also after the last iteration (result unused by the closed loop)
        self.cur["iterations"] = last_iter
        self.cur["success"] = False if until_conv else last_success
        return dict(self.cur)

    def reset_warm_start(self):
        """SCP_SLS_jit.py:500-551."""
        N = self.N
        X, U = self.cur["primal_x"], self.cur["primal_u"]
        Xn, Un = X.copy(), U.copy()
        Xn[:, :N] = X[:, 1:N + 1]
        if N >= 2:
            Un[:, :N - 1] = U[:, 1:N]
        Un[:, N - 1] = U[:, N - 1]
        Xn[:, N] = self.m.ddyn(X[:, N], U[:, N - 1])
        self.cur = {}
        self.fast.reset_solver_to_zeros()
        self.set_nominal(Xn, Un)


def nominal_sqp(m, N, Q, R, Qf, x0, iters=10, tol=1e-6):
    """Tube-free SQP for the first-step nominal trajectory (replaces IPOPT, solver/nlp.py:158-217 -- SURVEY 8f #2).

    Start: states interpolated linearly from x0 to the origin, zero inputs (inside the box because x0 is); each
    iteration linearises the multiple-shooting defects and solves the nominal QP with the un-tightened bounds, full
    step.  The CUDA path runs the same iteration (same start, same count), so both sides share the initial nominal."""
    x0 = np.asarray(x0, float).reshape(-1)
    nx, nu = m.nx, m.nu
    X = np.stack([x0 * (1 - k / N) for k in range(N + 1)], axis=1)
    U = np.zeros((nu, N))
    scp = SCPSLS(N, Q, R, m, Qf, rti=1, fast_sls_rti_steps=None)
    for _ in range(iters):
        scp.set_nominal(X, U)
        scp.update_jacobian()
        f = scp.fast
        if not f.forward_solve(X[:, 0] - x0):
            return X, U, False
        X = X + f.cur["primal_x"]
        U = U + f.cur["primal_u"]
        if np.max(np.abs(f.cur["primal_vec"])) < tol:
            break
    return X, U, True


def closed_loop(m, cfg, x0, W, steps, X0=None, U0=None, kernels=None, collect=None):
    """expe/main_rocket_compare_closed_loop.py:104-173 (``_run_robust``): MPC loop with pre-drawn disturbances W[i]
    (``None`` -> noise-free, as in the pendulum / quadrotor drivers)."""
    N = cfg["N"]
    solver = SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"],
                    rti=cfg["rti"], fast_sls_rti_steps=cfg["fast_sls_rti_steps"], kernels=kernels)
    if X0 is not None:
        solver.set_nominal(X0, U0)
    x = np.asarray(x0, float).copy()
    out = dict(X=[x.copy()], U=[], Xn=[], Un=[], backoff_x=[], backoff_u=[], skipped=[], K=[], success=[])
    for i in range(steps):
        if i > 0:
            solver.reset_warm_start()
        sol = solver.solve(x)
        out["success"].append(bool(sol.get("success", False)))
        if not sol.get("success", False):
            break
        out["Xn"].append(sol["primal_x"].copy())
        out["Un"].append(sol["primal_u"].copy())
        out["backoff_x"].append(sol["backoff_x"].copy())
        out["backoff_u"].append(sol["backoff_u"].copy())
        out["skipped"].append(not np.any(sol["backoff_x"]))   # last fast-SLS solve took the Q3 early return
        out["K"].append(None if sol.get("K") is None else sol["K"].copy())
        if collect is not None:
            collect(i, solver, sol)
        u0 = sol["primal_u"][:, 0]
        out["U"].append(u0.copy())
        x = m.ddyn(x, u0)
        if W is not None:
            x = x + m.E @ W[i]
        out["X"].append(x.copy())
    return out
