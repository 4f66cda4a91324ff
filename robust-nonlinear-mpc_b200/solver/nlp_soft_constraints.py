"""``NLPSoftConstraints``: the reference's soft-constrained NMPC comparator (solver/nlp_soft_constraints.py:8-276) on the CUDA library.

Same constructor, ``solve(x0, x_guess=None, u_guess=None, gamma_guess=None)`` signature and result keys (``success, primal_vec,
primal_x, primal_u, primal_gamma, cost``; no ``dual_vec``: the reference returns IPOPT's multipliers of its own constraint ordering).
The reference hands the NLP to IPOPT; here it is solved by full-step SQP on the GPU (``rnmpc_soft_nmpc``: linearisation kernel +
interior-point QP kernel with the slack variables eliminated entry-wise).  A local method needs a start: ``solve`` uses the guess it is
given, else the previous solution shifted by one stage (``SCP_SLS.reset_warm_start``'s rule), else -- first call -- the states
interpolated from x0 to the origin with zero inputs (the reference starts IPOPT from the zero vector at every call, :176-190).
One extra kwarg, ``batch``: independent instances per call (leading batch axis on every array)."""
from __future__ import annotations

import numpy as np

from ..engine import BatchSolver
from .ocp import OCP


class NLPSoftConstraints(OCP):
    def __init__(self, N, Q, R, m, Qf, rho_soft: float = 1e6, rho_soft_l1=None, batch: int = 1, device: int = 0, max_iter: int = 80,
                 tol: float = 1e-9):
        super().__init__(N, Q, R, m, Qf)
        self.nlp_solver_name = 'sqp (rnmpc_b200)'
        self.solution = {}
        self.rho_soft = float(rho_soft)
        self.rho_soft_l1 = float(rho_soft if rho_soft_l1 is None else rho_soft_l1)      # nlp_soft_constraints.py:22-23
        self.verbose = False
        self._mg, self._mgf = int(m.G.shape[0]), int(m.Gf.shape[0])
        self.batch, self.max_iter, self.tol = int(batch), int(max_iter), float(tol)
        if getattr(m, "model_name", None) not in ("pendulum", "quadrotor", "rocket"):
            raise TypeError("model must be one of this package's dyn.Pendulum / dyn.Quadrotor / dyn.Rocket")
        x_ub, x_lb, u_ub, u_lb, xf_ub, xf_lb = m.box()
        eye = lambda n: np.eye(n)
        self._engine = BatchSolver(m.model_name, N, self.Q, self.R, self.Qf, eye(m.nx), eye(m.nu), eye(m.nx), np.asarray(m.E, float), x_ub, x_lb,
                                   u_ub, u_lb, xf_ub, xf_lb, rti=1, sls_steps=1, capacity=self.batch, device=device)
        self._guess = None          # (X [B,N+1,nx], U [B,N,nu]) of the previous solution

    def _b(self, a, shape):
        return np.asarray(a, dtype=float).reshape((self.batch,) + shape)

    def _out(self, a):
        return a[0] if self.batch == 1 else a

    def set_initial_guess(self, x_guess=None, u_guess=None, gamma_guess=None):
        """nlp_soft_constraints.py:76-105 (the slack guess is not needed: the QP picks the slacks of its own iterate)."""
        nx, nu, N = self.m.nx, self.m.nu, self.N
        if x_guess is None and u_guess is None:
            return
        X = np.zeros((self.batch, N + 1, nx)) if x_guess is None else np.swapaxes(self._b(x_guess, (nx, N + 1)), 1, 2)
        U = np.zeros((self.batch, N, nu)) if u_guess is None else np.swapaxes(self._b(u_guess, (nu, N)), 1, 2)
        self._guess = (np.ascontiguousarray(X), np.ascontiguousarray(U))

    def solve(self, x0, x_guess=None, u_guess=None, gamma_guess=None):
        nx, nu, N, B = self.m.nx, self.m.nu, self.N, self.batch
        x0b = self._b(x0, (nx,))
        if x_guess is not None or u_guess is not None:
            self.set_initial_guess(x_guess, u_guess, gamma_guess)
            X, U = self._guess
        elif self._guess is not None:
            Xp, Up = self._guess
            xN = self._engine.ddyn(Xp[:, N], Up[:, N - 1]).cpu().numpy()
            X = np.concatenate([Xp[:, 1:], xN[:, None]], axis=1)
            U = np.concatenate([Up[:, 1:], Up[:, -1:]], axis=1)
        else:
            X = np.stack([x0b * (1.0 - k / N) for k in range(N + 1)], axis=1)
            U = np.zeros((B, N, nu))
        out = self._engine.soft_nmpc(x0b, X, U, self.rho_soft, self.rho_soft_l1, self.max_iter, self.tol)
        status = out["status"].cpu().numpy()
        ok = status == 0
        Xs, Us = out["X"].cpu().numpy(), out["U"].cpu().numpy()
        gam, gam_f = out["gamma"].cpu().numpy(), out["gamma_f"].cpu().numpy()
        self.solution = {'success': bool(ok[0]) if B == 1 else ok, 'status': self._out(status),
                         'sqp_iterations': self._out(out["sqp_iters"].cpu().numpy())}
        if not ok.any():
            return self.solution
        self._guess = (np.where(ok[:, None, None], Xs, X), np.where(ok[:, None, None], Us, U))
        primal_x, primal_u = np.swapaxes(Xs, 1, 2), np.swapaxes(Us, 1, 2)
        cost = np.einsum("bik,ij,bjk->b", primal_x[:, :, :N], self.Q, primal_x[:, :, :N]) + np.einsum("bik,ij,bjk->b", primal_u, self.R, primal_u) \
            + np.einsum("bi,ij,bj->b", primal_x[:, :, N], self.Qf, primal_x[:, :, N]) \
            + self.rho_soft * ((gam ** 2).sum(axis=(1, 2)) + (gam_f ** 2).sum(axis=1)) + self.rho_soft_l1 * gam.sum(axis=(1, 2))
        pack = np.stack([self.pack_primal_nominal(primal_x[b], primal_u[b]) for b in range(B)])
        o = self._out
        self.solution.update({'primal_vec': o(pack), 'primal_x': o(primal_x), 'primal_u': o(primal_u),
                              'primal_gamma': o(np.concatenate([gam.reshape(B, -1), gam_f], axis=1)), 'cost': o(cost),
                              'sqp_iterations': o(out["sqp_iters"].cpu().numpy()), 'status': o(status)})
        return self.solution
