"""Soft-constrained NMPC comparator of the reference: CPU restatement.

Follows (reference file:line):
  * ``solver/nlp_soft_constraints.py:175-276``  the NLP: multiple-shooting dynamics ``Z[:,k+1] = ddyn(Z[:,k], V[:,k])``, ``Z[:,0] = p``;
        every inequality row softened, ``G [x_k;u_k] - g <= gamma_k``, ``Gf x_N - gf <= gamma_f``, ``gamma >= 0``; objective
        ``sum x'Qx + u'Ru + x_N'Qf x_N + rho * (sumsqr(gamma) + sumsqr(gamma_f)) + rho_l1 * sum(gamma)`` (the L1 term on the STAGE slacks
        only, :226-232)
  * ``expe/main_rocket_compare_closed_loop.py:176-222``  ``_run_soft``: rho_soft = rho_soft_l1 = 1e6, closed loop with the pre-drawn
        disturbances, first input applied
The reference hands this NLP to IPOPT (bundled with CasADi, absent here: PARITY UNPINNED at that boundary).  IPOPT returns a local
minimiser; this file computes one by sequential quadratic programming: each iteration linearises the dynamics about the current
trajectory and solves the convex QP in (delta, gamma) exactly (oracle/qp.py: dense interior point + active-set refinement), full
steps, until the step vanishes -- at which point the QP's optimality conditions ARE the first-order conditions of the NLP
(``kkt_nlp`` checks them independently of how the point was found).  The reference starts IPOPT from the zero vector at every MPC step
(it never passes a guess, :176-190); a local method needs a start and this restatement (like the CUDA path) uses the shifted previous
solution, the interpolation of the nominal initialiser for the first step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np

from . import qp as oqp
from .scp import SCPSLS


class SoftQP:
    """Dense QP of one SQP iteration in the variables [delta (stage-wise); gamma_0 .. gamma_{N-1}; gamma_f]."""

    def __init__(self, m, N, Q, R, Qf, rho, rho1):
        self.m, self.N, self.rho, self.rho1 = m, N, float(rho), float(rho1)
        self.scp = SCPSLS(N, Q, R, m, Qf, rti=1, fast_sls_rti_steps=None)     # linearisation + the hard QP's matrices
        self.qpd = self.scp.fast.qp
        self.ng = N * m.ni + m.ni_f
        # the dense solver works in scaled slacks t = sc * gamma (rho gamma^2 = t^2 for sc = sqrt(rho)): with rho = 1e6 the unscaled
        # problem mixes unit and 1e6 curvatures and its interior-point iteration stalls
        self.sc = np.sqrt(self.rho) if self.rho > 0 else 1.0

    def solve(self, X, U, x0):
        """One QP about (X, U): returns dX, dU, gamma (N, ni), gamma_f, multipliers of the softened rows, cost of the QP."""
        m, N, qpd = self.m, self.N, self.qpd
        nx, nu, ni, nif = m.nx, m.nu, m.ni, m.ni_f
        nz, nv, ng = nx + nu, qpd.nv, self.ng
        self.scp.set_nominal(X, U)
        self.scp.update_jacobian()
        f = self.scp.fast
        Ac = qpd.Acon                                            # rows: per stage [dyn nx | G ni], terminal Gf, x0 rows
        n_rows = Ac.shape[0]
        A_ext = np.zeros((n_rows + ng, nv + ng))
        A_ext[:n_rows, :nv] = Ac
        # - gamma on the softened rows
        gcol = 0
        for k in range(N):
            r = k * (nx + ni) + nx
            A_ext[r:r + ni, nv + gcol:nv + gcol + ni] = -np.eye(ni) / self.sc
            gcol += ni
        r = N * (nx + ni)
        A_ext[r:r + nif, nv + gcol:nv + gcol + nif] = -np.eye(nif) / self.sc
        # gamma >= 0
        A_ext[n_rows:, nv:] = np.eye(ng)
        # bounds of the hard QP without the +1e-10 relaxation and without dropping anything: the slacks absorb every violation
        l = np.concatenate([np.asarray(f.lbg, float).ravel(), -(X[:, 0] - x0), np.zeros(ng)])
        u = np.concatenate([np.asarray(f.ubg, float).ravel(), -(X[:, 0] - x0), np.full(ng, oqp.BIG)])
        l = np.where(np.isneginf(l), -oqp.BIG, l)
        u = np.where(np.isposinf(u), oqp.BIG, u)
        u[qpd.ineq_rows] -= oqp.EPSILON                           # (update_ubg adds 1e-10 to the hard bounds; the soft NLP has none)
        H_ext = np.zeros((nv + ng, nv + ng))
        H_ext[:nv, :nv] = qpd.H
        H_ext[nv:, nv:] = (self.rho / self.sc ** 2) * np.eye(ng)
        q_ext = np.concatenate([f.q, np.concatenate([np.full(N * ni, self.rho1 / self.sc), np.zeros(nif)])])
        y, lam, info = oqp.solve_qp(H_ext, q_ext, A_ext, l, u)
        res = oqp.kkt_residuals(H_ext, q_ext, A_ext, l, u, y, lam)
        dX, dU = qpd.unpack(y[:nv])
        gam = y[nv:nv + N * ni].reshape(N, ni) / self.sc
        gam_f = y[nv + N * ni:] / self.sc
        mu = lam[qpd.ineq_rows]
        return dict(dX=dX, dU=dU, gamma=gam, gamma_f=gam_f, mu=mu[:N * ni].reshape(N, ni), mu_f=mu[N * ni:], kkt=res,
                    lam_dyn=lam[qpd.dyn_rows].reshape(N, nx), cost_qp=float(y @ H_ext @ y + q_ext @ y))


def soft_cost(m, Q, R, Qf, X, U, gamma, gamma_f, rho, rho1):
    """Objective of solver/nlp_soft_constraints.py:222-232."""
    N = U.shape[1]
    c = X[:, N] @ Qf @ X[:, N] + sum(X[:, k] @ Q @ X[:, k] + U[:, k] @ R @ U[:, k] for k in range(N))
    return float(c + rho * (np.sum(gamma ** 2) + np.sum(gamma_f ** 2)) + rho1 * np.sum(gamma))


def slacks_of(m, X, U):
    """The optimal slacks of a trajectory: gamma = max(0, G z - g) (the objective is increasing in every slack)."""
    N = U.shape[1]
    gam = np.stack([np.maximum(0.0, m.G @ np.concatenate([X[:, k], U[:, k]]) - m.g) for k in range(N)])
    gam_f = np.maximum(0.0, m.Gf @ X[:, N] - m.gf)
    return gam, gam_f


def soft_nlp_solve(m, N, Q, R, Qf, x0, X, U, rho=1e6, rho1=1e6, max_iter=40, tol=1e-9, sqp=None):
    """SQP from the guess (X, U) [shapes (nx, N+1), (nu, N)] -> dict(success, primal_x, primal_u, primal_gamma, cost, iterations, ...)."""
    x0 = np.asarray(x0, float).reshape(-1)
    sqp = sqp or SoftQP(m, N, Q, R, Qf, rho, rho1)
    X, U = np.array(X, float), np.array(U, float)
    step, it, out = np.inf, 0, None
    for it in range(max_iter):
        out = sqp.solve(X, U, x0)
        X, U = X + out["dX"], U + out["dU"]
        step = max(np.abs(out["dX"]).max(), np.abs(out["dU"]).max())
        if step < tol:
            break
    gam, gam_f = out["gamma"], out["gamma_f"]
    return dict(success=bool(step < 1e-6), primal_x=X, primal_u=U, primal_gamma=np.concatenate([gam.reshape(-1), gam_f]),
                gamma=gam, gamma_f=gam_f, cost=soft_cost(m, Q, R, Qf, X, U, gam, gam_f, rho, rho1), iterations=it + 1, last_step=float(step),
                mu=out["mu"], mu_f=out["mu_f"], lam_dyn=out["lam_dyn"], kkt_qp=out["kkt"])


def kkt_nlp(m, N, Q, R, Qf, x0, sol, rho=1e6, rho1=1e6):
    """First-order conditions of the NLP at `sol`, evaluated from scratch (complex-step Jacobians of the RK4 map): dynamics defect,
    stationarity in (x, u) with the returned multipliers, stationarity in the slacks (2 rho gamma + rho1 - mu = nu >= 0),
    complementarity.  Returns the worst residual of each group."""
    X, U, gam, gam_f, mu, mu_f, lam = sol["primal_x"], sol["primal_u"], sol["gamma"], sol["gamma_f"], sol["mu"], sol["mu_f"], sol["lam_dyn"]
    nx, nu = m.nx, m.nu
    defect = max(np.abs(m.ddyn(X[:, k], U[:, k]) - X[:, k + 1]).max() for k in range(N))
    defect = max(defect, np.abs(X[:, 0] - np.asarray(x0, float)).max())
    stat = 0.0
    for k in range(1, N + 1):
        gx = 2 * (Qf if k == N else Q) @ X[:, k]
        gx = gx + ((m.Gf.T @ mu_f) if k == N else (m.G[:, :nx].T @ mu[k]))
        gx = gx - lam[k - 1]                                              # row  A_{k-1} x + B_{k-1} u - x_k
        if k < N:
            A, _ = m.jacobians(X[:, k], U[:, k])
            gx = gx + A.T @ lam[k]
        stat = max(stat, np.abs(gx).max())
    for k in range(N):
        _, B = m.jacobians(X[:, k], U[:, k])
        gu = 2 * R @ U[:, k] + m.G[:, nx:].T @ mu[k] + B.T @ lam[k]
        stat = max(stat, np.abs(gu).max())
    nu_s = 2 * rho * gam + rho1 - mu                                      # multiplier of gamma >= 0
    nu_f = 2 * rho * gam_f - mu_f
    viol = np.stack([m.G @ np.concatenate([X[:, k], U[:, k]]) - m.g for k in range(N)]) - gam
    viol_f = m.Gf @ X[:, N] - m.gf - gam_f
    sign = max(0.0, -nu_s.min(), -nu_f.min(), -mu.min(), -mu_f.min(), -gam.min(), -gam_f.min(), viol.max(), viol_f.max())
    comp = max(np.abs(nu_s * gam).max(), np.abs(nu_f * gam_f).max(), np.abs(mu * viol).max(), np.abs(mu_f * viol_f).max())
    return dict(defect=float(defect), stationarity=float(stat), sign=float(sign), complementarity=float(comp))


def soft_closed_loop(m, cfg, x0, W, steps, X0, U0, rho=1e6, rho1=1e6):
    """expe/main_rocket_compare_closed_loop.py:176-222 (``_run_soft``) with the warm start described in the module docstring."""
    N, Q, R, Qf = cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"]
    sqp = SoftQP(m, N, Q, R, Qf, rho, rho1)
    x = np.asarray(x0, float).copy()
    X, U = np.array(X0, float), np.array(U0, float)
    out = dict(X=[x.copy()], U=[], Xn=[], Un=[], gamma=[], cost=[], iterations=[], success=[])
    for i in range(steps):
        if i > 0:       # shift (the same rule as SCP_SLS.reset_warm_start, SCP_SLS_jit.py:505-518)
            xN = m.ddyn(X[:, N], U[:, N - 1])
            X = np.concatenate([X[:, 1:], xN[:, None]], axis=1)
            U = np.concatenate([U[:, 1:], U[:, -1:]], axis=1)
        sol = soft_nlp_solve(m, N, Q, R, Qf, x, X, U, rho, rho1, sqp=sqp)
        out["success"].append(sol["success"])
        if not sol["success"]:
            break
        X, U = sol["primal_x"], sol["primal_u"]
        out["Xn"].append(X.copy()); out["Un"].append(U.copy()); out["gamma"].append(sol["primal_gamma"].copy())
        out["cost"].append(sol["cost"]); out["iterations"].append(sol["iterations"])
        u0 = U[:, 0].copy()
        out["U"].append(u0)
        x = m.ddyn(x, u0) + (m.E @ W[i] if W is not None else 0.0)
        out["X"].append(x.copy())
    return {k: (np.array(v) if k not in ("success",) else v) for k, v in out.items()}
