"""Nominal QP of the fast-SLS iteration: assembly in the reference's layout + an independent exact solver.

Follows (reference file:line):
  * ``solver/qp_jit.py:77-192, 194-273``  constraint matrix rows per stage ``[A_k B_k -I]`` then ``G``, terminal ``Gf``,
                                          then ``I * dx0`` (OSQP variant); cost ``P = 2 * blkdiag(Q,R,...,Qf)`` (quirk Q2)
  * ``solver/qp_jit.py:362-402``          bounds handed to OSQP: x0 rows ``-(x0) +- 1e-10``, +-inf -> +-1e20
  * ``solver/qp_jit.py:487-513``          unpacking of primal / dual vectors (``dual_mu`` is (ni, N))
  * ``solver/qp_jit.py:578-610``          update_ubg / offset_constraints bound bookkeeping
The reference solves this QP with OSQP 1.0.4 (``requirements.txt:31``; un-vendored C library, absent here), default
tolerances + polish (quirk Q1).  PARITY UNPINNED at that boundary: no OSQP output exists to compare with (the
published algorithm is restated in ``oracle/osqp_restated.py`` and lands on this file's solution).  The QP
is strictly convex (P = 2*diag(Q,R,..) > 0), its primal solution is unique, and OSQP's polished answer is that
solution.  This file therefore solves the same QP to high accuracy with an independent algorithm: a dense
Mehrotra primal-dual interior-point method followed by an exact active-set refinement; ``kkt_residuals`` lets the
tests assert <= 1e-9 on every fixture.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np

EPSILON = 1e-10  # solver/qp_jit.py:19
BIG = 1e20       # solver/qp_jit.py:382


class QPData:
    """Dense restatement of the matrices ``QP`` keeps in CSC form (solver/qp_jit.py:77-192)."""

    def __init__(self, N, nx, nu, Q, R, Qf, G, Gf):
        self.N, self.nx, self.nu = N, nx, nu
        self.ni, self.ni_f = G.shape[0], Gf.shape[0]
        self.G, self.Gf = np.asarray(G, float), np.asarray(Gf, float)
        nz = nx + nu
        self.nv = nz * N + nx
        self.m_main = N * (nx + self.ni) + self.ni_f      # rows of A_mat_csc
        self.m = self.m_main + nx                         # + x0 rows (A_mat_csc_osqp)
        H = np.zeros((self.nv, self.nv))
        for k in range(N):
            H[k * nz:k * nz + nx, k * nz:k * nz + nx] = Q
            H[k * nz + nx:(k + 1) * nz, k * nz + nx:(k + 1) * nz] = R
        H[N * nz:, N * nz:] = Qf
        self.H = H            # P_mat_csc ; OSQP sees P = 2 H
        self.Acon = None
        # row index helpers
        self.dyn_rows = np.concatenate([np.arange(k * (nx + self.ni), k * (nx + self.ni) + nx) for k in range(N)])
        self.ineq_rows = np.concatenate(
            [np.arange(k * (nx + self.ni) + nx, (k + 1) * (nx + self.ni)) for k in range(N)]
            + [np.arange(N * (nx + self.ni), self.m_main)])
        self.x0_rows = np.arange(self.m_main, self.m)

    def set_dynamics(self, A_list, B_list):
        """solver/qp_jit.py:194-249."""
        N, nx, nu, ni = self.N, self.nx, self.nu, self.ni
        nz = nx + nu
        Ac = np.zeros((self.m, self.nv))
        r = 0
        for k in range(N):
            Ac[r:r + nx, k * nz:k * nz + nx] = A_list[k]
            Ac[r:r + nx, k * nz + nx:(k + 1) * nz] = B_list[k]
            Ac[r:r + nx, (k + 1) * nz:(k + 1) * nz + nx] = -np.eye(nx)
            r += nx
            Ac[r:r + ni, k * nz:(k + 1) * nz] = self.G
            r += ni
        Ac[r:r + self.ni_f, N * nz:] = self.Gf
        r += self.ni_f
        Ac[r:r + nx, :nx] = np.eye(nx)
        self.Acon = Ac

    X0_TOL = 1e-3

    def bounds(self, lbg, ubg, x0):
        """solver/qp_jit.py:370-386: append the x0 box and sanitise infinities.

        Stage-0 state rows of G only restate the pinned x(0) = -x0: they are either redundant or make the QP
        infeasible.  OSQP with its default eps_abs = 1e-3 (quirk Q1) accepts violations of that size, an exact solver
        does not; both the oracle and the CUDA path therefore drop these rows when the violation is <= X0_TOL and keep
        them (infeasible QP -> success False) beyond."""
        l = np.concatenate([np.asarray(lbg, float).ravel(), -np.asarray(x0, float).ravel() - EPSILON])
        u = np.concatenate([np.asarray(ubg, float).ravel(), -np.asarray(x0, float).ravel() + EPSILON])
        nx, nz = self.nx, self.nx + self.nu
        d0 = -np.asarray(x0, float).ravel()
        up, lo = np.arange(nx, 2 * nx), np.arange(nx + nz, nx + nz + nx)     # rows of stage 0: [dyn nx | G upper nz | G lower nz]
        if np.all(d0 <= u[up] + self.X0_TOL) and np.all(-d0 <= u[lo] + self.X0_TOL):
            u[up], u[lo] = np.inf, np.inf
        l = np.where(np.isneginf(l), -BIG, np.where(np.isposinf(l), BIG, l))
        u = np.where(np.isposinf(u), BIG, np.where(np.isneginf(u), -BIG, u))
        return l, u

    def unpack(self, y):
        """solver/ocp.py:58-74."""
        N, nx, nu = self.N, self.nx, self.nu
        nz = nx + nu
        X = np.zeros((nx, N + 1))
        U = np.zeros((nu, N))
        for k in range(N):
            X[:, k] = y[k * nz:k * nz + nx]
            U[:, k] = y[k * nz + nx:(k + 1) * nz]
        X[:, N] = y[N * nz:]
        return X, U

    def pack(self, X, U):
        """solver/ocp.py:43-56."""
        parts = []
        for k in range(self.N):
            parts += [X[:, k], U[:, k]]
        parts.append(X[:, self.N])
        return np.concatenate(parts)

    def pack_solution(self, y, lam_full, cost):
        """solver/qp_jit.py:487-513 (osqp backend drops the last nx duals)."""
        N, nx, ni, ni_f = self.N, self.nx, self.ni, self.ni_f
        X, U = self.unpack(y)
        lam = lam_full[:-nx]
        mu_f = lam[-ni_f:]
        mu = lam[:-ni_f].reshape(N, nx + ni)[:, nx:].T
        return dict(success=True, primal_vec=y.copy(), primal_x=X, primal_u=U, dual_vec=lam.copy(),
                    dual_mu=mu.copy(), dual_mu_f=mu_f.copy(), cost=float(cost))


def kkt_residuals(H, q, Acon, l, u, x, y):
    """Residuals of  min x'Hx + q'x  s.t. l <= Acon x <= u  (OSQP form with P = 2H):
    stationarity ||2Hx + q + A'y||_inf, primal violation, and complementarity/sign violation of y."""
    Ax = Acon @ x
    stat = np.abs(2 * H @ x + q + Acon.T @ y).max()
    prim = max(np.maximum(Ax - u, 0).max(), np.maximum(l - Ax, 0).max())
    fin_u, fin_l = u < BIG / 10, l > -BIG / 10
    eq = fin_u & fin_l & (u - l <= 1e-8)
    yp, ym = np.maximum(y, 0), np.maximum(-y, 0)
    comp = 0.0
    ineq = ~eq
    if ineq.any():
        comp = max(np.abs(np.where(fin_u, yp * (u - Ax), yp))[ineq].max(),
                   np.abs(np.where(fin_l, ym * (Ax - l), ym))[ineq].max())
    return dict(stat=stat, prim=prim, comp=comp)


def solve_qp(H, q, Acon, l, u, tol=1e-11, max_iter=80):
    """Exact solve of  min x'Hx + q'x  s.t. l <= Acon x <= u.

    Rows with u - l <= 1e-8 are equalities at the interval midpoint (the reference pins dynamics and x0 rows to
    +-1e-10 intervals, solver/qp_jit.py:378-379,606-608); other rows are one- or two-sided inequalities.
    Returns (x, y, info) with y in OSQP's sign convention (y > 0 on an active upper bound)."""
    n = H.shape[0]
    P = 2.0 * H
    fin_u, fin_l = u < BIG / 10, l > -BIG / 10
    eq = fin_u & fin_l & (u - l <= 1e-8)
    E, e = Acon[eq], 0.5 * (l[eq] + u[eq])
    up, lo = fin_u & ~eq, fin_l & ~eq
    C = np.vstack([Acon[up], -Acon[lo]])
    d = np.concatenate([u[up], -l[lo]])
    me, mi = E.shape[0], C.shape[0]

    def kkt_solve(Wd, r1, r2):
        # [P + C'WC  E'; E  0] [dx; dy] = [r1; r2]
        M = P + (C.T * Wd) @ C
        KK = np.block([[M, E.T], [E, np.zeros((me, me))]])
        sol = np.linalg.solve(KK, np.concatenate([r1, r2]))
        return sol[:n], sol[n:]

    # --- start: equality-constrained solution, slacks pushed interior
    x, yE = kkt_solve(np.zeros(mi), -q, e)
    s = d - C @ x
    shift = max(1.0, 1.5 * max(0.0, -s.min())) if mi else 0.0
    s = np.maximum(s, 0) + shift if mi else s
    z = np.ones(mi)
    it = 0
    for it in range(max_iter):
        rd = P @ x + q + E.T @ yE + C.T @ z
        re = E @ x - e
        rc = C @ x + s - d
        mu = (s @ z) / max(mi, 1)
        if max(np.abs(rd).max(), np.abs(re).max() if me else 0, np.abs(rc).max() if mi else 0) < tol and mu < tol:
            break
        W = z / s
        # affine step
        rsz = -(s * z)
        rhs1 = -rd - C.T @ ((rsz + z * rc) / s)
        dx, dy = kkt_solve(W, rhs1, -re)
        ds = -rc - C @ dx
        dz = (rsz - z * ds) / s
        ap = min(1.0, (-s[ds < 0] / ds[ds < 0]).min()) if (ds < 0).any() else 1.0
        ad = min(1.0, (-z[dz < 0] / dz[dz < 0]).min()) if (dz < 0).any() else 1.0
        mu_aff = ((s + ap * ds) @ (z + ad * dz)) / max(mi, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        rsz = -(s * z) + sigma * mu - ds * dz
        rhs1 = -rd - C.T @ ((rsz + z * rc) / s)
        dx, dy = kkt_solve(W, rhs1, -re)
        ds = -rc - C @ dx
        dz = (rsz - z * ds) / s
        tau = 0.995
        ap = min(1.0, tau * (-s[ds < 0] / ds[ds < 0]).min()) if (ds < 0).any() else 1.0
        ad = min(1.0, tau * (-z[dz < 0] / dz[dz < 0]).min()) if (dz < 0).any() else 1.0
        x = x + ap * dx
        s = s + ap * ds
        yE = yE + ad * dy
        z = z + ad * dz
    info = dict(iters=it, mu=float((s @ z) / max(mi, 1)), polished=False)

    # --- exact active-set refinement (what OSQP's polish does, but verified)
    if mi:
        act = z > s
        for _ in range(10):
            Ca, da = C[act], d[act]
            ma = Ca.shape[0]
            KK = np.block([[P, E.T, Ca.T], [E, np.zeros((me, me + ma))], [Ca, np.zeros((ma, me + ma))]])
            rhs = np.concatenate([-q, e, da])
            sol, *_ = np.linalg.lstsq(KK, rhs, rcond=None)
            xa, ya, za = sol[:n], sol[n:n + me], sol[n + me:]
            # one step of refinement on the (possibly rank-deficient) system
            res = rhs - KK @ sol
            corr, *_ = np.linalg.lstsq(KK, res, rcond=None)
            sol = sol + corr
            xa, ya, za = sol[:n], sol[n:n + me], sol[n + me:]
            sa = d - C @ xa
            bad_primal = (~act) & (sa < -1e-11)
            zfull = np.zeros(mi)
            zfull[act] = za
            bad_dual = act & (zfull < -1e-11)
            if not bad_primal.any() and not bad_dual.any():
                x, yE, z, s = xa, ya, np.maximum(zfull, 0.0), sa
                info["polished"] = True
                break
            act = (act | bad_primal) & ~bad_dual
    y = np.zeros(Acon.shape[0])
    y[eq] = yE
    nu_, nl_ = int(up.sum()), int(lo.sum())
    yu = np.zeros(Acon.shape[0])
    yu[up] = z[:nu_]
    yl = np.zeros(Acon.shape[0])
    yl[lo] = z[nu_:nu_ + nl_]
    y = y + yu - yl
    info["n_active"] = int((z > 1e-12).sum()) if mi else 0
    return x, y, info
