"""OSQP restated: the algorithm the reference's nominal QP is handed to (``solver/qp_jit.py:537-548`` set-up with
``polishing=True, warm_starting=False``; ``solver/qp_jit.py:389-402`` solve, ``solved_inaccurate`` accepted).

OSQP 1.0.4 (``requirements.txt:31``) is an un-vendored C library that is absent from ``/root/reference``, from this
image and from the GPU image (``profiles/r2_probe_refdeps.log``).  This file restates its PUBLISHED algorithm --
B. Stellato, G. Banjac, P. Goulart, A. Bemporad, S. Boyd, "OSQP: an operator splitting solver for quadratic programs",
Math. Prog. Comp. 12 (2020): Algorithm 1 (ADMM), section 5.1 (Ruiz equilibration + cost scaling), 5.2 (rho per
constraint, adaptive rho), 3.4 (termination), 4 (polish) -- with the library's default settings:

    rho 0.1 (x 1e3 on equality rows, i.e. u - l < 1e-4), sigma 1e-6, alpha 1.6, scaling 10 passes,
    eps_abs = eps_rel 1e-3, check_termination every 25 iterations, max_iter 4000,
    adaptive rho every 50 iterations with tolerance 5, polish delta 1e-6 with 3 refinement steps,
    bounds beyond +-1e30 clipped (the reference's +-1e20 stand-ins for infinity are FINITE bounds to OSQP).

It is NOT the library: the adaptive-rho schedule of the C code depends on set-up timing unless an interval is
configured, its linear algebra is sparse LDL' (QDLDL), and so the ADMM iterates of the real OSQP differ from these in
the last digits and possibly in the iteration count.  What does not depend on those details -- and what this file is
for -- is the CLASS of answer the reference works with:

  * un-polished: a point within eps = 1e-3 (relative) of the solution,
  * polished: the solution of the equality-constrained QP on the active set the ADMM iterate identifies, which for
    a strictly convex QP with a correctly identified active set IS the exact solution that ``oracle/qp.py`` and the
    CUDA interior point compute.

``tests/test_oracle_cpu.py::test_restated_osqp_*`` measure both gaps on the committed QP fixtures; ``DESIGN.md``
section 4 records them.  PARITY at the QP boundary is therefore anchored on a restatement of the published algorithm,
not on outputs of the library itself ("parity unpinned" against OSQP's binaries stays true and is said so).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

OSQP_INFTY = 1e30
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3
MIN_SCALING, MAX_SCALING = 1e-4, 1e4

DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, scaling=10, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4,
                eps_dual_inf=1e-4, max_iter=4000, check_termination=25, adaptive_rho=True, adaptive_rho_interval=50,
                adaptive_rho_tolerance=5.0, polishing=True, delta=1e-6, polish_refine_iter=3)


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _ruiz(P, q, A, passes):
    """Section 5.1: equilibrate the columns of [[P, A'], [A, 0]], then scale the cost."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    for _ in range(passes):
        nD = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        nE = np.abs(A).max(axis=1) if m else np.zeros(0)
        d, e = 1.0 / np.sqrt(_limit(nD)), 1.0 / np.sqrt(_limit(nE))
        P = d[:, None] * P * d[None, :]
        q = d * q
        A = e[:, None] * A * d[None, :]
        D, E = D * d, E * e
        cost_norm = max(float(_limit(np.array([np.abs(P).max(axis=0).mean()]))[0]),
                        float(_limit(np.array([np.abs(q).max() if n else 0.0]))[0]))
        g = 1.0 / cost_norm
        P, q, c = P * g, q * g, c * g
    return P, q, A, D, E, c


def _rho_vec(l, u, rho):
    loose = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
    eq = (u - l) < RHO_TOL
    return np.where(loose, RHO_MIN, np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho, rho))


class _KKT:
    """[[P + sigma I, A'], [A, -diag(1/rho)]], factorised densely (the library: sparse LDL')."""

    def __init__(self, P, A, sigma, rho_vec):
        n, m = P.shape[0], A.shape[0]
        K = np.zeros((n + m, n + m))
        K[:n, :n] = P + sigma * np.eye(n)
        K[:n, n:] = A.T
        K[n:, :n] = A
        K[n:, n:] = -np.diag(1.0 / rho_vec)
        self.lu = sla.lu_factor(K)
        self.n = n

    def solve(self, r):
        return sla.lu_solve(self.lu, r)


def solve(P, q, A, l, u, **settings):
    """min 1/2 x'Px + q'x  s.t. l <= Ax <= u  (P here is OSQP's P, i.e. 2 H of ``oracle/qp.py``).

    Returns (x, y, info); ``info`` carries status, iterations, the residuals of the returned point, whether the polish
    was adopted, and the un-polished ADMM point (``x_admm``, ``y_admm``)."""
    s = dict(DEFAULTS)
    s.update(settings)
    n, m = P.shape[0], A.shape[0]
    l = np.maximum(np.asarray(l, float), -OSQP_INFTY)
    u = np.minimum(np.asarray(u, float), OSQP_INFTY)
    P0, q0, A0 = np.asarray(P, float), np.asarray(q, float), np.asarray(A, float)
    Ps, qs, As, D, E, c = _ruiz(P0, q0, A0, s["scaling"])
    ls, us = E * l, E * u
    Dinv, Einv = 1.0 / D, 1.0 / E

    rho = s["rho"]
    rv = _rho_vec(ls, us, rho)
    kkt = _KKT(Ps, As, s["sigma"], rv)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    sigma, alpha = s["sigma"], s["alpha"]

    def residuals(x, z, y):
        Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
        r_prim = np.abs(Einv * (Ax - z)).max() if m else 0.0
        r_dual = np.abs(Dinv * (Px + qs + Aty)).max() / c
        n_prim = max(np.abs(Einv * Ax).max(), np.abs(Einv * z).max()) if m else 0.0
        n_dual = max(np.abs(Dinv * Px).max(), np.abs(Dinv * Aty).max(), np.abs(Dinv * qs).max()) / c
        return r_prim, r_dual, n_prim, n_dual

    status, it, rho_updates = "maximum iterations reached", 0, 0
    for it in range(1, s["max_iter"] + 1):
        x_prev, z_prev, y_prev = x, z, y
        sol = kkt.solve(np.concatenate([sigma * x_prev - qs, z_prev - y_prev / rv]))
        xt, nu = sol[:n], sol[n:]
        zt = z_prev + (nu - y_prev) / rv
        x = alpha * xt + (1 - alpha) * x_prev
        zr = alpha * zt + (1 - alpha) * z_prev
        z = np.clip(zr + y_prev / rv, ls, us)
        y = y_prev + rv * (zr - z)

        if it % s["check_termination"] == 0:
            r_prim, r_dual, n_prim, n_dual = residuals(x, z, y)
            if r_prim <= s["eps_abs"] + s["eps_rel"] * n_prim and r_dual <= s["eps_abs"] + s["eps_rel"] * n_dual:
                status = "solved"
                break
            # certificates (section 3.4), on the iterate differences
            dy, dx = y - y_prev, x - x_prev
            ndy = np.abs(E * dy).max() if m else 0.0
            if ndy > 1e-30:
                sup = us @ np.maximum(dy, 0) + ls @ np.minimum(dy, 0)
                if np.abs(Dinv * (As.T @ dy)).max() <= s["eps_prim_inf"] * ndy and sup <= s["eps_prim_inf"] * ndy:
                    status = "primal infeasible"
                    break
        if s["adaptive_rho"] and it % s["adaptive_rho_interval"] == 0:
            # section 5.2: balance the scaled residuals
            Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
            rp = np.abs(Ax - z).max() / max(np.abs(Ax).max(), np.abs(z).max(), 1e-10)
            rd = np.abs(Px + qs + Aty).max() / max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qs).max(), 1e-10)
            rho_new = float(np.clip(rho * np.sqrt(rp / max(rd, 1e-10)), RHO_MIN, RHO_MAX))
            if rho_new > rho * s["adaptive_rho_tolerance"] or rho_new < rho / s["adaptive_rho_tolerance"]:
                rho = rho_new
                rv = _rho_vec(ls, us, rho)
                kkt = _KKT(Ps, As, sigma, rv)
                rho_updates += 1
    if status == "maximum iterations reached":
        r_prim, r_dual, n_prim, n_dual = residuals(x, z, y)
        if r_prim <= 10 * (s["eps_abs"] + s["eps_rel"] * n_prim) and r_dual <= 10 * (s["eps_abs"] + s["eps_rel"] * n_dual):
            status = "solved inaccurate"

    xu, yu, zu = D * x, E * y / c, Einv * z
    info = dict(status=status, iters=it, rho=rho, rho_updates=rho_updates, polished=False, x_admm=xu.copy(), y_admm=yu.copy())

    def unscaled_res(xv, yv, zv):
        return (np.abs(A0 @ xv - zv).max() if m else 0.0), np.abs(P0 @ xv + q0 + A0.T @ yv).max()

    info["r_prim"], info["r_dual"] = unscaled_res(xu, yu, zu)
    if s["polishing"] and status in ("solved", "solved inaccurate"):
        # section 4, on the scaled problem like the library: active sets from the signs of y against the gaps
        low = (z - ls) < -y
        upp = (us - z) < y
        Aact = np.vstack([As[low], As[upp]])
        b = np.concatenate([ls[low], us[upp]])
        ma = Aact.shape[0]
        K = np.zeros((n + ma, n + ma))
        K[:n, :n] = Ps
        K[:n, n:] = Aact.T
        K[n:, :n] = Aact
        Kreg = K.copy()
        Kreg[:n, :n] += s["delta"] * np.eye(n)
        Kreg[n:, n:] -= s["delta"] * np.eye(ma)
        rhs = np.concatenate([-qs, b])
        try:
            lu = sla.lu_factor(Kreg)
            sol = sla.lu_solve(lu, rhs)
            for _ in range(s["polish_refine_iter"]):
                sol = sol + sla.lu_solve(lu, rhs - K @ sol)
            xp = sol[:n]
            yp = np.zeros(m)
            yp[low] = sol[n:n + int(low.sum())]
            yp[upp] = sol[n + int(low.sum()):]
            zp = np.clip(As @ xp, ls, us)
            xpu, ypu, zpu = D * xp, E * yp / c, Einv * zp
            rp, rd = unscaled_res(xpu, ypu, zpu)
            info["n_active"] = ma
            # adopted only when both residuals improve (polish.c); else the ADMM point is returned
            if (rp < info["r_prim"] and rd < info["r_dual"]) or (rp < info["r_prim"] and info["r_dual"] < 1e-10) \
                    or (rd < info["r_dual"] and info["r_prim"] < 1e-10):
                xu, yu = xpu, ypu
                info.update(polished=True, r_prim=rp, r_dual=rd)
        except (sla.LinAlgError, ValueError):
            pass
    info["obj_val"] = float(0.5 * xu @ P0 @ xu + q0 @ xu)
    return xu, yu, info
