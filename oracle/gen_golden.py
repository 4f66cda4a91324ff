"""Generate the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

    python -m oracle.gen_golden [sls] [jac] [qp] [loops]

  sls_*.npz    outputs of the reference's OWN Numba kernels (solver/fast_SLS_jit.py:65-188, imported under stubs by
               oracle/ref_kernels.py) on seeded inputs -- pins oracle/sls.py and the CUDA fast-SLS kernel.
  jac_*.npz    A, B, c at fixed points from SymPy differentiation of the ODE right-hand side chained through RK4 --
               pins oracle/models.py (complex-step) and the CUDA linearisation kernel (dual numbers).
  qp_*.npz     nominal QPs taken from SCP iterations with their oracle solutions (KKT residual asserted <= 1e-9).
  loop_*.npz   closed loops of the three reference drivers computed by the oracle (with the reference kernels in
               the loop where available).
Inputs are regenerated from seeds by ``sls_inputs`` so the fixtures only hold outputs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import models as om
from . import qp as oqp
from . import ref_kernels, scp, sls

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sls_inputs(name, seed=0):
    """Seeded synthetic fast-SLS inputs (recipe of BASELINE.md section 2)."""
    m, cfg = om.CONFIGS[name]()
    N, nx, nu = cfg["N"], m.nx, m.nu
    rng = np.random.default_rng(seed)
    A = np.eye(nx) + 0.05 * rng.standard_normal((N, nx, nx))
    B = 0.05 * rng.standard_normal((N, nx, nu))
    eta = 5e4 * np.abs(rng.standard_normal((N, N, m.ni))) * (rng.random((N, N, m.ni)) < 0.1)
    eta_f = 5e4 * np.abs(rng.standard_normal((N + 1, m.ni_f))) * (rng.random((N + 1, m.ni_f)) < 0.1)
    return m, cfg, A, B, eta, eta_f


def gen_sls():
    ref = ref_kernels.load()
    for name in om.MODELS:
        m, cfg, A, B, eta, eta_f = sls_inputs(name)
        N, nx, nu = cfg["N"], m.nx, m.nu
        E = np.ascontiguousarray(np.stack([m.E] * (N + 1)))
        S, K = ref._backward_solve_numba(N, nx, nu, A, B, np.ascontiguousarray(m.G), np.ascontiguousarray(m.Gf), eta, eta_f,
                                         cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"])
        Px, Pu = ref._propagate(A, B, E, K)
        beta, beta_f, bo, bof = ref._backoff_from_phi(Px, Pu, np.ascontiguousarray(m.G[:, :nx]),
                                                      np.ascontiguousarray(m.G[:, nx:]), np.ascontiguousarray(m.Gf), 1e-10)
        np.savez_compressed(os.path.join(GOLDEN, f"sls_{name}.npz"), K=K, beta=beta, beta_f=beta_f, backoff=bo, backoff_f=bof,
                            S_diag=np.einsum("kjii->kji", S), S_sum=S.sum(axis=(2, 3)), Phi_x_last=Px[N], Phi_x_sum=Px.sum(axis=(2, 3)),
                            Phi_u_sum=Pu.sum(axis=(2, 3)), Phi_x_abs=np.abs(Px).sum(), Phi_u_abs=np.abs(Pu).sum(), S_abs=np.abs(S).sum(),
                            source="reference solver/fast_SLS_jit.py Numba kernels, numba " + __import__("numba").__version__)
        print("sls", name, "K", np.abs(K).max(), "backoff", bo.max())


def sympy_jacobians(name, x, u, h=0.05):
    """A, B of the RK4 map from SymPy's symbolic Jacobian of the ODE + the RK4 chain rule."""
    import sympy as sp

    m = om.MODELS[name]()
    xs = sp.symbols(f"x0:{m.nx}")
    us = sp.symbols(f"u0:{m.nu}")

    class SymMath:
        sin, cos, arctan, sqrt = staticmethod(sp.sin), staticmethod(sp.cos), staticmethod(sp.atan), staticmethod(sp.sqrt)
        @staticmethod
        def array(v):
            return list(v)

    # evaluate the NumPy ODE on SymPy symbols by temporarily swapping the math functions it uses
    saved = {k: getattr(np, k) for k in ("sin", "cos", "arctan", "sqrt", "array")}
    try:
        for k in saved:
            setattr(np, k, getattr(SymMath, k))
        f = m.ode(list(xs), list(us))
    finally:
        for k, v in saved.items():
            setattr(np, k, v)
    f = sp.Matrix(f)
    Jx = sp.lambdify((xs, us), f.jacobian(sp.Matrix(xs)), "numpy")
    Ju = sp.lambdify((xs, us), f.jacobian(sp.Matrix(us)), "numpy")
    fn = sp.lambdify((xs, us), f, "numpy")
    ode = lambda xx: np.asarray(fn(xx, u), float).reshape(-1)
    jx = lambda xx: np.asarray(Jx(xx, u), float)
    ju = lambda xx: np.asarray(Ju(xx, u), float)
    I = np.eye(m.nx)
    k1 = ode(x); x2 = x + 0.5 * h * k1
    k2 = ode(x2); x3 = x + 0.5 * h * k2
    k3 = ode(x3); x4 = x + h * k3
    k4 = ode(x4)
    d1x, d1u = jx(x), ju(x)
    d2x, d2u = jx(x2) @ (I + 0.5 * h * d1x), jx(x2) @ (0.5 * h * d1u) + ju(x2)
    d3x, d3u = jx(x3) @ (I + 0.5 * h * d2x), jx(x3) @ (0.5 * h * d2u) + ju(x3)
    d4x, d4u = jx(x4) @ (I + h * d3x), jx(x4) @ (h * d3u) + ju(x4)
    A = I + (h / 6) * (d1x + 2 * d2x + 2 * d3x + d4x)
    B = (h / 6) * (d1u + 2 * d2u + 2 * d3u + d4u)
    xp = x + (1 / 6) * (k1 + 2 * k2 + 2 * k3 + k4) * h
    return A, B, xp


def jac_points(name, n=4, seed=7):
    m = om.MODELS[name]()
    rng = np.random.default_rng(seed)
    xs = [rng.uniform(m.x_lb / 3, m.x_ub / 3) for _ in range(n)]
    mid = 0.5 * (m.u_lb + m.u_ub)
    us = [rng.uniform(mid - 0.3, mid + 0.3) for _ in range(n)]
    if name == "rocket":
        xs[0] = om.ROCKET_X0.copy()
    if name == "pendulum":
        xs[0] = om.PENDULUM_X0.copy()
    return np.array(xs), np.array(us)


def gen_jac():
    for name in om.MODELS:
        xs, us = jac_points(name)
        As, Bs, xps = [], [], []
        for x, u in zip(xs, us):
            A, B, xp = sympy_jacobians(name, x, u)
            As.append(A); Bs.append(B); xps.append(xp)
        np.savez_compressed(os.path.join(GOLDEN, f"jac_{name}.npz"), x=xs, u=us, A=np.array(As), B=np.array(Bs), xp=np.array(xps),
                            source="SymPy jacobian of the ODE + RK4 chain rule")
        m = om.MODELS[name]()
        err = max(np.abs(m.jacobians(x, u)[0] - A).max() for x, u, A in zip(xs, us, As))
        print("jac", name, "complex-step vs sympy", err)


def gen_qp():
    """QPs met along a tube-free SQP from the fixed rocket x0 and along the first closed-loop steps."""
    for name, x0 in (("rocket", om.ROCKET_X0), ("quadrotor", om.quadrotor_x0(0)), ("pendulum", om.PENDULUM_X0)):
        m, cfg = om.CONFIGS[name]()
        N = cfg["N"]
        solver = scp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=cfg["rti"],
                            fast_sls_rti_steps=cfg["fast_sls_rti_steps"])
        Xs, Us, x0s, bos, bofs, tight, sols = [], [], [], [], [], [], []
        orig = solver.fast.forward_solve

        def logged(xarg, _orig=orig):
            ok = _orig(xarg)
            c = solver.fast.cur
            Xs.append(solver.cur["primal_x"].T.copy()); Us.append(solver.cur["primal_u"].T.copy())
            x0s.append(solver.cur["primal_x"][:, 0] - xarg)
            is_tight = "K" in c and c["iteration_number"] > 0 and np.abs(c["backoff_x"]).max() > 0
            tight.append(bool(is_tight))
            bos.append(np.array(c["backoff"])); bofs.append(np.array(c["backoff_f"]))
            sols.append((c["primal_x"].T.copy(), c["primal_u"].T.copy(), c["dual_mu"].T.copy(), c["dual_mu_f"].copy(), c["cost_nominal"]))
            return ok

        solver.fast.forward_solve = logged
        X0, U0, ok = scp.nominal_sqp(m, N, cfg["Q"], cfg["R"], cfg["Qf"], x0, iters=10)
        solver.set_nominal(X0, U0)
        x = np.array(x0, float)
        rng = np.random.default_rng(5)
        for i in range(2):
            if i > 0:
                solver.reset_warm_start()
            sol = solver.solve(x)
            x = m.ddyn(x, sol["primal_u"][:, 0]) + m.E @ rng.uniform(-1, 1, m.nx)
        keep = list(range(len(Xs)))[:8]
        np.savez_compressed(os.path.join(GOLDEN, f"qp_{name}.npz"),
                            X=np.array([Xs[i] for i in keep]), U=np.array([Us[i] for i in keep]), x0=np.array([x0s[i] for i in keep]),
                            tight=np.array([tight[i] for i in keep]), backoff=np.array([bos[i] for i in keep]),
                            backoff_f=np.array([bofs[i] for i in keep]), dX=np.array([sols[i][0] for i in keep]),
                            dU=np.array([sols[i][1] for i in keep]), mu=np.array([sols[i][2] for i in keep]),
                            mu_f=np.array([sols[i][3] for i in keep]), cost=np.array([sols[i][4] for i in keep]),
                            source="oracle/qp.py interior point + active-set refinement")
        print("qp", name, len(keep), "tight", [tight[i] for i in keep])


def _run_loop(name, x0, W, steps, kernels, tag):
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    t0 = time.time()
    X0, U0, ok = scp.nominal_sqp(m, N, cfg["Q"], cfg["R"], cfg["Qf"], x0, iters=10)
    assert ok
    out = scp.closed_loop(m, cfg, x0, W, steps, X0=X0, U0=U0, kernels=kernels)
    n = len(out["U"])
    Kz = np.zeros((N, N + 1, m.nu, m.nx))
    np.savez_compressed(os.path.join(GOLDEN, f"loop_{tag}.npz"), x0=x0, W=np.zeros((0,)) if W is None else W, X0=X0.T, U0=U0.T,
                        states=np.array(out["X"]), inputs=np.array(out["U"]), Xn=np.array([a.T for a in out["Xn"]]),
                        Un=np.array([a.T for a in out["Un"]]), backoff_x=np.array(out["backoff_x"]), backoff_u=np.array(out["backoff_u"]),
                        skipped=np.array(out["skipped"]), K_first=np.array([Kz if k is None else k for k in out["K"]])[:, :3, 0],
                        K_abs=np.array([0.0 if k is None else np.abs(k).sum() for k in out["K"]]),
                        success=np.array(out["success"]), source="oracle/scp.py closed loop")
    print("loop", tag, "steps", n, "skipped", int(np.sum(out["skipped"])), "success", all(out["success"]), "%.1fs" % (time.time() - t0))


def gen_loops(only=None):
    kernels = ref_kernels.load() if ref_kernels.available() else None
    if only == "small":
        _run_loop("pendulum", om.PENDULUM_X0, None, 60, kernels, "pendulum")
        _run_loop("quadrotor", om.quadrotor_x0(0), None, 30, kernels, "quadrotor")
        return
    # g4: rocket, 29 steps, W = -0.8 (expe/main_rocket_compare_closed_loop.py:241) and default_rng(123) draws (:236-240)
    _run_loop("rocket", om.ROCKET_X0, -0.8 * np.ones((29, 17)), 29, kernels, "rocket_w08")
    _run_loop("rocket", om.ROCKET_X0, np.random.default_rng(123).uniform(-1.0, 1.0, size=(29, 17)), 29, kernels, "rocket_rng123")
    # g5: pendulum 60 steps noise free (expe/main_pendulum_robust_closed_loop.py:42-97)
    _run_loop("pendulum", om.PENDULUM_X0, None, 60, kernels, "pendulum")
    # config 2: quadrotor 30 steps noise free, seeded x0
    _run_loop("quadrotor", om.quadrotor_x0(0), None, 30, kernels, "quadrotor")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    what = sys.argv[1:] or ["sls", "jac", "qp", "loops"]
    if "sls" in what:
        gen_sls()
    if "jac" in what:
        gen_jac()
    if "qp" in what:
        gen_qp()
    if "loops" in what:
        gen_loops()
    if "loops_small" in what:
        gen_loops("small")
