"""Gap between "OSQP with default settings + polish" (restated: oracle/osqp_restated.py) and the exact solution of the
nominal QP (oracle/qp.py, what the CUDA interior point computes) on the committed QP fixtures, with the bounds exactly
as the reference hands them over (solver/qp_jit.py:370-386: +-1e20 for infinity, x0 rows +-1e-10, stage-0 state rows
of G kept).  CPU only:  python -m oracle.osqp_gap > profiles/r3f_osqp_restated_gap.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import models as om, scp as oscp, osqp_restated as osr  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def fixture_qps(name):
    """Yields (index, tightened?, P, q, A, l, u, exact primal, exact inequality multipliers) per stored QP."""
    g = np.load(os.path.join(GOLD, f"qp_{name}.npz"))
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    solver = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=1)
    for i in range(len(g["X"])):
        X, U = g["X"][i].T, g["U"][i].T
        solver.set_nominal(X, U)
        solver.update_jacobian()
        f = solver.fast
        if g["tight"][i]:
            tight = np.stack(f.g_list[:-1]) - g["backoff"][i]
            cmat = np.stack(f.c_list, axis=1)
            stage = np.vstack([-cmat, tight.T]).reshape(-1, order="F")
            f.ubg = np.concatenate([stage, f.gf_model - g["backoff_f"][i]])
        x0 = X[:, 0] - g["x0"][i]
        l = np.concatenate([np.asarray(f.lbg, float).ravel(), -x0 - 1e-10])
        u = np.concatenate([np.asarray(f.ubg, float).ravel(), -x0 + 1e-10])
        l = np.where(np.isneginf(l), -1e20, l)
        u = np.where(np.isposinf(u), 1e20, u)
        y = f.qp.pack(g["dX"][i].T, g["dU"][i].T)
        mu = np.concatenate([g["mu"][i].ravel(), g["mu_f"][i]])
        yield i, int(g["tight"][i]), f.qp, 2 * f.qp.H, np.asarray(f.q, float).ravel(), f.qp.Acon, l, u, y, mu


def gaps(qp, x, lam, y, mu):
    d = qp.pack_solution(x, lam, 0.0)
    mu_s = np.concatenate([d["dual_mu"].T.ravel(), d["dual_mu_f"]])
    return np.abs(x - y).max() / max(1.0, np.abs(y).max()), np.abs(mu_s - mu).max() / max(1.0, np.abs(mu).max())


def main():
    print("fixture  tight  eps    status             iters  polish  |x-x*|/max(1,|x*|)  ADMM point  |mu-mu*| rel  wrong-sign duals  s")
    for name in ("pendulum", "quadrotor", "rocket"):
        for i, tight, qp, P, q, A, l, u, y, mu in fixture_qps(name):
            for eps in (1e-3, 1e-4, 1e-5):
                t = time.time()
                x, lam, info = osr.solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps)
                Ax = A @ x
                bad = int(np.sum((lam > 1e-9) & (u - Ax > 1e-7)) + np.sum((lam < -1e-9) & (Ax - l > 1e-7)))
                dx, dmu = gaps(qp, x, lam, y, mu)
                dxa, _ = gaps(qp, info["x_admm"], info["y_admm"], y, mu)
                print(f"{name:9s}{i}  {tight}      {eps:.0e}  {info['status']:18s} {info['iters']:5d}  {str(info['polished']):6s}  "
                      f"{dx:.2e}            {dxa:.2e}    {dmu:.2e}      {bad:3d}               {time.time() - t:.2f}")
                if info["polished"] and bad == 0 and dx < 1e-6:
                    break


if __name__ == "__main__":
    main()
