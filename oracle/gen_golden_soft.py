"""Golden fixtures of the soft-constrained NMPC comparator (tests/golden/soft_*.npz) from the CPU restatement oracle/soft.py.

    python -m oracle.gen_golden_soft

rocket: the compare driver's scenario (expe/main_rocket_compare_closed_loop.py:226-241: fixed x0, W = -0.8 on every state, rho = rho_l1 =
1e6), first 3 MPC steps; pendulum: 4 steps from a state outside its box (active slacks), with disturbances.  The dense SQP needs ~2.5 min
per rocket step.  TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np

from . import models as om, scp as oscp, soft as osoft

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run(name, m, cfg, x0, W, steps, x0_nom=None):
    N, Q, R, Qf = cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"]
    X0, U0, ok = oscp.nominal_sqp(m, N, Q, R, Qf, x0 if x0_nom is None else x0_nom, iters=10)
    assert ok
    out = osoft.soft_closed_loop(m, cfg, x0, W, steps, X0, U0)
    assert all(out["success"]) and len(out["Xn"]) == steps
    np.savez(os.path.join(GOLDEN, f"soft_{name}.npz"), x0=x0, W=W, X0=X0.T, U0=U0.T, states=out["X"], inputs=out["U"],
             Xn=np.swapaxes(out["Xn"], 1, 2), Un=np.swapaxes(out["Un"], 1, 2), gamma=out["gamma"], cost=out["cost"],
             iterations=out["iterations"], rho=1e6, rho1=1e6)
    print(name, "iterations", out["iterations"], "cost", out["cost"], "max slack", [float(g.max()) for g in out["gamma"]])


if __name__ == "__main__":
    m, cfg = om.pendulum_config()
    x0 = om.PENDULUM_X0.copy()
    x0[0] = 1.05 * m.x_ub[0]                                   # outside the box: the slacks carry the violation
    W = np.random.default_rng(3).uniform(-1, 1, size=(4, m.nx))
    run("pendulum", m, cfg, x0, W, 4, x0_nom=om.PENDULUM_X0)
    m, cfg = om.rocket_config()
    run("rocket", m, cfg, om.ROCKET_X0.copy(), -0.8 * np.ones((3, m.nx)), 3)
