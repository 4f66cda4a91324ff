"""What "OSQP with default settings" (restated: oracle/osqp_restated.py) does when the measured state x(0) lies outside the
state box by d, i.e. when the stage-0 state row of G contradicts the pinned x(0) (solver/qp_jit.py:370-386 keeps both).
The exact solvers of this repository (oracle/qp.py, the CUDA interior point) drop those rows for d <= 1e-3 and fail the
instance beyond (X0_TOL).  CPU only:  python -m oracle.x0_policy_gap > profiles/r3f_x0_policy_gap.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import osqp_restated as osr, qp as oqp  # noqa: E402
from oracle.osqp_gap import fixture_qps  # noqa: E402

print("fixture     d       status             iters  polish  excess over the stage-0 row  x0-row residual  |x - x*(rows dropped)| rel")
for name, idx in (("pendulum", 0), ("quadrotor", 0), ("rocket", 3)):
    for i, tight, qp, P, q, A, l, u, y, mu in fixture_qps(name):
        if i != idx:
            continue
        j = 0
        row_up, x0row = qp.nx + j, qp.x0_rows[j]
        pinned = 0.5 * (l[x0row] + u[x0row])
        for d in (0.0, 1e-4, 5e-4, 1e-3, 2e-3, 1e-2, 1e-1):
            u2 = u.copy()
            if d > 0:
                u2[row_up] = pinned - d                 # the box ends d below the measured state
            x, lam, info = osr.solve(P, q, A, l, u2)
            Ax = A @ x
            # the exact answer of this repository's policy: the same QP without the stage-0 state rows
            u3 = u2.copy()
            u3[qp.nx:2 * qp.nx] = oqp.BIG
            u3[qp.nx + qp.nx + qp.nu:qp.nx + qp.nx + qp.nu + qp.nx] = oqp.BIG
            xs, _, _ = oqp.solve_qp(P / 2, q, A, l, u3)
            print(f"{name:9s}{i}  {d:.0e}  {info['status']:18s} {info['iters']:5d}  {str(info['polished']):6s}  "
                  f"{Ax[row_up] - u2[row_up]:+.2e}                    {abs(Ax[x0row] - pinned):.2e}         "
                  f"{np.abs(x - xs).max() / max(1.0, np.abs(xs).max()):.2e}")
