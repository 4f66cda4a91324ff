"""CPU-only (oracle): is the first nominal QP of a given quadrotor scenario degenerate, i.e. are the gradients of its active
constraints (dynamics + pinned x0 + active bounds) linearly dependent so that the multipliers are not unique?
    python tools/degeneracy_check.py [scenario index ...]      (population of tools/failure_parity.py: x0 = quadrotor_x0(5 + index))"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import models as om, scp as oscp
from oracle.qp import BIG

for sc in [int(a) for a in sys.argv[1:]] or [1669, 0]:
    m, cfg = om.quadrotor_config()
    x0 = om.quadrotor_x0(5 + sc)
    X0, U0, ok = oscp.nominal_sqp(m, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], x0, iters=10)
    s = oscp.SCPSLS(cfg["N"], cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=cfg["rti"],
                    fast_sls_rti_steps=cfg["fast_sls_rti_steps"])
    s.fast.keep_qp_log = True
    s.set_nominal(X0, U0)
    sol = s.solve(x0)
    log = s.fast.qp_log[0]
    Acon, l, u, x, y = s.fast.qp.Acon, log["l"], log["u"], log["x"], log["y"]
    Ax = Acon @ x
    fin_u, fin_l = u < BIG / 10, l > -BIG / 10
    eq = fin_u & fin_l & (u - l <= 1e-8)
    act_u = fin_u & ~eq & (np.abs(Ax - u) < 1e-7)
    act_l = fin_l & ~eq & (np.abs(Ax - l) < 1e-7)
    act = eq | act_u | act_l
    J = Acon[act]
    sv = np.linalg.svd(J, compute_uv=False)
    rank = int((sv > 1e-9 * sv[0]).sum())
    weak = int(((act_u | act_l) & (np.abs(y) < 1e-8)).sum())
    print("scenario %d (nominal SQP ok %s, solve ok %s): %d variables | active rows %d = %d equalities + %d bounds | rank %d -> %d dependent "
          "| smallest singular values %s | weakly active (multiplier < 1e-8) %d | largest multiplier %.1f" % (
              sc, ok, sol.get("success", True), Acon.shape[1], J.shape[0], int(eq.sum()), int((act_u | act_l).sum()), rank, J.shape[0] - rank,
              np.array2string(sv[-3:], precision=2), weak, np.abs(y[act_u | act_l]).max() if (act_u | act_l).any() else 0.0))
