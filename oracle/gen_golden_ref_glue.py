"""Writes tests/golden/refglue_{pendulum,quadrotor,rocket}.npz: two consecutive passes of the reference's own fast-SLS glue
(oracle/ref_glue.py) on the linearisation and the QP multipliers of the oracle at each model's driver configuration -- the
first pass from the epsilon back-offs (state after initialize_backoff), the second from the first pass's beta and the
multipliers of the tightened QP.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden_ref_glue
"""
import os

import numpy as np

from . import models as om
from . import ref_glue
from . import scp as oscp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def scenario(name):
    """Oracle solver of `name` at its driver configuration, with a nominal trajectory and a start state."""
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=2)
    x0 = {"pendulum": np.array([0.5, 0.5, 0.0, 0.0]), "rocket": om.ROCKET_X0}.get(name)
    if x0 is None:
        rng = np.random.default_rng(0)
        x0 = rng.uniform(m.x_lb / 2, m.x_ub / 2)
        q = rng.normal(size=4)
        x0[6:10] = q / np.linalg.norm(q)
    X, U = oscp.nominal_sqp(m, N, cfg["Q"], cfg["R"], cfg["Qf"], x0)[:2]
    return m, cfg, s, x0, X, U


def main():
    for name in ("pendulum", "quadrotor", "rocket"):
        m, cfg, s, x0, X, U = scenario(name)
        N = cfg["N"]
        s.set_nominal(X, U)
        s.update_jacobian()
        f = s.fast
        rec = {}
        pvecs = []
        for p in range(2):
            assert f.forward_solve(X[:, 0] - x0)
            c = f.cur
            pvecs.append(c["primal_vec"].copy())
            out = ref_glue.reference_pass(N, m, cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], f.A, f.B, f.E, f.g_list, f.c_list,
                                          c["dual_mu"], c["dual_mu_f"], c["beta"], c["beta_f"],
                                          primal_vecs=pvecs + [pvecs[-1] + 5e-4, pvecs[-1] + 2e-3] if p == 1 else ())
            rec.update({f"p{p}_mu": c["dual_mu"].copy(), f"p{p}_mu_f": c["dual_mu_f"].copy(), f"p{p}_beta_in": c["beta"].copy(),
                        f"p{p}_beta_f_in": c["beta_f"].copy()})
            rec.update({f"p{p}_{k}": v for k, v in out.items()})
            # the oracle's own pass moves its state on (and must agree: asserted in the tests, not here)
            f.evaluate_dual_eta(); f.backward_solve(); f.update_tightening()
        rec["q3_inputs"] = np.array(pvecs + [pvecs[-1] + 5e-4, pvecs[-1] + 2e-3])
        np.savez(os.path.join(OUT, f"refglue_{name}.npz"), A=np.array(f.A), B=np.array(f.B), E=np.array(f.E),
                 g=np.array(f.g_list[:-1]), g_N=np.array(f.g_list[-1]), c=np.array(f.c_list), x0=x0, X=X, U=U,
                 source="reference fast_SLS.evaluate_dual_eta / backward_solve / update_tightening / check_convergence_socp "
                        "executed through oracle/ref_glue.py", **rec)
        print(name, "written; pass-1 max backoff", rec["p0_backoff"].max(), "pass-2", rec["p1_backoff"].max(), "q3", rec["p1_q3"])


if __name__ == "__main__":
    main()
