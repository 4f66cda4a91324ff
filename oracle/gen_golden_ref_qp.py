"""Writes tests/golden/refqp_{pendulum,quadrotor,rocket}.npz: what the reference's own ``QP`` class (oracle/ref_qp.py) hands to
OSQP -- P, A, q, l, u of the nominal and of the tightened QP of one fast-SLS pass -- and the dictionaries its ``solve`` returns
when the recorded problem is solved by the restated OSQP (eps 1e-6, rocket 1e-5, so that the polish identifies the active set; the settings the
reference passes are recorded beside).  Call order as in fast_SLS.update_dynamics_list / _step (fast_SLS_jit.py:250-327).

    python -m oracle.gen_golden_ref_qp
"""
import os

import numpy as np

from . import models as om
from . import ref_qp
from .gen_golden_ref_glue import scenario

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


EPS = dict(pendulum=1e-6, quadrotor=1e-6, rocket=1e-5)     # rocket: ADMM needs > 20000 iterations for 1e-6 (profiles/r3f_osqp_restated_gap.txt)


def main():
    for name in ("pendulum", "quadrotor", "rocket"):
        m, cfg, s, x0, X, U = scenario(name)
        N = cfg["N"]
        # the measured state sits off the nominal trajectory (2 % of the box, seeded), so the QPs have a non-trivial solution
        x0 = x0 + 0.02 * (m.x_ub - m.x_lb) / 2 * np.random.default_rng(5).uniform(-1, 1, m.nx)
        s.set_nominal(X, U)
        s.update_jacobian()
        f = s.fast
        rec = ref_qp._Recorder()
        qp = ref_qp.reference_qp(N, m, cfg["Q"], cfg["R"], cfg["Qf"], rec, dict(eps_abs=EPS[name], eps_rel=EPS[name], max_iter=20000))
        g_list = [np.array(g, float).reshape(-1, 1) for g in f.g_list]
        c_list = [np.array(c, float).reshape(-1, 1) for c in f.c_list]
        qp.update_dynamics([np.array(a) for a in f.A], [np.array(b) for b in f.B], [np.array(e) for e in f.E], g_list)
        qp.offset_constraints(np.hstack(c_list))
        qp.update_q_cost_lin(np.array(f.q, float).reshape(-1, 1))
        dx0 = X[:, 0] - x0
        sol0 = qp.solve(dx0.reshape(-1, 1))
        assert sol0["success"], name
        # the oracle's pass produces the tightened bounds (its agreement with the reference's update_tightening: refglue_*.npz)
        assert f.forward_solve(dx0)
        f.evaluate_dual_eta(); f.backward_solve(); f.update_tightening()
        qp.update_ubg(np.array(f.ubg, float))
        sol1 = qp.solve(dx0.reshape(-1, 1))
        assert sol1["success"], name
        out = dict(X=X, U=U, x0=x0, backoff=f.cur["backoff"].copy(), backoff_f=f.cur["backoff_f"].copy(), setup_settings=str(rec.setups[-1]["settings"]), n_setups=len(rec.setups))
        for p, (sv, sol) in enumerate(zip(rec.solves, (sol0, sol1))):
            out.update({f"p{p}_P": sv["P"], f"p{p}_A": sv["A"], f"p{p}_q": sv["q"], f"p{p}_l": sv["l"], f"p{p}_u": sv["u"],
                        f"p{p}_polished": sv["polished"], f"p{p}_status": sv["status"]})
            out.update({f"p{p}_{k}": np.asarray(sol[k], float) for k in ("primal_vec", "primal_x", "primal_u", "dual_vec", "dual_mu", "dual_mu_f", "cost")})
        np.savez_compressed(os.path.join(OUT, f"refqp_{name}.npz"), source="reference solver/qp_jit.py QP class driven through oracle/ref_qp.py; "
                            f"solver = oracle/osqp_restated.py at eps {EPS[name]:g}", **out)
        print(name, "setups", len(rec.setups), rec.setups[-1]["settings"], "|", [(sv["status"], sv["polished"]) for sv in rec.solves],
              "A", rec.solves[0]["A"].shape)


if __name__ == "__main__":
    main()
