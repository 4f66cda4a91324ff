"""Soft-constrained comparator on the GPU: per-step SQP iterations / status for the rocket compare scenario."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rnmpc_b200.solver.nlp_soft_constraints import NLPSoftConstraints
from rnmpc_b200.workloads import ROCKET_X0, rocket_config
m, cfg = rocket_config()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "soft_rocket.npz"))
for mode in ("golden guess", "interpolation"):
    s = NLPSoftConstraints(cfg["N"], cfg["Q"], cfg["R"], m, cfg["Qf"], rho_soft=1e6, rho_soft_l1=1e6, max_iter=int(os.environ.get("MAXIT", "80")))
    x = ROCKET_X0.copy()
    for i in range(steps):
        sol = s.solve(x, x_guess=g["X0"].T, u_guess=g["U0"].T) if (i == 0 and mode == "golden guess") else s.solve(x)
        st = s._engine.buffer("status_tmp", 1) if False else None
        print(mode, "step", i, "success", sol["success"], "sqp iters", sol.get("sqp_iterations"), "status", sol.get("status"),
              "cost", sol.get("cost"), "gmax", None if "primal_gamma" not in sol else sol["primal_gamma"].max(),
              "err vs golden", None if (i >= g["Xn"].shape[0] or "primal_x" not in sol) else np.abs(sol["primal_x"] - g["Xn"][i].T).max())
        if not sol["success"]:
            break
        x = np.array(m.ddyn(x, sol["primal_u"][:, 0])).reshape(-1) + m.E @ (-0.8 * np.ones(17))
