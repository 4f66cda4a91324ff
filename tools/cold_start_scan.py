"""Interior-point iterations per solve for the cold-start variants (option qp_mu0) on the three model families."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.helpers import make_solver, rocket_scenarios
for name, B, steps in (("rocket", 2048, 29), ("quadrotor", 1024, 30), ("pendulum", 1024, 60)):
    m, cfg = om.CONFIGS[name]()
    rng = np.random.default_rng(5)
    if name == "rocket": x0, W = rocket_scenarios(B, seed=5, steps=steps)
    else:
        x0 = np.repeat(om.PENDULUM_X0[None], B, 0) if name == "pendulum" else np.stack([om.quadrotor_x0(5 + b) for b in range(B)])
        W = rng.uniform(-1, 1, size=(B, steps, m.nx))
    for mu0 in [float(a) for a in sys.argv[1:]] or [0, 1, 2, 10, -1, -0.3]:
        _, _, s = make_solver(name, B)
        s.set_option("sticky_fail", 1); s.set_option("qp_mu0", mu0)
        s.init_nominal(x0, 10)
        x = s.dev(x0).clone(); its, mx = [], []
        torch.cuda.synchronize(); import time; t0 = time.perf_counter()
        for i in range(steps):
            if i > 0: s.reset_warm_start(B)
            s.solve(x)
            ok = (s.buffer("flags") & 1) == 0
            q = s.buffer("qp_iters")[ok]; its.append(q.mean()); mx.append(q.max())
            s.plant_step(x, W[:, i])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%-9s qp_mu0 %5g: iterations per solve mean %.1f, mean of per-step max %.1f, max %d | failed scenarios %d | loop %.1f ms" % (
            name, mu0, np.mean(its), np.mean(mx), np.max(mx), int((s.buffer("flags") & 1).sum()), dt * 1e3))
        s.close()
