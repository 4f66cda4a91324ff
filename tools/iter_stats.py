"""QP iteration / failure statistics along the bench closed loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = 29
x0, W = rocket_scenarios(B, steps=steps)
m, cfg, s = make_solver("rocket", B)
s.enable_timing(True)
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
for i in range(steps):
    if i > 0: s.reset_warm_start(B)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.solve(x); s.plant_step(x, W[:, i])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = s.buffer("qp_iters"); fl = s.buffer("flags"); st = s.buffer("status")
    tm = s.timings()
    print("step %2d %.1f ms qp_last %.1f ms | iters/solve mean %.1f p99 %d max %d | fails %d status %s" % (
        i, dt * 1e3, tm["t_qp_ms"], it.mean(), np.percentile(it, 99), it.max(), int((fl & 1).sum()), np.bincount(st, minlength=4)))
