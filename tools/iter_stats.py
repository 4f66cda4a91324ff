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
for a in sys.argv[2:]:
    k, v = a.split('='); s.set_option(k, float(v))
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
tot = 0.0
for i in range(steps):
    if i > 0: s.reset_warm_start(B)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.solve(x); s.plant_step(x, W[:, i])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = s.buffer("qp_iters"); fl = s.buffer("flags"); st = s.buffer("status")
    tm = s.timings(); tot += dt
    print("step %2d %.1f ms qp_last %.1f ms | iters/solve mean %.1f p99 %d max %d | fails %d status %s" % (
        i, dt * 1e3, tm["t_qp_ms"], it.mean(), np.percentile(it, 99), it.max(), int((fl & 1).sum()), np.bincount(st, minlength=4)))
print("TOTAL %.1f ms for %d steps -> %.0f solves/s" % (tot * 1e3, steps, B * steps / tot))
