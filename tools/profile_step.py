"""One closed-loop MPC step of the bench workload (for ncu): python tools/profile_step.py [B] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x0, W = rocket_scenarios(B, steps=steps + 1)
m, cfg, s = make_solver("rocket", B)
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
s.enable_timing(True)
for a in sys.argv[3:]:
    k_, v_ = a.split("="); s.set_option(k_, float(v_))
import time
for i in range(steps):
    if i > 0:
        s.reset_warm_start(B)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.solve(x)
    s.plant_step(x, W[:, i])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("step %d: %.2f ms -> %.0f solves/s" % (i, dt * 1e3, B / dt), s.timings())
torch.cuda.synchronize()
print("fails", int((s.buffer("flags") & 1).sum()), "qp iters mean", s.buffer("qp_iters").mean())
