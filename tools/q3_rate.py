"""How many instances of the bench workload take the Q3 early return (fast_SLS_jit.py:283-296) at each MPC step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = 29
x0, W = rocket_scenarios(B, seed=2025, steps=steps + 1)
m, cfg, s = make_solver("rocket", B)
s.set_option("sticky_fail", 1)
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
for i in range(steps):
    if i > 0: s.reset_warm_start(B)
    s.solve(x); s.plant_step(x, W[:, i])
    f = s.buffer("flags")
    print("step %2d: Q3 %5d  tight %5d  failed %4d  iters/solve %.1f" % (i, int(((f & 2) != 0).sum()), int(((f & 4) != 0).sum()), int((f & 1).sum()), s.buffer("qp_iters").mean()))
