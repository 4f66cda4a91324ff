import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = 4096
x0, W = rocket_scenarios(B, steps=16)
m, cfg, s = make_solver("rocket", B)
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
for i in range(15):
    if i > 0: s.reset_warm_start(B)
    s.solve(x); s.plant_step(x, W[:, i])
    if i in (1, 5, 10, 14):
        it = s.buffer("qp_iters"); fl = s.buffer("flags")
        ok = (fl & 1) == 0
        print("step", i, "hist(ok)", np.bincount(it[ok], minlength=60)[10:60].tolist(), "failed iters", sorted(it[~ok].tolist())[-8:])
