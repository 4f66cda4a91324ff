"""Kernel times of the solve at small batch sizes (per-warp latency of the sequential chains)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
opts = [a.split("=") for a in sys.argv[1:] if "=" in a]
for B in [int(a) for a in sys.argv[1:] if "=" not in a] or [1, 9, 456, 4096]:
    x0, W = rocket_scenarios(B, steps=8)
    m, cfg, s = make_solver("rocket", B)
    s.enable_timing(True)
    for k_, v_ in opts: s.set_option(k_, float(v_))
    s.init_nominal(x0, 10)
    x = s.dev(x0).clone()
    for i in range(6):
        if i > 0: s.reset_warm_start(B)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.solve(x)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        s.plant_step(x, W[:, i])
        it = s.buffer("qp_iters"); tm = s.timings()
        if i >= 3:
            print("B %5d step %d: solve %.3f ms | %s | iters/solve mean %.1f max %d" % (B, i, dt * 1e3, {k: round(v, 3) for k, v in tm.items()}, it.mean(), it.max()))
