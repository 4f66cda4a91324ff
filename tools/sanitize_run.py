"""Small runs of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py        (the sanitizer is closed on the GPU pool of this project: only the plain run
    of this script -- every path executes, no failures -- has been done there)
block QP kernel (batches of 4), instance-per-slot kernel with many hand-offs (batch 1100, hand-off after 5 iterations), warm path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.helpers import make_solver, rocket_scenarios
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def loop(name, B, steps, opts):
    m, cfg = om.CONFIGS[name]()
    if name == "rocket": x0, W = rocket_scenarios(B, seed=3, steps=steps)
    else:
        x0 = np.repeat(om.PENDULUM_X0[None], B, 0) if name == "pendulum" else np.stack([om.quadrotor_x0(b) for b in range(B)])
        W = np.random.default_rng(3).uniform(-1, 1, size=(B, steps, m.nx))
    _, _, s = make_solver(name, B)
    for k, v in opts.items(): s.set_option(k, v)
    s.init_nominal(x0, 3)
    x = s.dev(x0).clone()
    for i in range(steps):
        if i > 0: s.reset_warm_start(B)
        s.solve(x); s.plant_step(x, W[:, i])
    torch.cuda.synchronize()
    print(name, B, opts, "failed", int((s.buffer("flags") & 1).sum()), "iters mean %.1f" % s.buffer("qp_iters").mean(), flush=True)
    s.close()
if which in ("all", "block"):
    loop("rocket", 4, 2, {}); loop("quadrotor", 4, 1, {}); loop("pendulum", 4, 2, {})
if which in ("all", "handoff"):
    loop("rocket", 1100, 1, {"qp_handoff_iter": 5})
if which in ("all", "slot"):
    loop("rocket", 6, 1, {"qp_block_max": 0, "qp_handoff_iter": 0}); loop("rocket", 6, 2, {"qp_block_max": 0, "warm_mask": 3})
