import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.helpers import make_solver, rocket_scenarios
B, steps = 1024, 29
m, cfg, s = make_solver("rocket", B)
for sc in (0.05, 0.02, 0.01, 0.005, 0.002, 0.0):
    x0, W = rocket_scenarios(B, steps=steps, x0_scale=sc)
    for its in (10, 20):
        s.init_nominal(x0, its)
        print("x0_scale", sc, "sqp iters", its, "init fails", int((s.buffer("flags") & 1).sum()), "status", np.bincount(s.buffer("status"), minlength=4),
              "max|dX| of last sqp step %.2e" % np.abs(s.buffer("dX")[(s.buffer("flags") & 1) == 0]).max())
x0, W = rocket_scenarios(B, steps=steps, x0_scale=0.0)
nx, nu = m.nx, m.nu
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
failed = np.zeros(B, bool)
for i in range(steps):
    if i > 0: s.reset_warm_start(B)
    xh = x.cpu().numpy()
    s.solve(x)
    fl = s.buffer("flags"); st = s.buffer("status"); bo = s.buffer("backoff")
    newf = ((fl & 1) != 0) & ~failed
    # stage-0 violation of the original box and of the previous tube
    v_box = np.maximum(xh - m.x_ub, m.x_lb - xh).max(axis=1)
    v_t = np.maximum(xh + bo[:, 0, :nx] - m.x_ub, m.x_lb - (xh - bo[:, 0, :nx])).max(axis=1)
    print("step", i, "new fails", int(newf.sum()), "status hist", np.bincount(st[newf], minlength=4), "viol box max %.2e, tube0 median(newfail) %.2e max %.2e; n(v_t>0)=%d n(v_t>1e-3)=%d" % (
        v_box.max(), np.median(v_t[newf]) if newf.any() else 0, v_t[newf].max() if newf.any() else 0, int((v_t[~failed] > 0).sum()), int((v_t[~failed] > 1e-3).sum())),
        "iters mean %.1f" % s.buffer("qp_iters")[~failed].mean())
    failed |= newf
    s.plant_step(x, W[:, i])
print("total failed", failed.sum())
