"""First QP of the first solve on the quadrotor population with both QP kernels: primal and dual agreement per scenario (looking for
the source of the rare 1e-5-level disagreements between implementations: non-unique multipliers at degenerate solutions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.helpers import make_solver
name, B, seed = "quadrotor", 2048, 5
m, cfg = om.CONFIGS[name]()
x0 = np.stack([om.quadrotor_x0(seed + b) for b in range(B)])
_, _, s0 = make_solver(name, B); s0.init_nominal(x0, 10)
X, U = s0.buffer("X").copy(), s0.buffer("U").copy(); s0.close()
np_ = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
res = {}
for label, opts in (("slot", {"qp_handoff_iter": 0, "qp_block_max": 0}), ("block", {"qp_block_max": 100000})):
    _, _, s = make_solver(name, B)
    for k, v in opts.items(): s.set_option(k, v)
    A, Bm, c = s.linearize(X, U)
    out = s.qp(A, Bm, c, X, U, x0)
    res[label] = {k: np_(v).copy() for k, v in out.items()}; s.close()
a, b = res["slot"], res["block"]
ok = (a["status"] == 0) & (b["status"] == 0)
d = np.abs(a["dU"] - b["dU"]).reshape(B, -1).max(axis=1); dx = np.abs(a["dX"] - b["dX"]).reshape(B, -1).max(axis=1)
dm = np.abs(a["mu"] - b["mu"]).reshape(B, -1).max(axis=1)
print("first QP of the first solve, slot vs block kernel, %d instances solved by both:" % ok.sum())
print("  max |dU diff| %.2e (scenario %d), quantiles 50%% %.1e 99%% %.1e 99.9%% %.1e" % (d[ok].max(), np.argmax(np.where(ok, d, 0)), *np.quantile(d[ok], [0.5, 0.99, 0.999])))
print("  max |dX diff| %.2e, max |mu diff| %.2e (mu scale %.1e)" % (dx[ok].max(), dm[ok].max(), np.abs(a["mu"]).max()))
for sc in (1937, 1805, 1591, 1669):
    print("  scenario %d: |dU diff| %.2e |dX diff| %.2e |mu diff| %.2e iterations %d / %d  max|mu| %.1e" % (sc, d[sc], dx[sc], dm[sc], a["iters"][sc], b["iters"][sc], np.abs(a["mu"][sc]).max()))
