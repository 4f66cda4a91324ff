"""Are the scenarios on which two implementations disagree the ones whose Q3 test (|QP solution - previous| <= 1e-3, fast_SLS_jit.py
check_convergence_socp) is decided within rounding of its threshold?   python tools/q3_sensitivity.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.helpers import make_solver
name, B, seed = "quadrotor", 2048, 5
m, cfg = om.CONFIGS[name]()
x0 = np.stack([om.quadrotor_x0(seed + b) for b in range(B)])
def run(opts, nominal=None):
    _, _, s = make_solver(name, B)
    for k, v in opts.items(): s.set_option(k, v)
    if nominal is None: s.init_nominal(x0, 10)
    else: s.set_nominal(*nominal)
    nom = (s.buffer("X").copy(), s.buffer("U").copy())
    s.set_option("q3_margin_reset", 1)
    s.solve(s.dev(x0))
    r = (s.buffer("U")[:, 0].copy(), s.buffer("q3_margin").copy(), nom); s.close(); return r
slot = {"qp_handoff_iter": 0, "qp_block_max": 0}; block = {"qp_block_max": 100000}
for eps in (1e-3, 0.0):
    a = run(dict(slot, q3_eps=eps)); b = run(dict(block, q3_eps=eps), nominal=a[2])
    du = np.abs(a[0] - b[0]).max(axis=1)
    bad = np.where(du > 1e-7)[0]
    print("q3_eps = %g: slot vs block kernel from the same nominal, scenarios whose first input differs by > 1e-7: %s %s" % (eps, bad, du[bad]))
    if eps > 0:
        mg = a[1]
        print("   distance of the Q3 test quantity from its threshold (min over the checks of the solve): those scenarios %s | population quantiles 0.1%% %.1e, 1%% %.1e, 50%% %.1e" % (
            mg[bad], np.quantile(mg, 0.001), np.quantile(mg, 0.01), np.quantile(mg, 0.5)))
        print("   scenarios closest to the threshold:", np.argsort(mg)[:5], mg[np.argsort(mg)[:5]])
