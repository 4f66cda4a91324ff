import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = 4096
x0, W = rocket_scenarios(B, steps=14)
m, cfg, s = make_solver("rocket", B)
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
for i in range(13):
    if i > 0: s.reset_warm_start(B)
    s.solve(x); s.plant_step(x, W[:, i])
s.reset_warm_start(B)
X, U = s.buffer("X"), s.buffer("U")
A, Bm, c = s.linearize(X, U)
xh = x.cpu().numpy()
ref = None
for tol, mi in ((1e-14, 60), (1e-13, 60), (1e-12, 60), (1e-11, 60), (1e-10, 60)):
    _, _, s2 = make_solver("rocket", B, qp_tol_mu=tol, qp_max_iter=mi)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = s2.qp(A, Bm, c, X, U, xh)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    o = {k: v.cpu().numpy() for k, v in out.items()}
    it = o["iters"]; ok = o["status"] == 0
    if ref is None: ref = o
    okb = ok & (ref["status"] == 0)
    dmu = np.abs(o["mu"][okb] - ref["mu"][okb]).max(axis=(1, 2)); dx = np.abs(o["dX"][okb] - ref["dX"][okb]).max(axis=(1, 2))
    print("tol %.0e: %.1f ms iters mean %.1f p50 %d p90 %d p99 %d max %d | status %s | vs tol1e-14: dmu p50 %.1e p99 %.1e max %.1e, dX max %.1e" % (
        tol, dt * 1e3, it.mean(), np.median(it), np.percentile(it, 90), np.percentile(it, 99), it.max(), np.bincount(o["status"], minlength=4), np.median(dmu), np.percentile(dmu, 99), dmu.max(), dx.max()))
    print("   hist", np.bincount(it)[5:45])
    s2.close()
