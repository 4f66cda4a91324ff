"""Complementarity tolerance of the interior point (option qp_tol_mu, default 1e-14) against the 29-step rocket loop at the
bench batch: interior-point iterations, time per MPC step, failure set and the distance of the closed-loop states,
nominal trajectories, gains and back-offs from the run with the default tolerance."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STEPS = 29
x0, W = rocket_scenarios(B, steps=STEPS)
ref = None
for tol in (1e-14, 1e-13, 1e-12, 1e-11, 1e-10):
    m, cfg, s = make_solver("rocket", B, qp_tol_mu=tol)
    s.set_option("sticky_fail", 1)
    s.init_nominal(x0, 10)
    x = s.dev(x0).clone()
    Wd = s.dev(W)
    states, iters = [], 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(STEPS):
        if i > 0: s.reset_warm_start(B)
        s.solve(x); s.plant_step(x, Wd[:, i].contiguous())
        states.append(x.clone())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    fl = s.buffer("flags"); dead = (fl & 1) != 0
    out = dict(states=torch.stack(states, 1).cpu().numpy(), X=s.buffer("X"), U=s.buffer("U"), K=s.buffer("K"),
               backoff=s.buffer("backoff"), dead=dead, it=s.buffer("qp_iters"))
    if ref is None: ref = out
    ok = ~(dead | ref["dead"])
    def d(k):
        a, b = out[k][ok], ref[k][ok]
        return np.abs(a - b).max() / max(1.0, np.abs(b).max())
    print("tol %.0e: %.2f ms/step, qp_iters(last step, both QPs) mean %.2f | dead %d (same set as 1e-14: %s) | rel diff vs 1e-14: states %.1e X %.1e U %.1e K %.1e backoff %.1e" % (
        tol, dt / STEPS * 1e3, out["it"][~dead].mean(), dead.sum(), bool((dead == ref["dead"]).all()), d("states"), d("X"), d("U"), d("K"), d("backoff")), flush=True)
