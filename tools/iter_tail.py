"""Interior-point iteration counts per QP launch over the bench loop, split by outcome: how long do successful solves take at
most, and how long do the launches wait for solves that end up failing?   python tools/iter_tail.py [B] [key=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = 29
x0, W = rocket_scenarios(B, steps=steps)
m, cfg, s = make_solver("rocket", B)
s.set_option("sticky_fail", 1)
for a in sys.argv[2:]:
    k, v = a.split("="); s.set_option(k, float(v))
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
dead = np.zeros(B, bool)
ok1, ok2, bad = [], [], []
for i in range(steps):
    if i > 0: s.reset_warm_start(B)
    s.solve(x)
    fl = (s.buffer("flags") & 1) == 1
    tot, last, st = s.buffer("qp_iters"), s.buffer("iters_tmp"), s.buffer("status")
    new_fail = fl & ~dead
    good = ~fl
    ok2.append(last[good]); ok1.append((tot - last)[good])
    for b in np.where(new_fail)[0]:
        bad.append((i, int(tot[b]), int(last[b]), int(st[b])))
    dead |= fl
    s.plant_step(x, W[:, i])
ok1, ok2 = np.concatenate(ok1), np.concatenate(ok2)
q = lambda a: "mean %.1f  p50 %d  p99 %d  p99.9 %d  max %d" % (a.mean(), *np.percentile(a, [50, 99, 99.9]).astype(int), a.max())
print("successful solves: first QP (cold):       ", q(ok1))
print("successful solves: tightened QP (warm):   ", q(ok2))
print("histogram of the tightened QP's iterations (>= 16):", {int(k): int((ok2 == k).sum()) for k in np.unique(ok2) if k >= 16})
print("histogram of the first QP's iterations (>= 16):    ", {int(k): int((ok1 == k).sum()) for k in np.unique(ok1) if k >= 16})
bad = np.array(bad)
print("failing solves: %d; iterations spent in the QP launch that failed: %s" % (len(bad), q(bad[:, 2]) if len(bad) else "-"))
print("   by status:", {int(k): int((bad[:, 3] == k).sum()) for k in np.unique(bad[:, 3])} if len(bad) else "-",
      "| failed in the first QP (total == last):", int((bad[:, 1] == bad[:, 2]).sum()) if len(bad) else 0)
print("   iterations histogram of failing launches:", {int(k): int((bad[:, 2] == k).sum()) for k in np.unique(bad[:, 2])} if len(bad) else "-")
