"""Per-phase cycle counts of qp_mma_kernel (option qp_profile: logical thread 0 of instance 0, clock64 between the phase marks):
    python tools/qp_phase_profile.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from tests.helpers import make_solver, rocket_scenarios
from rnmpc_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x0, W = rocket_scenarios(B, steps=8)
m, cfg, s = make_solver("rocket", B)
for a_ in sys.argv[2:]:
    k_, v_ = a_.split("="); s.set_option(k_, float(v_))
s.init_nominal(x0, 10)
x = s.dev(x0).clone()
names = ["init", "S_N + first fragments", "A: W = S Z", "B: M = Z'[W p]", "C: Cholesky, K, d", "D: S', A+BK, p", "corr: rhs, r'", "corr: backward chain",
         "corr: t, d", "w = B d + rp", "forward chain", "inputs + ratio test + reduce", "update + reduce", "outputs"]
for i in range(4):
    if i > 0: s.reset_warm_start(B)
    if i == 3: s.set_option("qp_profile", 1)
    s.solve(x); s.plant_step(x, W[:, i])
torch.cuda.synchronize()
v = np.zeros(16, dtype=np.int64)
_lib.check(s.lib.rnmpc_copy_buffer(s._h, b"qp_prof", C.c_void_p(v.ctypes.data), 128))
it = int(s.buffer("qp_iters", 1)[0])
print("batch", B, ": interior-point iterations of instance 0 in this solve (2 QPs):", it, " total cycles", int(v.sum()))
for n, c in zip(names, v):
    print("  %-32s %9d cycles  %5.1f %%   %8.0f per iteration" % (n, c, 100.0 * c / max(v.sum(), 1), c / max(it, 1)))
