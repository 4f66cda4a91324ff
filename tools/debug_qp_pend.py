import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import make_solver, load
from tests.test_gpu_parity import set_qp_kernel, np_
np.set_printoptions(linewidth=200, precision=4)
for name in ("pendulum", "rocket"):
    g = load(f"qp_{name}.npz")
    for kern in ("tmem", "warp"):
        m, cfg, s = make_solver(name, 8)
        set_qp_kernel(s, kern)
        A, B, c = s.linearize(g["X"], g["U"])
        idx = np.where(g["tight"] == False)[0]
        ti = torch.as_tensor(idx, device=A.device)
        out = s.qp(A[ti], B[ti], c[ti], g["X"][idx], g["U"][idx], g["x0"][idx], None, None)
        out = {k: np_(v) for k, v in out.items()}
        print(name, kern, "status", out["status"], "iters", out["iters"])
        print(" mu[0][:3]", out["mu"][0][:3])
        print(" ref     ", g["mu"][idx[0]][:3])
        print(" dX err", np.abs(out["dX"][0] - g["dX"][idx[0]]).max(), "mu_f", out["mu_f"][0])
        s.close()
import ctypes as C
from rnmpc_b200 import _lib
name = "pendulum"
g = load(f"qp_{name}.npz")
m, cfg, s = make_solver(name, 8)
set_qp_kernel(s, "tmem"); s.set_option("qp_profile", 1)
A, B, c = s.linearize(g["X"], g["U"])
idx = np.where(g["tight"] == False)[0]
ti = torch.as_tensor(idx, device=A.device)
out = s.qp(A[ti], B[ti], c[ti], g["X"][idx], g["U"][idx], g["x0"][idx], None, None)
torch.cuda.synchronize()
v = np.zeros(16, dtype=np.int64)
_lib.check(s.lib.rnmpc_copy_buffer(s._h, b"qp_prof", C.c_void_p(v.ctypes.data), 128))
print("DBG", v.view(np.float64), "nzmask", v[7])
