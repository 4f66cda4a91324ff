"""GPU path vs the C++ CPU port (oracle) on the SAME disturbed rocket closed loops: which scenarios fail, at which step, and
how far the states of the surviving ones are apart.   python tools/failure_parity.py [B] [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om, scp as oscp, cpu_port
from tests.helpers import make_solver, rocket_scenarios

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2025
name = next((a for a in sys.argv[3:] if "=" not in a), "rocket")
opts = [a.split("=") for a in sys.argv[3:] if "=" in a]          # key=value solver options (rnmpc_set_option)
steps = {"rocket": 29, "pendulum": 60, "quadrotor": 30}[name]     # lengths of the reference drivers' loops
m, cfg = om.CONFIGS[name]()
_, _, s = make_solver(name, B)
s.set_option("sticky_fail", 1)
for k_, v_ in opts: s.set_option(k_, float(v_))
rng = np.random.default_rng(seed)
if name == "rocket":
    x0, W = rocket_scenarios(B, seed=seed, steps=steps)
    X0, U0, ok = oscp.nominal_sqp(m, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], x0[0], iters=10)
    X0b, U0b = np.repeat(X0.T[None], B, 0), np.repeat(U0.T[None], B, 0)
    s.set_nominal(X0b, U0b)
else:
    # pendulum: the driver's x0, disturbance draws U[-1,1]^nx through E; quadrotor: seeded x0 draws of the driver + disturbances.
    # The nominal initialiser is not the subject: the GPU's is used and handed to the CPU port.
    x0 = np.repeat(om.PENDULUM_X0[None], B, 0) if name == "pendulum" else np.stack([om.quadrotor_x0(seed + b) for b in range(B)])
    W = rng.uniform(-1, 1, size=(B, steps, m.nx))
    s.init_nominal(x0, 10)
    X0b, U0b = s.buffer("X", B).copy(), s.buffer("U", B).copy()
st_c, in_c, nf_c, ok_c = cpu_port.closed_loop(m, cfg, x0, W, steps, X0b, U0b, 0, return_ok=True)
ff_c = np.where((ok_c == 0).any(1), (ok_c == 0).argmax(1), steps)
x = s.dev(x0).clone()
st_g = np.empty((B, steps + 1, m.nx)); ff_g = np.full(B, steps); its_g = np.zeros((B, steps), dtype=int)
for i in range(steps):
    st_g[:, i] = torch.as_tensor(x).cpu().numpy()
    if i > 0: s.reset_warm_start(B)
    s.solve(x)
    fl = s.buffer("flags") & 1
    its_g[:, i] = s.buffer("qp_iters")
    ff_g = np.where((fl == 1) & (ff_g == steps), i, ff_g)
    s.plant_step(x, W[:, i])
st_g[:, steps] = torch.as_tensor(x).cpu().numpy()

same = ff_c == ff_g
print("scenarios %d | failing on CPU port %d, on GPU %d | same first-failure step (or both never): %d | differ: %d" % (
    B, (ff_c < steps).sum(), (ff_g < steps).sum(), same.sum(), (~same).sum()))
for b in np.where(~same)[0][:20]:
    print("  scenario %d: CPU port fails at step %s, GPU at step %s" % (b, ff_c[b] if ff_c[b] < steps else "-", ff_g[b] if ff_g[b] < steps else "-"))
# states up to (and including) the step of the first failure of either side are produced by successful solves on both sides
scale = np.maximum(1.0, np.abs(st_c).max(axis=(0, 1)))
err = np.zeros(B)
for b in range(B):
    k = min(ff_c[b], ff_g[b])            # states 0..k come from k successful solves
    err[b] = (np.abs(st_c[b, :k + 1] - st_g[b, :k + 1]) / scale).max()
print("closed-loop states before any failure: max scaled error %.3e (tolerance 1e-5), 99.9 %% quantile %.3e" % (err.max(), np.quantile(err, 0.999)))
print("worst scenarios:", np.argsort(-err)[:5], err[np.argsort(-err)[:5]])
print("per-step scaled error of the worst scenarios (first failure step in brackets; '-' = never):")
for b in np.argsort(-err)[:6]:
    k = min(ff_c[b], ff_g[b])
    prof = (np.abs(st_c[b, :k + 1] - st_g[b, :k + 1]) / scale).max(axis=1)
    print("  %5d [%s]: %s" % (b, k if k < steps else "-", " ".join("%.0e" % v for v in prof)))
    print("        GPU interior-point iterations per solve: %s" % " ".join("%d" % v for v in its_g[b]))
surv = (ff_c == steps) & (ff_g == steps)
print("scenarios that never fail: %d, max scaled error among them %.3e; among those that fail later %.3e" % (surv.sum(), err[surv].max(), err[~surv].max() if (~surv).any() else 0.0))
