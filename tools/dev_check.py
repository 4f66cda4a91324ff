"""Developer smoke check (GPU): CUDA kernels vs the CPU oracle on a handful of instances."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om, sls as osls, scp as oscp, qp as oqp
from rnmpc_b200.engine import BatchSolver, fp64_peak_tflops

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

def make(name, cap, **kw):
    m, cfg = om.CONFIGS[name]()
    s = BatchSolver(name, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E,
                    m.x_ub, m.x_lb, m.u_ub, m.u_lb, rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=cap, **kw)
    return m, cfg, s

print("fp64 peak TFLOP/s", fp64_peak_tflops())
for name in ["pendulum", "quadrotor", "rocket"]:
    m, cfg, s = make(name, 8)
    N, nx, nu = cfg["N"], m.nx, m.nu
    rng = np.random.default_rng(0)
    B = 5
    X = rng.uniform(m.x_lb / 4, m.x_ub / 4, size=(B, N + 1, nx))
    mid = 0.5 * (m.u_lb + m.u_ub)
    U = rng.uniform(mid - 0.2, mid + 0.2, size=(B, N, nu))
    A, Bm, c = s.linearize(X, U)
    A, Bm, c = A.cpu().numpy(), Bm.cpu().numpy(), c.cpu().numpy()
    eA = eB = ec = 0
    for b in range(B):
        for k in range(N):
            a_, b_ = m.jacobians(X[b, k], U[b, k])
            eA = max(eA, np.abs(a_ - A[b, k]).max()); eB = max(eB, np.abs(b_ - Bm[b, k]).max())
            ec = max(ec, np.abs(m.ddyn(X[b, k], U[b, k]) - X[b, k + 1] - c[b, k]).max())
    print(name, "linearize abs err A %.2e B %.2e c %.2e" % (eA, eB, ec))
    # SLS
    nz = nx + nu; ni = 2 * nz; nif = 2 * nx
    As = np.eye(nx) + 0.05 * rng.standard_normal((B, N, nx, nx)); Bs = 0.05 * rng.standard_normal((B, N, nx, nu))
    eta = 5e4 * np.abs(rng.standard_normal((B, N, N, ni))) * (rng.random((B, N, N, ni)) < 0.1)
    eta_f = 5e4 * np.abs(rng.standard_normal((B, N + 1, nif))) * (rng.random((B, N + 1, nif)) < 0.1)
    out = s.sls(As, Bs, eta, eta_f, want_S=True, want_phi=True)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for b in range(B):
        E = np.stack([m.E] * (N + 1))
        S, K = osls.backward_solve(N, nx, nu, As[b], Bs[b], m.G, m.Gf, eta[b], eta_f[b], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"])
        Px, Pu = osls.propagate(As[b], Bs[b], E, K)
        be, bef, bo, bof = osls.backoff_from_phi(Px, Pu, m.G[:, :nx], m.G[:, nx:], m.Gf, 1e-10)
        if b == 0:
            print(name, "sls rel err S %.2e K %.2e Px %.2e Pu %.2e beta %.2e beta_f %.2e bo %.2e bof %.2e" % (
                rel(out["S"][b], S), rel(out["K"][b], K), rel(out["Phi_x"][b], Px), rel(out["Phi_u"][b], Pu),
                rel(out["beta"][b], be), rel(out["beta_f"][b], bef), rel(out["backoff"][b], bo), rel(out["backoff_f"][b], bof)))
    s.close()

# QP + closed loop on the rocket
m, cfg, s = make("rocket", 4)
N, nx, nu = cfg["N"], m.nx, m.nu
x0 = om.ROCKET_X0
t = time.time(); Xo, Uo, ok = oscp.nominal_sqp(m, N, cfg["Q"], cfg["R"], cfg["Qf"], x0, iters=10); print("oracle sqp", time.time() - t, ok)
s.init_nominal(np.stack([x0] * 2), iters=10)
Xg, Ug = s.buffer("X", 2), s.buffer("U", 2)
print("init_nominal rel err X %.2e U %.2e  status" % (rel(Xg[0], Xo.T), rel(Ug[0], Uo.T)), s.buffer("status", 2), s.buffer("qp_iters", 2))
# one QP against the oracle
scp = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=1)
scp.set_nominal(Xo, Uo); scp.update_jacobian()
assert scp.fast.forward_solve(Xo[:, 0] - x0)
A_, B_, c_ = s.linearize(Xo.T[None], Uo.T[None])
q = s.qp(A_, B_, c_, Xo.T[None], Uo.T[None], x0[None])
q = {k: v.cpu().numpy() for k, v in q.items()}
print("qp status", q["status"], "iters", q["iters"], "rel err dX %.2e dU %.2e mu %.2e (abs %.2e, max %.2e) mu_f %.2e cost %.2e" % (
    rel(q["dX"][0], scp.fast.cur["primal_x"].T), rel(q["dU"][0], scp.fast.cur["primal_u"].T), rel(q["mu"][0], scp.fast.cur["dual_mu"].T),
    np.abs(q["mu"][0] - scp.fast.cur["dual_mu"].T).max(), np.abs(scp.fast.cur["dual_mu"]).max(),
    np.abs(q["mu_f"][0] - scp.fast.cur["dual_mu_f"]).max(), rel(q["cost"][0], scp.fast.cur["cost_nominal"])))
# closed loop, 4 steps, W = -0.8
steps = 4
W = -0.8 * np.ones((steps, nx))
t = time.time(); ref = oscp.closed_loop(m, cfg, x0, W, steps, X0=Xo, U0=Uo); print("oracle closed loop", time.time() - t)
s.set_nominal(Xo.T[None], Uo.T[None])
x = s.dev(x0[None]).clone()
for i in range(steps):
    if i > 0: s.reset_warm_start(1)
    s.solve(x)
    Xn, Un, K, bo, fl = s.buffer("X", 1)[0], s.buffer("U", 1)[0], s.buffer("K", 1)[0], s.buffer("backoff", 1)[0], s.buffer("flags", 1)[0]
    print("step", i, "flags", fl, "qp_iters", s.buffer("qp_iters", 1)[0], "rel err Xn %.2e Un %.2e K %.2e backoff_x %.2e backoff_u %.2e" % (
        rel(Xn, ref["Xn"][i].T), rel(Un, ref["Un"][i].T), rel(K, ref["K"][i]), rel(bo[:, :nx], ref["backoff_x"][i][:N]), rel(bo[:, nx:nx + nu], ref["backoff_u"][i])))
    s.plant_step(x, W[i][None])
    print("     plant rel err %.2e" % rel(x.cpu().numpy()[0], ref["X"][i + 1]))
B = 4096
m, cfg, s2 = make("rocket", B)
xs = np.stack([x0] * B)
torch.cuda.synchronize(); t = time.time(); s2.init_nominal(xs, 10); torch.cuda.synchronize(); print("init_nominal B=%d: %.3f s" % (B, time.time() - t))
xd = s2.dev(xs)
for r in range(3):
    torch.cuda.synchronize(); t = time.time(); s2.solve(xd); torch.cuda.synchronize(); dt = time.time() - t
    print("solve B=%d: %.4f s -> %.0f solves/s; fails %d" % (B, dt, B / dt, int((s2.buffer("flags") & 1).sum())))
    s2.reset_warm_start(B)
