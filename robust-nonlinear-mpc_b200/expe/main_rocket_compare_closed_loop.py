"""Robust arm of the reference's expe/main_rocket_compare_closed_loop.py (:41-173, 225-289) on the GPU path, batched over
disturbance sequences -- BASELINE.json configs[3]: the fixed initial state, N = 15, 29 MPC steps, scenario 0 with the driver's
W = -0.8 (:241), the others with default_rng(123).uniform(-1, 1) draws (the draw the driver has commented out, :240).

    python -m rnmpc_b200.expe.main_rocket_compare_closed_loop --run [--batch 4096] [--keep 1] [--out DIR]

The result file has the reference's `r_*` / `s_*` / common / `Jr_*` / `Js_*` keys with a leading scenario axis (`r_state_trajectory
[B, nx, T]`, ...); nominal and back-off trajectories are stored for the first `--keep` scenarios only (a full set is 63 KB per scenario
and step).  The soft-constrained NMPC arm (`_run_soft`, :176-222; solver/nlp_soft_constraints.py) runs for the first `--soft` scenarios
(default 1: the driver's W = -0.8 sequence) on `solver.nlp_soft_constraints.NLPSoftConstraints` of this package."""
import os
from datetime import datetime

import numpy as np

from ..engine import BatchSolver
from ..workloads import ROCKET_X0, rocket_config

folder_name = "rockETH_compare_closed_loop"


def closed_loop_cost(X_all, U_all, Q, R, Qf):
    """_compute_closed_loop_cost (:94-101), vectorised over scenarios: X_all [B, nx, T], U_all [B, nu, T-1]."""
    Js = np.einsum("bit,ij,bjt->b", X_all[:, :, :-1], Q, X_all[:, :, :-1]) + np.einsum("bit,ij,bjt->b", U_all, R, U_all)
    JT = np.einsum("bi,ij,bj->b", X_all[:, :, -1], Qf, X_all[:, :, -1])
    return Js, JT, Js + JT


def run_robust(x0, W, keep=1, device=0):
    """_run_robust (:104-173) for B scenarios at once: x0 [nx] or [B, nx], W [B, steps, nx]."""
    m, cfg = rocket_config()
    W = np.asarray(W, float)
    B, steps, nx = W.shape
    nu, N = m.nu, cfg["N"]
    x0 = np.broadcast_to(np.asarray(x0, float).reshape(-1, nx), (B, nx)).copy()
    x_ub, x_lb, u_ub, u_lb, _, _ = m.box()
    s = BatchSolver("rocket", N, cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E, x_ub, x_lb, u_ub, u_lb,
                    rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=B, device=device)
    s.set_option("sticky_fail", 1)            # the reference's loop would stop at a failed solve; the scenario is frozen instead
    keep = min(keep, B)
    X = np.zeros((B, nx, steps + 1)); U = np.zeros((B, nu, steps))
    Xn = np.zeros((keep, nx, N + 1, steps)); Un = np.zeros((keep, nu, N, steps))
    box = np.zeros((keep, nx, N + 1, steps)); bou = np.zeros((keep, nu, N, steps))
    ok = np.ones(B, dtype=bool)
    x = s.dev(x0).clone()
    s.init_nominal(x, 10)
    X[:, :, 0] = x0
    for i in range(steps):
        if i > 0:
            s.reset_warm_start(B)
        s.solve(x)
        ok &= (s.buffer("flags", B) & 1) == 0
        Ui = s.buffer("U", B)
        U[:, :, i] = Ui[:, 0]
        Xn[:, :, :, i] = np.swapaxes(s.buffer("X", keep), 1, 2)
        Un[:, :, :, i] = np.swapaxes(Ui[:keep], 1, 2)
        bo, bof = s.buffer("backoff", keep), s.buffer("backoff_f", keep)
        box[:, :, :N, i] = np.swapaxes(bo[:, :, :nx], 1, 2); box[:, :, N, i] = bof[:, :nx]
        bou[:, :, :, i] = np.swapaxes(bo[:, :, nx:nx + nu], 1, 2)
        s.plant_step(x, W[:, i])
        X[:, :, i + 1] = x.cpu().numpy()
    s.close()
    return dict(state_trajectory=X, input_trajectory=U, nominal_trajectory_x=Xn, nominal_trajectory_u=Un, backoff_trajectory_x=box,
                backoff_trajectory_u=bou, success=ok, g=m.g, nx=nx, nu=nu, dt=m.dt, N=N)


def run_soft(x0, W, device=0):
    """_run_soft (:176-222) for B scenarios at once: x0 [nx] or [B, nx], W [B, steps, nx].  rho_soft = rho_soft_l1 = 1e6 (:180).  A
    scenario whose NLP solve fails stops (the reference raises); its `success` entry is False."""
    from ..solver.nlp_soft_constraints import NLPSoftConstraints
    m, cfg = rocket_config()
    W = np.asarray(W, float)
    B, steps, nx = W.shape
    nu, N = m.nu, cfg["N"]
    x = np.broadcast_to(np.asarray(x0, float).reshape(-1, nx), (B, nx)).copy()
    solver = NLPSoftConstraints(N, cfg["Q"], cfg["R"], m, cfg["Qf"], rho_soft=1e6, rho_soft_l1=1e6, batch=B, device=device)
    X = np.zeros((B, nx, steps + 1)); U = np.zeros((B, nu, steps))
    Xn = np.zeros((B, nx, N + 1, steps)); Un = np.zeros((B, nu, N, steps))
    ok = np.ones(B, dtype=bool)
    done = np.zeros(B, dtype=int)              # MPC steps completed per scenario
    X[:, :, 0] = x
    eng = solver._engine
    for i in range(steps):
        sol = solver.solve(x if B > 1 else x[0])
        succ = np.atleast_1d(sol['success'])
        ok &= succ
        done += ok
        if not ok.any():
            break
        px = sol['primal_x'] if B > 1 else sol['primal_x'][None]
        pu = sol['primal_u'] if B > 1 else sol['primal_u'][None]
        Xn[:, :, :, i], Un[:, :, :, i] = px, pu
        U[:, :, i] = pu[:, :, 0]
        xn = eng.ddyn(x, pu[:, :, 0], W[:, i]).cpu().numpy()            # x+ = ddyn(x, u0) + E w  (:208)
        x = np.where(ok[:, None], xn, x)
        X[:, :, i + 1] = x
    eng.close()
    return dict(state_trajectory=X, input_trajectory=U, nominal_trajectory_x=Xn, nominal_trajectory_u=Un,
                backoff_trajectory_x=np.zeros((B, nx, N + 1, steps)), backoff_trajectory_u=np.zeros((B, nu, N, steps)), success=ok,
                steps_completed=done, g=m.g, nx=nx, nu=nu, dt=m.dt, N=N)


def generate(batch=4096, keep=1, out_dir=None, device=0, soft=1):
    m, cfg = rocket_config()
    T = 30
    W = np.random.default_rng(123).uniform(-1.0, 1.0, size=(batch, T - 1, m.nx))
    W[0] = -0.8                                # the sequence the reference driver runs
    res = run_robust(ROCKET_X0, W, keep, device)
    ok = res["success"]
    Js, JT, J = closed_loop_cost(res["state_trajectory"], res["input_trajectory"], cfg["Q"], cfg["R"], cfg["Qf"])
    soft = max(0, min(int(soft), batch))
    soft_keys = {}
    if soft > 0:
        sres = run_soft(ROCKET_X0, W[:soft], device)
        sJs, sJT, sJ = closed_loop_cost(sres["state_trajectory"], sres["input_trajectory"], cfg["Q"], cfg["R"], cfg["Qf"])
        soft_keys = {"s_" + k: sres[k] for k in ("state_trajectory", "input_trajectory", "nominal_trajectory_x", "nominal_trajectory_u",
                                                 "backoff_trajectory_x", "backoff_trajectory_u", "success", "steps_completed")}
        soft_keys.update(Js_stage=sJs, Js_terminal=sJT, Js_total=sJ)
        for b in np.where(~sres["success"])[0][:4]:
            print(f"scenario {b}: the soft-constrained arm stopped after {sres['steps_completed'][b]} of {T - 1} MPC steps (its NLP solve failed)")
        # the reference's cost table (:248-253, 403-406) for the scenarios both arms ran
        both = ok[:soft] & sres["success"]
        for b in np.where(both)[0][:4]:
            print(f"scenario {b}: robust  stage {Js[b]:.3f} terminal {JT[b]:.3f} total {J[b]:.3f} | soft  stage {sJs[b]:.3f} terminal {sJT[b]:.3f} total {sJ[b]:.3f}")
    out_dir = out_dir or folder_name
    os.makedirs(out_dir, exist_ok=True)
    filename = os.path.join(out_dir, f"rockETH_compare_closed_loop_{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}.npz")
    np.savez(filename, **{"r_" + k: res[k] for k in ("state_trajectory", "input_trajectory", "nominal_trajectory_x", "nominal_trajectory_u",
                                                      "backoff_trajectory_x", "backoff_trajectory_u", "success")},
             dt=res["dt"], g=res["g"], nx=res["nx"], nu=res["nu"], simulation_time_steps=T, N=res["N"], x0=ROCKET_X0, W=W,
             Jr_stage=Js, Jr_terminal=JT, Jr_total=J, **soft_keys)
    print(f"Saved robust closed loops of {batch} scenarios ({int(ok.sum())} without a failed solve; mean total cost of those "
          f"{J[ok].mean():.3f}) to {filename}")
    return filename


if __name__ == "__main__":
    import argparse
    p = argparse.ArgumentParser(description="Rocket robust closed loops over disturbance draws (GPU path)")
    p.add_argument("--run", action="store_true")
    p.add_argument("--batch", type=int, default=4096)
    p.add_argument("--keep", type=int, default=1, help="scenarios whose nominal / back-off trajectories are stored")
    p.add_argument("--out", default=None)
    p.add_argument("--soft", type=int, default=1, help="scenarios (from the first) that also run the soft-constrained NMPC arm")
    a = p.parse_args()
    if a.run:
        generate(a.batch, a.keep, a.out, soft=a.soft)
    else:
        print("plotting is the reference's: its plot() reads the r_* / s_* keys this driver writes (take scenario 0 of the leading axis)")
