"""The closed-loop rollout and result file shared by the drivers of this package: the loop and the np.savez key set of the
reference's expe/main_*_robust_closed_loop.py (rocket :144-206, pendulum :63-120, quadrotor :108-160)."""
import os
from datetime import datetime

import numpy as np


def rollout_and_save(tag, solver, m, N, x0, simulation_time_steps, noise, out_dir, folder_name, file_prefix, verbose=True):
    """noise(i) -> w_i (nx,) or None.  Returns (file name, all solves succeeded)."""
    x0 = np.array(x0, float).reshape(-1)
    state_trajectory = np.zeros((m.nx, simulation_time_steps))
    state_trajectory[:, 0] = x0
    input_trajectory = np.zeros((m.nu, simulation_time_steps - 1))
    backoff_trajectory_x = np.zeros((m.nx, N + 1, simulation_time_steps))
    backoff_trajectory_u = np.zeros((m.nu, N, simulation_time_steps))
    nominal_trajectory_x = np.zeros((m.nx, N + 1, simulation_time_steps))
    nominal_trajectory_u = np.zeros((m.nu, N, simulation_time_steps))
    t_jac = np.zeros((simulation_time_steps, 1))
    t_qp = np.zeros((simulation_time_steps, 1))
    t_riccati = np.zeros((simulation_time_steps, 1))
    ok = True
    for i in range(simulation_time_steps):
        if i > 0:
            solver.reset_warm_start()
        solution = solver.solve(x0)
        if not solution.get("success", True):
            print(f"[{tag}] Step {i + 1}/{simulation_time_steps}: solver failed, stopping")
            ok = False
            break
        t_jac[i] = float(solution.get("t_jac_ms", np.nan))
        t_qp[i] = float(solution.get("t_qp_ms", np.nan))
        t_riccati[i] = float(solution.get("t_backward_ms", np.nan))
        if verbose:
            print(f"[{tag}] Step {i + 1}/{simulation_time_steps}: jac={t_jac[i, 0]:.2f} ms, qp={t_qp[i, 0]:.2f} ms, "
                  f"riccati={t_riccati[i, 0]:.2f} ms")
        backoff_trajectory_x[:, :, i] = solution["backoff_x"].T
        backoff_trajectory_u[:, :, i] = solution["backoff_u"].T
        nominal_trajectory_x[:, :, i] = solution["primal_x"]
        nominal_trajectory_u[:, :, i] = solution["primal_u"]
        state_trajectory[:, i] = solution["primal_x"][:, 0]
        if i < simulation_time_steps - 1:
            input_trajectory[:, i] = solution["primal_u"][:, 0]
        w = noise(i)
        x0 = np.asarray(m.ddyn(x0, solution["primal_u"][:, 0])).reshape(-1)
        if w is not None:
            x0 = x0 + m.E @ w
    out_dir = out_dir or folder_name
    os.makedirs(out_dir, exist_ok=True)
    filename = os.path.join(out_dir, f"{file_prefix}_{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}.npz")
    np.savez(filename, state_trajectory=state_trajectory, input_trajectory=input_trajectory,
             nominal_trajectory_x=nominal_trajectory_x, nominal_trajectory_u=nominal_trajectory_u,
             backoff_trajectory_x=backoff_trajectory_x, backoff_trajectory_u=backoff_trajectory_u,
             dt=m.dt, g=m.g, nx=m.nx, nu=m.nu, simulation_time_steps=simulation_time_steps, N=N,
             t_jac=t_jac, t_qp=t_qp, t_riccati=t_riccati)
    print(f"Results saved to {filename}")
    return filename, ok


def cli(description, generate):
    import argparse
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument("--run", action="store_true", help="Run the closed-loop simulation (generate data)")
    parser.add_argument("--N", type=int, default=None, help="Horizon length (overrides default if provided)")
    parser.add_argument("--out", default=None, help="output directory")
    args = parser.parse_args()
    if args.run:
        generate(args.N, args.out)
    else:
        print("plotting is the reference's: run its driver of the same name without --run on the saved file")
