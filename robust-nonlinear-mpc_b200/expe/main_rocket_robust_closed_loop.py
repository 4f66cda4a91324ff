"""Rocket robust closed loop on the GPU path, same command line and result schema as the reference's driver
(expe/main_rocket_robust_closed_loop.py: configuration :33-93, initial state :110-126, loop :144-183, np.savez keys :189-206,
--run / --N :457-466), so that the reference's plot functions read the file this script writes.

    python -m rnmpc_b200.expe.main_rocket_robust_closed_loop --run [--N 15] [--out DIR]

Differences from the reference driver: the first-step nominal trajectory comes from the package's tube-free SQP initialiser
instead of IPOPT; plotting is left to the reference (`plot()` there needs matplotlib only)."""
import argparse
import os
from datetime import datetime

import numpy as np

from ..dyn.rocket import Rocket
from ..solver.SCP_SLS_jit import SCP_SLS
from ..workloads import ROCKET_X0

folder_name = "rockETH_robust_closed_loop"


def generate(N=None, out_dir=None, seed=0, simulation_time_steps=30, verbose=True):
    """One 30-step closed loop from the drivers' fixed initial state with disturbances w_t = 2 rand(nx) - 1 (np.random.seed(0));
    returns the path of the .npz file."""
    np.random.seed(seed)
    m = Rocket()
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [1.0] * 4 + [1.0, 5.0, 5.0] + [1.0] * 4)
    R = np.eye(4)
    Qf = 10 * Q
    N = int(N) if N is not None else 15
    m.dt = 0.05
    q_vec_std = 0.5 * np.deg2rad(2.0)
    q_w_std = 0.1 * q_vec_std
    m.E = m.dt * np.diag([0.20] * 3 + [0.20] * 3 + [q_vec_std] * 3 + [q_w_std] + [0.2] * 3 + [0.8, 0.2, 0.04, 0.04])
    solver = SCP_SLS(N, Q, R, m, Qf, Q_reg=1e4 * np.eye(m.nx), R_reg=1e4 * np.eye(m.nu), Q_reg_f=1e4 * np.eye(m.nx),
                     rti=1, fast_sls_rti_steps=1)
    solver.fast_SLS_solver.solver_forward.export_standard_qp = False
    solver.verbose = verbose
    solver.fast_SLS_solver.verbose = verbose
    solver.fast_SLS_solver.solver_forward.verbose = False
    solver.fast_SLS_solver.CONV_EPS = 1e-4
    solver.epsilon_convergence = 1e-3

    x0 = ROCKET_X0.copy()
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
    success = np.ones(simulation_time_steps, dtype=bool)

    for i in range(simulation_time_steps):
        if i > 0:
            solver.reset_warm_start()
        solution = solver.solve(x0)
        if not solution.get("success", True):
            success[i:] = False
            print(f"[Rocket] Step {i + 1}/{simulation_time_steps}: solver failed, stopping")
            break
        t_jac[i] = float(solution.get("t_jac_ms", np.nan))
        t_qp[i] = float(solution.get("t_qp_ms", np.nan))
        t_riccati[i] = float(solution.get("t_backward_ms", np.nan))
        if verbose:
            print(f"[Rocket] Step {i + 1}/{simulation_time_steps}: jac={t_jac[i, 0]:.2f} ms, qp={t_qp[i, 0]:.2f} ms, "
                  f"riccati={t_riccati[i, 0]:.2f} ms")
        backoff_trajectory_x[:, :, i] = solution["backoff_x"].T
        backoff_trajectory_u[:, :, i] = solution["backoff_u"].T
        nominal_trajectory_x[:, :, i] = solution["primal_x"]
        nominal_trajectory_u[:, :, i] = solution["primal_u"]
        state_trajectory[:, i] = solution["primal_x"][:, 0]
        if i < simulation_time_steps - 1:
            input_trajectory[:, i] = solution["primal_u"][:, 0]
        w = 2 * np.random.rand(m.nx) - 1
        x0 = np.asarray(m.ddyn(x0, solution["primal_u"][:, 0])).reshape(-1) + m.E @ w

    out_dir = out_dir or folder_name
    os.makedirs(out_dir, exist_ok=True)
    filename = os.path.join(out_dir, f"rockETH_robust_closed_loop_{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}.npz")
    np.savez(filename, state_trajectory=state_trajectory, input_trajectory=input_trajectory,
             nominal_trajectory_x=nominal_trajectory_x, nominal_trajectory_u=nominal_trajectory_u,
             backoff_trajectory_x=backoff_trajectory_x, backoff_trajectory_u=backoff_trajectory_u,
             dt=m.dt, g=m.g, nx=m.nx, nu=m.nu, simulation_time_steps=simulation_time_steps, N=N,
             t_jac=t_jac, t_qp=t_qp, t_riccati=t_riccati)
    print(f"Results saved to {filename}")
    return filename, bool(success.all())


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Rocket robust closed-loop experiment (GPU path)")
    parser.add_argument("--run", action="store_true", help="Run the closed-loop simulation (generate data)")
    parser.add_argument("--N", type=int, default=None, help="Horizon length (overrides default if provided)")
    parser.add_argument("--out", default=None, help="output directory (default: rockETH_robust_closed_loop/)")
    args = parser.parse_args()
    if args.run:
        generate(args.N, args.out)
    else:
        print("plotting is the reference's: run its expe/main_rocket_robust_closed_loop.py without --run on the saved file")
