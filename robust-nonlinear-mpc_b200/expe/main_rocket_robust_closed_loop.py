"""Rocket robust closed loop on the GPU path, same command line and result schema as the reference's driver
(expe/main_rocket_robust_closed_loop.py: configuration :33-93, initial state :110-126, loop :144-183, np.savez keys :189-206,
--run / --N :457-466), so that the reference's plot functions read the file this script writes.

    python -m rnmpc_b200.expe.main_rocket_robust_closed_loop --run [--N 15] [--out DIR]

Differences from the reference driver: the first-step nominal trajectory comes from the package's tube-free SQP initialiser
instead of IPOPT; plotting is left to the reference (`plot()` there needs matplotlib only)."""
import numpy as np

from ..dyn.rocket import Rocket
from ..solver.SCP_SLS_jit import SCP_SLS
from ..workloads import ROCKET_X0
from ._closed_loop import cli, rollout_and_save

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

    noise = lambda i: 2 * np.random.rand(m.nx) - 1          # main_rocket_robust_closed_loop.py:180
    return rollout_and_save("Rocket", solver, m, N, ROCKET_X0.copy(), simulation_time_steps, noise, out_dir, folder_name,
                            "rockETH_robust_closed_loop", verbose)


if __name__ == "__main__":
    cli("Rocket robust closed-loop experiment (GPU path)", generate)
