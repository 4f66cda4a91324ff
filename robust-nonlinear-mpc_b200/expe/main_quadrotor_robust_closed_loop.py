"""Quadrotor robust closed loop on the GPU path: command line, configuration and result schema of the reference's
expe/main_quadrotor_robust_closed_loop.py (:35-160; BASELINE.json configs[1]): 30 steps, rti = 3, fast_sls_rti_steps = 2, no noise.
The reference draws its initial state from an unseeded generator (:82-90); here the same recipe is seeded (`--seed`, default 0).

    python -m rnmpc_b200.expe.main_quadrotor_robust_closed_loop --run [--N 15] [--out DIR]"""
import numpy as np

from ..dyn.quadrotor import Quadrotor
from ..solver.SCP_SLS_jit import SCP_SLS
from ..workloads import quadrotor_x0
from ._closed_loop import cli, rollout_and_save

folder_name = "quadrotor_robust_closed_loop"


def generate(N=None, out_dir=None, seed=0, verbose=True):
    m = Quadrotor()
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [1.0] * 4 + [2.0] * 3)
    R = np.eye(4)
    Qf = 10 * Q
    N = int(N) if N is not None else 15
    m.dt = 0.05
    q_vec_std = 0.5 * np.deg2rad(2.0)
    q_w_std = 0.1 * q_vec_std
    m.E = m.dt * 5 * np.diag([0.10] * 3 + [0.15] * 3 + [q_w_std] + [q_vec_std] * 3 + [0.2] * 3)
    solver = SCP_SLS(N, Q, R, m, Qf, Q_reg=1e4 * np.eye(m.nx), R_reg=1e4 * np.eye(m.nu), Q_reg_f=1e4 * np.eye(m.nx),
                     rti=3, fast_sls_rti_steps=2)
    solver.fast_SLS_solver.solver_forward.export_standard_qp = False
    solver.verbose = verbose
    solver.fast_SLS_solver.verbose = verbose
    solver.fast_SLS_solver.solver_forward.verbose = False
    solver.fast_SLS_solver.CONV_EPS = 1e-4
    solver.epsilon_convergence = 1e-3
    return rollout_and_save("Quadrotor", solver, m, N, quadrotor_x0(seed), 30, lambda i: None, out_dir, folder_name,
                            "quadrotor_robust_closed_loop", verbose)


if __name__ == "__main__":
    cli("Quadrotor robust closed-loop experiment (GPU path)", generate)
