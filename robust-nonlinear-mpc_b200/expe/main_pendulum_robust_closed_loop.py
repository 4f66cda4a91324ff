"""Pendulum robust closed loop on the GPU path: command line, configuration and result schema of the reference's
expe/main_pendulum_robust_closed_loop.py (:24-120; BASELINE.json configs[0]): x0 = [0.5, 0.5, 0, 0], E = 0.003 I, bounds
+-10 / +-5, Q = R = I, Qf = 10 I, Q_reg = R_reg = 1e3 I, Q_reg_f = 1e4 I, rti = 3, fast_sls_rti_steps = 2, 60 steps, no noise.

    python -m rnmpc_b200.expe.main_pendulum_robust_closed_loop --run [--N 15] [--out DIR]"""
import numpy as np

from ..dyn.pendulum import Pendulum
from ..solver.SCP_SLS_jit import SCP_SLS
from ._closed_loop import cli, rollout_and_save

folder_name = "pendulum_robust_closed_loop"


def generate(N=None, out_dir=None, verbose=True):
    m = Pendulum()
    Q, R, Qf = np.eye(m.nx), np.eye(m.nu), 10 * np.eye(m.nx)
    N = int(N) if N is not None else 15
    m.E = 0.003 * np.eye(m.nx)
    m.dt = 0.05
    m.replace_constraints(10 * np.ones(m.nx), -10 * np.ones(m.nx), 5 * np.ones(m.nu), -5 * np.ones(m.nu),
                          10 * np.ones(m.nx), -10 * np.ones(m.nx))
    solver = SCP_SLS(N, Q, R, m, Qf, Q_reg=1e3 * np.eye(m.nx), R_reg=1e3 * np.eye(m.nu), Q_reg_f=1e4 * np.eye(m.nx),
                     rti=3, fast_sls_rti_steps=2)
    solver.epsilon_convergence = 1e-10
    solver.verbose = verbose
    solver.fast_SLS_solver.verbose = verbose
    solver.fast_SLS_solver.solver_forward.verbose = False
    x0 = np.array([0.5, 0.5, 0.0, 0.0])
    return rollout_and_save("Pendulum", solver, m, N, x0, 60, lambda i: None, out_dir, folder_name, "pendulum_robust_closed_loop",
                            verbose)


if __name__ == "__main__":
    cli("Pendulum robust closed-loop experiment (GPU path)", generate)
