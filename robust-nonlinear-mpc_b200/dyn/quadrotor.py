"""Quadrotor model, mirror of dyn/quadrotor.py:30-107 (ODE :108-174 in csrc/rnmpc_models.cuh)."""
import numpy as np

from .model import Model


class Quadrotor(Model):
    model_name = "quadrotor"

    def __init__(self):
        self.params = {'m': 1.0, 'g': 9.81, 'l': 0.15, 'Jx': 0.02, 'Jy': 0.02, 'Jz': 0.04, 'kM': 0.01}
        self.state_names = ['x', 'y', 'z', 'vx', 'vy', 'vz', 'qw', 'qx', 'qy', 'qz', 'wx', 'wy', 'wz']
        self.control_names = ['f1', 'f2', 'f3', 'f4']
        self.nx, self.nu, self.nw = 13, 4, 13
        self.dt = 0.05
        f_hover = self.params['m'] * self.params['g'] / 4.0
        self.neutral_state = np.array([0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], float)
        self.neutral_input = np.full(4, f_hover)
        x_ub = np.array([20.0] * 3 + [10.0] * 3 + [1.5] * 4 + [20.0] * 3)
        self._set_box(-x_ub, x_ub, np.zeros(4), 20.0 * np.ones(4))
        self.E = np.diag([0.05] * 3 + [0.1] * 3 + [0.02, 0.02, 0.02, 0.01] + [0.2] * 3)
