"""Cart-pole model, mirror of dyn/pendulum.py:7-24 (the ODE, :26-44, lives in csrc/rnmpc_models.cuh)."""
import numpy as np

from .model import Model


class Pendulum(Model):
    model_name = "pendulum"

    def __init__(self):
        self.nx, self.nu, self.nw = 4, 1, 4
        self.dt = 0.05
        x_max, u_max = 10.0 * np.ones(4), np.array([5.0])
        self._set_box(-x_max, x_max, -u_max, u_max)
        self.E = 0.1 * np.eye(4)

    def replace_constraints(self, x_max, x_min, u_max, u_min, x_max_f, x_min_f):
        """dyn/pendulum.py:142-144."""
        self.g = np.hstack((x_max, u_max, -np.asarray(x_min), -np.asarray(u_min))).astype(float)
        self.gf = np.hstack((x_max_f, -np.asarray(x_min_f))).astype(float)
