"""``OCP`` base class: sizes, default regularisers and the stage-wise vector layout (mirror of solver/ocp.py:7-74)."""
from __future__ import annotations

import numpy as np


class OCP:
    def __init__(self, N, Q, R, m, Qf, Q_reg=None, R_reg=None, Q_reg_f=None):
        self.N, self.m = N, m
        self.Q, self.R, self.Qf = np.asarray(Q, float), np.asarray(R, float), np.asarray(Qf, float)
        self.xf = np.zeros((m.nx, 1))
        # defaults of solver/ocp.py:15-26
        self.Q_reg = np.eye(self.Q.shape[0]) if Q_reg is None else np.asarray(Q_reg, float)
        self.R_reg = np.eye(self.R.shape[0]) if R_reg is None else np.asarray(R_reg, float)
        self.Q_reg_f = np.eye(self.Qf.shape[0]) if Q_reg_f is None else np.asarray(Q_reg_f, float)
        self.CONV_EPS = 1e-6

    def pack_primal_nominal(self, X: np.ndarray, U: np.ndarray) -> np.ndarray:
        """y = [x0; u0; ...; x_{N-1}; u_{N-1}; xN] (solver/ocp.py:43-56)."""
        nx, nu, N = self.m.nx, self.m.nu, self.N
        assert X.shape == (nx, N + 1), f"X must be {(nx, N + 1)}, got {X.shape}"
        assert U.shape == (nu, N), f"U must be {(nu, N)}, got {U.shape}"
        return np.concatenate([np.vstack([X[:, :N], U]).reshape(-1, order="F"), X[:, N]])

    def unpack_primal_nominal(self, y: np.ndarray):
        """Inverse of ``pack_primal_nominal`` (solver/ocp.py:58-74)."""
        nx, nu, N = self.m.nx, self.m.nu, self.N
        yv = np.asarray(y).reshape(-1)
        expected = (nx + nu) * N + nx
        assert yv.size == expected, f"y must be size {expected}, got {yv.size}"
        Z = yv[:(nx + nu) * N].reshape(nx + nu, N, order="F")
        return np.column_stack([Z[:nx], yv[(nx + nu) * N:]]), Z[nx:].copy()
