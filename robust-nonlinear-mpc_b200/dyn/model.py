"""``Model`` base class with the reference's attribute set (dyn/model.py:4-45); ``ddyn`` runs on the GPU."""
from __future__ import annotations

import numpy as np


class Model:
    """Mirror of dyn/model.py:4-45.

    Attributes (same names as the reference): ``nx, nu, nw, ni, ni_f, dt, G, g, Gf, gf, E``.  ``ddyn(x, u, h=.05)`` is the
    classical RK4 step of the model's ODE (dyn/model.py:15-34), evaluated by the CUDA library (``rnmpc_ddyn``); there
    is no Python/CPU implementation of the ODE in the product path.
    """

    model_name = None          # "pendulum" | "quadrotor" | "rocket": selects the compiled ODE in the library
    discretization_method = "rk4"

    def __init__(self):
        self.dt = None
        self.nx = self.nu = self.nw = self.ni = self.ni_f = None

    # -- box helper used by the concrete models -------------------------------------------------------------------
    def _set_box(self, x_lb, x_ub, u_lb, u_ub):
        nz = self.nx + self.nu
        self.G = np.vstack([np.eye(nz), -np.eye(nz)])
        self.g = np.concatenate([x_ub, u_ub, -np.asarray(x_lb), -np.asarray(u_lb)]).astype(float)
        self.ni = 2 * nz
        self.Gf = np.vstack([np.eye(self.nx), -np.eye(self.nx)])
        self.gf = np.concatenate([x_ub, -np.asarray(x_lb)]).astype(float)
        self.ni_f = 2 * self.nx

    def box(self):
        """(x_ub, x_lb, u_ub, u_lb, xf_ub, xf_lb) recovered from g / gf; raises if G, Gf are not [I; -I]."""
        nx, nu = self.nx, self.nu
        nz = nx + nu
        G, Gf = np.asarray(self.G, float), np.asarray(self.Gf, float)
        if G.shape != (2 * nz, nz) or not np.array_equal(G, np.vstack([np.eye(nz), -np.eye(nz)])) \
                or Gf.shape != (2 * nx, nx) or not np.array_equal(Gf, np.vstack([np.eye(nx), -np.eye(nx)])):
            raise NotImplementedError("rnmpc_b200 supports box constraints G = [I; -I], Gf = [I; -I] only "
                                      "(the structure of every reference model, dyn/rocket.py:129-139)")
        g, gf = np.asarray(self.g, float).reshape(-1), np.asarray(self.gf, float).reshape(-1)
        return g[:nx], -g[nz:nz + nx], g[nx:nz], -g[nz + nx:], gf[:nx], -gf[nx:]

    _ddyn_engine = None

    def _engine(self):
        from ..engine import BatchSolver   # deferred: importing the models must not require CUDA
        cls = type(self)
        if cls._ddyn_engine is None:
            I = np.eye
            cls._ddyn_engine = BatchSolver(self.model_name, 1, I(self.nx), I(self.nu), I(self.nx), I(self.nx), I(self.nu),
                                           I(self.nx), I(self.nx), np.ones(self.nx), -np.ones(self.nx), np.ones(self.nu),
                                           -np.ones(self.nu), capacity=1)
        return cls._ddyn_engine

    def ddyn(self, x, u, h=.05):
        if getattr(self, "discretization_method", "rk4") != "rk4":
            raise NotImplementedError("only the RK4 discretisation of dyn/model.py:26-32 is compiled into the CUDA library")
        self.dt = h        # the reference overwrites dt on every call (dyn/model.py:17)
        x = np.asarray(x, dtype=float).reshape(-1)
        u = np.asarray(u, dtype=float).reshape(-1)
        xp = self._engine().ddyn(x[None], u[None], h=float(h))
        return xp.cpu().numpy()[0]

    def remove_constraints(self):
        raise NotImplementedError("unconstrained models are outside the ported path (dyn/model.py:36-45)")
