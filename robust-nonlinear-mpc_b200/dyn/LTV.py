"""LTV container, mirror of dyn/LTV.py:17-32 (all-ones placeholders; only the sizes and constraint copies matter)."""
import numpy as np

from .model import Model


class LTV(Model):
    def __init__(self, m, N):
        self.N = N
        self.nx, self.nu, self.nw = m.nx, m.nu, m.nw
        self.G, self.ni, self.Gf, self.gf, self.ni_f = m.G, m.ni, m.Gf, m.gf, m.ni_f
        self.model_name = m.model_name
        self.A_list = [np.ones((self.nx, self.nx)) for _ in range(N)]
        self.B_list = [np.ones((self.nx, self.nu)) for _ in range(N)]
        self.E_list = [np.ones((self.nx, self.nw)) for _ in range(N + 1)]
        self.g_list = [np.ones(self.ni) for _ in range(N)] + [np.ones(self.ni_f)]
