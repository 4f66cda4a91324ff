"""Shared helpers of the test-suite (oracle-side and GPU-side)."""
import os

import numpy as np

from oracle import models as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    """max |a-b| / max |b|  (norm-wise relative error)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def relf(a, b, floor=1.0):
    """max |a-b| / max(floor, max |b|): relative error with an absolute floor (quantities that converge to zero)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def make_solver(name, capacity, **kw):
    """BatchSolver configured like the reference driver of `name` (oracle.models.CONFIGS)."""
    from rnmpc_b200.engine import BatchSolver
    m, cfg = om.CONFIGS[name]()
    args = dict(rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=capacity)
    args.update(kw)
    s = BatchSolver(name, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E,
                    m.x_ub, m.x_lb, m.u_ub, m.u_lb, **args)
    return m, cfg, s


def rocket_scenarios(B, seed=2025, steps=29, x0_scale=0.0):
    """Rocket Monte-Carlo scenarios.  x0_scale = 0: BASELINE.json configs[3] (the fixed x0 of
    expe/main_rocket_compare_closed_loop.py:75-91 for every scenario, disturbance draws U[-1,1]^nx per step);
    x0_scale > 0: configs[4]: x0 perturbed by x0_scale * half-range * U[-1,1], quaternion renormalised."""
    m, _ = om.rocket_config()
    rng = np.random.default_rng(seed)
    x0 = om.ROCKET_X0[None, :] + x0_scale * (m.x_ub - m.x_lb)[None, :] / 2 * rng.uniform(-1, 1, size=(B, m.nx))
    qn = np.linalg.norm(om.ROCKET_X0[6:10])
    x0[:, 6:10] *= (qn / np.linalg.norm(x0[:, 6:10], axis=1))[:, None]
    W = rng.uniform(-1, 1, size=(B, steps, m.nx))
    return x0, W
