"""Workload definitions of the reference's experiment drivers, product side (no dependency on the test oracle).

rocket: expe/main_rocket_robust_closed_loop.py:33-85,110-126 / expe/main_rocket_compare_closed_loop.py:41-70,75-91,107-114
(model, disturbance map E, costs, horizon, RTI settings, the fixed initial state) and the Monte-Carlo scenario generator of
BASELINE.json configs[3] / [4] (disturbance draws U[-1,1]^nx per step, optionally perturbed initial states)."""
import numpy as np

from .dyn.quadrotor import Quadrotor
from .dyn.rocket import Rocket

ROCKET_X0 = np.array([1.75729, 4.15951, 4.72757, -0.18913, -0.38367, -0.08697, -0.79487, 0.00768, -0.21110,
                      -0.56883, -0.12752, -0.58026, -0.76542, 0.20555, 0.54610, -0.40116, -0.35401])


def rocket_config():
    """(model, settings) of the rocket drivers: N = 15, rti = 1, fast_sls_rti_steps = 1."""
    m = Rocket()
    q_vec = 0.5 * np.deg2rad(2.0)
    m.E = m.dt * np.diag([0.2] * 6 + [q_vec] * 3 + [0.1 * q_vec] + [0.2] * 3 + [0.8, 0.2, 0.04, 0.04])
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [1.0] * 4 + [1.0, 5.0, 5.0] + [1.0] * 4)
    cfg = dict(N=15, Q=Q, R=np.eye(4), Qf=10 * Q, Q_reg=1e4 * np.eye(17), R_reg=1e4 * np.eye(4), Q_reg_f=1e4 * np.eye(17),
               rti=1, fast_sls_rti_steps=1)
    return m, cfg


def rocket_scenarios(B, seed=2025, steps=29, x0_scale=0.0):
    """x0 [B, nx], W [B, steps, nx].  x0_scale = 0: the drivers' fixed x0 for every scenario, disturbance draws only;
    x0_scale > 0: x0 perturbed by x0_scale * half-range * U[-1,1], quaternion renormalised."""
    m, _ = rocket_config()
    x_ub, x_lb, _, _, _, _ = m.box()
    rng = np.random.default_rng(seed)
    x0 = ROCKET_X0[None, :] + x0_scale * (x_ub - x_lb)[None, :] / 2 * rng.uniform(-1, 1, size=(B, m.nx))
    qn = np.linalg.norm(ROCKET_X0[6:10])
    x0[:, 6:10] *= (qn / np.linalg.norm(x0[:, 6:10], axis=1))[:, None]
    W = rng.uniform(-1, 1, size=(B, steps, m.nx))
    return x0, W


def quadrotor_x0(seed=0):
    """Seeded form of the unseeded draw at expe/main_quadrotor_robust_closed_loop.py:82-90: uniform in the half range of the
    state box, quaternion slots 6:10 <- a normalised normal(4) draw."""
    m = Quadrotor()
    x_ub, x_lb, _, _, _, _ = m.box()
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(x_lb / 2, x_ub / 2)
    q = rng.normal(size=4)
    x0[6:10] = q / np.linalg.norm(q)
    return x0
