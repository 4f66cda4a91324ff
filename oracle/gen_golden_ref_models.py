"""Writes tests/golden/refdyn_{pendulum,quadrotor,rocket}.npz: outputs of the reference's own ``ode`` / ``ddyn`` statements
(``dyn/*.py`` imported unmodified through ``oracle/ref_models.py``) at seeded points inside the state / input boxes, plus the
constant tables of the reference constructors.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden_ref_models
"""
import os

import numpy as np

from . import models as om
from . import ref_models as rm

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def points(name, n=32, seed=11):
    """Seeded states / inputs: half of them well inside the boxes, half across the whole boxes (the gimbal map of the rocket
    and the quaternion terms are exercised far from the linearisation points); the rocket's fixed x0 comes first."""
    m = om.MODELS[name]()
    rng = np.random.default_rng(seed)
    xs = np.array([rng.uniform(m.x_lb * s, m.x_ub * s) for s in ([0.3] * (n // 2) + [1.0] * (n - n // 2))])
    us = np.array([rng.uniform(m.u_lb, m.u_ub) for _ in range(n)])
    if name == "rocket":
        xs[0] = om.ROCKET_X0
        us[0] = 0.0
    return xs, us


def main():
    classes = rm.load()
    for name in ("pendulum", "quadrotor", "rocket"):
        ref = classes[name]()
        xs, us = points(name)
        ode = np.array([rm.ode(ref, x, u) for x, u in zip(xs, us)])
        xp = np.array([rm.ddyn(ref, x, u, 0.05) for x, u in zip(xs, us)])
        xp02 = np.array([rm.ddyn(ref, x, u, 0.02) for x, u in zip(xs, us)])
        AB = [rm.jacobians(ref, x, u, 0.05) for x, u in zip(xs, us)]
        ref.discretization_method = "euler"
        xe = np.array([rm.ddyn(ref, x, u, 0.05) for x, u in zip(xs, us)])
        np.savez(os.path.join(OUT, f"refdyn_{name}.npz"), x=xs, u=us, ode=ode, xp=xp, xp_h002=xp02, xp_euler=xe, A=np.array([a for a, _ in AB]), B=np.array([b for _, b in AB]),
                 G=np.asarray(ref.G, float), g=np.asarray(ref.g, float).ravel(), Gf=np.asarray(ref.Gf, float),
                 gf=np.asarray(ref.gf, float).ravel(), E=np.asarray(ref.E, float),
                 source=f"reference dyn/{name}.py + dyn/model.py executed through oracle/ref_models.py (numeric CasADi stand-in)")
        print(name, "written:", len(xs), "points, max |ode|", np.abs(ode).max())


if __name__ == "__main__":
    main()
