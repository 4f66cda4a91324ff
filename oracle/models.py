"""NumPy restatement of the reference's nonlinear models and their RK4 discretisation.

Follows (reference file:line):
  * ``dyn/model.py:15-34``      Model.ddyn  -- classical RK4, default step h = 0.05 (quirk Q6: the SCP path always
                                 calls ``ddyn`` with the default h, whatever ``m.dt`` is)
  * ``dyn/pendulum.py:7-44``    cart-pole
  * ``dyn/quadrotor.py:30-174`` 13-state quaternion quadrotor
  * ``dyn/rocket.py:21-254``    17-state rockETH rocket with gimbal map
The ODEs are written for arrays of any dtype (float64 or complex128) so that Jacobians can be taken by the
complex-step method (exact to rounding for these analytic right-hand sides), which is independent of the
forward-mode dual numbers used by the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np

H_DEFAULT = 0.05  # dyn/model.py:15


def _quat_rot(qw, qx, qy, qz):
    # dyn/rocket.py:195-208, dyn/quadrotor.py:125-138  (body -> world rotation from [w,x,y,z])
    return (
        (1 - 2 * qy**2 - 2 * qz**2, 2 * qx * qy - 2 * qz * qw, 2 * qx * qz + 2 * qy * qw),
        (2 * qx * qy + 2 * qz * qw, 1 - 2 * qx**2 - 2 * qz**2, 2 * qy * qz - 2 * qx * qw),
        (2 * qx * qz - 2 * qy * qw, 2 * qy * qz + 2 * qx * qw, 1 - 2 * qx**2 - 2 * qy**2),
    )


def _quat_kin(qw, qx, qy, qz, wx, wy, wz):
    # dyn/rocket.py:214-223: q_dot = 0.5 * Omega(omega) q, q = [w,x,y,z]
    return (
        0.5 * (-wx * qx - wy * qy - wz * qz),
        0.5 * (wx * qw + wz * qy - wy * qz),
        0.5 * (wy * qw - wz * qx + wx * qz),
        0.5 * (wz * qw + wy * qx - wx * qy),
    )


class BoxModel:
    """Common data of a reference model: sizes, box constraints G=[I;-I], disturbance map E."""

    name = "base"

    def _set_box(self, x_lb, x_ub, u_lb, u_ub):
        # dyn/rocket.py:129-139, dyn/quadrotor.py:96-101, dyn/pendulum.py:13-21
        self.x_lb, self.x_ub = np.asarray(x_lb, float), np.asarray(x_ub, float)
        self.u_lb, self.u_ub = np.asarray(u_lb, float), np.asarray(u_ub, float)
        nz = self.nx + self.nu
        self.G = np.vstack([np.eye(nz), -np.eye(nz)])
        self.g = np.concatenate([self.x_ub, self.u_ub, -self.x_lb, -self.u_lb])
        self.ni = 2 * nz
        self.Gf = np.vstack([np.eye(self.nx), -np.eye(self.nx)])
        self.gf = np.concatenate([self.x_ub, -self.x_lb])
        self.ni_f = 2 * self.nx

    def ode(self, x, u):
        raise NotImplementedError

    def ddyn(self, x, u, h=H_DEFAULT):
        # dyn/model.py:26-32
        x = np.asarray(x).reshape(-1)
        u = np.asarray(u).reshape(-1)
        k1 = self.ode(x, u)
        k2 = self.ode(x + 0.5 * h * k1, u)
        k3 = self.ode(x + 0.5 * h * k2, u)
        k4 = self.ode(x + h * k3, u)
        return x + (1 / 6) * (k1 + 2 * k2 + 2 * k3 + k4) * h

    def jacobians(self, x, u, h=H_DEFAULT):
        """A = d ddyn/dx, B = d ddyn/du (what CasADi's ``jacobian`` yields, solver/SCP_SLS_jit.py:203-205),
        by complex-step differentiation."""
        x = np.asarray(x, float).reshape(-1)
        u = np.asarray(u, float).reshape(-1)
        eps = 1e-30
        A = np.empty((self.nx, self.nx))
        B = np.empty((self.nx, self.nu))
        for i in range(self.nx):
            xc = x.astype(complex)
            xc[i] += 1j * eps
            A[:, i] = self.ddyn(xc, u.astype(complex), h).imag / eps
        for i in range(self.nu):
            uc = u.astype(complex)
            uc[i] += 1j * eps
            B[:, i] = self.ddyn(x.astype(complex), uc, h).imag / eps
        return A, B


class Pendulum(BoxModel):
    """dyn/pendulum.py:7-44 (cart-pole)."""

    name = "pendulum"
    model_id = 0

    def __init__(self):
        self.nx, self.nu, self.nw = 4, 1, 4
        self.dt = 0.05
        self._set_box(-10 * np.ones(4), 10 * np.ones(4), [-5.0], [5.0])
        self.E = 0.1 * np.eye(4)

    def ode(self, X, u):
        m1, m2, l, g = 1.0, 0.1, 0.5, 9.81  # dyn/pendulum.py:31-34
        _, x_dot, th, th_dot = X
        uu = u[0]
        s, c = np.sin(th), np.cos(th)
        den = m1 + m2 * (1 - c**2)
        x_dd = (uu + m2 * l * th_dot**2 * s - m2 * g * s * c) / den
        th_dd = (-uu * c - m2 * l * th_dot**2 * s * c + (m1 + m2) * g * s) / (l * den)
        return np.array([x_dot, x_dd, th_dot, th_dd])


class Quadrotor(BoxModel):
    """dyn/quadrotor.py:30-174."""

    name = "quadrotor"
    model_id = 1

    def __init__(self):
        self.nx, self.nu, self.nw = 13, 4, 13
        self.dt = 0.05
        x_ub = np.array([20.0] * 3 + [10.0] * 3 + [1.5] * 4 + [20.0] * 3)
        self._set_box(-x_ub, x_ub, np.zeros(4), 20.0 * np.ones(4))
        self.E = np.diag([0.05] * 3 + [0.1] * 3 + [0.02, 0.02, 0.02, 0.01] + [0.2] * 3)

    def ode(self, X, u):
        m, g, l, Jx, Jy, Jz, kM = 1.0, 9.81, 0.15, 0.02, 0.02, 0.04, 0.01  # dyn/quadrotor.py:32-40
        vx, vy, vz = X[3], X[4], X[5]
        qw, qx, qy, qz = X[6], X[7], X[8], X[9]
        wx, wy, wz = X[10], X[11], X[12]
        f1, f2, f3, f4 = u
        Fz = f1 + f2 + f3 + f4
        R = _quat_rot(qw, qx, qy, qz)
        ax, ay, az = R[0][2] * Fz / m, R[1][2] * Fz / m, R[2][2] * Fz / m - g
        qd = _quat_kin(qw, qx, qy, qz, wx, wy, wz)
        tx, ty, tz = l * (f2 - f4), l * (f3 - f1), kM * (f1 - f2 + f3 - f4)
        # J wdot = tau - w x (J w)
        wdx = (tx - (wy * Jz * wz - wz * Jy * wy)) / Jx
        wdy = (ty - (wz * Jx * wx - wx * Jz * wz)) / Jy
        wdz = (tz - (wx * Jy * wy - wy * Jx * wx)) / Jz
        return np.array([vx, vy, vz, ax, ay, az, qd[0], qd[1], qd[2], qd[3], wdx, wdy, wdz])


class Rocket(BoxModel):
    """dyn/rocket.py:21-254 (rockETH)."""

    name = "rocket"
    model_id = 2

    P = dict(mass=1.16, grav=9.81, Jxx=0.0021, Jyy=0.1, Jzz=0.1, off=0.42, tau_T=0.06, tau_s=0.10,
             ga=5.0, gb=35.2, gc=33.0, gd=28.0, ge=35.2, hover=11.3796)  # dyn/rocket.py:23-37,173

    def __init__(self):
        self.nx, self.nu, self.nw = 17, 4, 17
        self.dt = 0.05
        x_ub = np.array([10.0] * 3 + [1.0] * 3 + [1.5] * 4 + [2.0] * 3 + [50.0, 2.0, 1.0, 1.0])
        u_ub = np.array([50.0, 2.0, 1.0, 1.0])
        self._set_box(-x_ub, x_ub, -u_ub, u_ub)
        q_vec = 0.5 * np.deg2rad(2.0)
        # dyn/rocket.py:152-162 (default E; the drivers overwrite it)
        self.E = np.diag([0.03] * 3 + [0.08] * 3 + [q_vec] * 3 + [0.1 * q_vec] + [0.10] * 3 + [0.8, 0.2, 0.04, 0.04])

    def gimbal(self, servo, tilt):
        # dyn/rocket.py:246-254
        p = self.P
        iv1 = p["gd"] + p["ga"] * np.cos(servo)
        iv2 = p["ge"] - p["ga"] * np.sin(servo)
        uu = p["gb"] ** 2 - p["gc"] ** 2 - iv1**2 - iv2**2
        vv = 2 * p["gc"] * np.cos(tilt) * iv2
        ww = -2 * p["gc"] * iv1
        iv3 = ww**2 + vv**2 - uu**2
        return 2 * np.arctan((vv - np.sqrt(iv3)) / (uu + ww))

    def ode(self, X, u):
        p = self.P
        vx, vy, vz = X[3], X[4], X[5]
        qw, qx, qy, qz = X[6], X[7], X[8], X[9]
        wx, wy, wz = X[10], X[11], X[12]
        T = X[13] + p["hover"]
        tq, sa1, sa2 = X[14], X[15], X[16]
        g1 = self.gimbal(sa1, 0.0 * sa1)
        g2 = self.gimbal(sa2, g1)
        Fx = -T * np.sin(g1) * np.cos(g2)
        Fy = T * np.sin(g2)
        Fz = T * np.cos(g1) * np.cos(g2)
        R = _quat_rot(qw, qx, qy, qz)
        ax = (R[0][0] * Fx + R[0][1] * Fy + R[0][2] * Fz) / p["mass"]
        ay = (R[1][0] * Fx + R[1][1] * Fy + R[1][2] * Fz) / p["mass"]
        az = (R[2][0] * Fx + R[2][1] * Fy + R[2][2] * Fz) / p["mass"] - p["grav"]
        qd = _quat_kin(qw, qx, qy, qz, wx, wy, wz)
        # torque = [0,0,-off] x F ; J wdot = torque - w x (J w)     (dyn/rocket.py:226-229)
        tx, ty, tz = p["off"] * Fy, -p["off"] * Fx, 0.0
        Jx, Jy, Jz = p["Jxx"], p["Jyy"], p["Jzz"]
        wdx = (tx - (wy * Jz * wz - wz * Jy * wy)) / Jx
        wdy = (ty - (wz * Jx * wx - wx * Jz * wz)) / Jy
        wdz = (tz - (wx * Jy * wy - wy * Jx * wx)) / Jz
        Td = (u[0] + p["hover"] - T) / p["tau_T"]
        tqd = (u[1] - tq) / p["tau_T"]
        s1d = (u[2] - sa1) / p["tau_s"]
        s2d = (u[3] - sa2) / p["tau_s"]
        return np.array([vx, vy, vz, ax, ay, az, qd[0], qd[1], qd[2], qd[3], wdx, wdy, wdz, Td, tqd, s1d, s2d])


ROCKET_X0 = np.array([1.75729, 4.15951, 4.72757, -0.18913, -0.38367, -0.08697, -0.79487, 0.00768, -0.21110,
                      -0.56883, -0.12752, -0.58026, -0.76542, 0.20555, 0.54610, -0.40116, -0.35401])
"""Fixed rocket initial state, expe/main_rocket_robust_closed_loop.py:110-126."""


def rocket_config():
    """Model + costs of expe/main_rocket_robust_closed_loop.py:33-85 / main_rocket_compare_closed_loop.py:41-70,107-114."""
    m = Rocket()
    q_vec = 0.5 * np.deg2rad(2.0)
    m.E = m.dt * np.diag([0.2] * 6 + [q_vec] * 3 + [0.1 * q_vec] + [0.2] * 3 + [0.8, 0.2, 0.04, 0.04])
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [1.0] * 4 + [1.0, 5.0, 5.0] + [1.0] * 4)
    R = np.eye(4)
    cfg = dict(N=15, Q=Q, R=R, Qf=10 * Q, Q_reg=1e4 * np.eye(17), R_reg=1e4 * np.eye(4), Q_reg_f=1e4 * np.eye(17),
               rti=1, fast_sls_rti_steps=1)
    return m, cfg


def quadrotor_config():
    """expe/main_quadrotor_robust_closed_loop.py:35-69."""
    m = Quadrotor()
    q_vec = 0.5 * np.deg2rad(2.0)
    m.E = m.dt * 5 * np.diag([0.10] * 3 + [0.15] * 3 + [0.1 * q_vec] + [q_vec] * 3 + [0.2] * 3)
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [1.0] * 4 + [2.0] * 3)
    cfg = dict(N=15, Q=Q, R=np.eye(4), Qf=10 * Q, Q_reg=1e4 * np.eye(13), R_reg=1e4 * np.eye(4),
               Q_reg_f=1e4 * np.eye(13), rti=3, fast_sls_rti_steps=2)
    return m, cfg


def quadrotor_x0(seed=0):
    """Seeded replacement of the unseeded draw at expe/main_quadrotor_robust_closed_loop.py:82-90."""
    m = Quadrotor()
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(m.x_lb / 2, m.x_ub / 2)
    q = rng.normal(size=4)
    x0[6:10] = q / np.linalg.norm(q)
    return x0


def pendulum_config():
    """expe/main_pendulum_robust_closed_loop.py:24-50."""
    m = Pendulum()
    m.E = 0.003 * np.eye(4)
    cfg = dict(N=15, Q=np.eye(4), R=np.eye(1), Qf=10 * np.eye(4), Q_reg=1e3 * np.eye(4), R_reg=1e3 * np.eye(1),
               Q_reg_f=1e4 * np.eye(4), rti=3, fast_sls_rti_steps=2)
    return m, cfg


PENDULUM_X0 = np.array([0.5, 0.5, 0.0, 0.0])  # expe/main_pendulum_robust_closed_loop.py:58

MODELS = {"pendulum": Pendulum, "quadrotor": Quadrotor, "rocket": Rocket}
CONFIGS = {"pendulum": pendulum_config, "quadrotor": quadrotor_config, "rocket": rocket_config}
