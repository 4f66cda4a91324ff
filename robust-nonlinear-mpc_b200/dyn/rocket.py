"""rockETH rocket model, mirror of dyn/rocket.py:21-163 (ODE + gimbal map :165-254 in csrc/rnmpc_models.cuh)."""
import numpy as np

from .model import Model


class Rocket(Model):
    model_name = "rocket"

    def __init__(self):
        self.params = {
            'mass': 1.16, 'gravity_constant': 9.81, 'inertia_xx': 0.00210, 'inertia_yy': 0.10000, 'inertia_zz': 0.10000,
            'thrust_cog_offset': 0.42000, 'thrust_magnitude_time_constant': 0.06000, 'servo_angle_time_constant': 0.10,
            'gimbal_a': 5.0, 'gimbal_b': 35.2, 'gimbal_c': 33.0, 'gimbal_d': 28.0, 'gimbal_e': 35.2,
        }
        self.state_names = ['x', 'y', 'z', 'v_x', 'v_y', 'v_z', 'quat_x', 'quat_y', 'quat_z', 'quat_w',
                            'angular_vx', 'angular_vy', 'angular_vz', 'thrust_magnitude', 'torque_x', 'servo_angle_1',
                            'servo_angle_2']
        self.control_names = ['thrust_magnitude_u', 'torque_u', 'servo_angle_1_u', 'servo_angle_2_u']
        self.state_index = {n: i for i, n in enumerate(self.state_names)}
        self.control_index = {n: i for i, n in enumerate(self.control_names)}
        self.state_groups = {'pos': slice(0, 3), 'vel': slice(3, 6), 'quat': slice(6, 10), 'omega': slice(10, 13),
                             'act': slice(13, 17)}
        self.neutral_state = np.array([0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], float)
        self.nx, self.nu, self.nw = 17, 4, 17
        self.dt = 0.05
        x_ub = np.array([10.0] * 3 + [1.0] * 3 + [1.5] * 4 + [2.0] * 3 + [50.0, 2.0, 1.0, 1.0])
        u_ub = np.array([50.0, 2.0, 1.0, 1.0])
        self._set_box(-x_ub, x_ub, -u_ub, u_ub)
        cv = self.state_names + self.control_names
        self.constraint_names = [f"{n}_ub" for n in cv] + [f"{n}_lb" for n in cv]
        self.constraint_names_f = [f"{n}_ub" for n in self.state_names] + [f"{n}_lb" for n in self.state_names]
        q_vec_std = 0.5 * np.deg2rad(2.0)
        self.E = np.diag([0.03] * 3 + [0.08] * 3 + [q_vec_std] * 3 + [0.1 * q_vec_std] + [0.10] * 3 + [0.8, 0.2, 0.04, 0.04])
