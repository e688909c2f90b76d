"""IMU sensor (mirror of gym_quadruped/sensors/imu.py:20-139).

When `QuadrupedEnv(sensors=(IMU,), sensors_kwargs=(...,))` is used, the accelerometer / gyro truth signals, the white
noise and the random-walk biases are produced inside the fused step kernel (counter-based Philox instead of the global
`np.random` stream -- statistically, not bit-wise, equivalent); this class only exposes the six observation names and
reads them out of the env's packed observation tensor.
"""
from __future__ import annotations

from .base_sensor import Sensor

LIN_ACC_OBS = ('imu_acc', 'imu_acc_noise', 'imu_acc_bias')
GYRO_OBS = ('imu_gyro', 'imu_gyro_noise', 'imu_gyro_bias')


class IMU(Sensor):
    ALL_OBS = LIN_ACC_OBS + GYRO_OBS
    fused_in_kernel = True

    def __init__(self, mj_model=None, mj_data=None, accel_name=None, gyro_name=None, imu_site_name=None,
                 accel_noise: float = 0.01, gyro_noise: float = 0.01, accel_bias_rate: float = 0.01, gyro_bias_rate: float = 0.01,
                 env=None):
        super().__init__(mj_model, mj_data)
        self._accel_name, self._gyro_name, self._site = accel_name, gyro_name, imu_site_name
        self.noise = (accel_noise, gyro_noise, accel_bias_rate, gyro_bias_rate)
        self._env = env if env is not None else mj_data

    def step(self):
        """No host work: the kernel already advanced bias and noise for this step."""

    def get_observation(self, obs_name):
        if obs_name not in self.ALL_OBS:
            raise ValueError(f'Invalid observation name {obs_name}')
        return self._env._sensor_obs(obs_name)

    @staticmethod
    def available_observations():
        return IMU.ALL_OBS

    @property
    def linear_acceleration(self):
        return tuple(self.get_observation(n) for n in LIN_ACC_OBS)

    @property
    def angular_velocity(self):
        return tuple(self.get_observation(n) for n in GYRO_OBS)
