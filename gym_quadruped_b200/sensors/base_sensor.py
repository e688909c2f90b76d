"""Sensor plug-in protocol (mirror of gym_quadruped/sensors/base_sensor.py:4-41)."""
from __future__ import annotations

import numpy as np


class Sensor:
    """Base class of env sensors: constructed as `cls(mj_model=..., mj_data=..., **kwargs)`, stepped once per env step."""

    def __init__(self, mj_model=None, mj_data=None, **kwargs):
        self._mj_model = mj_model
        self._mj_data = mj_data

    def step(self, **kwargs) -> None:
        raise NotImplementedError

    def get_observation(self, obs_name: str) -> np.ndarray:
        raise NotImplementedError

    @staticmethod
    def available_observations() -> list:
        raise NotImplementedError
