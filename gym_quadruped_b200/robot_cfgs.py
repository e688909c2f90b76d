"""Robot name -> model tables and per-robot constants (mirror of gym_quadruped/robot_cfgs.py:8-60).

Only the robots BASELINE.json's configs name are compiled so far (mini_cheetah, aliengo, go2, hyqreal1); the substring /
exact matching rules of the reference are kept, including its quirk that plain "hyqreal" is rejected (:49-58).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Optional

import numpy as np

_LEGS = ('FL', 'FR', 'RL', 'RR')


def _default_leg_joints():
    return {leg: [f'{leg}_hip_joint', f'{leg}_thigh_joint', f'{leg}_calf_joint'] for leg in _LEGS}


@dataclass
class RobotConfig:
    mjcf_filename: str
    hip_height: float
    qpos0_js: Optional[Iterable] = None
    feet_geom_names: dict = field(default_factory=lambda: {leg: leg for leg in _LEGS})
    leg_joints: dict = field(default_factory=_default_leg_joints)
    accel_name: Optional[str] = None
    gyro_name: Optional[str] = None
    imu_site_name: Optional[str] = None
    tables: str = ''  # name of the compiled table set under gym_quadruped_b200/assets


# pegasus: listed by the reference's robot_cfgs.py but its MJCF is not in the checkout
_NOT_BUILT = {'pegasus': 0.5}


def get_robot_config(robot_name: str) -> RobotConfig:
    name = robot_name.lower()
    if 'mini_cheetah' in name:
        return RobotConfig('mini_cheetah/mini_cheetah.xml', 0.225, qpos0_js=[0, -np.pi / 2, 0] * 2 + [0, np.pi / 2, 0] * 2,
                           tables='mini_cheetah')
    if name == 'go1':
        return RobotConfig('go1/go1.xml', 0.3, tables='go1')
    if name == 'go2':
        return RobotConfig('go2/go2.xml', 0.28, tables='go2')
    if name == 'aliengo':
        return RobotConfig('aliengo/aliengo.xml', 0.35, tables='aliengo')
    if 'hyqreal1' in name:
        return RobotConfig('hyqreal1/hyqreal1.xml', 0.498, tables='hyqreal1')
    if 'hyqreal2' in name:
        return RobotConfig('hyqreal2/hyqreal2.xml', 0.498, tables='hyqreal2')
    if name == 'b2':
        return RobotConfig('b2/b2.xml', 0.485, tables='b2')
    if 'spot' in name:
        return RobotConfig('spot/spot.xml', 0.46, tables='spot')
    for key in _NOT_BUILT:
        if name == key:
            raise NotImplementedError(f'robot {robot_name!r} is known to the reference but its tables are not compiled yet '
                                      f'(SURVEY.md section 8f, rank 4)')
    raise ValueError(f'Unknown robot name: {robot_name}')
