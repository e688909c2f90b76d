"""Leg containers and the observation-space builder (mirror of gym_quadruped/utils/quadruped_utils.py:17-325)."""
from __future__ import annotations

import operator
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

from ..spaces import Box, Dict


@dataclass
class LegsAttr:
    """Per-leg attribute container with fields FR, FL, RR, RL and default iteration order FL, FR, RL, RR
    (quadruped_utils.py:17-129)."""

    FR: Any
    FL: Any
    RR: Any
    RL: Any

    order = ['FL', 'FR', 'RL', 'RR']

    def to_list(self, order=None):
        return [getattr(self, leg) for leg in (order if order is not None else self.order)]

    def __getitem__(self, key):
        assert key in self.order, f'Key {key} is not a valid leg label. Expected any of {self.order}'
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __iter__(self):
        return iter(self.to_list())

    def _zip(self, other, op, symbol, scalar_types=None):
        if isinstance(other, LegsAttr):
            return LegsAttr(**{leg: op(getattr(self, leg), getattr(other, leg)) for leg in ('FR', 'FL', 'RR', 'RL')})
        if isinstance(other, scalar_types or type(self.FR)):
            return LegsAttr(**{leg: op(getattr(self, leg), other) for leg in ('FR', 'FL', 'RR', 'RL')})
        raise TypeError(f"Unsupported operand type for {symbol}: 'LegsAttr' and '{type(other)}'")

    def __add__(self, other):
        return self._zip(other, operator.add, '+')

    def __sub__(self, other):
        return self._zip(other, operator.sub, '-')

    def __truediv__(self, other):
        if isinstance(other, LegsAttr):
            raise TypeError("Unsupported operand type for /: 'LegsAttr' and 'LegsAttr'")
        return self._zip(other, operator.truediv, '/', (type(self.FR), int, float))

    def __matmul__(self, other):
        return self._zip(other, operator.matmul, '@')

    def __str__(self):
        return ', '.join(f'{leg}={getattr(self, leg)}' for leg in self.order)

    __repr__ = __str__


@dataclass
class JointInfo:
    """Joint-space bookkeeping record (quadruped_utils.py:133-162)."""

    name: str
    type: int
    body_id: int
    nq: int
    nv: int
    qpos_idx: tuple
    qvel_idx: tuple
    range: list
    tau_idx: tuple = field(default_factory=tuple)
    actuator_id: int = -1


def extract_joint_info(tables: dict) -> 'OrderedDict[str, JointInfo]':
    """Joint name -> JointInfo from the compiled tables (replaces extract_mj_joint_info, quadruped_utils.py:165-232)."""
    info = OrderedDict()
    info['root'] = JointInfo('root', 0, 1, 7, 6, tuple(range(7)), tuple(range(6)), [0.0, 0.0])
    for j, name in enumerate(tables['joint_names']):
        info[name] = JointInfo(name, 3, j + 2, 1, 1, (7 + j,), (6 + j,), list(tables['jnt_range'][j]), tau_idx=(j,), actuator_id=j)
    return info


# name -> dimension rules of configure_observation_space (quadruped_utils.py:253-311), in its matching order
def obs_dim(name: str, nq=19, nv=18, nu=12) -> int:
    if name == 'qpos':
        return nq
    if name == 'qvel':
        return nv
    if name == 'tau_ctrl_setpoint':
        return nu
    if name == 'qpos_js':
        return nq - 7
    if name == 'qvel_js':
        return nv - 6
    if name == 'base_pos' or any(k in name for k in ('base_lin_vel', 'base_lin_acc', 'base_ang_vel', 'base_ori_euler_xyz')):
        return 3
    if name == 'base_ori_quat_wxyz':
        return 4
    if name == 'base_ori_SO3':
        return 9
    if 'feet_pos' in name or 'feet_vel' in name:
        return 12
    if name == 'contact_state':
        return 4
    if 'contact_forces' in name:
        return 12
    if 'gravity_vector' in name or 'imu' in name:
        return 3
    if name in ('work', 'kinetic_energy'):
        return 1
    raise ValueError(f'Invalid observation name: {name}')


def configure_observation_space(tables: dict, obs_names: Sequence[str]):
    """gym Dict space with one float32 Box per observation name; joint-limit bounds where the reference sets them."""
    jr = np.asarray(tables['jnt_range'], dtype=np.float64)
    lim = np.asarray(tables['jnt_limited'], dtype=bool)
    # MuJoCo reports range (0, 0) for unlimited joints; the reference copies jnt_range verbatim (:249-262)
    lo = np.where(lim, jr[:, 0], 0.0)
    hi = np.where(lim, jr[:, 1], 0.0)
    cr = np.asarray(tables['act_ctrlrange'], dtype=np.float64)
    spaces = OrderedDict()
    for name in obs_names:
        d = obs_dim(name)
        low, high = np.full(d, -np.inf), np.full(d, np.inf)
        if name == 'qpos':
            low[7:], high[7:] = lo, hi
        elif name == 'qpos_js':
            low, high = lo.copy(), hi.copy()
        elif name == 'tau_ctrl_setpoint':
            low, high = cr[:, 0].copy(), cr[:, 1].copy()
        elif name == 'contact_state':
            low, high = np.zeros(d), np.ones(d)
        spaces[name] = Box(low=low, high=high, shape=(d,), dtype=np.float32)
    return Dict(spaces)
