"""`QuadrupedEnv`: the reference's gym.Env surface on top of the batched B200 step.

Mirror of gym_quadruped/quadruped_env.py:71-1432 for the hot path (constructor kwargs :85-99, `step` :251-307,
`reset` :309-406, `ALL_OBS` :35-81, the accessors controllers use :488-1044).  `num_envs=1` (default) reproduces the
reference's return types -- dict of `(dim,)` float64 NumPy arrays, Python float / bool flags, info dict; `num_envs>1`
returns dict-of-`[N, dim]` torch views into one packed `[N, D]` CUDA tensor and tensor flags.

Everything physical happens in libqstep's fused kernel (see csrc/qs_env.cuh); this file only slices tensors.
Viewer / rendering (`render`, ghost robots, key callbacks) are out of scope.
"""
from __future__ import annotations

import copy
import logging
import math
from typing import Any

import numpy as np
import torch

from . import backend
from .backend import BatchSim, command_mode_bits
from .model import QS_NOBS_BASE, Model
from .robot_cfgs import RobotConfig, get_robot_config
from .sensors.base_sensor import Sensor
from .sensors.heightmap import HeightMap
from .sensors.imu import IMU
from .spaces import Box, Env
from .utils.math_utils import _process_range
from .utils.quadruped_utils import LegsAttr, configure_observation_space, extract_joint_info

log = logging.getLogger(__name__)

BASE_OBS = ['base_pos', 'base_lin_vel', 'base_lin_vel_err', 'base_lin_acc', 'base_ang_vel', 'base_ang_vel_err',
            'base_ori_euler_xyz', 'base_ori_quat_wxyz', 'base_ori_SO3', 'gravity_vector:base']
BASE_OBS_BASE_FRAME = ['base_lin_vel:base', 'base_lin_vel_err:base', 'base_lin_acc:base', 'base_ang_vel:base',
                       'base_ang_vel_err:base']
GEN_COORDS_OBS = ['qpos', 'qvel', 'tau_ctrl_setpoint', 'qpos_js', 'qvel_js', 'kinetic_energy', 'work']
FEET_OBS = ['feet_pos', 'feet_pos:base', 'feet_vel', 'feet_vel_rel', 'feet_vel:base', 'feet_vel_rel:base', 'contact_state',
            'contact_forces', 'contact_forces:base']

# packed layout of the kernel's observation row: name -> (offset, dim); SURVEY.md section 8(a)
_OBS_DIMS = [3, 3, 3, 3, 3, 3, 3, 4, 9, 3, 3, 3, 3, 3, 3, 19, 18, 12, 12, 12, 1, 1, 12, 12, 12, 12, 12, 12, 4, 12, 12]
OBS_LAYOUT: dict[str, tuple[int, int]] = {}
_off = 0
for _name, _dim in zip(BASE_OBS + BASE_OBS_BASE_FRAME + GEN_COORDS_OBS + FEET_OBS, _OBS_DIMS):
    OBS_LAYOUT[_name] = (_off, _dim)
    _off += _dim
assert _off == QS_NOBS_BASE
IMU_LAYOUT = {n: (QS_NOBS_BASE + 3 * i, 3) for i, n in enumerate(IMU.ALL_OBS)}
_PER_LEG_OBS = {'feet_pos', 'feet_pos:base', 'feet_vel', 'feet_vel_rel', 'feet_vel:base', 'feet_vel_rel:base', 'contact_forces',
                'contact_forces:base'}
_MODEL_LEGS = ('FL', 'FR', 'RL', 'RR')


class QuadrupedEnv(Env):
    """Batched quadruped environment with the reference's constructor and step / reset surface."""

    _DEFAULT_OBS = ('qpos', 'qvel', 'tau_ctrl_setpoint', 'feet_pos:base', 'feet_vel:base')
    ALL_OBS = BASE_OBS + BASE_OBS_BASE_FRAME + GEN_COORDS_OBS + FEET_OBS
    metadata = {'render.modes': ['human'], 'version': 0}

    def __init__(
        self,
        robot: str,
        state_obs_names: tuple[str, ...] = _DEFAULT_OBS,
        scene: str = 'flat',
        sim_dt: float = 0.002,
        base_vel_command_type: str = 'forward',
        ref_base_lin_vel=0.5,
        ref_base_ang_vel=0.0,
        ground_friction_coeff=1.0,
        legs_order: tuple[str, str, str, str] = ('FL', 'FR', 'RL', 'RR'),
        sensors: tuple | None = None,
        sensors_kwargs: tuple[dict[str, Any]] | None = None,
        external_disturbances_kwargs: dict[str, Any] | None = None,
        *,
        num_envs: int = 1,
        device: str | int | torch.device = 'cuda:0',
        seed: int = 0,
        precision: str = 'fp32',
        env_id_offset: int = 0,
        auto_reset: bool = False,
        pipeline: bool = False,
    ):
        self._init_args = dict(robot=robot, state_obs_names=state_obs_names, scene=scene, sim_dt=sim_dt,
                               base_vel_command_type=base_vel_command_type, ref_base_lin_vel=ref_base_lin_vel,
                               ref_base_ang_vel=ref_base_ang_vel, ground_friction_coeff=ground_friction_coeff, legs_order=legs_order,
                               sensors=sensors, sensors_kwargs=sensors_kwargs, external_disturbances_kwargs=external_disturbances_kwargs)
        self.robot_name = robot
        self.robot_cfg: RobotConfig = get_robot_config(robot_name=robot)
        self.base_vel_command_type = base_vel_command_type
        self._command_mode = command_mode_bits(base_vel_command_type)
        self.base_lin_vel_range = _process_range(ref_base_lin_vel)
        self.base_ang_vel_range = _process_range(ref_base_ang_vel)
        if self.base_lin_vel_range is None or self.base_ang_vel_range is None:
            raise NotImplementedError('callable velocity references are not supported by the batched env')
        self.ground_friction_coeff_range = _process_range(ground_friction_coeff)
        self.legs_order = tuple(legs_order)
        assert sorted(self.legs_order) == sorted(_MODEL_LEGS), f'legs_order must be a permutation of {_MODEL_LEGS}'
        self.is_paused = False
        # batched extension: `step` resets, inside the same kernel launch, every env that terminates (same-step auto-reset: the
        # returned observation of such an env is its post-reset one, `terminated` still flags the episode end)
        self.auto_reset_on_step = bool(auto_reset)
        self.num_envs = int(num_envs)
        self.device = torch.device(device)

        self.model = Model(self.robot_cfg.tables, scene, sim_dt)
        self.terrain_limits = self.model.terrain_limits
        self.joint_info = extract_joint_info(self.model.tables)
        idx = {leg: [7 + 3 * k + j for j in range(3)] for k, leg in enumerate(_MODEL_LEGS)}
        self.legs_qpos_idx = LegsAttr(**idx)
        self.legs_qvel_idx = LegsAttr(**{leg: [i - 1 for i in v] for leg, v in idx.items()})
        self.legs_tau_idx = LegsAttr(**{leg: [i - 7 for i in v] for leg, v in idx.items()})
        self._feet_geom_id = LegsAttr(**{leg: self.model.tables['foot_geom'][k] for k, leg in enumerate(_MODEL_LEGS)})
        self._feet_body_id = LegsAttr(**{leg: 4 + 3 * k for k, leg in enumerate(_MODEL_LEGS)})

        # action space: unbounded, exactly like the reference (its limit logic is always-truthy, quadruped_env.py:219-225)
        self.action_space = Box(low=-np.inf, high=np.inf, shape=(12,), dtype=np.float32)

        # sensors: IMU is fused into the kernel; other Sensor subclasses are stepped on the host after the kernel
        self.sensors: list[Sensor] = []
        use_imu, imu_noise, hm_cfg = False, (0.01, 0.01, 0.01, 0.01), None
        sensors = sensors or ()
        sensors_kwargs = sensors_kwargs or tuple({} for _ in sensors)
        deferred = []
        for cls, kw in zip(sensors, sensors_kwargs):
            if isinstance(cls, type) and issubclass(cls, IMU):
                s = cls(mj_model=self.model, mj_data=self, **kw)
                use_imu, imu_noise = True, s.noise
                self.sensors.append(s)
            elif isinstance(cls, type) and issubclass(cls, HeightMap):
                # fused into the step kernel: rows*cols*3 extra observation columns named 'heightmap'
                hm_cfg = (int(kw['num_rows']), int(kw['num_cols']), float(kw['dist_x']), float(kw['dist_y']))
            else:
                deferred.append((cls, kw))

        hm_dim = 0 if hm_cfg is None else hm_cfg[0] * hm_cfg[1] * 3
        for name in state_obs_names:
            if name not in OBS_LAYOUT and not (use_imu and name in IMU_LAYOUT) and not (hm_cfg is not None and name == 'heightmap'):
                raise ValueError(f'Invalid observation name: {name}, available obs: {self.ALL_OBS}')
        self.observation_space = configure_observation_space(self.model.tables, [n for n in state_obs_names if n != 'heightmap'])
        if 'heightmap' in state_obs_names:
            self.observation_space.spaces['heightmap'] = Box(low=-np.inf, high=np.inf, shape=(hm_dim,), dtype=np.float32)
        self.state_obs_names = state_obs_names

        # pipeline=True lets back-to-back `step` launches overlap on the device (BatchSim / QsConfig.pipeline); only meaningful
        # when the actions do not depend on the previous observation (open-loop rollouts, replay)
        self.sim = BatchSim(self.model, self.num_envs, device=self.device, precision=0 if precision == 'fp32' else 1,
                            use_imu=use_imu, imu_noise=imu_noise, seed=seed, env_id_offset=env_id_offset, heightmap=hm_cfg,
                            pipeline=pipeline)
        for cls, kw in deferred:  # host-side plug-ins following the Sensor protocol (base_sensor.py:4-41)
            self.sensors.append(cls(mj_model=self.model, mj_data=self, **kw))
        self._layout = dict(OBS_LAYOUT)
        if use_imu:
            self._layout.update(IMU_LAYOUT)
        if hm_cfg is not None:
            self._layout['heightmap'] = (QS_NOBS_BASE + (18 if use_imu else 0), hm_dim)
        perm = [_MODEL_LEGS.index(leg) for leg in self.legs_order]
        self._leg_perm = None if perm == [0, 1, 2, 3] else torch.tensor(
            [3 * p + i for p in perm for i in range(3)], device=self.device, dtype=torch.long)
        self._leg_perm_np = None if self._leg_perm is None else np.array([3 * p + i for p in perm for i in range(3)])
        self._time_host = 0.0

        # In-episode schedules (:293-305): '+reset' command resampling and the external base wrench run inside the step kernel
        # (per-env counters in sim.cmd_count / cmd_limit / ext_count / ext_limit, wrench in sim.ext_wrench): no host round trip.
        # As in the reference the wrench is only ever applied for type == 'reset' (:299-305); it is drawn here, at construction (:240-242).
        self.external_disturbances_kwargs = external_disturbances_kwargs
        ext_on = external_disturbances_kwargs is not None and external_disturbances_kwargs.get('type') == 'reset'
        self.sim.set_schedule(command_mode=self._command_mode, lin_vel_range=self.base_lin_vel_range, ang_vel_range=self.base_ang_vel_range,
                              ext_ranges=external_disturbances_kwargs if external_disturbances_kwargs is not None else None,
                              ext_enabled=ext_on)
        self.sim.reset_options = self._reset_options(True, {})  # used by auto-reset before the first explicit reset()
        self._warned_status = 0
        self.viewer = None
        self.step_num = 0
        self._last_obs_tensor = self.sim.obs
        # single-env fast path: the step goes through qs_step_host with pinned one-row buffers (one launch, one synchronisation;
        # the kernel reads ctrl and writes the observation row / flags straight through the mapped host memory)
        self._host = None
        if self.num_envs == 1:
            pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()
            self._host = dict(ctrl=pin(1, 12), obs=pin(1, self.sim.obs_dim), rew=pin(1), term=pin(1, dtype=torch.uint8),
                              trunc=pin(1, dtype=torch.uint8))
            self._host['obs_np'] = self._host['obs'].numpy()
            self._host['ctrl_np'] = self._host['ctrl'].numpy()

    # ------------------------------------------------------------------ gym API
    def step(self, action):
        """Apply joint torques, advance one sim step, return (obs, reward, terminated, truncated, info); :251-307."""
        if self.num_envs == 1:
            return self._step_single(action)
        if self.auto_reset_on_step:
            obs_t, rew, term, trunc = self.sim.step_autoreset(action)
        else:
            obs_t, rew, term, trunc = self.sim.step(action)
        for s in self.sensors:
            s.step()
        info_invalid = self.sim.invalid_body_mask
        self.step_num += 1
        obs = self._obs_dict(obs_t)
        # `status` (per env): bit0 non-finite state, bit1 contact buffer overflow, bit2 solver iteration cap, bit3 reset could not clear contact
        info = {'time': self.sim.sim_time, 'step_num': self.step_num - 1, 'invalid_contacts': info_invalid, 'status': self.sim.status}
        return obs, rew, term.bool(), trunc.bool(), info

    def _step_single(self, action):
        """num_envs == 1: reference return types (dict of float64 arrays, Python flags) with one launch and one synchronisation."""
        h = self._host
        if isinstance(action, torch.Tensor):
            h['ctrl'].copy_(action.detach().reshape(1, 12).to('cpu', torch.float32))
        else:
            h['ctrl_np'][0, :] = np.asarray(action, dtype=np.float32).reshape(12)
        self.sim.step_host(h['ctrl'], h['obs'], h['rew'], h['term'], h['trunc'],  # returns after the stream is synchronised
                           auto_reset=self.sim.reset_options if self.auto_reset_on_step else None)
        self.sim.obs.copy_(h['obs'], non_blocking=True)  # keeps the device-side row (frame conversions of the accessors) current
        for s in self.sensors:
            s.step()
        self.step_num += 1
        obs = self._obs_dict(self.sim.obs, host_row=h['obs_np'][0])
        terminated = bool(h['term'][0])
        invalid = {}
        if terminated:  # a contact can only be invalid if the episode terminated (quadruped_env.py:283-285)
            m = self.sim.invalid_body_mask[0].tolist()
            mask = int(m[0]) | (int(m[1]) << 8)
            names = self.model.tables['body_names']
            invalid = {f'world:0_{names[b]}:{b}': None for b in range(1, 14) if mask >> b & 1}
        self._time_host += self.simulation_dt
        status = int(self.sim.status[0].item()) if terminated else 0  # read only when something happened: no extra sync per step
        if status & ~self._warned_status & 7:
            self._warned_status |= status
            log.warning('libqstep status 0x%x (bit0 non-finite state, bit1 contact buffer overflow, bit2 solver iteration cap)', status)
        info = {'time': self._time_host, 'step_num': self.step_num - 1, 'invalid_contacts': invalid, 'status': status}
        return obs, 0, terminated, False, info

    def reset(self, qpos=None, qvel=None, seed: int | None = None, random: bool = True, options: dict[str, Any] | None = None,
              env_mask: torch.Tensor | None = None):
        """Reset (all envs, or those selected by `env_mask`) and return the observation dict; :309-406."""
        options = {} if options is None else options
        self.step_num = 0
        if seed is not None:
            self.sim.set_seed(seed)  # np.random.seed(seed) (:337-338): new key, draw counters restarted, no buffer touched
        opt = self._reset_options(random, options)
        self.sim.reset_options = opt
        mask = None if env_mask is None else env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if qpos is not None or qvel is not None:
            assert qpos is not None and qvel is not None, 'qpos and qvel must be given together'
            q = torch.as_tensor(np.asarray(qpos) if not isinstance(qpos, torch.Tensor) else qpos, dtype=torch.float32, device=self.device)
            v = torch.as_tensor(np.asarray(qvel) if not isinstance(qvel, torch.Tensor) else qvel, dtype=torch.float32, device=self.device)
            obs_t = self.sim.reset(mask, q.reshape(-1, 19).expand(self.num_envs, 19).contiguous(),
                                   v.reshape(-1, 18).expand(self.num_envs, 18).contiguous(), opt)
        else:
            obs_t = self.sim.reset(mask, None, None, opt)
        if self.num_envs == 1 and (int(self.sim.status[0].item()) & 8):
            raise RuntimeError('Unable to initialize the robot without ground contact.')
        self._time_host = self.simulation_dt  # the reset ends with one step from time 0 (:333,:397)
        return self._obs_dict(obs_t)

    def _reset_options(self, random: bool, options: dict):
        return self.sim.make_reset_options(
            randomize=random, angle_sweep=options.get('angle_sweep', 20 * math.pi / 180),
            roll_sweep=options.get('roll_sweep', 10 * math.pi / 180), pitch_sweep=options.get('pitch_sweep', 10 * math.pi / 180),
            lin_vel_range=self.base_lin_vel_range, ang_vel_range=self.base_ang_vel_range,
            friction_range=self.ground_friction_coeff_range, command_mode=self._command_mode)

    def auto_reset(self):
        """Batched convenience: reset every env whose last `terminated` flag is set (one masked kernel launch)."""
        return self._obs_dict(self.sim.reset_done())

    def render(self, *args, **kwargs):
        raise NotImplementedError('rendering / viewer decorations are out of scope for the batched B200 env')

    def close(self):
        if getattr(self, 'sim', None) is not None:
            self.sim.close()

    # ------------------------------------------------------------------ observation plumbing
    def _obs_dict(self, obs_t: torch.Tensor, host_row=None):
        self._last_obs_tensor = obs_t
        if self.num_envs == 1:  # the reference's types: a dict of fresh float64 arrays (:1156-1204)
            flat = (obs_t[0].detach().cpu().numpy() if host_row is None else host_row).astype(np.float64)
            res = {}
            for name in self.state_obs_names:
                off, dim = self._layout[name]
                a = flat[off:off + dim]
                res[name] = a[self._leg_perm_np] if (self._leg_perm is not None and name in _PER_LEG_OBS) else a.copy()
            return res
        out = {}
        for name in self.state_obs_names:
            off, dim = self._layout[name]
            v = obs_t[:, off:off + dim]
            if self._leg_perm is not None and name in _PER_LEG_OBS:
                v = v.index_select(1, self._leg_perm)
            out[name] = v
        return out

    def _sensor_obs(self, name):
        off, dim = self._layout[name]
        v = self._last_obs_tensor[:, off:off + dim]
        return v[0].detach().cpu().numpy().astype(np.float64) if self.num_envs == 1 else v

    def _np(self, t: torch.Tensor):
        return t[0].detach().cpu().numpy().astype(np.float64) if self.num_envs == 1 else t

    # ------------------------------------------------------------------ state tensors (write-then-step semantics of mjData)
    @property
    def qpos(self) -> torch.Tensor:
        """[N, 19] fp32 view of the generalized positions.  Joint angles / orientation may be edited in place (write-then-step, like
        `env.mjData.qpos[...] = ...`); the BASE POSITION is mastered in fp64 (`sim.base_pos64`, resets scatter envs over +-10 km) and
        `qpos[:, :3]` is only its fp32 image, refreshed every step -- move the base with `set_state`."""
        return self.sim.qpos

    @property
    def qvel(self) -> torch.Tensor:
        return self.sim.qvel

    def set_state(self, qpos, qvel, env_ids=None):
        """`env.mjData.qpos[...] = ...` equivalent (examples/aliengo_with_heightmap.py:32-34).  The packed observation row (and the
        accessors that slice it: base velocities, feet positions, `frame='base'` conversions) still describes the state before the
        write until the next `step`; the `qs_forward`-based accessors (Jacobians, mass matrix, com, hip positions) see it at once."""
        self.sim.set_state(torch.as_tensor(qpos), torch.as_tensor(qvel), env_ids)

    # ------------------------------------------------------------------ accessors (quadruped_env.py:488-1044)
    def _slice(self, name):
        off, dim = OBS_LAYOUT[name]
        return self._np(self.sim.obs[:, off:off + dim])

    def _frame_name(self, base, frame):
        if frame == 'world':
            return base
        if frame == 'base':
            return base + ':base'
        raise ValueError(f"Invalid frame: {frame} != 'world' or 'base'")

    def _legs(self, arr, width=3):
        a = arr.reshape(*arr.shape[:-1], 4, width) if self.num_envs > 1 else arr.reshape(4, width)
        pick = (lambda k: a[..., k, :]) if self.num_envs > 1 else (lambda k: a[k])
        return LegsAttr(**{leg: pick(k) for k, leg in enumerate(_MODEL_LEGS)})

    def target_base_vel(self, frame='world'):
        cmd = self.sim.command
        yaw = self.sim.obs[:, 20]
        c, s = torch.cos(yaw), torch.sin(yaw)
        lin = torch.stack([c * cmd[:, 0] - s * cmd[:, 1], s * cmd[:, 0] + c * cmd[:, 1], cmd[:, 2]], dim=1)
        ang = torch.stack([torch.zeros_like(yaw), torch.zeros_like(yaw), cmd[:, 3]], dim=1)
        if frame == 'base':
            R = self.sim.obs[:, 25:34].reshape(-1, 3, 3)
            lin, ang = torch.einsum('nji,nj->ni', R, lin), torch.einsum('nji,nj->ni', R, ang)
        elif frame != 'world':
            raise ValueError(f"Invalid frame: {frame} != 'world' or 'base'")
        return self._np(lin), self._np(ang)

    def base_lin_vel(self, frame='world'):
        return self._slice(self._frame_name('base_lin_vel', frame))

    def base_lin_vel_err(self, frame='world'):
        return self._slice(self._frame_name('base_lin_vel_err', frame))

    def base_ang_vel(self, frame='world'):
        return self._slice(self._frame_name('base_ang_vel', frame))

    def base_ang_vel_err(self, frame='world'):
        return self._slice(self._frame_name('base_ang_vel_err', frame))

    def base_lin_acc(self, frame='world'):
        return self._slice(self._frame_name('base_lin_acc', frame))

    def feet_pos(self, frame='world') -> LegsAttr:
        return self._legs(self._slice(self._frame_name('feet_pos', frame)))

    def feet_vel(self, frame='world', relative=False) -> LegsAttr:
        return self._legs(self._slice(self._frame_name('feet_vel_rel' if relative else 'feet_vel', frame)))

    def feet_contact_state(self, frame='world', ground_reaction_forces=False):
        cs = self._slice('contact_state')
        state = LegsAttr(**{leg: (bool(cs[k]) if self.num_envs == 1 else cs[:, k] > 0.5) for k, leg in enumerate(_MODEL_LEGS)})
        if not ground_reaction_forces:
            return state, None
        return state, None, self._legs(self._slice(self._frame_name('contact_forces', frame)))

    def _tables(self, field):
        self.sim.forward()
        return self.sim.get(field)

    def _jac_tables(self, fields, frame):
        self.sim.forward()
        if frame == 'base':
            R = self.sim.obs[:, 25:34].reshape(-1, 3, 3)
        elif frame != 'world':
            raise ValueError(f"Invalid frame: {frame} != 'world' or 'base'")
        out = []
        for f in fields:
            J = self.sim.get(f)
            if frame == 'base':
                J = torch.einsum('nji,nljd->nlid', R, J)
            out.append(LegsAttr(**{leg: self._np(J[:, k]) for k, leg in enumerate(_MODEL_LEGS)}))
        return out

    def feet_jacobians(self, frame='world', return_rot_jac=False):
        """Foot Jacobians (per leg 3 x nv, batched [N, 3, nv]) at the current state: translational, and with
        `return_rot_jac` also the rotational one of the calf body (mj_jac, :681-740)."""
        fields = (backend.FIELD_FEET_JACP, backend.FIELD_FEET_JACR) if return_rot_jac else (backend.FIELD_FEET_JACP,)
        res = self._jac_tables(fields, frame)
        return tuple(res) if return_rot_jac else res[0]

    def feet_jacobians_dot(self, frame='world', return_rot_jac=False):
        """Time derivative of the foot Jacobians at the current (qpos, qvel) (mj_jacDot, :742-797)."""
        fields = (backend.FIELD_FEET_JACP_DOT, backend.FIELD_FEET_JACR_DOT) if return_rot_jac else (backend.FIELD_FEET_JACP_DOT,)
        res = self._jac_tables(fields, frame)
        return tuple(res) if return_rot_jac else res[0]

    @property
    def mass_matrix(self):
        return self._np(self._tables(backend.FIELD_MASS_MATRIX))

    @property
    def legs_mass_matrix(self):
        M = self._tables(backend.FIELD_MASS_MATRIX)
        out = {}
        for k, leg in enumerate(_MODEL_LEGS):
            i = 6 + 3 * k
            out[leg] = self._np(M[:, i:i + 3, i:i + 3])
        return LegsAttr(**out)

    def get_base_inertia(self):
        return self._np(self._tables(backend.FIELD_MASS_MATRIX)[:, 3:6, 3:6])

    @property
    def legs_qfrc_bias(self):
        b = self._tables(backend.FIELD_QFRC_BIAS)
        return LegsAttr(**{leg: self._np(b[:, 6 + 3 * k:9 + 3 * k]) for k, leg in enumerate(_MODEL_LEGS)})

    @property
    def legs_qfrc_passive(self):
        b = self._tables(backend.FIELD_QFRC_PASSIVE)
        return LegsAttr(**{leg: self._np(b[:, 6 + 3 * k:9 + 3 * k]) for k, leg in enumerate(_MODEL_LEGS)})

    @property
    def com(self):
        return self._np(self._tables(backend.FIELD_COM))

    def hip_positions(self, frame='world') -> LegsAttr:
        x = self._tables(backend.FIELD_XPOS)  # bodies 1..13
        if frame == 'base':  # the reference applies R^T without subtracting the base position (:581-595)
            R = self.sim.obs[:, 25:34].reshape(-1, 3, 3)
            x = torch.einsum('nji,nbj->nbi', R, x)
        elif frame != 'world':
            raise ValueError(f"Invalid frame: {frame} != 'world' or 'base'")
        return LegsAttr(**{leg: self._np(x[:, 1 + 3 * k]) for k, leg in enumerate(_MODEL_LEGS)})

    @property
    def base_configuration(self):
        R = self.sim.obs[:, 25:34].reshape(-1, 3, 3)
        X = torch.eye(4, device=self.device).repeat(self.num_envs, 1, 1)
        X[:, :3, :3] = R
        X[:, :3, 3] = self.sim.base_pos64.to(torch.float32)
        return self._np(X)

    @property
    def joint_space_state(self):
        return self._np(self.sim.qpos[:, 7:]), self._np(self.sim.qvel[:, 6:])

    @property
    def base_pos(self):
        return self._np(self.sim.base_pos64)

    @property
    def base_ori_euler_xyz(self):
        return self._slice('base_ori_euler_xyz')

    @property
    def heading_orientation_SO3(self):
        yaw = self.sim.obs[:, 20]
        c, s, z, o = torch.cos(yaw), torch.sin(yaw), torch.zeros_like(yaw), torch.ones_like(yaw)
        return self._np(torch.stack([c, -s, z, s, c, z, z, z, o], dim=1).reshape(-1, 3, 3))

    @property
    def torque_ctrl_setpoint(self):
        return self._slice('tau_ctrl_setpoint')

    @property
    def gravity_vector(self):
        return self._slice('gravity_vector:base')

    @property
    def kinetic_energy(self):
        return self._slice('kinetic_energy')

    @property
    def work(self):
        return self._slice('work')

    @property
    def simulation_dt(self):
        return self.model.c.timestep

    @property
    def simulation_time(self):
        return float(self.sim.sim_time[0].item()) if self.num_envs == 1 else self.sim.sim_time

    def get_hyperparameters(self):
        return copy.copy(self._init_args)

    def __str__(self):
        msg = f'robot={self._init_args["robot"]} terrain={self._init_args["scene"]} task={self.base_vel_command_type} num_envs={self.num_envs}'
        if self.base_vel_command_type != 'human':
            msg += (f' lin_vel_range=({self.base_lin_vel_range[0]:.3f}, {self.base_lin_vel_range[1]:.3f})'
                    f' ang_vel_range=({self.base_ang_vel_range[0]:.3f}, {self.base_ang_vel_range[1]:.3f})'
                    f' lat_friction_range=({self.ground_friction_coeff_range[0]:.1e}, {self.ground_friction_coeff_range[1]:.1e})')
        return msg
