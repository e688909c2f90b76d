"""Thin host layer over libqstep (include/qstep.h): loads the C-ABI with ctypes, owns the per-env device buffers as
torch tensors and launches the fused kernels on torch's current CUDA stream.

There is deliberately no CPU path here: if the CUDA extension is missing or no GPU is visible this module raises.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np
import torch

from .model import (QS_CONTACT_STRIDE, QS_NOBS_BASE, QS_NOBS_IMU, Model, QsBuffers, QsConfig, QsModel, QsResetOptions, QsSchedule)

import os

CSRC = Path(__file__).resolve().parent / 'csrc'
# QSTEP_LIB selects an alternative build of the same library (diagnostic builds such as -DQS_PROF); never a CPU path
LIB_PATH = Path(os.environ['QSTEP_LIB']).resolve() if os.environ.get('QSTEP_LIB') else CSRC / 'libqstep.so'

FIELD_MASS_MATRIX, FIELD_QFRC_BIAS, FIELD_QFRC_PASSIVE, FIELD_FEET_JACP, FIELD_FEET_POS, FIELD_COM, FIELD_CONTACTS, \
    FIELD_QFRC_SMOOTH, FIELD_QFRC_CONSTRAINT, FIELD_XPOS, FIELD_SENSOR_IMU, FIELD_FEET_JACR, FIELD_FEET_JACP_DOT, \
    FIELD_FEET_JACR_DOT = range(14)
_FIELD_SHAPE = {
    FIELD_MASS_MATRIX: (18, 18), FIELD_QFRC_BIAS: (18,), FIELD_QFRC_PASSIVE: (18,), FIELD_FEET_JACP: (4, 3, 18),
    FIELD_FEET_POS: (4, 3), FIELD_COM: (3,), FIELD_QFRC_SMOOTH: (18,), FIELD_QFRC_CONSTRAINT: (18,), FIELD_XPOS: (13, 3),
    FIELD_SENSOR_IMU: (6,), FIELD_FEET_JACR: (4, 3, 18), FIELD_FEET_JACP_DOT: (4, 3, 18), FIELD_FEET_JACR_DOT: (4, 3, 18),
}

CMD_FORWARD, CMD_RANDOM, CMD_ROTATE, CMD_RESET = 1, 2, 4, 8

_lib = None


def load_library() -> C.CDLL:
    """Load libqstep.so (built in-tree by `__graft_entry__.build()`); fail loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f'{LIB_PATH} not found: the CUDA extension has not been built. Run '
                           f'`python -c "import __graft_entry__ as g; g.build()"` (needs nvcc). There is no CPU fallback.')
    L = C.CDLL(str(LIB_PATH))
    vp, fp, u8p = C.c_void_p, C.c_void_p, C.c_void_p
    L.qs_abi_version.restype = C.c_int
    for name in ('qs_model_sizeof', 'qs_config_sizeof', 'qs_buffers_sizeof'):
        getattr(L, name).restype = C.c_int
    L.qs_obs_dim.argtypes = [C.POINTER(QsConfig)]
    L.qs_create.argtypes = [C.POINTER(QsModel), C.POINTER(QsConfig), C.POINTER(vp)]
    L.qs_destroy.argtypes = [vp]
    L.qs_last_error.argtypes = [vp]
    L.qs_last_error.restype = C.c_char_p
    L.qs_bind.argtypes = [vp, C.POINTER(QsBuffers)]
    L.qs_set_schedule.argtypes = [vp, C.POINTER(QsSchedule), vp]
    L.qs_set_seed.argtypes = [vp, C.c_uint64, vp]
    L.qs_step_variant.argtypes = [vp]
    L.qs_step_variant.restype = C.c_char_p
    L.qs_step.argtypes = [vp, fp, fp, fp, u8p, u8p, vp]
    L.qs_step_host.argtypes = [vp, fp, C.POINTER(QsResetOptions), fp, fp, u8p, u8p, vp]
    L.qs_step_host_strided.argtypes = [vp, fp, C.POINTER(QsResetOptions), fp, C.c_size_t, fp, u8p, u8p, vp]
    L.qs_step_k.argtypes = [vp, C.c_int, fp, C.POINTER(QsResetOptions), fp, C.c_size_t, fp, u8p, u8p, vp]
    L.qs_step_autoreset.argtypes = [vp, fp, C.POINTER(QsResetOptions), fp, fp, u8p, u8p, vp]
    L.qs_reset.argtypes = [vp, u8p, fp, fp, C.POINTER(QsResetOptions), fp, vp]
    L.qs_reset_done.argtypes = [vp, u8p, C.POINTER(QsResetOptions), fp, vp]
    L.qs_forward.argtypes = [vp, vp]
    L.qs_get.argtypes = [vp, C.c_int, fp, vp]
    L.qs_max_contacts.argtypes = [vp]
    L.qs_raycast_heightmap.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, fp, vp]
    L.qs_launch_count.argtypes = [vp]
    L.qs_launch_count.restype = C.c_int64
    if L.qs_model_sizeof() != C.sizeof(QsModel) or L.qs_config_sizeof() != C.sizeof(QsConfig) \
            or L.qs_buffers_sizeof() != C.sizeof(QsBuffers):
        raise RuntimeError('libqstep.so was built against a different include/qstep.h (struct size mismatch)')
    _lib = L
    return L


def command_mode_bits(base_vel_command_type: str) -> int:
    """quadruped_env.py:1049-1070: substring matching on the command type string."""
    t = base_vel_command_type
    bits = 0
    if 'forward' in t:
        bits |= CMD_FORWARD
    elif 'random' in t:
        bits |= CMD_RANDOM
    elif 'human' in t:
        pass
    else:
        raise ValueError(f'Invalid base linear velocity command type: {t}')
    if 'rotate' in t:
        bits |= CMD_ROTATE
    if 'reset' in t:
        bits |= CMD_RESET
    return bits


class BatchSim:
    """N independent environments of one robot+scene on one GPU."""

    def __init__(self, model: Model, num_envs: int, device: int | str | torch.device = 0, precision: int = 0,
                 use_imu: bool = False, imu_noise=(0.01, 0.01, 0.01, 0.01), seed: int = 0, env_id_offset: int = 0,
                 solver_max_iter: int = 0, heightmap: tuple | None = None, pipeline: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError('gym_quadruped_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
        self.L = load_library()
        self.model = model
        self.device = torch.device(device if not isinstance(device, int) else f'cuda:{device}')
        if self.device.type != 'cuda':
            raise RuntimeError(f'device must be a CUDA device, got {self.device}')
        self.N = int(num_envs)
        cfg = QsConfig()
        cfg.num_envs, cfg.device, cfg.precision = self.N, self.device.index or 0, int(precision)
        cfg.use_imu = int(bool(use_imu))
        if use_imu and not model.c.has_imu:
            raise ValueError(f'robot {model.robot} has no accelerometer/gyro pair in its model')
        cfg.imu_accel_noise, cfg.imu_gyro_noise, cfg.imu_accel_bias_rate, cfg.imu_gyro_bias_rate = [float(x) for x in imu_noise]
        cfg.seed, cfg.env_id_offset, cfg.solver_max_iter = int(seed), int(env_id_offset), int(solver_max_iter)
        # pipeline=True: back-to-back step launches overlap on the device (see QsConfig.pipeline in include/qstep.h for the contract)
        cfg.pipeline = int(bool(pipeline))
        if heightmap is not None:  # (rows, cols, dx, dy): appended to every observation row
            cfg.hm_rows, cfg.hm_cols, cfg.hm_dx, cfg.hm_dy = int(heightmap[0]), int(heightmap[1]), float(heightmap[2]), float(heightmap[3])
        self.cfg = cfg
        self.obs_dim = self.L.qs_obs_dim(C.byref(cfg))
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.L.qs_create(C.byref(model.c), C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f'qs_create failed ({rc}): {self.L.qs_last_error(None).decode()}')
        N, dev = self.N, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        key = torch.tensor(list(model.c.key_qpos), dtype=torch.float64)
        self.qpos = key.to(torch.float32).to(dev).repeat(N, 1).contiguous()
        self.qvel = torch.zeros(N, 18, **f32)
        self.qacc = torch.zeros(N, 18, **f32)
        self.qacc_warmstart = torch.zeros(N, 18, **f32)
        self.base_pos64 = key[:3].to(dev).repeat(N, 1).contiguous()
        self.qfrc_applied = torch.zeros(N, 6, **f32)
        self.command = torch.zeros(N, 4, **f32)
        self.friction = torch.full((N, 2), -1.0, **f32)
        self.sim_time = torch.zeros(N, **f32)
        self.step_count = torch.zeros(N, dtype=torch.int32, device=dev)
        self.imu_bias = torch.zeros(N, 6, **f32)
        self.status = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.ncon = torch.zeros(N, dtype=torch.int32, device=dev)
        self.solver_iter = torch.zeros(N, dtype=torch.int32, device=dev)
        self.invalid_body_mask = torch.zeros(N, 2, dtype=torch.uint8, device=dev)
        # in-episode schedules (quadruped_env.py:293-305), advanced inside the step kernel
        self.cmd_count = torch.zeros(N, dtype=torch.int32, device=dev)
        self.cmd_limit = torch.zeros(N, dtype=torch.int32, device=dev)
        self.ext_count = torch.zeros(N, dtype=torch.int32, device=dev)
        self.ext_limit = torch.zeros(N, dtype=torch.int32, device=dev)
        self.ext_wrench = torch.zeros(N, 6, **f32)
        b = QsBuffers()
        for name in ('qpos', 'qvel', 'qacc', 'qacc_warmstart', 'base_pos64', 'qfrc_applied', 'command', 'friction', 'sim_time',
                     'step_count', 'imu_bias', 'status', 'ncon', 'solver_iter', 'invalid_body_mask', 'cmd_count', 'cmd_limit',
                     'ext_count', 'ext_limit', 'ext_wrench'):
            setattr(b, name, getattr(self, name).data_ptr())
        self._buffers = b
        self._check(self.L.qs_bind(self.h, C.byref(b)))
        # outputs
        self.obs = torch.zeros(N, self.obs_dim, **f32)
        self.reward = torch.zeros(N, **f32)
        self.terminated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.reset_options = self.make_reset_options()

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f'libqstep error {rc}: {self.L.qs_last_error(self.h).decode()}')

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            torch.cuda.synchronize(self.device)
            self.L.qs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.L.qs_launch_count(self.h))

    @property
    def step_variant(self) -> str:
        """Name of the compiled kernel variant that `step` launches (csrc/qs_variants.h)."""
        return self.L.qs_step_variant(self.h).decode()

    def set_seed(self, seed: int):
        """`np.random.seed(seed)` equivalent: new key, per-env draw counters restarted; no buffer is touched."""
        self._check(self.L.qs_set_seed(self.h, C.c_uint64(int(seed)), self._stream()))

    def set_schedule(self, command_mode: int = 0, lin_vel_range=(0.0, 0.0), ang_vel_range=(0.0, 0.0), ext_ranges: dict | None = None,
                     ext_enabled: bool = False):
        """Install the in-kernel schedules: '+reset' command resampling (quadruped_env.py:293-296) and the external base wrench
        (external_disturbances_kwargs, :299-305).  `ext_ranges` maps 'x','y','z','roll','pitch','yaw' to a 1- or 2-element range."""
        sc = QsSchedule()
        sc.command_mode, sc.ext_enabled = int(command_mode), int(bool(ext_enabled))
        sc.lin_vel_range[:] = [float(x) for x in lin_vel_range]
        sc.ang_vel_range[:] = [float(x) for x in ang_vel_range]
        for k, key in enumerate(('x', 'y', 'z', 'roll', 'pitch', 'yaw')):
            r = (ext_ranges or {}).get(key)
            lo, hi = (0.0, 0.0) if r is None else ((float(r[0]), float(r[0])) if len(r) == 1 else (float(r[0]), float(r[1])))
            sc.ext_lo[k], sc.ext_hi[k] = lo, hi
        self._check(self.L.qs_set_schedule(self.h, C.byref(sc), self._stream()))

    def make_reset_options(self, randomize=True, angle_sweep=20 * math.pi / 180, roll_sweep=10 * math.pi / 180,
                           pitch_sweep=10 * math.pi / 180, lin_vel_range=(0.5, 0.5), ang_vel_range=(0.0, 0.0),
                           friction_range=(1.0, 1.0), command_mode=CMD_FORWARD) -> QsResetOptions:
        o = QsResetOptions()
        o.angle_sweep, o.vel_sweep, o.roll_sweep, o.pitch_sweep = float(angle_sweep), 0.5, float(roll_sweep), float(pitch_sweep)
        o.hip_height = self.model.hip_height
        o.lin_vel_range[:] = [float(x) for x in lin_vel_range]
        o.ang_vel_range[:] = [float(x) for x in ang_vel_range]
        o.friction_range[:] = [float(x) for x in friction_range]
        o.command_mode, o.randomize = int(command_mode), int(bool(randomize))
        return o

    # ------------------------------------------------------------------ state access (write-then-step semantics)
    def set_state(self, qpos: torch.Tensor, qvel: torch.Tensor, env_ids=None):
        """Write generalized coordinates; keeps the fp64 base-position master copy in sync."""
        idx = slice(None) if env_ids is None else env_ids
        qpos = torch.as_tensor(qpos, device=self.device)
        self.base_pos64[idx] = qpos[..., :3].to(torch.float64)
        self.qpos[idx] = qpos.to(torch.float32)
        self.qvel[idx] = torch.as_tensor(qvel, device=self.device).to(torch.float32)

    # ------------------------------------------------------------------ hot path
    def _as_ctrl(self, ctrl) -> torch.Tensor:
        if (not isinstance(ctrl, torch.Tensor) or ctrl.device != self.device or ctrl.dtype != torch.float32 or not ctrl.is_contiguous()
                or ctrl.shape != (self.N, 12)):
            ctrl = torch.as_tensor(ctrl, dtype=torch.float32, device=self.device).reshape(self.N, 12).contiguous()
        return ctrl

    def step(self, ctrl: torch.Tensor):
        """One fused kernel launch: ctrl [N,12] (cuda fp32, contiguous) -> obs [N,D], reward, terminated, truncated."""
        ctrl = self._as_ctrl(ctrl)
        self._check(self.L.qs_step(self.h, ctrl.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                   self.terminated.data_ptr(), self.truncated.data_ptr(), self._stream()))
        return self.obs, self.reward, self.terminated, self.truncated

    def step_autoreset(self, ctrl: torch.Tensor, options: QsResetOptions | None = None, obs_out: torch.Tensor | None = None):
        """`step` plus, in the same launch, a random reset of every env that just terminated (post-reset obs/state returned).
        `obs_out`: write the observation rows there instead of `self.obs` (double-buffered consumers, see distributed.ObsGather)."""
        o = options or self.reset_options
        ctrl = self._as_ctrl(ctrl)
        obs = self.obs if obs_out is None else obs_out
        self._check(self.L.qs_step_autoreset(self.h, ctrl.data_ptr(), C.byref(o), obs.data_ptr(), self.reward.data_ptr(),
                                             self.terminated.data_ptr(), self.truncated.data_ptr(), self._stream()))
        return obs, self.reward, self.terminated, self.truncated

    def step_k(self, ctrl_seq: torch.Tensor, options: QsResetOptions | None = None, auto_reset: bool = True,
               obs_ring: torch.Tensor | None = None, terminated_ring: torch.Tensor | None = None):
        """K consecutive steps from one call (`qs_step_k`): ctrl_seq [K, N, 12]; bit-identical to K calls of `step_autoreset` /
        `step`, one launch per step, launches overlapped on the device.  `obs_ring` [K, N, D] keeps every step's observation rows
        (default: only the last step's rows, in `self.obs`); `terminated_ring` [K, N] uint8 keeps every step's flags."""
        K = int(ctrl_seq.shape[0])
        if ctrl_seq.device != self.device or ctrl_seq.dtype != torch.float32 or not ctrl_seq.is_contiguous() or ctrl_seq.shape[1:] != (self.N, 12):
            ctrl_seq = torch.as_tensor(ctrl_seq, dtype=torch.float32, device=self.device).reshape(K, self.N, 12).contiguous()
        o = (options or self.reset_options) if auto_reset else None
        obs, stride = (self.obs, 0) if obs_ring is None else (obs_ring, self.N * self.obs_dim)
        if obs_ring is not None:
            assert obs_ring.shape == (K, self.N, self.obs_dim) and obs_ring.is_contiguous() and obs_ring.dtype == torch.float32
        if terminated_ring is None:
            terminated_ring = torch.empty(K, self.N, dtype=torch.uint8, device=self.device)
        assert terminated_ring.shape == (K, self.N) and terminated_ring.dtype == torch.uint8 and terminated_ring.is_contiguous()
        self._check(self.L.qs_step_k(self.h, K, ctrl_seq.data_ptr(), C.byref(o) if o is not None else None, obs.data_ptr(), stride,
                                     None, terminated_ring.data_ptr(), None, self._stream()))
        self.terminated.copy_(terminated_ring[-1])
        return obs, terminated_ring

    def step_host(self, ctrl_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  terminated_host: torch.Tensor, truncated_host: torch.Tensor, auto_reset: QsResetOptions | None = None):
        """Same step through HOST buffers (pinned CPU tensors): H2D ctrl, kernel, D2H results, stream-synchronised.
        `obs_host` may be a [N, D] view of a wider pinned tensor (padded rows, e.g. `torch.empty(N, 256).pin_memory()[:, :D]`)."""
        stride = int(obs_host.stride(0)) if obs_host.dim() == 2 else self.obs_dim
        assert obs_host.dim() != 2 or obs_host.stride(1) == 1
        self._check(self.L.qs_step_host_strided(self.h, ctrl_host.data_ptr(), C.byref(auto_reset) if auto_reset is not None else None,
                                                obs_host.data_ptr(), stride, reward_host.data_ptr(),
                                                terminated_host.data_ptr(), truncated_host.data_ptr(), self._stream()))

    def reset(self, mask: torch.Tensor | None = None, qpos: torch.Tensor | None = None, qvel: torch.Tensor | None = None,
              options: QsResetOptions | None = None):
        o = options or self.reset_options
        mp = mask.data_ptr() if mask is not None else None
        if qpos is not None:
            qpos = torch.as_tensor(qpos, dtype=torch.float32, device=self.device).reshape(self.N, 19).contiguous()
            qvel = torch.as_tensor(qvel, dtype=torch.float32, device=self.device).reshape(self.N, 18).contiguous()
        self._check(self.L.qs_reset(self.h, mp, qpos.data_ptr() if qpos is not None else None,
                                    qvel.data_ptr() if qvel is not None else None, C.byref(o), self.obs.data_ptr(), self._stream()))
        return self.obs

    def reset_done(self, options: QsResetOptions | None = None):
        """Auto-reset every env whose `terminated` flag is set (one masked launch)."""
        o = options or self.reset_options
        self._check(self.L.qs_reset_done(self.h, self.terminated.data_ptr(), C.byref(o), self.obs.data_ptr(), self._stream()))
        return self.obs

    def forward(self):
        self._check(self.L.qs_forward(self.h, self._stream()))

    def get(self, field: int) -> torch.Tensor:
        if field == FIELD_CONTACTS:
            nmax = self.L.qs_max_contacts(self.h)
            out = torch.empty(self.N, nmax, QS_CONTACT_STRIDE, dtype=torch.float32, device=self.device)
        else:
            out = torch.empty(self.N, *_FIELD_SHAPE[field], dtype=torch.float32, device=self.device)
        self._check(self.L.qs_get(self.h, field, out.data_ptr(), self._stream()))
        return out
