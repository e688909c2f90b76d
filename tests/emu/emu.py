"""ctypes wrapper for tests/emu/build/libqsemu.so (host warp emulator of the kernel source; test infrastructure)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from gym_quadruped_b200.model import Model, QsModel

HERE = Path(__file__).resolve().parent
LIB = HERE / 'build' / 'libqsemu.so'
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [HERE / 'emu_main.cpp', HERE / 'emu_kernel.cpp', HERE / 'warp_emu.h', HERE / 'kernel_emu.h', HERE / 'Makefile'] + list((HERE.parents[1] / 'gym_quadruped_b200' / 'csrc').glob('qs_*'))
        if not LIB.exists() or LIB.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
            subprocess.run(['make', '-C', str(HERE), '-B'], check=True, capture_output=True)
        L = C.CDLL(str(LIB))
        dp = C.POINTER(C.c_double)
        L.emu_step.argtypes = [C.POINTER(QsModel), C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_int, C.c_double, C.c_int, C.c_int]
        L.emu_step.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def emu_step(model: Model, qpos, qvel, warm, ctrl, mu_floor=-1.0, mu_feet=-1.0, command=(0, 0, 0, 0), applied=(0,) * 6,
             precision=1, max_iter=100, tol=1e-8, mode=1, specialised=False):
    """Returns dict(qpos, qvel, qacc, obs, misc...) after one emulated step (mode=1) or forward pass (mode=0)."""
    qpos = np.array(qpos, dtype=np.float64); qvel = np.array(qvel, dtype=np.float64); warm = np.array(warm, dtype=np.float64)
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    envp = np.array([mu_floor, mu_feet, *command, *applied], dtype=np.float64)
    obs = np.zeros(232); misc = np.zeros(1024)
    rc = lib().emu_step(C.byref(model.c), precision, _p(qpos), _p(qvel), _p(warm), _p(ctrl), _p(envp), _p(obs), _p(misc), max_iter, tol, mode, int(specialised))
    assert rc == 0
    ncon = int(misc[5])
    return {
        'qpos': qpos, 'qvel': qvel, 'qacc': misc[8:26].copy(), 'obs': obs[:227].copy(), 'iters': int(misc[0]), 'maxed': bool(misc[1]),
        'contact_mask': int(misc[2]), 'invalid_mask': int(misc[3]), 'oob': bool(misc[4]), 'ncon': ncon, 'overflow': bool(misc[6]), 'feat': int(misc[7]),
        'bias': misc[26:44].copy(), 'fsm': misc[44:62].copy(), 'qacc_smooth': misc[62:80].copy(), 'fcon': misc[80:98].copy(),
        'M': misc[98:422].reshape(18, 18).copy(), 'contacts': misc[422:422 + 20 * ncon].reshape(ncon, 20).copy(), 'imu': misc[742:748].copy(), 'heightmap': misc[748:823].reshape(5, 5, 3).copy(),
    }


# ---------------------------------------------------------------------------------------------- the whole kernel on the emulator
from gym_quadruped_b200.model import QsBuffers, QsResetOptions, QsSchedule  # noqa: E402


class EmuLaunch(C.Structure):
    _fields_ = [
        ('precision', C.c_int), ('mode', C.c_int), ('num_envs', C.c_int), ('use_imu', C.c_int), ('hm_rows', C.c_int), ('hm_cols', C.c_int),
        ('hm_dx', C.c_double), ('hm_dy', C.c_double),
        ('max_iter', C.c_int), ('env_id_offset', C.c_int), ('auto_reset', C.c_int), ('pad0', C.c_int),
        ('tol', C.c_double), ('seed', C.c_uint64),
        ('imu_an', C.c_double), ('imu_gn', C.c_double), ('imu_abr', C.c_double), ('imu_gbr', C.c_double),
        ('buf', QsBuffers),
        ('episode', C.c_void_p), ('tick', C.c_void_p), ('cmd_epoch', C.c_void_p), ('ext_epoch', C.c_void_p),
        ('sched', QsSchedule),
        ('ctrl', C.c_void_p), ('obs', C.c_void_p), ('reward', C.c_void_p), ('terminated', C.c_void_p), ('truncated', C.c_void_p),
        ('mask', C.c_void_p), ('in_qpos', C.c_void_p), ('in_qvel', C.c_void_p),
        ('ro', QsResetOptions),
        ('aux', C.c_void_p),
        ('gather_world', C.c_int), ('gather_rank', C.c_int), ('gather_row_stride', C.c_int), ('gather_seq', C.c_int),
        ('gather_peers', C.c_void_p * 8), ('gather_flags', C.c_void_p * 8),
    ]


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 as in csrc/qs_math.cuh (independent restatement for the tests)."""
    M = 0xffffffff
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


def u32_to_unit(x):
    return np.float32(x >> 8) * np.float32(1.0 / 16777216.0)


class EmuSim:
    """The C-ABI's buffer set on the host (numpy) plus launches of the kernel body on the warp emulator: the CPU twin of
    `gym_quadruped_b200.backend.BatchSim` for tests (generic kernel variant, plain stream order)."""

    def __init__(self, model: Model, n: int, precision=1, seed=0, use_imu=False, heightmap=None, env_id_offset=0, max_iter=None):
        self.model, self.N, self.precision, self.seed, self.use_imu = model, n, precision, seed, bool(use_imu)
        self.hm = heightmap or (0, 0, 0.0, 0.0)
        self.obs_dim = 227 + (18 if use_imu else 0) + self.hm[0] * self.hm[1] * 3
        self.env_id_offset = env_id_offset
        self.max_iter = max_iter or (50 if precision == 0 else 100)
        f32 = np.float32
        self.qpos = np.zeros((n, 19), f32); self.qvel = np.zeros((n, 18), f32); self.qacc = np.zeros((n, 18), f32)
        self.qacc_warmstart = np.zeros((n, 18), f32); self.base_pos64 = np.zeros((n, 3)); self.qfrc_applied = np.zeros((n, 6), f32)
        self.command = np.zeros((n, 4), f32); self.friction = np.full((n, 2), -1.0, f32); self.sim_time = np.zeros(n, f32)
        self.step_count = np.zeros(n, np.int32); self.imu_bias = np.zeros((n, 6), f32); self.status = np.zeros(n, np.uint8)
        self.ncon = np.zeros(n, np.int32); self.solver_iter = np.zeros(n, np.int32); self.invalid_body_mask = np.zeros((n, 2), np.uint8)
        self.cmd_count = np.zeros(n, np.int32); self.cmd_limit = np.full(n, 0x7fffffff, np.int32)
        self.ext_count = np.zeros(n, np.int32); self.ext_limit = np.full(n, 0x7fffffff, np.int32); self.ext_wrench = np.zeros((n, 6), f32)
        self.episode = np.zeros(n, np.uint32); self.tick = np.zeros(n, np.uint32)
        self.cmd_epoch = np.zeros(n, np.uint32); self.ext_epoch = np.zeros(n, np.uint32)
        self.obs = np.zeros((n, self.obs_dim), f32); self.reward = np.zeros(n, f32)
        self.terminated = np.zeros(n, np.uint8); self.truncated = np.zeros(n, np.uint8)
        self.sched = QsSchedule()
        self.imu_noise = (0.0, 0.0, 0.0, 0.0)  # accel noise, gyro noise, accel bias rate, gyro bias rate (QsConfig.imu_*)
        q = np.array(model.c.qpos0, dtype=np.float64)
        self.qpos[:] = q.astype(f32); self.base_pos64[:] = q[:3]
        self.aux = None
        self.gather = None  # dict(world, rank, stride, seq, peers=[float32 arrays [world*N, stride]], flags=[uint32 arrays [8]])

    def set_state(self, qpos, qvel):
        qpos = np.asarray(qpos, dtype=np.float64).reshape(self.N, 19)
        self.base_pos64[:] = qpos[:, :3]; self.qpos[:] = qpos.astype(np.float32); self.qvel[:] = np.asarray(qvel, dtype=np.float32).reshape(self.N, 18)

    def _launch(self, mode, ctrl=None, auto_reset=None, mask=None, in_qpos=None, in_qvel=None, ro=None):
        L = EmuLaunch()
        L.precision, L.mode, L.num_envs, L.use_imu = self.precision, mode, self.N, int(self.use_imu)
        L.hm_rows, L.hm_cols, L.hm_dx, L.hm_dy = self.hm
        L.max_iter, L.env_id_offset, L.auto_reset = self.max_iter, self.env_id_offset, int(auto_reset is not None)
        L.tol = 1e-6 if self.precision == 0 else 1e-8
        L.seed = self.seed
        L.imu_an, L.imu_gn, L.imu_abr, L.imu_gbr = self.imu_noise
        for name, _ in QsBuffers._fields_:
            setattr(L.buf, name, getattr(self, name).ctypes.data)
        for name in ('episode', 'tick', 'cmd_epoch', 'ext_epoch', 'obs', 'reward', 'terminated', 'truncated'):
            setattr(L, name, getattr(self, name).ctypes.data)
        L.sched = self.sched
        keep = []
        for name, arr, dt in (('ctrl', ctrl, np.float32), ('mask', mask, np.uint8), ('in_qpos', in_qpos, np.float32), ('in_qvel', in_qvel, np.float32)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dtype=dt); keep.append(a)
                setattr(L, name, a.ctypes.data)
        if ro is not None or auto_reset is not None:
            L.ro = ro if ro is not None else auto_reset
        if self.gather is not None and mode == 0:
            g = self.gather
            L.gather_world, L.gather_rank, L.gather_row_stride, L.gather_seq = g['world'], g['rank'], g['stride'], g['seq']
            for k in range(g['world']):
                L.gather_peers[k] = g['peers'][k].ctypes.data
                L.gather_flags[k] = g['flags'][k].ctypes.data
        if mode == 2:
            stride = lib().emu_aux_stride()
            self.aux = np.zeros((self.N, stride), np.float32)
            L.aux = self.aux.ctypes.data
        lib().emu_kernel.argtypes = [C.POINTER(QsModel), C.POINTER(EmuLaunch)]
        assert lib().emu_launch_sizeof() == C.sizeof(EmuLaunch)
        rc = lib().emu_kernel(C.byref(self.model.c), C.byref(L))
        assert rc == 0, rc

    def step(self, ctrl):
        self._launch(0, ctrl=np.asarray(ctrl, dtype=np.float32).reshape(self.N, 12))

    def step_autoreset(self, ctrl, opt):
        self._launch(0, ctrl=np.asarray(ctrl, dtype=np.float32).reshape(self.N, 12), auto_reset=opt)

    def reset(self, opt, mask=None, qpos=None, qvel=None):
        self._launch(1, mask=mask, in_qpos=qpos, in_qvel=qvel, ro=opt)

    def forward(self):
        self._launch(2)
