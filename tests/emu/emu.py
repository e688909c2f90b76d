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
        srcs = [HERE / 'emu_main.cpp', HERE / 'warp_emu.h'] + list((HERE.parents[1] / 'gym_quadruped_b200' / 'csrc').glob('qs_*'))
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
