"""ctypes wrapper around oracle/build/liboracle.so (fp64 CPU restatement of the hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package never imports this module.  PARITY UNPINNED (see qstep_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from gym_quadruped_b200.model import QS_CONTACT_STRIDE, QS_NOBS_BASE, Model, QsModel

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / 'build' / 'liboracle.so'

F_M, F_BIAS, F_PASSIVE, F_FEET_JACP, F_FEET_POS, F_COM, F_CONTACTS, F_SMOOTH, F_CONSTRAINT, F_XPOS, F_IMU, \
    F_QACC_SMOOTH, F_EFC, F_FLAGS, F_FEET_JACR, F_FEET_JACP_DOT, F_FEET_JACR_DOT, F_EFC_FULL = range(18)
EFC_FULL_STRIDE = 34


def build(force: bool = False) -> Path:
    src = HERE / 'qstep_oracle.c'
    hdr = HERE.parent / 'include' / 'qstep.h'
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(['make', '-C', str(HERE), '-B'], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(QsModel)]
        L.orc_destroy.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.orc_set_state.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.orc_set_env.argtypes = [C.c_void_p, C.c_double, C.c_double, dp, dp]
        L.orc_set_time.argtypes = [C.c_void_p, C.c_double]
        L.orc_get_time.restype = C.c_double
        L.orc_get_time.argtypes = [C.c_void_p]
        L.orc_forward.argtypes = [C.c_void_p, dp]
        L.orc_step.argtypes = [C.c_void_p, dp, dp]
        L.orc_step.restype = C.c_int
        L.orc_lift.argtypes = [C.c_void_p]
        L.orc_lift.restype = C.c_int
        L.orc_get.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_get.restype = C.c_int
        L.orc_rollout.argtypes = [C.c_void_p, dp, C.c_int, dp]
        L.orc_rollout.restype = C.c_int
        L.orc_heightmap.argtypes = [C.c_void_p, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, dp]
        L.orc_rollout_autoreset.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int)]
        L.orc_rollout_autoreset.restype = C.c_int
        assert L.orc_model_sizeof() == C.sizeof(QsModel), 'QsModel layout mismatch between ctypes and C'
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One fp64 environment."""

    def __init__(self, model: Model):
        self.model = model
        self.L = lib()
        self.h = self.L.orc_create(C.byref(model.c))
        self.obs_dim = QS_NOBS_BASE + (6 if model.c.has_imu else 0)

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_state(self, qpos=None, qvel=None, warm=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, warm)]
        self.L.orc_set_state(self.h, _p(a[0]), _p(a[1]), _p(a[2]))

    def get_state(self):
        qpos, qvel, qacc, warm = np.zeros(19), np.zeros(18), np.zeros(18), np.zeros(18)
        self.L.orc_get_state(self.h, _p(qpos), _p(qvel), _p(qacc), _p(warm))
        return qpos, qvel, qacc, warm

    def set_env(self, mu_floor=-1.0, mu_feet=-1.0, command=None, qfrc_applied=None):
        c = None if command is None else np.ascontiguousarray(command, dtype=np.float64)
        f = None if qfrc_applied is None else np.ascontiguousarray(qfrc_applied, dtype=np.float64)
        self.L.orc_set_env(self.h, float(mu_floor), float(mu_feet), _p(c), _p(f))

    def forward(self, ctrl=None):
        c = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        self.L.orc_forward(self.h, _p(c))

    def step(self, ctrl):
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        obs = np.zeros(self.obs_dim)
        term = self.L.orc_step(self.h, _p(c), _p(obs))
        return obs, bool(term)

    def rollout(self, ctrl_table):
        c = np.ascontiguousarray(ctrl_table, dtype=np.float64)
        return self.L.orc_rollout(self.h, _p(c), len(c), None)

    def rollout_autoreset(self, ctrl_table, reset_states, cursor=0):
        c = np.ascontiguousarray(ctrl_table, dtype=np.float64)
        r = np.ascontiguousarray(reset_states, dtype=np.float64)
        cur = C.c_int(cursor)
        n = self.L.orc_rollout_autoreset(self.h, _p(c), len(c), _p(r), len(r), C.byref(cur))
        return n, cur.value

    def heightmap(self, center, yaw, rows, cols, dx, dy):
        c = np.ascontiguousarray(center, dtype=np.float64)
        out = np.zeros((rows, cols, 3))
        self.L.orc_heightmap(self.h, _p(c), float(yaw), rows, cols, float(dx), float(dy), _p(out))
        return out

    def lift(self):
        return self.L.orc_lift(self.h)

    def get(self, field):
        buf = np.zeros(131072)
        n = self.L.orc_get(self.h, field, _p(buf))
        if field == F_EFC_FULL:
            return buf[:EFC_FULL_STRIDE * n].reshape(n, EFC_FULL_STRIDE).copy()
        if field == F_M:
            return buf[:324].reshape(18, 18).copy()
        if field in (F_FEET_JACP, F_FEET_JACR, F_FEET_JACP_DOT, F_FEET_JACR_DOT):
            return buf[:n].reshape(4, 3, 18).copy()
        if field == F_FEET_POS:
            return buf[:12].reshape(4, 3).copy()
        if field == F_CONTACTS:
            return buf[:n * QS_CONTACT_STRIDE].reshape(n, QS_CONTACT_STRIDE).copy()
        if field == F_XPOS:
            return buf[:39].reshape(13, 3).copy()
        if field == F_EFC:
            return buf[:6 * n].reshape(n, 6).copy()
        return buf[:n].copy()

    def flags(self):
        f = self.get(F_FLAGS)
        return {'contact_state': f[:4].astype(bool), 'invalid_contact': bool(f[4]), 'out_of_bounds': bool(f[5]),
                'ncon': int(f[6]), 'nefc': int(f[7]), 'solver_iter': int(f[8]), 'overflow': bool(f[9]),
                'invalid_body_mask': int(f[10])}
