"""Host-side model tables: ctypes mirrors of `include/qstep.h` and the loader that turns the committed compiler output
(`assets/<robot>.json`) plus a scene description into a `QsModel`.

Replaces `MjModel.from_xml_path` (gym_quadruped/quadruped_env.py:170) and `generate_terrain`
(gym_quadruped/utils/mujoco/terrain.py:309-365) for the scenes BASELINE.json names (flat, perlin, random_boxes).
"""
from __future__ import annotations

import ctypes as C
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from .terrain import generate_terrain

ASSET_DIR = Path(__file__).resolve().parent / 'assets'

QS_ABI_VERSION = 5
QS_NBODY, QS_NJNT, QS_NQ, QS_NV, QS_NU, QS_NLEG = 14, 12, 19, 18, 12, 4
QS_MAXGEOM, QS_MAXBOX = 48, 128
QS_NOBS_BASE, QS_NOBS_IMU = 227, 18
QS_CONTACT_STRIDE = 20

d, i32 = C.c_double, C.c_int32


class QsGeomParams(C.Structure):
    _fields_ = [('friction', d * 3), ('solref', d * 2), ('solimp', d * 5), ('solmix', d), ('margin', d), ('gap', d),
                ('condim', i32), ('priority', i32)]


class QsModel(C.Structure):
    _fields_ = [
        ('abi_version', i32),
        ('timestep', d), ('gravity', d * 3), ('impratio', d), ('tolerance', d), ('ls_tolerance', d), ('meaninertia', d),
        ('cone', i32), ('iterations', i32), ('ls_iterations', i32), ('noslip_pad', i32),
        ('body_parent', i32 * QS_NBODY),
        ('body_pos', d * 3 * QS_NBODY), ('body_quat', d * 4 * QS_NBODY), ('body_ipos', d * 3 * QS_NBODY),
        ('body_iquat', d * 4 * QS_NBODY), ('body_mass', d * QS_NBODY), ('body_inertia', d * 3 * QS_NBODY),
        ('body_invweight0', d * 2 * QS_NBODY),
        ('jnt_pos', d * 3 * QS_NJNT), ('jnt_axis', d * 3 * QS_NJNT), ('jnt_range', d * 2 * QS_NJNT),
        ('jnt_solref', d * 2 * QS_NJNT), ('jnt_solimp', d * 5 * QS_NJNT), ('jnt_margin', d * QS_NJNT),
        ('jnt_limited', i32 * QS_NJNT),
        ('qpos0', d * QS_NQ), ('key_qpos', d * QS_NQ),
        ('dof_damping', d * QS_NV), ('dof_armature', d * QS_NV), ('dof_frictionloss', d * QS_NV),
        ('dof_invweight0', d * QS_NV), ('dof_solref', d * 2 * QS_NV), ('dof_solimp', d * 5 * QS_NV),
        ('act_ctrlrange', d * 2 * QS_NU), ('act_forcerange', d * 2 * QS_NU),
        ('act_ctrllimited', i32 * QS_NU), ('act_forcelimited', i32 * QS_NU),
        ('ngeom', i32),
        ('geom_type', i32 * QS_MAXGEOM), ('geom_body', i32 * QS_MAXGEOM), ('geom_foot_leg', i32 * QS_MAXGEOM),
        ('geom_vertadr', i32 * QS_MAXGEOM), ('geom_vertnum', i32 * QS_MAXGEOM),
        ('geom_pos', d * 3 * QS_MAXGEOM), ('geom_quat', d * 4 * QS_MAXGEOM), ('geom_size', d * 3 * QS_MAXGEOM),
        ('geom_bcenter', d * 3 * QS_MAXGEOM), ('geom_rbound', d * QS_MAXGEOM),
        ('geom_par', QsGeomParams * QS_MAXGEOM),
        ('foot_geom', i32 * QS_NLEG),
        ('nvert', i32), ('pad0', i32), ('vert', C.POINTER(d)),
        ('terrain_type', i32), ('nbox', i32),
        ('floor_par', QsGeomParams), ('terrain_limits', d * 4),
        ('hf_nrow', i32), ('hf_ncol', i32), ('hf_size', d * 4), ('hf_pos', d * 3), ('hf_data', C.POINTER(C.c_float)),
        ('hf_par', QsGeomParams),
        ('box_pos', d * 3 * QS_MAXBOX), ('box_quat', d * 4 * QS_MAXBOX), ('box_half', d * 3 * QS_MAXBOX),
        ('box_friction', d * 3 * QS_MAXBOX), ('box_par', QsGeomParams),
        ('has_imu', i32), ('pad1', i32), ('imu_pos', d * 3), ('imu_quat', d * 4),
    ]


class QsConfig(C.Structure):
    _fields_ = [
        ('num_envs', i32), ('device', i32), ('precision', i32), ('use_imu', i32), ('hm_rows', i32), ('hm_cols', i32),
        ('hm_dx', d), ('hm_dy', d),
        ('imu_accel_noise', d), ('imu_gyro_noise', d), ('imu_accel_bias_rate', d), ('imu_gyro_bias_rate', d),
        ('seed', C.c_uint64), ('env_id_offset', i32), ('solver_max_iter', i32), ('pipeline', i32), ('pad0', i32),
    ]


class QsBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'qpos', 'qvel', 'qacc', 'qacc_warmstart', 'base_pos64', 'qfrc_applied', 'command', 'friction', 'sim_time',
        'step_count', 'imu_bias', 'status', 'ncon', 'solver_iter', 'invalid_body_mask', 'cmd_count', 'cmd_limit', 'ext_count',
        'ext_limit', 'ext_wrench')]


class QsSchedule(C.Structure):
    _fields_ = [('command_mode', i32), ('ext_enabled', i32), ('lin_vel_range', d * 2), ('ang_vel_range', d * 2),
                ('ext_lo', d * 6), ('ext_hi', d * 6)]


class QsResetOptions(C.Structure):
    _fields_ = [
        ('angle_sweep', d), ('vel_sweep', d), ('roll_sweep', d), ('pitch_sweep', d), ('hip_height', d),
        ('lin_vel_range', d * 2), ('ang_vel_range', d * 2), ('friction_range', d * 2),
        ('command_mode', i32), ('randomize', i32),
    ]


DEFAULT_GEOM = dict(friction=(1.0, 0.005, 0.0001), solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0),
                    solmix=1.0, margin=0.0, gap=0.0, condim=3, priority=0)


def _fill_par(par: QsGeomParams, src: dict):
    par.friction[:] = [float(x) for x in src['friction']]
    par.solref[:] = [float(x) for x in src['solref']]
    par.solimp[:] = [float(x) for x in src['solimp']]
    par.solmix, par.margin, par.gap = float(src['solmix']), float(src['margin']), float(src['gap'])
    par.condim, par.priority = int(src['condim']), int(src['priority'])


def _set2d(field, rows):
    for i, row in enumerate(rows):
        for j, v in enumerate(row):
            field[i][j] = float(v)


@lru_cache(maxsize=None)
def load_robot_tables(robot: str) -> dict:
    """Committed compiler output for `robot` (see compiler/compile.py)."""
    path = ASSET_DIR / f'{robot}.json'
    if not path.exists():
        raise ValueError(f'Unknown robot name: {robot}')
    return json.loads(path.read_text())


class Model:
    """A `QsModel` plus the numpy arrays that keep its pointers alive."""

    def __init__(self, robot: str, scene: str = 'flat', sim_dt: float = 0.002):
        t = load_robot_tables(robot)
        self.tables = t
        self.robot, self.scene = robot, scene
        self.hip_height = float(t['hip_height'])
        m = QsModel()
        m.abi_version = QS_ABI_VERSION
        m.timestep = sim_dt
        m.gravity[:] = [0.0, 0.0, -9.81]
        m.impratio = float(t['impratio'])
        m.tolerance, m.ls_tolerance = 1e-8, 0.01
        m.meaninertia = float(t['meaninertia'])
        m.cone = 1 if t['cone'] == 'elliptic' else 0
        m.iterations, m.ls_iterations = 100, 50
        m.body_parent[:] = t['body_parent']
        _set2d(m.body_pos, t['body_pos']); _set2d(m.body_quat, t['body_quat'])
        _set2d(m.body_ipos, t['body_ipos']); _set2d(m.body_iquat, t['body_iquat'])
        m.body_mass[:] = t['body_mass']
        _set2d(m.body_inertia, t['body_inertia']); _set2d(m.body_invweight0, t['body_invweight0'])
        _set2d(m.jnt_pos, t['jnt_pos']); _set2d(m.jnt_axis, t['jnt_axis']); _set2d(m.jnt_range, t['jnt_range'])
        _set2d(m.jnt_solref, t['jnt_solref']); _set2d(m.jnt_solimp, t['jnt_solimp'])
        m.jnt_margin[:] = t['jnt_margin']
        m.jnt_limited[:] = t['jnt_limited']
        m.qpos0[:] = t['qpos0']
        m.key_qpos[:] = t['key_qpos']
        m.dof_damping[:] = t['dof_damping']; m.dof_armature[:] = t['dof_armature']
        m.dof_frictionloss[:] = t['dof_frictionloss']; m.dof_invweight0[:] = t['dof_invweight0']
        _set2d(m.dof_solref, t['dof_solref']); _set2d(m.dof_solimp, t['dof_solimp'])
        _set2d(m.act_ctrlrange, t['act_ctrlrange']); _set2d(m.act_forcerange, t['act_forcerange'])
        m.act_ctrllimited[:] = t['act_ctrllimited']; m.act_forcelimited[:] = t['act_forcelimited']
        geoms = t['geoms']
        assert len(geoms) <= QS_MAXGEOM
        m.ngeom = len(geoms)
        for g, e in enumerate(geoms):
            m.geom_type[g], m.geom_body[g], m.geom_foot_leg[g] = e['type'], e['body'], e['foot_leg']
            m.geom_vertadr[g], m.geom_vertnum[g] = e['vertadr'], e['vertnum']
            m.geom_pos[g][:] = e['pos']; m.geom_quat[g][:] = e['quat']; m.geom_size[g][:] = e['size']
            m.geom_bcenter[g][:] = e['bcenter']; m.geom_rbound[g] = e['rbound']
            _fill_par(m.geom_par[g], e)
        m.foot_geom[:] = t['foot_geom']
        self._vert = np.ascontiguousarray(np.asarray(t['vert'], dtype=np.float64).reshape(-1, 3))
        m.nvert = len(self._vert)
        m.vert = self._vert.ctypes.data_as(C.POINTER(d)) if m.nvert else None

        # ---- scene (terrain.py:309-365); the flat floor plane is present in every procedural scene
        terr = generate_terrain(scene, self.hip_height)
        self.terrain = terr
        _fill_par(m.floor_par, DEFAULT_GEOM)
        _fill_par(m.hf_par, DEFAULT_GEOM)
        _fill_par(m.box_par, DEFAULT_GEOM)
        m.terrain_limits[:] = terr['terrain_limits']
        m.terrain_type = {'flat': 0, 'hfield': 1, 'boxes': 2}[terr['type']]
        self._hf = None
        if terr['type'] == 'hfield':
            self._hf = np.ascontiguousarray(terr['data'], dtype=np.float32)
            m.hf_nrow, m.hf_ncol = self._hf.shape
            m.hf_size[:] = terr['size']
            m.hf_pos[:] = terr['pos']
            m.hf_data = self._hf.ctypes.data_as(C.POINTER(C.c_float))
        elif terr['type'] == 'boxes':
            n = len(terr['box_pos'])
            assert n <= QS_MAXBOX
            m.nbox = n
            _set2d(m.box_pos, terr['box_pos']); _set2d(m.box_quat, terr['box_quat']); _set2d(m.box_half, terr['box_half'])
            fri = terr.get('box_friction')
            _set2d(m.box_friction, fri if fri is not None else np.tile(DEFAULT_GEOM['friction'], (n, 1)))
            m.box_par.priority = int(terr.get('box_priority', 0))

        imu = t.get('imu')
        m.has_imu = 1 if imu else 0
        if imu:
            m.imu_pos[:] = imu['pos']; m.imu_quat[:] = imu['quat']
        else:
            m.imu_quat[:] = [1.0, 0, 0, 0]
        self.c = m

    @property
    def terrain_limits(self):
        return tuple(self.c.terrain_limits)

    # convenient numpy views used by tests / the env
    def array(self, name):
        return np.ctypeslib.as_array(getattr(self.c, name)).copy()
