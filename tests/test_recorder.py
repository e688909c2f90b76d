"""Rollout recorder: the reference's `H5Writer` layout (gym_quadruped/utils/data/h5py.py:90-172) round-tripped
through the `.npz` container, and the HDF5 converter checked against an in-memory stand-in for h5py (h5py is not in this image)."""
import json

import numpy as np
import pytest

from gym_quadruped_b200.sensors import IMU
from gym_quadruped_b200.spaces import Box
from gym_quadruped_b200.utils.data.recorder import RolloutReader, RolloutWriter, to_hdf5


class _Spaces:
    def __init__(self, d):
        self.spaces = d


class _FakeEnv:
    """The three things the writer needs from a QuadrupedEnv: hyper-parameters, observation space, action space."""
    num_envs = 1

    def __init__(self):
        self.observation_space = _Spaces({'qpos': Box(-np.inf, np.inf, (19,), np.float64), 'base_ori_SO3': Box(-np.inf, np.inf, (3, 3), np.float64),
                                          'contact_state': Box(-np.inf, np.inf, (4,), np.float64)})
        self.action_space = Box(-np.inf, np.inf, (12,), np.float32)

    def get_hyperparameters(self):
        return dict(robot='mini_cheetah', scene='flat', sim_dt=0.002, ref_base_lin_vel=(0.5, 1.0), ground_friction_coeff=0.8,
                    state_obs_names=('qpos', 'base_ori_SO3', 'contact_state'), sensors=(IMU,),
                    sensors_kwargs=({'accel_name': 'Body_Acc', 'gyro_name': 'Body_Gyro'},), external_disturbances_kwargs=None,
                    legs_order=('FL', 'FR', 'RL', 'RR'))


def _traj(rng, T):
    return ({'qpos': rng.randn(T, 19), 'base_ori_SO3': rng.randn(T, 3, 3), 'contact_state': (rng.rand(T, 4) > 0.5).astype(float),
             'action': rng.randn(T, 12)}, (0.002 * np.arange(1, T + 1))[:, None])


def test_round_trip_matches_reference_layout(tmp_path):
    rng = np.random.RandomState(0)
    w = RolloutWriter(tmp_path / 'run.npz', _FakeEnv())
    r0 = RolloutReader(tmp_path / 'run.npz')
    assert r0.len() == 0 and r0.recordings['qpos'].shape == (0, 0, 19)      # empty file right after construction (h5py.py:107-117)
    trajs = [_traj(rng, 50) for _ in range(3)]
    for obs, t in trajs:
        w.append_trajectory(obs, t)
    w.append_trajectories({k: np.stack([trajs[0][0][k], trajs[1][0][k]]) for k in trajs[0][0]}, np.stack([trajs[0][1], trajs[1][1]]))
    r = RolloutReader(tmp_path / 'run.npz')
    assert r.len() == 5
    assert r.recordings['time'].shape == (5, 50, 1) and r.recordings['base_ori_SO3'].shape == (5, 50, 3, 3) and r.recordings['action'].shape == (5, 50, 12)
    assert all(v.dtype == np.float64 for v in r.recordings.values())
    for i, (obs, t) in enumerate(trajs):
        time, data = r.get_trajectory(i)
        assert np.array_equal(time, t) and set(data) == {'qpos', 'base_ori_SO3', 'contact_state', 'action'}
        for k in obs:
            assert np.array_equal(data[k], obs[k])
    assert np.array_equal(r.get_trajectory(4)[1]['qpos'], trajs[1][0]['qpos'])
    hp = r.env_hparams
    assert hp['robot'] == 'mini_cheetah' and hp['sim_dt'] == 0.002 and hp['ref_base_lin_vel'] == [0.5, 1.0]
    assert hp['sensors'] == [IMU] and 'external_disturbances_kwargs' not in hp               # class refs restored, None skipped (:43-46)
    with pytest.raises(ValueError):
        w.append_trajectory({'qpos': np.zeros((7, 19))}, np.zeros((9, 1)))


class _FakeH5:
    """Minimal in-memory h5py stand-in: File / Group / Dataset / attrs, enough to check the tree the converter builds."""

    class Node(dict):
        def __init__(self):
            super().__init__()
            self.attrs = {}

        def create_group(self, k):
            self[k] = _FakeH5.Node()
            return self[k]

        def require_group(self, k):
            return self[k] if k in self else self.create_group(k)

        def create_dataset(self, k, shape, maxshape, dtype):
            self[k] = _FakeH5.Dataset(shape, maxshape, dtype)
            return self[k]

    class Dataset:
        def __init__(self, shape, maxshape, dtype):
            self.shape, self.maxshape, self.dtype, self.data = shape, maxshape, dtype, np.zeros(shape)

        def __setitem__(self, idx, v):
            self.data[idx] = v

    files = {}

    class File(Node):
        def __init__(self, path, mode):
            super().__init__()
            _FakeH5.files[path] = self

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False


def test_hdf5_converter_builds_the_reference_tree(tmp_path):
    rng = np.random.RandomState(1)
    w = RolloutWriter(tmp_path / 'run.npz', _FakeEnv())
    obs, t = _traj(rng, 20)
    w.append_trajectory(obs, t)
    to_hdf5(tmp_path / 'run.npz', tmp_path / 'run.h5', h5py_module=_FakeH5)
    f = _FakeH5.files[str(tmp_path / 'run.h5')]
    assert set(f) == {'env_hparams', 'recordings'}
    hp = f['env_hparams'].attrs
    assert hp['robot'] == 'mini_cheetah' and json.loads(hp['ref_base_lin_vel']) == [0.5, 1.0]
    assert json.loads(hp['sensors']) == ['TYPE:gym_quadruped_b200.sensors.imu.IMU']                 # h5py.py:36-40
    rec = f['recordings']
    assert set(rec) == {'time', 'qpos', 'base_ori_SO3', 'contact_state', 'action'}
    assert rec['qpos'].shape == (1, 20, 19) and rec['qpos'].maxshape == (None, None, 19) and rec['qpos'].dtype == 'float64'
    assert rec['time'].shape == (1, 20, 1) and np.array_equal(rec['base_ori_SO3'].data[0], obs['base_ori_SO3'])
