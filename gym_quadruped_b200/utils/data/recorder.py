"""Rollout recorder in the reference's on-disk layout (gym_quadruped/utils/data/h5py.py:90-216).

Layout (identical tree, names, shapes and dtypes as `H5Writer`):
    env_hparams/            constructor arguments of the env (`get_hyperparameters()`): scalars / strings as-is, lists and tuples
                            as JSON strings, class references as "TYPE:<module>.<name>" inside the JSON list, nested dicts as sub-groups
    recordings/time         float64 [traj, time, 1]
    recordings/<obs>        float64 [traj, time, *obs_shape]   one dataset per observation of the env's observation space
    recordings/action       float64 [traj, time, 12]

h5py is not part of this image, so the container written here is a NumPy `.npz` whose keys are the HDF5 paths above
(`recordings/qpos`, ..., `env_hparams` as one JSON document); `to_hdf5()` converts it to a real HDF5 file wherever h5py is
installed, producing exactly what the reference's `H5Writer` would have written, readable by its `H5Reader` /
`ProprioceptiveDataset`.  A batched GPU rollout of N envs over T steps becomes N trajectories of length T in ONE append.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np


def _encode(value):
    """JSON-able image of one hyper-parameter, following save_dict_to_h5 (h5py.py:23-48)."""
    if isinstance(value, dict):
        return {'__group__': {k: _encode(v) for k, v in value.items() if v is not None}}
    if isinstance(value, (list, tuple)):
        try:
            return {'__json__': json.dumps(value)}
        except TypeError:
            if value and isinstance(value[0], type):
                return {'__json__': json.dumps([f'TYPE:{v.__module__}.{v.__name__}' for v in value])}
            if value and isinstance(value[0], dict):  # sensors_kwargs: tuple of dicts
                return {'__json__': json.dumps([{k: (list(x) if isinstance(x, tuple) else x) for k, x in d.items()} for d in value])}
            raise NotImplementedError(f'Need to define how to store {type(value[0])} objects')
    if isinstance(value, np.ndarray):
        return {'__array__': value.tolist()}
    if isinstance(value, (str, int, float, bool)):
        return value
    raise TypeError(f'Cannot save type {type(value)}')


class RolloutWriter:
    """`H5Writer` counterpart (same constructor / `append_trajectory` surface) writing the `.npz` container."""

    def __init__(self, file_path, env, extra_obs: dict | None = None):
        self.file_path = Path(file_path)
        self.file_path.parent.mkdir(parents=True, exist_ok=True)
        self.hparams = {k: _encode(v) for k, v in env.get_hyperparameters().items() if v is not None}
        self.shapes = {k: tuple(sp.shape) for k, sp in env.observation_space.spaces.items()}
        self.shapes['action'] = tuple(env.action_space.shape)
        for k, shp in (extra_obs or {}).items():
            self.shapes[k] = tuple(shp)
        self.data = {k: [] for k in self.shapes}
        self.time = []
        self.num_steps = None
        self._flush()

    def append_trajectory(self, state_obs_traj: dict, time):
        """One trajectory: values of shape (T, *obs_shape), time (T, 1)  (h5py.py:129-172)."""
        self.append_trajectories({k: np.asarray(v)[None] for k, v in state_obs_traj.items()}, np.asarray(time)[None])

    def append_trajectories(self, state_obs_traj: dict, time):
        """A batch of trajectories at once: values of shape (n_traj, T, *obs_shape), time (n_traj, T, 1); torch tensors welcome
        (a [T, N, dim] GPU rollout is passed as `x.permute(1, 0, 2)`)."""
        def to_np(x):
            if hasattr(x, 'detach'):
                x = x.detach().cpu().numpy()
            return np.asarray(x, dtype=np.float64)
        time = to_np(time)
        n, T = time.shape[:2]
        if self.num_steps is not None and T != self.num_steps:
            # the reference resizes every dataset to the NEW trajectory length (h5py.py:153,163): keep the same semantics
            for k in self.data:
                self.data[k] = [self._fit(a, T) for a in self.data[k]]
            self.time = [self._fit(a, T) for a in self.time]
        self.num_steps = T
        for key, value in state_obs_traj.items():
            value = to_np(value)
            if value.shape[1] != T:
                raise ValueError(f'Observation {key} has inconsistent time steps: {value.shape[1]} vs {T} in time array.')
            if key not in self.shapes:
                raise KeyError(f'{key} is not part of the recording (observation space / extra_obs)')
            if tuple(value.shape[2:]) != self.shapes[key]:
                raise ValueError(f'Error appending {key} traj of shape={value.shape}, expected (*, {T}, {self.shapes[key]})')
            self.data[key].append(value)
        self.time.append(time.reshape(n, T, 1))
        self._flush()

    @staticmethod
    def _fit(a, T):
        if a.shape[1] >= T:
            return a[:, :T]
        pad = np.zeros((a.shape[0], T - a.shape[1]) + a.shape[2:])
        return np.concatenate([a, pad], axis=1)

    def _flush(self):
        T = self.num_steps or 0
        out = {'env_hparams': np.array(json.dumps(self.hparams))}
        out['recordings/time'] = np.concatenate(self.time, axis=0) if self.time else np.zeros((0, 0, 1))
        ntraj = out['recordings/time'].shape[0]
        for k, shp in self.shapes.items():
            out[f'recordings/{k}'] = np.concatenate(self.data[k], axis=0) if self.data[k] else np.zeros((ntraj, T) + shp)
        with open(self.file_path, 'wb') as f:
            np.savez(f, **out)


class RolloutReader:
    """`H5Reader` counterpart (h5py.py:175-216) for the `.npz` container."""

    def __init__(self, file_path):
        file_path = Path(file_path)
        assert file_path.exists(), f'File not found: {file_path.absolute()}'
        z = np.load(file_path, allow_pickle=False)
        self.recordings = {k.split('/', 1)[1]: z[k] for k in z.files if k.startswith('recordings/')}
        self.env_hparams = _decode_group(json.loads(str(z['env_hparams'])))
        self.n_trajectories = self.recordings['time'].shape[0]

    def len(self):
        return self.n_trajectories

    def get_trajectory(self, traj_idx):
        time = self.recordings['time'][traj_idx]
        return time, {k: v[traj_idx] for k, v in self.recordings.items() if k != 'time'}

    def close(self):
        pass


def _import_class(ref: str):
    import importlib
    assert ref.startswith('TYPE:'), f'Invalid class reference: {ref}'
    module_name, class_name = ref.split(':', 1)[1].rsplit('.', 1)
    return getattr(importlib.import_module(module_name), class_name)


def _decode_group(enc: dict) -> dict:
    out = {}
    for k, v in enc.items():
        if isinstance(v, dict) and '__group__' in v:
            out[k] = _decode_group(v['__group__'])
        elif isinstance(v, dict) and '__json__' in v:
            val = json.loads(v['__json__'])
            if isinstance(val, list):
                val = [_import_class(e) if isinstance(e, str) and e.startswith('TYPE:') else e for e in val]
            out[k] = val
        elif isinstance(v, dict) and '__array__' in v:
            out[k] = np.array(v['__array__'])
        else:
            out[k] = v
    return out


def to_hdf5(npz_path, h5_path, h5py_module=None):
    """Convert the `.npz` container into the reference's HDF5 file (needs h5py; pass a module object to inject one)."""
    h5py = h5py_module
    if h5py is None:
        import h5py  # noqa: F811  (not available in the build image; present wherever the reference itself runs)
    z = np.load(npz_path, allow_pickle=False)
    enc = json.loads(str(z['env_hparams']))

    def put(group, d):
        for k, v in d.items():
            if isinstance(v, dict) and '__group__' in v:
                put(group.require_group(k), v['__group__'])
            elif isinstance(v, dict) and '__json__' in v:
                group.attrs[k] = v['__json__']
            elif isinstance(v, dict) and '__array__' in v:
                group.attrs[k] = np.array(v['__array__'])
            else:
                group.attrs[k] = v

    with h5py.File(str(h5_path), 'w') as hf:
        put(hf.create_group('env_hparams'), enc)
        rec = hf.create_group('recordings')
        for k in z.files:
            if k.startswith('recordings/'):
                a = z[k]
                ds = rec.create_dataset(k.split('/', 1)[1], shape=a.shape, maxshape=(None, None) + a.shape[2:], dtype='float64')
                if a.size:
                    ds[...] = a
    return h5_path


class RolloutRecorder:
    """Records a batched rollout straight from the env's packed GPU tensors: call `record(obs, action)` after every `env.step`,
    `flush()` appends the N trajectories (one per env) to the writer.  Episodes are NOT split at resets (the reference's
    `H5Writer` stores whatever trajectory the caller hands over); `terminated` is stored as an extra observation when asked."""

    def __init__(self, env, file_path, with_terminated: bool = False):
        self.env = env
        extra = {'terminated': (1,)} if with_terminated else None
        self.writer = RolloutWriter(file_path, env, extra_obs=extra)
        self.with_terminated = with_terminated
        self._obs, self._act, self._time, self._term = [], [], [], []

    def record(self, obs: dict, action, terminated=None):
        import torch
        self._obs.append({k: v.detach().clone() if hasattr(v, 'detach') else torch.as_tensor(np.asarray(v)) for k, v in obs.items()})
        self._act.append(torch.as_tensor(action).detach().clone().reshape(self.env.num_envs, -1))
        t = self.env.sim.sim_time.detach().clone()
        self._time.append(t)
        if self.with_terminated:
            self._term.append(torch.as_tensor(terminated).detach().clone().to(torch.float32))

    def flush(self):
        import torch
        if not self._obs:
            return
        n = self.env.num_envs
        traj = {k: torch.stack([o[k].reshape(n, -1) for o in self._obs], dim=1) for k in self._obs[0]}  # [N, T, dim]
        traj = {k: v.reshape(n, len(self._obs), *self.writer.shapes[k]) for k, v in traj.items()}
        traj['action'] = torch.stack(self._act, dim=1).to(torch.float64)
        if self.with_terminated:
            traj['terminated'] = torch.stack(self._term, dim=1).reshape(n, -1, 1)
        time = torch.stack(self._time, dim=1).reshape(n, -1, 1)
        self.writer.append_trajectories(traj, time)
        self._obs, self._act, self._time, self._term = [], [], [], []
