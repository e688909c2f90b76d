"""The C-ABI library exports what include/qstep.h declares; host-side mirrors of the reference's containers; env sharding."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared_functions():
    text = (ROOT / 'include' / 'qstep.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qs_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ROOT / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so'
    if not lib.exists():
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run(['nm', '-D', '--defined-only', str(lib)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (qs_[a-z_0-9]+)', out))
    declared = _declared_functions()
    assert len(declared) >= 15
    missing = [f for f in declared if f not in exported]
    assert not missing, f'declared in qstep.h but not exported: {missing}'


def test_library_loads_and_struct_sizes_agree():
    from gym_quadruped_b200 import backend
    from gym_quadruped_b200.model import QsBuffers, QsConfig, QsModel
    L = backend.load_library()  # dlopen only: no CUDA call is made without a GPU
    assert L.qs_abi_version() == 5
    assert L.qs_model_sizeof() == ctypes.sizeof(QsModel) and L.qs_config_sizeof() == ctypes.sizeof(QsConfig)
    assert L.qs_buffers_sizeof() == ctypes.sizeof(QsBuffers)
    cfg = QsConfig(); cfg.use_imu = 1
    assert L.qs_obs_dim(ctypes.byref(cfg)) == 227 + 18


def test_product_refuses_to_run_without_cuda():
    """No CPU fallback: the env must fail loudly when no GPU is visible (this test only runs on the CPU-only host)."""
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from gym_quadruped_b200.quadruped_env import QuadrupedEnv
    with pytest.raises(RuntimeError, match='CUDA'):
        QuadrupedEnv('mini_cheetah')


def test_product_does_not_import_the_oracle():
    for py in (ROOT / 'gym_quadruped_b200').rglob('*.py'):
        assert not re.search(r'^\s*(from|import)\s+oracle\b', py.read_text(), flags=re.M), f'{py} imports the oracle'
    for src in (ROOT / 'gym_quadruped_b200' / 'csrc').glob('*'):
        if src.suffix in ('.cu', '.cuh', '.h'):
            assert 'oracle' not in src.read_text().lower().replace('fp64 oracle', ''), f'{src} references the oracle'


def test_legs_attr_and_observation_space():
    from gym_quadruped_b200.model import load_robot_tables
    from gym_quadruped_b200.quadruped_env import OBS_LAYOUT, QuadrupedEnv
    from gym_quadruped_b200.utils.quadruped_utils import LegsAttr, configure_observation_space
    a = LegsAttr(FR=np.array([1.0, 0, 0]), FL=np.array([2.0, 0, 0]), RR=np.array([3.0, 0, 0]), RL=np.array([4.0, 0, 0]))
    assert [x[0] for x in a.to_list()] == [2, 1, 4, 3]                      # default order FL, FR, RL, RR
    assert [x[0] for x in a.to_list(order=['FR', 'FL', 'RR', 'RL'])] == [1, 2, 3, 4]
    assert (a + a).FL[0] == 4 and (a - a).RR[0] == 0 and (a / 2).RL[0] == 2 and a['FR'][0] == 1
    a['FR'] = np.zeros(3)
    assert a.FR[0] == 0 and len(list(iter(a))) == 4
    with pytest.raises(TypeError):
        a + 'x'
    space = configure_observation_space(load_robot_tables('aliengo'), QuadrupedEnv.ALL_OBS)
    assert len(space.keys()) == 31 and sum(s.shape[0] for s in space.values()) == 227
    assert OBS_LAYOUT['contact_state'] == (199, 4) and OBS_LAYOUT['qpos'] == (52, 19) and OBS_LAYOUT['contact_forces:base'] == (215, 12)
    assert space['qpos_js'].high[0] == np.float32(1.22173) and np.isinf(space['qvel'].high).all()
    with pytest.raises(ValueError):
        configure_observation_space(load_robot_tables('aliengo'), ['not_an_obs'])


def test_command_modes_robot_cfgs_and_action_space():
    from gym_quadruped_b200.backend import CMD_FORWARD, CMD_RANDOM, CMD_RESET, CMD_ROTATE, command_mode_bits
    from gym_quadruped_b200.robot_cfgs import get_robot_config
    from gym_quadruped_b200.spaces import Box
    assert command_mode_bits('forward+rotate') == CMD_FORWARD | CMD_ROTATE
    assert command_mode_bits('random+reset') == CMD_RANDOM | CMD_RESET and command_mode_bits('human') == 0
    with pytest.raises(ValueError):
        command_mode_bits('sideways')
    assert get_robot_config('mini_cheetah').hip_height == 0.225 and get_robot_config('hyqreal1').tables == 'hyqreal1'
    with pytest.raises(ValueError):
        get_robot_config('hyqreal')
    with pytest.raises(NotImplementedError):
        get_robot_config('pegasus')
    assert [get_robot_config(r).tables for r in ('b2', 'go1', 'go2', 'hyqreal2', 'aliengo', 'spot')] == ['b2', 'go1', 'go2', 'hyqreal2', 'aliengo', 'spot']
    box = Box(low=-np.inf, high=np.inf, shape=(12,), dtype=np.float32)
    s = box.sample()
    assert s.shape == (12,) and s.dtype == np.float32 and np.abs(s).max() < 10  # unbounded Box samples N(0,1) (App. B.6)


def test_env_sharding_plan():
    from gym_quadruped_b200.distributed import shard_range
    for total, world in ((16384, 4), (65536, 8), (10, 3)):
        spans = [shard_range(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gym_quadruped_b200.distributed import gather_rows, shard_range
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
a, b = shard_range(r, 2, 8)
local = torch.arange(a, b, dtype=torch.float32).unsqueeze(1).repeat(1, 3)       # rows carry their global env id
full = gather_rows(local)
assert full.shape == (8, 3) and torch.equal(full[:, 0], torch.arange(8, dtype=torch.float32)), full
dist.barrier(); dist.destroy_process_group()
print('ok', r)
'''


def test_obs_gather_two_ranks_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err[-2000:]
        assert 'ok' in out


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / 'bench.py'), '--impl', 'reference', '--steps', '3', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'env-steps/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('mini_cheetah/flat') and line['n_gpus'] == 1 and line['steps'] == 3


def test_lane_role_table_of_the_kernel_source():
    """qs_env.cuh tabulates, per lane, the triangle coordinates of its Hessian entries and its dof role; recompute them here."""
    src = (Path(__file__).resolve().parents[1] / 'gym_quadruped_b200' / 'csrc' / 'qs_env.cuh').read_text()
    table = [int(x) for x in re.search(r'kLaneRoles\[32\] = \{([^}]*)\}', src).group(1).split(',')]
    assert len(table) == 32
    for lane, v in enumerate(table):
        (i0, j0), (i1, j1) = [next((i, e - i * (i + 1) // 2) for i in range(12) if i * (i + 1) // 2 <= e < (i + 1) * (i + 2) // 2)
                              for e in (lane, lane + 32)]
        dl, dk = ((lane - 6) // 3, (lane - 6) % 3) if 6 <= lane < 18 else (0, 0)
        assert v == i0 | (j0 << 4) | (i1 << 8) | (j1 << 12) | (dl << 16) | (dk << 18) | ((lane // 6) << 20) | ((lane % 6) << 23)


def test_bench_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) must work on a CPU-only host and print one
    JSON line whose `config` has the keys of the GPU arm's (the driver compares the two objects)."""
    import json
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1', '--envs', '32'],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['unit'] == 'env-steps/s' and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    sys.path.insert(0, str(ROOT))
    import bench
    bench.select_workload('cfg2')
    assert set(line['config']) == set(bench.workload_config(1, 32))
