"""Soak test of the pipelined step launches: long back-to-back runs on several workloads, checked against a serialized twin."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model

for name, steps in (('cfg2', 20000), ('cfg4', 3000), ('cfg3', 3000), ('cfg5', 3000)):
    wl = bench.WORKLOADS[name]
    m = Model(wl['robot'], wl['scene'])
    a = BatchSim(m, wl['envs'], device=0, seed=1, use_imu=wl['imu'], heightmap=wl['hm'], pipeline=True)
    b = BatchSim(m, wl['envs'], device=0, seed=1, use_imu=wl['imu'], heightmap=wl['hm'], pipeline=False)
    opt = a.make_reset_options(**bench.RESET_KW)
    a.reset(options=opt); b.reset(options=opt)
    act = torch.randn(64, wl['envs'], 12, device='cuda:0', generator=torch.Generator(device='cuda:0').manual_seed(0)) * 50
    t0 = time.perf_counter()
    for i in range(steps):
        a.step_autoreset(act[i % 64], opt)
    torch.cuda.synchronize()
    ta = time.perf_counter() - t0
    check = min(steps, 1500)
    a2 = BatchSim(m, wl['envs'], device=0, seed=1, use_imu=wl['imu'], heightmap=wl['hm'], pipeline=True)
    a2.reset(options=opt)
    for i in range(check):
        a2.step_autoreset(act[i % 64], opt); b.step_autoreset(act[i % 64], opt)
    torch.cuda.synchronize()
    same = all(torch.equal(getattr(a2, n), getattr(b, n)) for n in ('qpos', 'qvel', 'obs', 'terminated', 'command', 'friction', 'step_count'))
    print(name, f'{steps} pipelined steps in {ta:.2f} s ({wl["envs"] * steps / ta / 1e6:.1f} M env-steps/s), finite={bool(torch.isfinite(a.obs).all())}, '
          f'status_or={int(a.status.max())}, pipelined == serialized over {check} steps: {same}', flush=True)
    assert same and torch.isfinite(a.obs).all()
