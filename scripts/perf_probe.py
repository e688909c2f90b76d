"""Quick device-side timing of the step kernel variants / launch modes (not the bench): env-steps/s for each BASELINE workload,
serialized vs pipelined launches, specialised vs generic kernel.  Usage: python scripts/perf_probe.py [cfg2 cfg3 ...] [--steps K]"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import RESET_KW, WORKLOADS  # noqa: E402
from gym_quadruped_b200.backend import BatchSim  # noqa: E402
from gym_quadruped_b200.model import Model  # noqa: E402


def run(name, pipeline, generic, steps, warmup, envs=None):
    wl = WORKLOADS[name]
    if generic:
        os.environ['QSTEP_GENERIC'] = '1'
    else:
        os.environ.pop('QSTEP_GENERIC', None)
    n = envs or wl['envs']
    sim = BatchSim(Model(wl['robot'], wl['scene']), n, device=0, seed=0, use_imu=wl['imu'], heightmap=wl['hm'], pipeline=pipeline)
    opt = sim.make_reset_options(**RESET_KW)
    sim.reset(options=opt)
    ring = 64
    actions = torch.randn(ring, n, 12, device='cuda:0', generator=torch.Generator(device='cuda:0').manual_seed(0)) * 50
    for i in range(warmup):
        sim.step_autoreset(actions[i % ring], opt)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            sim.step_autoreset(actions[i % ring], opt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = ms if best is None else min(best, ms)
    it = sim.solver_iter.cpu() & 0xff
    res = dict(workload=name, variant=sim.step_variant, pipeline=pipeline, envs=n, ms_per_step=round(best, 5), msteps_per_s=round(n / best / 1e3, 2),
               mean_iter=round(float(it.float().mean()), 2), max_iter=int(it.max()), mean_ncon=round(float(sim.ncon.float().mean()), 2),
               status_or=int(sim.status.max()))
    sim.close()
    return res


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if a in WORKLOADS]
    steps = int(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else 500
    names = args or ['cfg2', 'cfg3', 'cfg4', 'cfg5']
    for name in names:
        for pipeline, generic in ((False, True), (False, False), (True, False)):
            print(json.dumps(run(name, pipeline, generic, steps, 100)), flush=True)
