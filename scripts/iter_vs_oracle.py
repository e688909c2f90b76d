"""Diagnostic: Newton iterations of the fp32 kernel vs the fp64 oracle on identical states of a random-action rollout."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
from oracle.oracle import Oracle
import bench
robot = os.environ.get('QS_ROBOT', 'hyqreal1')
m = Model(robot, 'flat'); n = 256
sim = BatchSim(m, n, device=0); opt = sim.make_reset_options(**bench.RESET_KW); sim.reset(options=opt)
g = torch.Generator(device='cuda').manual_seed(0)
scale = float(os.environ.get('QS_SCALE', '50'))
for t in range(150): sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * scale, opt)
gi, oi = [], []
orc = [Oracle(m) for _ in range(n)]
for t in range(6):
    q0 = sim.qpos.cpu().numpy().astype(np.float64); q0[:, :3] = sim.base_pos64.cpu().numpy()
    v0 = sim.qvel.cpu().numpy().astype(np.float64); w0 = sim.qacc_warmstart.cpu().numpy().astype(np.float64)
    fr = sim.friction.cpu().numpy(); cmd = sim.command.cpu().numpy()
    a = torch.randn(n, 12, device='cuda', generator=g) * scale
    sim.step(a); ac = a.cpu().numpy().astype(np.float64)
    it = (sim.solver_iter & 255).cpu().numpy()
    for i, o in enumerate(orc):
        o.set_state(q0[i], v0[i], w0[i]); o.set_env(float(fr[i, 0]), float(fr[i, 1]), cmd[i].astype(np.float64)); o.step(ac[i])
        gi.append(int(it[i])); oi.append(o.flags()['solver_iter'])
    sim.reset_done(opt)
gi, oi = np.array(gi), np.array(oi)
print(robot, 'mean iters gpu', gi.mean(), 'oracle', oi.mean(), 'max', gi.max(), oi.max())
print('hist gpu   ', np.bincount(gi, minlength=12)[:18])
print('hist oracle', np.bincount(oi, minlength=12)[:18])
print('gpu - oracle:', np.bincount(np.clip(gi - oi + 5, 0, 10), minlength=11), '(index 5 = equal)')
