"""Diagnostic: step time of robot/scene variants with and without auto-reset."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench

def run(robot, scene, n, autoreset, steps=200, **kw):
    m = Model(robot, scene)
    sim = BatchSim(m, n, device=0, **kw)
    opt = sim.make_reset_options(**bench.RESET_KW)
    sim.reset(options=opt)
    g = torch.Generator(device='cuda').manual_seed(0)
    acts = torch.randn(64, n, 12, device='cuda', generator=g) * 50
    for i in range(50):
        sim.step_autoreset(acts[i % 64], opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if autoreset:
            sim.step_autoreset(acts[i % 64], opt)
        else:
            sim.step(acts[i % 64])
    e1.record(); torch.cuda.synchronize()
    it = (sim.solver_iter & 255).float()
    print(f'{robot:13s} {scene:13s} n={n} autoreset={autoreset}: {e0.elapsed_time(e1)/steps*1e3:8.1f} us/step  mean iters {it.mean():.2f} max {it.max():.0f} '
          f'ncon mean {sim.ncon.float().mean():.2f} max {sim.ncon.max().item()} status!=0 {(sim.status != 0).sum().item()} lib warps? ', flush=True)

for robot, scene in (('mini_cheetah', 'flat'), ('go2', 'flat'), ('go2', 'random_boxes'), ('aliengo', 'random_boxes'), ('aliengo', 'perlin'), ('hyqreal1', 'flat')):
    for ar in (True, False):
        run(robot, scene, 4096, ar)
