import sys, os
sys.path.insert(0,'/root/repo')
import torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench
for robot, scene in (('mini_cheetah','flat'), ('aliengo','perlin')):
    for n in (2048, 4096, 8192, 16384):
        m = Model(robot, scene); sim = BatchSim(m, n, device=0)
        opt = sim.make_reset_options(**bench.RESET_KW); sim.reset(options=opt)
        g = torch.Generator(device='cuda').manual_seed(0)
        acts = torch.randn(32, n, 12, device='cuda', generator=g) * 50
        for i in range(60): sim.step_autoreset(acts[i % 32], opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): sim.step_autoreset(acts[i % 32], opt)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        print(f'{os.environ.get("QSTEP_LIB","default")[-12:]} {robot:12s} {scene:7s} n={n:6d}: {us:8.1f} us/step  {n/us:6.2f} M env-steps/s', flush=True)
        sim.close()
