"""Diagnostic: distribution of Newton iterations / contacts per env-step in the benchmark workload."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench

model = Model('mini_cheetah', 'flat')
n = 4096
sim = BatchSim(model, n, device=0)
opt = sim.make_reset_options(**bench.RESET_KW)
sim.reset(options=opt)
g = torch.Generator(device='cuda').manual_seed(0)
hist = torch.zeros(64, dtype=torch.long, device='cuda'); lhist = torch.zeros(128, dtype=torch.long, device='cuda'); chist = torch.zeros(32, dtype=torch.long, device='cuda')
maxed = 0; term = 0
T = 600
for t in range(T):
    a = torch.randn(n, 12, device='cuda', generator=g) * 50
    sim.step_autoreset(a, opt)
    if t >= 100:
        hist += torch.bincount((sim.solver_iter & 255).clamp(0, 63), minlength=64); lhist += torch.bincount((sim.solver_iter >> 8).clamp(0, 127), minlength=128)
        chist += torch.bincount(sim.ncon.clamp(0, 31), minlength=32)
        maxed += int((sim.status & 4).ne(0).sum()); term += int(sim.terminated.sum())
tot = hist.sum().item()
print('solver_iter histogram (fraction):', {i: round(h / tot, 5) for i, h in enumerate(hist.tolist()) if h})
print('mean iters', (hist * torch.arange(64, device='cuda')).sum().item() / tot)
print('ls evals per env-step:', {i: round(h / tot, 5) for i, h in enumerate(lhist.tolist()) if h})
print('ncon histogram:', {i: round(h / tot, 5) for i, h in enumerate(chist.tolist()) if h})
print('maxed fraction', maxed / tot, 'terminated fraction', term / tot)
