"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): step + auto-reset + forward + raycast, 3 robots."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gym_quadruped_b200.backend import BatchSim, FIELD_CONTACTS
from gym_quadruped_b200.model import Model

for robot, scene in (('mini_cheetah', 'flat'), ('go2', 'random_boxes'), ('aliengo', 'perlin')):
    m = Model(robot, scene)
    sim = BatchSim(m, 96, device=0, use_imu=bool(m.c.has_imu), heightmap=(3, 3, 0.1, 0.1))
    opt = sim.make_reset_options(friction_range=(0.2, 1.5))
    sim.reset(options=opt)
    if scene != 'flat':
        q = sim.qpos.clone(); q[:, 0] = 2.0; q[:, 1] = -1.0 if scene == 'random_boxes' else 2.0; q[:, 2] = 0.5 if scene == 'random_boxes' else 0.9
        sim.set_state(q, sim.qvel)
    g = torch.Generator(device='cuda').manual_seed(0)
    for t in range(12):
        sim.step_autoreset(torch.randn(96, 12, device='cuda', generator=g) * 30, opt)
    sim.forward(); sim.get(FIELD_CONTACTS)
    torch.cuda.synchronize()
    print(robot, scene, 'ok', float(sim.obs.abs().max()), int(sim.ncon.max()))
