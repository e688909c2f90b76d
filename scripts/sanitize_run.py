"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): specialised + generic step kernels, pipelined launches
(finish-order queues, programmatic dependent launch), auto-reset, in-kernel schedules, mesh-vs-terrain collider, forward, raycast."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gym_quadruped_b200.backend import BatchSim, FIELD_CONTACTS
from gym_quadruped_b200.model import Model

cases = (('mini_cheetah', 'flat', True), ('go2', 'random_boxes', True), ('aliengo', 'perlin', False), ('mini_cheetah', 'perlin', True),
         ('hyqreal1', 'random_boxes', False))
for robot, scene, pipeline in cases:
    m = Model(robot, scene)
    flat_fast = scene == 'flat'
    sim = BatchSim(m, 96, device=0, use_imu=bool(m.c.has_imu) and not flat_fast, heightmap=None if flat_fast else (3, 3, 0.1, 0.1), pipeline=pipeline)
    sim.set_schedule(command_mode=1 | 4 | 8, lin_vel_range=(0.3, 0.9), ang_vel_range=(-0.4, 0.4), ext_enabled=True, ext_ranges={'x': (-30, 30), 'z': (5,)})
    opt = sim.make_reset_options(friction_range=(0.2, 1.5), command_mode=1 | 4 | 8)
    sim.reset(options=opt)
    sim.cmd_limit[:] = 3; sim.ext_limit[:] = 4
    if scene != 'flat':
        q = sim.qpos.clone(); q[:, 0] = 2.0; q[:, 1] = -1.0 if scene == 'random_boxes' else 2.0; q[:, 2] = 0.5 if scene == 'random_boxes' else 0.9
        sim.set_state(q, sim.qvel)
    g = torch.Generator(device='cuda').manual_seed(0)
    T = int(os.environ.get('SAN_STEPS', '16'))
    ctrl = torch.randn(T, 96, 12, device='cuda', generator=g) * 30
    torch.cuda.synchronize()
    for t in range(T):
        sim.step_autoreset(ctrl[t], opt)  # back-to-back launches: chained when pipeline=True
    sim.forward(); sim.get(FIELD_CONTACTS)
    torch.cuda.synchronize()
    print(robot, scene, sim.step_variant, 'pipeline' if pipeline else 'serialized', 'ok', float(sim.obs.abs().max()), int(sim.ncon.max()), flush=True)
