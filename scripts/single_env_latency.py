"""BASELINE configs[0]: mini_cheetah / flat / ONE env, random actions, the loop of the reference's tests/env_test.py:29-51
(step, reset when terminated).  Prints the steps per second of
  * QuadrupedEnv.step through the single-env path (qs_step_host with pinned one-row buffers, reference return types),
  * the bare C-ABI call (qs_step_host) underneath it,
  * the kernel alone (device-resident buffers, CUDA events),
  * the fp64 oracle port on one host core (what a CPU engine costs for one env; not MuJoCo).
A batch engine cannot win this configuration: one warp runs ~7.7 k dependent instructions and the call synchronises every step."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gym_quadruped_b200.model import Model  # noqa: E402
from gym_quadruped_b200.quadruped_env import QuadrupedEnv  # noqa: E402

ALL_OBS = QuadrupedEnv.ALL_OBS if hasattr(QuadrupedEnv, 'ALL_OBS') else None
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.RandomState(0)
acts = rng.randn(STEPS, 12) * 50.0

env = QuadrupedEnv('mini_cheetah', scene='flat', **({'state_obs_names': tuple(ALL_OBS)} if ALL_OBS else {}))
env.reset(random=True)
for a in acts[:200]:
    _, _, term, _, _ = env.step(a)
    if term:
        env.reset(random=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
resets = 0
for a in acts:
    _, _, term, _, _ = env.step(a)
    if term:
        env.reset(random=True); resets += 1
torch.cuda.synchronize()
t_env = (time.perf_counter() - t0) / STEPS

sim, h = env.sim, env._host
t0 = time.perf_counter()
for a in acts:
    h['ctrl_np'][0, :] = a
    sim.step_host(h['ctrl'], h['obs'], h['rew'], h['term'], h['trunc'], auto_reset=sim.reset_options)
t_abi = (time.perf_counter() - t0) / STEPS

ctrl = torch.as_tensor(acts, dtype=torch.float32, device='cuda').reshape(STEPS, 1, 12)
for i in range(50):
    sim.step_autoreset(ctrl[i], sim.reset_options)
torch.cuda.synchronize()
a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a_.record()
for i in range(STEPS):
    sim.step_autoreset(ctrl[i], sim.reset_options)
b_.record()
torch.cuda.synchronize()
t_kernel = a_.elapsed_time(b_) / STEPS * 1e-3

from oracle.oracle import Oracle, build  # noqa: E402  (diagnostic script: the oracle is the CPU yardstick here)
build()
m = Model('mini_cheetah', 'flat')
o = Oracle(m)
q = np.array(m.c.key_qpos); q[2] = m.hip_height
o.set_state(q, np.zeros(18), np.zeros(18)); o.lift()
start = o.get_state()
t0 = time.perf_counter()
for a in acts:
    _, term = o.step(a)
    if term:
        o.set_state(start[0], np.zeros(18), np.zeros(18))
t_orc = (time.perf_counter() - t0) / STEPS

print(json.dumps({'config': 'BASELINE configs[0]: mini_cheetah / flat / 1 env / random actions x50', 'steps': STEPS, 'resets_in_env_loop': resets,
                  'QuadrupedEnv.step_us': round(t_env * 1e6, 1), 'QuadrupedEnv.steps_per_s': round(1 / t_env),
                  'qs_step_host_us': round(t_abi * 1e6, 1), 'kernel_back_to_back_us': round(t_kernel * 1e6, 1),
                  'oracle_fp64_one_core_us': round(t_orc * 1e6, 1), 'oracle_steps_per_s': round(1 / t_orc)}))
