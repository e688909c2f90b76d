import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from gym_quadruped_b200.quadruped_env import QuadrupedEnv
env = QuadrupedEnv('mini_cheetah', scene='flat', state_obs_names=tuple(QuadrupedEnv.ALL_OBS), base_vel_command_type='forward+rotate',
                   ref_base_lin_vel=(0.5, 1.0), ground_friction_coeff=(0.2, 1.5))
env.reset()
for _ in range(50):
    obs, r, term, trunc, info = env.step(env.action_space.sample() * 50)
    if term: env.reset()
t0 = time.perf_counter(); n = 2000; resets = 0
for _ in range(n):
    obs, r, term, trunc, info = env.step(env.action_space.sample() * 50)
    if term: env.reset(); resets += 1
dt = time.perf_counter() - t0
print(f'single-env QuadrupedEnv.step: {n/dt:.0f} steps/s ({dt/n*1e6:.0f} us/step), resets {resets}')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    obs, r, term, trunc, info = env.step(env.action_space.sample() * 50)
    if term: env.reset()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
