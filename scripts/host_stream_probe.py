"""Diagnostic: open-loop rollout with HOST-resident actions and observations streamed by qs_step_k (one call, K overlapped launches;
the kernels read ctrl[K,N,12] from and write obs[K,N,D] to pinned host memory through their mapped addresses).  Every step's
inputs and outputs cross PCIe, but step t+1 computes while step t's rows are still in flight -- the PCIe-bound limit of the path,
next to the synchronous single-step `e2e` of bench.py (each step waits for its own rows)."""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model

n, K = 4096, 64
sim = BatchSim(Model('mini_cheetah', 'flat'), n, device=0)
opt = sim.make_reset_options(**bench.RESET_KW)
sim.reset(options=opt)
act = torch.randn(64, n, 12, device='cuda') * 50
for i in range(300):
    sim.step_autoreset(act[i % 64], opt)
torch.cuda.synchronize()
D = sim.obs_dim
ctrl_h = (torch.randn(K, n, 12) * 50).pin_memory()
obs_h = torch.empty(K, n, D).pin_memory()
term_h = torch.empty(K, n, dtype=torch.uint8).pin_memory()
L = sim.L
stream = sim._stream()


def run():
    sim._check(L.qs_step_k(sim.h, K, C.c_void_p(ctrl_h.data_ptr()), C.byref(opt), C.c_void_p(obs_h.data_ptr()), C.c_size_t(n * D), None,
                           C.c_void_p(term_h.data_ptr()), None, stream))
    torch.cuda.synchronize()


run(); run()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
t = sorted(ts)[len(ts) // 2]
print(f'host-streamed qs_step_k: K={K}, {n} envs: {t / K * 1e6:.1f} us/step, {n * K / t / 1e6:.2f} M env-steps/s, '
      f'{(n * D * 4 + n * 48) * K / t / 1e9:.1f} GB/s over PCIe; finite={bool(torch.isfinite(obs_h).all())}, last row max {float(obs_h[-1].abs().max()):.1f}')
