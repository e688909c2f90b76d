"""Diagnostic: where does the end-to-end (host buffer) step time go?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model

n = 4096
sim = BatchSim(Model('mini_cheetah', 'flat'), n, device=0)
opt = sim.make_reset_options(**bench.RESET_KW)
sim.reset(options=opt)
dev = torch.device('cuda:0')
act = torch.randn(64, n, 12, device=dev) * 50
for i in range(300):
    sim.step_autoreset(act[i % 64], opt)  # steady-state rollout
def timeit(fn, k=300):
    for i in range(20): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
print('device step, async            us', timeit(lambda i: sim.step_autoreset(act[i % 64], opt)))
print('device step + sync every step us', timeit(lambda i: (sim.step_autoreset(act[i % 64], opt), torch.cuda.synchronize())))
for pinned in (True, False):
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    ctrl_h = [mk(torch.randn(n, 12) * 50) for _ in range(8)]
    obs_h, rew_h = mk(torch.empty(n, sim.obs_dim)), mk(torch.empty(n))
    term_h, trunc_h = mk(torch.empty(n, dtype=torch.uint8)), mk(torch.empty(n, dtype=torch.uint8))
    print('step_host pinned=%s           us' % pinned, timeit(lambda i: sim.step_host(ctrl_h[i % 8], obs_h, rew_h, term_h, trunc_h, auto_reset=opt)))
ctrl_h = [(torch.randn(n, 12) * 50).pin_memory() for _ in range(8)]
rew_h = torch.empty(n).pin_memory(); term_h = torch.empty(n, dtype=torch.uint8).pin_memory(); trunc_h = torch.empty(n, dtype=torch.uint8).pin_memory()
for width in (sim.obs_dim, 228, 256):
    wide = torch.empty(n, width).pin_memory()
    view = wide[:, :sim.obs_dim]
    print('step_host pinned, row stride %d floats us' % width, timeit(lambda i: sim.step_host(ctrl_h[i % 8], view, rew_h, term_h, trunc_h, auto_reset=opt)))
# kernel duration when the obs rows are written straight into pinned host memory (zero-copy) vs device memory
import ctypes as C
obs_h = torch.empty(n, sim.obs_dim).pin_memory()
def zc(i):
    sim._check(sim.L.qs_step_autoreset(sim.h, act[i % 64].data_ptr(), C.byref(opt), obs_h.data_ptr(), sim.reward.data_ptr(),
                                      sim.terminated.data_ptr(), sim.truncated.data_ptr(), sim._stream()))
for name, fn in (('device obs', lambda i: sim.step_autoreset(act[i % 64], opt)), ('zero-copy obs', zc)):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for i, (a, b) in enumerate(ev):
        a.record(); fn(i); b.record()
    torch.cuda.synchronize()
    print(name, 'kernel us', sum(a.elapsed_time(b) for a, b in ev) / len(ev) * 1e3)
# plain D2H copy of the obs tensor for reference
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
for a, b in ev:
    a.record(); obs_h.copy_(sim.obs, non_blocking=True); b.record()
torch.cuda.synchronize()
print('cudaMemcpyAsync D2H obs us', sum(a.elapsed_time(b) for a, b in ev) / len(ev) * 1e3)
