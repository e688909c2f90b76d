"""Diagnostic: per-warp cycle marks of the step kernel in the benchmark workload (needs a -DQS_PROF build of libqstep).

    nvcc ... -DQS_PROF -DQS_ONLY_F3 -o gym_quadruped_b200/csrc/libqstep_prof.so gym_quadruped_b200/csrc/qstep.cu
    QSTEP_LIB=gym_quadruped_b200/csrc/libqstep_prof.so python scripts/warp_timeline.py
"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench

model = Model('mini_cheetah', 'flat')
n = 4096
sim = BatchSim(model, n, device=0)
opt = sim.make_reset_options(**bench.RESET_KW)
sim.reset(options=opt)
prof = torch.zeros(n * 24, dtype=torch.int32, device='cuda')
sim.L.qs_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device='cuda').manual_seed(0)
for t in range(300):
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
sim.L.qs_debug_set_prof(sim.h, C.c_void_p(prof.data_ptr()))
names = ['load', 'position', 'vel+M+constraints', 'solve', 'integrate+obs+writeback', 'reset pass']
agg = []; sagg = []
for t in range(20):
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
    torch.cuda.synchronize()
    raw = prof.cpu().numpy().astype(np.int64) & 0xffffffff
    P = raw[:n * 16].reshape(n, 16); S = raw[n * 16:].reshape(n, 8)
    t1, t2, t3, t4, t5, t6 = (P[:, k] for k in (1, 2, 3, 4, 5, 6))
    it, ls, nc, sm = P[:, 8], P[:, 9], P[:, 10], P[:, 11]
    agg.append(P.copy()); sagg.append(S.copy())
    if t < 4:
        print(f'--- step {t}: makespan(max end) {t6.max()} cycles; mean end {t6.mean():.0f}; p50 {np.percentile(t6,50):.0f} p90 {np.percentile(t6,90):.0f} p99 {np.percentile(t6,99):.0f}')
        ph = [t1, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5]
        for nm, d in zip(names, ph):
            print(f'   {nm:28s} mean {d.mean():9.0f}  p99 {np.percentile(d,99):9.0f}  max {d.max():9.0f}')
        order = np.argsort(-t6)[:8]
        for i in order:
            print(f'   slow env {i}: end {t6[i]} pass0 {t5[i]} solve {t4[i]-t3[i]} iters {it[i]} ls {ls[i]} ncon {nc[i]} sm {sm[i]} reset {int(t6[i]-t5[i] > 2000)}')
        # per-SM makespan
        smax = np.array([t6[sm == s].max() for s in np.unique(sm)])
        print(f'   per-SM end: mean {smax.mean():.0f} min {smax.min()} max {smax.max()}')
P = np.concatenate(agg)
it = P[:, 8]; solve = P[:, 4] - P[:, 3]
A = np.vstack([np.ones_like(it), it, P[:, 9], P[:, 10]]).T.astype(float)
coef, *_ = np.linalg.lstsq(A, solve.astype(float), rcond=None)
print('solve cycles ~ %.0f + %.0f*iters + %.0f*ls_evals + %.0f*ncon' % tuple(coef))
for k in range(0, 10):
    sel = it == k
    if sel.sum():
        print(f'iters={k}: frac {sel.mean():.4f} solve mean {solve[sel].mean():.0f} end-of-pass0 mean {P[sel,5].mean():.0f}')
S = np.concatenate(sagg)
snames = ['pre-loop (M factor, warm start)', 'update+grad+convergence', 'build_hessian', 'factor_H', 'solve_H', 'pre line search', 'line search', 'move']
print('solver sub-phases, mean cycles per env-step | for envs with iters>=4:')
sel = it >= 4
for k, nm in enumerate(snames):
    print(f'   {nm:34s} {S[:, k].mean():9.0f} | {S[sel, k].mean():9.0f}  per-iter {S[sel, k].sum() / max(1, it[sel].sum()):8.0f}')
print('   ls evals per iteration (heavy):', P[sel, 9].sum() / max(1, it[sel].sum()))
