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

import os
model = Model(os.environ.get('QS_ROBOT', 'mini_cheetah'), os.environ.get('QS_SCENE', 'flat'))
import os
n = int(os.environ.get('QS_N', '4096'))
sim = BatchSim(model, n, device=0)
opt = sim.make_reset_options(**bench.RESET_KW)
sim.reset(options=opt)
prof = torch.zeros(n, 32, dtype=torch.int32, device='cuda')
sim.L.qs_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device='cuda').manual_seed(0)
for t in range(300):
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
sim.L.qs_debug_set_prof(sim.h, C.c_void_p(prof.data_ptr()))
names = ['load', 'position', 'vel+M+constraints', 'solve', 'integrate', 'pack_obs', 'writeback', 'reset pass']
agg = []
for t in range(20):
    prof.zero_()
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
    torch.cuda.synchronize()
    P = prof.cpu().numpy().astype(np.int64) & 0xffffffff
    agg.append(P.copy())
    t1, t2, t3, t4, t5, t6, t7, tend = (P[:, k] for k in (1, 2, 3, 4, 5, 6, 7, 15))
    it, ls, nc, sm = P[:, 24], P[:, 25], P[:, 26], P[:, 27]
    early = t5 < t3  # early-out envs (terminated by the collision stage): mark 5 is set right after mark 2
    if t < 3:
        print(f'--- step {t}: makespan {tend.max()} cycles; mean end {tend.mean():.0f}; p50 {np.percentile(tend,50):.0f} p90 {np.percentile(tend,90):.0f} '
              f'p99 {np.percentile(tend,99):.0f}; early-out envs {early.sum()}')
        ok = ~early
        ph = [t1, t2 - t1, t3 - t2, t4 - t3, t6 - t4, t7 - t6, t5 - t7]
        for nm, d in zip(names, ph):
            d = d[ok]
            print(f'   {nm:28s} mean {d.mean():9.0f}  p99 {np.percentile(d,99):9.0f}  max {d.max():9.0f}')
        if early.any():
            print(f'   reset pass (early-out envs)  mean {(tend - t5)[early].mean():9.0f}  max {(tend - t5)[early].max():9.0f}')
        for i in np.argsort(-tend)[:6]:
            print(f'   slow env {i}: end {tend[i]} solve {t4[i]-t3[i]} iters {it[i]} ls {ls[i]} ncon {nc[i]} sm {sm[i]} early-out {int(early[i])}')
        lifts = P[:, 29]
        if (lifts > 0).any():
            sel = lifts > 0
            print(f'   lift loops: {int(sel.sum())} envs, iterations mean {lifts[sel].mean():.1f} max {lifts.max()}, cycles per lift iteration ~ {((tend - t5)[sel & early] / np.maximum(1, lifts[sel & early])).mean() if (sel & early).any() else 0:.0f}')
        smax = np.array([tend[sm == s].max() for s in np.unique(sm)])
        print(f'   per-SM end: mean {smax.mean():.0f} min {smax.min()} max {smax.max()}')
P = np.concatenate(agg)
ok = P[:, 5] > P[:, 3]
P = P[ok]
it = P[:, 24]; solve = P[:, 4] - P[:, 3]
A = np.vstack([np.ones_like(it), it, P[:, 25], P[:, 26]]).T.astype(float)
coef, *_ = np.linalg.lstsq(A, solve.astype(float), rcond=None)
print('solve cycles ~ %.0f + %.0f*iters + %.0f*ls_evals + %.0f*ncon' % tuple(coef))
for k in range(0, 10):
    sel = it == k
    if sel.sum():
        print(f'iters={k}: frac {sel.mean():.4f} solve mean {solve[sel].mean():.0f} end-of-pass0 mean {P[sel,5].mean():.0f}')
S = P[:, 16:24]
snames = ['pre-loop (M factor, warm start)', 'update+grad+convergence', 'build_hessian', 'factor_H', 'solve_H', 'pre line search', 'line search', 'move']
print('solver sub-phases, mean cycles per env-step | for envs with iters>=4:')
sel = it >= 4
for k, nm in enumerate(snames):
    print(f'   {nm:34s} {S[:, k].mean():9.0f} | {S[sel, k].mean():9.0f}  per-iter {S[sel, k].sum() / max(1, it[sel].sum()):8.0f}')
print('   ls evals per iteration (heavy):', P[sel, 25].sum() / max(1, it[sel].sum()))
