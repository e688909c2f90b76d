import ctypes as C, sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench
import os
model = Model(os.environ.get('QS_ROBOT', 'mini_cheetah'), 'flat'); n=4096
sim = BatchSim(model, n, device=0); opt = sim.make_reset_options(**bench.RESET_KW); sim.reset(options=opt)
prof = torch.zeros(n * 32, dtype=torch.int32, device='cuda')
sim.L.qs_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device='cuda').manual_seed(0)
for t in range(300): sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
sim.L.qs_debug_set_prof(sim.h, C.c_void_p(prof.data_ptr()))
rows=[]
for t in range(30):
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt); torch.cuda.synchronize()
    raw = prof.cpu().numpy()
    P = raw.reshape(n,32); S = P[:,16:24].copy().view(np.float32); P = P[:, 16:].copy(); P[:, 8] = raw.reshape(n,32)[:,24]
    it = P[:,8]
    for i in np.where(it>=4)[0]: rows.append((it[i], S[i].copy()))
print('envs with iters>=4:', len(rows))
for it_, s_ in rows[:40]:
    print(it_, ' '.join('%.2e'%x for x in s_[:it_]))
# how many iterations would be saved with larger tolerance
for tol in (1e-6, 1e-5, 1e-4, 1e-3):
    saved=0; tot=0
    for it_, s_ in rows:
        tot+=it_
        k = next((j for j in range(min(it_,8)) if s_[j] < tol), it_)
        saved += it_- (k+1) if k<it_ else 0
    print('tol',tol,'iterations total',tot,'saved',saved)
