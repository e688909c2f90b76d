import ctypes as C, sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from gym_quadruped_b200.backend import BatchSim
from gym_quadruped_b200.model import Model
import bench
model = Model('aliengo', 'random_boxes'); n=4096
sim = BatchSim(model, n, device=0); opt = sim.make_reset_options(**bench.RESET_KW); sim.reset(options=opt)
prof = torch.zeros(n * 32, dtype=torch.int32, device='cuda')
sim.L.qs_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device='cuda').manual_seed(0)
for t in range(100): sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt)
sim.L.qs_debug_set_prof(sim.h, C.c_void_p(prof.data_ptr()))
for t in range(5):
    prof.zero_()
    sim.step_autoreset(torch.randn(n, 12, device='cuda', generator=g) * 50, opt); torch.cuda.synchronize()
    P = prof.cpu().numpy().reshape(n,32)
    lift = P[:,29]; pen = P[:,30].copy().view(np.float32)
    tend = P[:,15].astype(np.int64) & 0xffffffff
    idx = np.where(lift>0)[0]
    print('step',t,'resets with lift iterations:',len(idx),'iters',lift[idx][:20],'last pen',pen[idx][:8],'end cycles of those',tend[idx][:10], 'makespan',tend.max(), 'terminated', int(sim.terminated.sum()))
    print('   slowest envs', np.argsort(-tend)[:5], tend[np.argsort(-tend)[:5]], 'their lift', lift[np.argsort(-tend)[:5]])
