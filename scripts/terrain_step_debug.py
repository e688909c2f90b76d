import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from gym_quadruped_b200.backend import BatchSim, FIELD_CONTACTS
from gym_quadruped_b200.model import Model
from oracle.oracle import Oracle, F_CONTACTS
import os
robot, scene = os.environ.get('QS_ROBOT','aliengo'), os.environ.get('QS_SCENE','stairs')
xy, z0 = (float(os.environ.get('QS_X','1.6')), float(os.environ.get('QS_Y','0.0'))), float(os.environ.get('QS_Z','0.85'))
m = Model(robot, scene); n, T = 6, 220
rng = np.random.RandomState(3); key = np.array(m.c.key_qpos)
qpos = np.tile(key,(n,1)); qvel=np.zeros((n,18))
for i in range(n):
    qpos[i,0:2] = np.array(xy)+rng.uniform(-0.6,0.6,2); qpos[i,2]=z0; qpos[i,7:]+=rng.uniform(-0.15,0.15,12)
    o=Oracle(m); o.set_state(qpos[i],np.zeros(18),np.zeros(18)); assert o.lift()>=0; qpos[i]=o.get_state()[0]
orc=[Oracle(m) for _ in range(n)]
for o in orc: o.set_env(0.8,0.8,[0.5,0,0,0])
sim=BatchSim(m,n,device=0); sim.set_state(torch.tensor(qpos),torch.tensor(qvel)); sim.friction[:]=0.8
sim.command[:]=torch.tensor([0.5,0,0,0],device='cuda')
for t in range(T):
    q0=sim.qpos.cpu().numpy().astype(np.float64); q0[:,:3]=sim.base_pos64.cpu().numpy()
    v0=sim.qvel.cpu().numpy().astype(np.float64); w0=sim.qacc_warmstart.cpu().numpy().astype(np.float64)
    ctrl=(40*(key[7:]-q0[:,7:])-2*v0[:,6:]+rng.randn(n,12)*2).astype(np.float32)
    sim.forward(); gc_all=sim.get(FIELD_CONTACTS).cpu().numpy(); nc=sim.ncon.cpu().numpy().copy()
    sim.step(torch.tensor(ctrl,device='cuda'))
    q1=sim.qpos.cpu().numpy(); v1=sim.qvel.cpu().numpy()
    for i,o in enumerate(orc):
        o.set_state(q0[i],v0[i],w0[i]); o.step(ctrl[i].astype(np.float64))
        qo,vo,_,_=o.get_state()
        err=max(np.abs(q1[i]-qo).max(),0.1*np.abs(v1[i]-vo).max())
        if err>5e-5:
            oc=o.get(F_CONTACTS); gc=gc_all[i,:nc[i]]
            print('step',t,'env',i,'err',err,'ncon',nc[i],len(oc),'iters gpu',int(sim.solver_iter[i])&255,'orc',o.flags()['solver_iter'])
            ks=lambda c: np.lexsort((np.round(c[:,3],5),np.round(c[:,2],5),np.round(c[:,1],5),c[:,16]))
            oc=oc[ks(oc)]; gc=gc[ks(gc)]
            print(' geoms', oc[:,16].astype(int), gc[:,16].astype(int))
            print(' max pos diff', np.abs(oc[:,:4]-gc[:,:4]).max(), 'dist orc', np.round(oc[:,0],5), 'gpu', np.round(gc[:,0],5))
            print(' force orc', np.round(oc[:,13],2), 'gpu', np.round(gc[:,13],2))
            print(' fric orc', oc[:,18], 'gpu', gc[:,18], 'dim orc', oc[:,19], 'gpu', gc[:,19])
            print(' ft orc', np.round(oc[:,14:16],3).tolist(), 'gpu', np.round(gc[:,14:16],3).tolist())
            raise SystemExit
print('no big error')
