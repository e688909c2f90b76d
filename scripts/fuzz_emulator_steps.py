"""Randomised single-step comparison of the kernel source on the host emulator against the fp64 oracle (CPU only).

    python scripts/fuzz_emulator_steps.py <seed> <cases> [precision: 1 = fp64 arithmetic (default), 0 = fp32]

Random robot / scene / pose (arbitrary orientation, base from buried to airborne) / velocity / ctrl / friction; compares contact
count, body masks, out-of-bounds flag and the state (1e-7 in fp64, 2e-3 in fp32).  DESIGN.md quotes its results."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gym_quadruped_b200.model import Model
from oracle.oracle import Oracle, F_CONTACTS
from tests.emu.emu import emu_step

robots=['mini_cheetah','aliengo','go2','hyqreal1','hyqreal2','go1','b2','spot']
scenes=['flat','random_boxes','perlin','stairs']
rng=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; total=0; t0=time.time(); maxerr=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 200):
    robot=robots[rng.randint(len(robots))]; scene=scenes[rng.randint(len(scenes))]
    m=Model(robot,scene)
    key=np.array(m.c.key_qpos)
    q=key.copy()
    q[7:]+=rng.uniform(-0.6,0.6,12)
    if scene!='flat':
        q[0:2]=rng.uniform(-4,4,2)
    # random orientation: often upright-ish, sometimes anything
    if rng.rand()<0.6:
        ax=rng.randn(3); ax/=np.linalg.norm(ax); ang=rng.uniform(0,0.5)
    else:
        ax=rng.randn(3); ax/=np.linalg.norm(ax); ang=rng.uniform(0,np.pi)
    q[3:7]=np.r_[np.cos(ang/2), np.sin(ang/2)*ax]
    q[2]=rng.uniform(0.02, 1.2*m.hip_height+ (0.7 if scene in('perlin','stairs','random_boxes') else 0))
    v=rng.uniform(-2,2,18)
    ctrl=rng.randn(12)*30
    mu=rng.uniform(0.2,1.5)
    o=Oracle(m); o.set_state(q,v,np.zeros(18)); o.set_env(mu,mu,[0.3,0,0,0.1])
    obs,term=o.step(ctrl)
    f=o.flags()
    qo,vo,_,_=o.get_state()
    if not np.isfinite(qo).all(): continue
    e=emu_step(m,q,v,np.zeros(18),ctrl,mu,mu,[0.3,0,0,0.1],precision=int(sys.argv[3]) if len(sys.argv)>3 else 1,mode=1)
    total+=1
    if f['ncon']>16:
        if not e['overflow']: print('OVERFLOW FLAG MISMATCH',robot,scene,it); bad+=1
        # masks must still agree
        if e['invalid_mask']!=f['invalid_body_mask']: print('MASK MISMATCH under overflow',robot,scene,it,e['invalid_mask'],f['invalid_body_mask']); bad+=1
        continue
    err=max(np.abs(e['qpos']-qo).max(), np.abs(e['qvel']-vo).max())
    ok = e['ncon']==f['ncon'] and e['invalid_mask']==f['invalid_body_mask'] and e['contact_mask']==sum(int(b)<<i for i,b in enumerate(f['contact_state'])) and err<(1e-7 if len(sys.argv)<=3 or sys.argv[3]=="1" else 2e-3) and bool(e['oob'])==bool(f['out_of_bounds'])
    maxerr=max(maxerr,err if e['ncon']==f['ncon'] else 0)
    if not ok:
        bad+=1
        print('MISMATCH',robot,scene,'it',it,'ncon',e['ncon'],f['ncon'],'masks',e['invalid_mask'],f['invalid_body_mask'],'err',err,'iters',e['iters'],f['solver_iter'] if 'solver_iter' in f else None, 'z',q[2])
print('done',total,'cases, bad',bad,'maxerr',maxerr,'time',round(time.time()-t0,1))
