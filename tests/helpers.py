"""Shared helpers for the parity tests: seeded initial states and oracle rollouts."""
from __future__ import annotations

import numpy as np

from gym_quadruped_b200.model import Model
from oracle.oracle import Oracle


def seeded_states(model: Model, n: int, seed: int = 0, joint_noise=0.3, vel_noise=0.5, tilt=0.1, lift=True):
    """n initial states around the home keyframe, lifted out of ground contact by the oracle (quadruped_env.py:376-388)."""
    rng = np.random.RandomState(seed)
    key = np.array(model.c.key_qpos)
    qpos = np.tile(key, (n, 1))
    qvel = np.zeros((n, 18))
    o = Oracle(model)
    for i in range(n):
        qpos[i, 7:] += rng.uniform(-joint_noise, joint_noise, 12)
        qpos[i, 0:2] = rng.uniform(-2, 2, 2)
        r, p, y = rng.uniform(-tilt, tilt), rng.uniform(-tilt, tilt), rng.uniform(-np.pi, np.pi)
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
        qpos[i, 3:7] = [cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy]
        qvel[i, 6:] = rng.uniform(-vel_noise, vel_noise, 12)
        if lift:
            o.set_state(qpos[i], np.zeros(18), np.zeros(18))
            assert o.lift() >= 0
            qpos[i] = o.get_state()[0]
    # make every value exactly representable in fp32 so oracle and kernel start from identical numbers
    return qpos.astype(np.float32).astype(np.float64), qvel.astype(np.float32).astype(np.float64)


def oracle_rollout(model: Model, qpos, qvel, ctrl_seq, mu=(-1.0, -1.0), command=(0, 0, 0, 0)):
    """ctrl_seq [T, n, 12] -> dict of per-step arrays from n independent oracle envs."""
    T, n, _ = ctrl_seq.shape
    out = {k: [] for k in ('qpos', 'qvel', 'obs', 'term', 'cstate', 'invalid', 'ncon')}
    envs = []
    for i in range(n):
        o = Oracle(model)
        o.set_state(qpos[i], qvel[i], np.zeros(18))
        o.set_env(mu[0], mu[1], command)
        envs.append(o)
    for t in range(T):
        rec = {k: [] for k in out}
        for i, o in enumerate(envs):
            obs, term = o.step(ctrl_seq[t, i])
            qp, qv, _, _ = o.get_state()
            f = o.flags()
            rec['qpos'].append(qp); rec['qvel'].append(qv); rec['obs'].append(obs[:227]); rec['term'].append(term)
            rec['cstate'].append(f['contact_state']); rec['invalid'].append(f['invalid_body_mask']); rec['ncon'].append(f['ncon'])
        for k in out:
            out[k].append(np.array(rec[k]))
    return {k: np.array(v) for k, v in out.items()}
