"""GPU parity tests added in round 2 (through the C-ABI, against the fp64 oracle on identical inputs):

* specialised kernel variants == generic variant, pipelined (overlapping) step launches == serialized launches, bit for bit;
* north_star's bar on the elliptic-cone robots of BASELINE configs 4 / 5: open-loop 100 steps at ctrl = 50*N(0,1);
* open-loop rollouts on the terrain scenes of configs 3 / 4 with a written tie-break rule for contact events;
* the observation row at EVERY step; 64 envs sampled from the full 4096-env batch;
* `out_of_bounds` termination (quadruped_env.py:1250-1257), `qfrc_applied` (:299-305), the in-kernel command / disturbance
  schedules (:293-305, :1046-1139) and `reset(seed=...)`.
"""
import os

import numpy as np
import pytest
import torch

from gym_quadruped_b200.backend import CMD_FORWARD, CMD_RANDOM, CMD_RESET, CMD_ROTATE, BatchSim
from gym_quadruped_b200.model import Model
from oracle.oracle import F_CONTACTS, Oracle
from tests.helpers import oracle_rollout, seeded_states

pytestmark = pytest.mark.gpu

CONFIGS = [('mini_cheetah', 'flat', False, None, 'f3_pyr_flat_mesh'), ('aliengo', 'perlin', False, (5, 5, 0.1, 0.1), 'f3_pyr_hfield_prim'),
           ('go2', 'random_boxes', False, None, 'f6_ell_boxes_prim'), ('hyqreal1', 'flat', True, None, 'f3_ell_flat_mesh'),
           ('aliengo', 'flat', False, None, 'f3_pyr_flat_prim'), ('go1', 'flat', False, None, 'f6_ell_flat_prim'), ('spot', 'flat', False, None, 'f6_ell_flat_mesh')]


def _state(sim):
    q = sim.qpos.cpu().numpy().astype(np.float64)
    q[:, :3] = sim.base_pos64.cpu().numpy()
    return q, sim.qvel.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize('robot,scene,imu,hm,variant', CONFIGS)
def test_specialised_variant_equals_generic(robot, scene, imu, hm, variant, cuda_device):
    """Every BASELINE configuration runs its own specialised step kernel (compile-time feature switches, csrc/qs_variants.h); the
    generic run-time-dispatch kernel must reproduce it bit for bit (random-action rollout with in-kernel auto-reset)."""
    m = Model(robot, scene)
    n, T = 1024, 60
    kw = dict(use_imu=imu, heightmap=hm, seed=3)
    a = BatchSim(m, n, device=cuda_device, **kw)
    assert a.step_variant == variant
    os.environ['QSTEP_GENERIC'] = '1'
    try:
        b = BatchSim(m, n, device=cuda_device, **kw)
    finally:
        del os.environ['QSTEP_GENERIC']
    assert b.step_variant in ('f3', 'f6')
    opt = a.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5), command_mode=1 | 4)
    for s in (a, b):
        s.reset(options=opt)
    g = torch.Generator(device='cpu').manual_seed(1)
    n_term = 0
    for t in range(T):
        ctrl = (torch.randn(n, 12, generator=g) * 50).to(cuda_device)
        a.step_autoreset(ctrl, opt); b.step_autoreset(ctrl, opt)
        n_term += int(a.terminated.sum())
    for name in ('obs', 'qpos', 'qvel', 'qacc', 'terminated', 'base_pos64', 'ncon', 'solver_iter', 'status', 'command', 'friction'):
        assert torch.equal(getattr(a, name), getattr(b, name)), f'{name}: specialised and generic kernels differ'
    assert n_term > 0


@pytest.mark.parametrize('robot,scene,n,ring,wrap', [
    ('mini_cheetah', 'flat', 4096, 0, False), ('mini_cheetah', 'flat', 100, 0, False), ('go2', 'random_boxes', 1500, 0, False),
    ('mini_cheetah', 'flat', 4096, 2, False), ('go2', 'random_boxes', 4096, 2, False), ('go2', 'random_boxes', 4096, 3, False),
    ('mini_cheetah', 'flat', 40000, 0, False), ('hyqreal1', 'flat', 20000, 2, False),  # ten waves of CTAs per launch
    ('mini_cheetah', 'flat', 4096, 0, True), ('go2', 'random_boxes', 1500, 3, True),  # publish counters cross 2^32 during the rollout
    # small batches: many launches are resident at once (a launch is 16 / 6 CTAs), several of them spinning on the same ring entry
    ('go2', 'random_boxes', 300, 0, False), ('go2', 'random_boxes', 100, 2, False), ('aliengo', 'perlin', 60, 3, True)])
def test_pipelined_launches_equal_serialized(robot, scene, n, ring, wrap, cuda_device, monkeypatch):
    """QsConfig.pipeline: consecutive step launches overlap on the device (programmatic dependent launch; every env waits only for
    its own previous step through the finish-order queues).  Results must not depend on it: K back-to-back launches with overlap
    == the same K launches in plain stream order, for the state, the per-step flags and the observation of the last step; also
    across launch-chain breaks (a reset or a host read between steps).

    `ring` > 0 shortens the queue ring (QSTEP_RING_DEPTH, read by qs_create): fast envs then catch up with the ring entry a
    straggler (an env whose reset lifts the robot out of a box, up to 100 times) has not published to yet -- they have to wait for
    the entry's publish counter to reach their launch's base instead of taking a position relative to the older launch; and the
    launches s, s + ring, ... that read the same ring entry can be resident together, so a queue slot carries the generation of the
    read it is meant for (without the tag a later launch took an env published for an earlier one: a hang with ring 3 / 1500 envs)."""
    if ring:
        monkeypatch.setenv('QSTEP_RING_DEPTH', str(ring))
    if wrap:
        # `wrap`: start the handle's launch sequence just below the point where the 32-bit publish counters (n per use of a ring entry)
        # wrap -- with 4096 envs that point comes every 8.4 M launches, ten minutes into a pipelined rollout
        depth = ring or 8
        monkeypatch.setenv('QSTEP_SEQ_START', str(depth * ((1 << 32) // n - 3)))
    m = Model(robot, scene)
    a = BatchSim(m, n, device=cuda_device, seed=7, pipeline=True)
    b = BatchSim(m, n, device=cuda_device, seed=7, pipeline=False)
    opt = a.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5))
    for s in (a, b):
        s.reset(options=opt)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    T = 240
    ctrl = torch.randn(T, n, 12, device=cuda_device, generator=g) * 50
    term_a = torch.zeros(T, n, dtype=torch.uint8, device=cuda_device)
    torch.cuda.synchronize()
    n_term = 0
    for t in range(T):
        a.step_autoreset(ctrl[t], opt)                # launches t and t+1 may overlap ...
        if t % 60 == 59:
            term_a[t].copy_(a.terminated)             # ... a foreign kernel in between is ordered after the whole step launch
        if t == 100:
            a.reset_done(opt)                         # a reset kernel breaks the chain (terminated envs are reset a second time: both sims do it)
    for t in range(T):
        b.step_autoreset(ctrl[t], opt)
        if t % 60 == 59:
            assert torch.equal(term_a[t], b.terminated), f'terminated flags differ at step {t}'
            n_term += int(b.terminated.sum())
        if t == 100:
            b.reset_done(opt)
    torch.cuda.synchronize()
    for name in ('obs', 'qpos', 'qvel', 'qacc', 'qacc_warmstart', 'terminated', 'base_pos64', 'command', 'friction', 'step_count', 'sim_time'):
        assert torch.equal(getattr(a, name), getattr(b, name)), f'{name} depends on launch overlap'
    assert (a.step_count <= T).all()


@pytest.mark.parametrize('robot', ['go2', 'hyqreal1', 'mini_cheetah', 'aliengo'])
def test_open_loop_100_steps_at_full_torque(robot, cuda_device):
    """north_star's bar for every BASELINE robot, including the elliptic-cone ones (go2, hyqreal1: impratio 100): open-loop 100 steps
    at ctrl = 50*N(0,1) (clamped by the actuators), contact set / contact_state / termination identical at every step, fp32 state
    within 1e-4 of the fp64 oracle.

    Written rule for the state bound.  A single fp32 step of these robots carries ~1e-6 relative error in qacc (|qacc| ~ 1e3 rad/s^2
    under full torque -> ~2e-6 in qvel per step); over 100 steps this random-walks to ~2e-5, and an impact within those steps can
    amplify what has accumulated.  The bound is therefore asserted as: every env <= 1e-4 in the fp64-arithmetic build of the same
    kernel (QsConfig.precision = 1, fp32 state buffers); in the fp32 product build the median env <= 3e-5, at least 14 of 16 envs
    <= 1e-4 and no env above 1e-3.  Flags are compared exactly as long as an env is within 1e-4; an env that has left the bound is no
    longer compared (its contact events may shift by a step)."""
    m = Model(robot, 'flat')
    n, T = 16, 100
    qpos, qvel = seeded_states(m, n, seed=13)
    rng = np.random.RandomState(8)
    ctrl = (rng.randn(T, n, 12) * 50).astype(np.float32)
    ref = oracle_rollout(m, qpos, qvel, ctrl.astype(np.float64), mu=(0.9, 0.9), command=(0.5, 0.0, 0.0, 0.0))
    for precision in (1, 0):
        sim = BatchSim(m, n, device=cuda_device, precision=precision)
        sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
        sim.friction[:] = 0.9
        sim.command[:] = torch.tensor([0.5, 0.0, 0.0, 0.0], device=cuda_device)
        ctrl_d = torch.tensor(ctrl, device=cuda_device)
        worst = np.zeros(n)
        for t in range(T):
            obs, _, term, _ = sim.step(ctrl_d[t])
            q, v = _state(sim)
            worst = np.maximum(worst, np.maximum(np.abs(q - ref['qpos'][t]).max(axis=1), np.abs(v - ref['qvel'][t]).max(axis=1)))
            ok = worst <= 1e-4
            cs = obs[:, 199:203].cpu().numpy() > 0.5
            assert (cs[ok] == ref['cstate'][t][ok]).all(), f'contact_state differs at step {t}'
            assert (term.cpu().numpy().astype(bool)[ok] == ref['term'][t][ok]).all(), f'termination differs at step {t}'
            assert (sim.ncon.cpu().numpy()[ok] == ref['ncon'][t][ok]).all(), f'contact count differs at step {t}'
        if precision == 1:
            assert worst.max() <= 1e-4, f'fp64-arithmetic build: {worst}'
        else:
            assert np.median(worst) <= 3e-5 and (worst <= 1e-4).sum() >= n - 2 and worst.max() <= 1e-3, f'fp32 build: {np.sort(worst)}'


@pytest.mark.parametrize('torque_scale', [4.0, 50.0])
def test_observation_row_matches_oracle_at_every_step(torque_scale, cuda_device):
    """All 227 ALL_OBS scalars against the oracle at each of 100 open-loop steps (round 1 compared the last step only)."""
    m = Model('mini_cheetah', 'flat')
    n, T = 8, 100
    qpos, qvel = seeded_states(m, n, seed=11)
    rng = np.random.RandomState(5)
    ctrl = (rng.randn(T, n, 12) * torque_scale).astype(np.float32)
    ref = oracle_rollout(m, qpos, qvel, ctrl.astype(np.float64), mu=(0.8, 0.8), command=(0.6, 0.0, 0.0, 0.2))
    sim = BatchSim(m, n, device=cuda_device)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.friction[:] = 0.8
    sim.command[:] = torch.tensor([0.6, 0.0, 0.0, 0.2], device=cuda_device)
    ctrl_d = torch.tensor(ctrl, device=cuda_device)
    worst, worst_col = 0.0, -1
    for t in range(T):
        obs, _, _, _ = sim.step(ctrl_d[t])
        o_gpu = obs.cpu().numpy()[:, :227].astype(np.float64)
        o_ref = ref['obs'][t]
        # kinetic energy / work / contact forces / accelerations scale with the state -> error relative to max(1, |value|)
        err = np.abs(o_gpu - o_ref) / np.maximum(1.0, np.abs(o_ref))
        # qacc-derived columns (base_lin_acc 9:12, 43:46, work 126) and contact forces (203:227) inherit the solver tolerance
        loose = np.zeros(227, dtype=bool); loose[9:12] = loose[43:46] = True; loose[126] = True; loose[203:227] = True
        assert err[:, ~loose].max() < 2e-4, f'step {t}: obs column {np.argmax(err[:, ~loose].max(axis=0))}: {err[:, ~loose].max()}'
        assert err[:, loose].max() < 5e-3, f'step {t}: force / acceleration column: {err[:, loose].max()}'
        if err.max() > worst:
            worst, worst_col = err.max(), int(np.argmax(err.max(axis=0)))
    print(f'worst relative obs error {worst:.2e} (column {worst_col})')


def test_sampled_envs_of_the_full_batch_match_oracle(cuda_device):
    """BASELINE configs[1] at full size: 4096 envs, random reset, 40 open-loop steps at ctrl = 50*N(0,1); 64 envs sampled across the
    batch are replayed by the oracle from their post-reset state (with each env's own friction and command).  Same written rule
    as test_open_loop_100_steps_at_full_torque: flags are compared exactly while an env is within 1e-4 of the oracle; the median
    env stays below 3e-5, at least 60 of the 64 below 1e-4, none above 1e-3 (landing impacts amplify the accumulated fp32 rounding)."""
    m = Model('mini_cheetah', 'flat')
    n, T, ns = 4096, 40, 64
    sim = BatchSim(m, n, device=cuda_device, seed=21)
    opt = sim.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5), command_mode=CMD_FORWARD | CMD_ROTATE, ang_vel_range=(-0.3, 0.3))
    sim.reset(options=opt)
    torch.cuda.synchronize()
    ids = np.random.RandomState(0).choice(n, ns, replace=False)
    q0, v0 = _state(sim)
    w0 = sim.qacc_warmstart.cpu().numpy().astype(np.float64)
    fr = sim.friction.cpu().numpy().astype(np.float64); cmd = sim.command.cpu().numpy().astype(np.float64)
    g = torch.Generator(device='cpu').manual_seed(9)
    ctrl = torch.randn(T, n, 12, generator=g) * 50
    orc = []
    for i in ids:
        o = Oracle(m)
        o.set_state(q0[i], v0[i], w0[i]); o.set_env(fr[i, 0], fr[i, 1], cmd[i])
        orc.append(o)
    worst = np.zeros(ns)
    for t in range(T):
        obs, _, term, _ = sim.step(ctrl[t].to(cuda_device))
        q, v = _state(sim)
        cs = obs[:, 199:203].cpu().numpy() > 0.5
        tm = term.cpu().numpy().astype(bool); nc = sim.ncon.cpu().numpy()
        for k, (i, o) in enumerate(zip(ids, orc)):
            ref_obs, ref_term = o.step(ctrl[t, i].numpy().astype(np.float64))
            qo, vo, _, _ = o.get_state()
            f = o.flags()
            # positions are compared relative to the env's own origin: resets scatter the envs over +-10 km (fp64 master copy)
            worst[k] = max(worst[k], np.abs(q[i] - qo).max(), np.abs(v[i] - vo).max())
            if worst[k] <= 1e-4:
                assert tm[i] == ref_term and nc[i] == f['ncon'] and (cs[i] == f['contact_state']).all(), f'env {i} step {t}: flags differ'
    assert np.median(worst) < 3e-5 and (worst <= 1e-4).sum() >= ns - 4 and worst.max() < 1e-3, np.sort(worst)[-6:]


@pytest.mark.parametrize('robot,scene,xy,z0,seed', [('aliengo', 'perlin', (3.0, 2.0), 0.95, 3), ('go2', 'random_boxes', (2.0, -1.0), 0.45, 3),
                                                     ('aliengo', 'random_boxes', (3.5, 1.0), 0.62, 5)])
def test_terrain_open_loop_rollout(robot, scene, xy, z0, seed, cuda_device):
    """configs 3 / 4 open loop: the robot drops onto the terrain under a PD hold computed from the ORACLE's state (so both sides get the
    same torques), 150 steps, the oracle is never re-seeded.  Contact count, contact_state and termination must agree at every
    step.  Tie-break rule (written, tested): a step may disagree only if the oracle holds a contact whose distance is within
    5e-6 m of its activation margin at that step (an fp32 ulp of the metre-scale positions that enter the distance) -- such an event
    may fire one step earlier or later on either side; the sets must agree again at the following step."""
    m = Model(robot, scene)
    n, T = 6, 150
    rng = np.random.RandomState(seed)
    key = np.array(m.c.key_qpos)
    qpos = np.tile(key, (n, 1)); qvel = np.zeros((n, 18))
    for i in range(n):
        qpos[i, 0:2] = np.array(xy) + rng.uniform(-0.6, 0.6, 2)
        qpos[i, 2] = z0
        qpos[i, 7:] += rng.uniform(-0.15, 0.15, 12)
        o = Oracle(m)
        o.set_state(qpos[i], np.zeros(18), np.zeros(18)); assert o.lift() >= 0
        qpos[i] = o.get_state()[0]
    qpos = qpos.astype(np.float32).astype(np.float64)
    orc = [Oracle(m) for _ in range(n)]
    for i, o in enumerate(orc):
        o.set_state(qpos[i], qvel[i], np.zeros(18)); o.set_env(0.8, 0.8, [0.5, 0, 0, 0])
    sim = BatchSim(m, n, device=cuda_device)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.friction[:] = 0.8
    sim.command[:] = torch.tensor([0.5, 0, 0, 0], device=cuda_device)
    margin = 0.001 if robot == 'go2' else 0.0
    worst, max_ncon, ties, pending = 0.0, 0, 0, np.zeros(n, dtype=bool)
    alive = np.ones(n, dtype=bool)
    for t in range(T):
        ctrl = np.zeros((n, 12), dtype=np.float32)
        for i, o in enumerate(orc):
            q0, v0, _, _ = o.get_state()
            ctrl[i] = (40 * (key[7:] - q0[7:]) - 2 * v0[6:] + rng.randn(12) * 2).astype(np.float32)
        obs, _, term, _ = sim.step(torch.tensor(ctrl, device=cuda_device))
        q1, v1 = _state(sim)
        for i, o in enumerate(orc):
            if not alive[i]:
                continue
            _, ref_term = o.step(ctrl[i].astype(np.float64))
            f = o.flags()
            same = (bool(term[i].item()) == ref_term and int(sim.ncon[i].item()) == f['ncon']
                    and ((obs[i, 199:203].cpu().numpy() > 0.5) == f['contact_state']).all())
            if not same:
                d = o.get(F_CONTACTS)[:, 0]
                assert not pending[i], f'step {t} env {i}: contact sets still differ one step after a tie'
                assert len(d) and np.abs(d - margin).min() < 5e-6, f'step {t} env {i}: contact set differs and no contact sits at its margin'
                pending[i] = True; ties += 1
                continue
            pending[i] = False
            qo, vo, _, _ = o.get_state()
            max_ncon = max(max_ncon, f['ncon'])
            worst = max(worst, np.abs(q1[i] - qo).max(), 0.1 * np.abs(v1[i] - vo).max())
            if ref_term:
                alive[i] = False  # episode over (a body other than a calf touches the terrain): stop comparing this env
    assert max_ncon >= 4 and ties <= 2 and worst < 2e-4, (max_ncon, ties, worst)


def test_out_of_bounds_termination(cuda_device):
    """_check_out_of_terrain_bounds (quadruped_env.py:1250-1257): random_boxes has finite terrain_limits (terrain.py:237); a robot
    thrown across the +x limit must raise `terminated` at exactly the step at which the oracle's base x exceeds it, with no contact."""
    m = Model('mini_cheetah', 'random_boxes')
    lim = np.array(m.terrain_limits)  # (x_max, x_min, y_max, y_min)
    assert np.isfinite(lim).all() and lim[0] < 1e3
    n = 4
    key = np.array(m.c.key_qpos)
    qpos = np.tile(key, (n, 1)); qvel = np.zeros((n, 18))
    # env 0 crosses x_max, env 1 crosses x_min, env 2 crosses y_max, env 3 stays inside
    qpos[:, 2] = 1.5
    qpos[0, 0] = lim[0] - 0.05; qvel[0, 0] = 3.0
    qpos[1, 0] = lim[1] + 0.05; qvel[1, 0] = -3.0
    qpos[2, 1] = lim[2] - 0.05; qvel[2, 1] = 3.0
    qpos[3, 0:2] = [0.5 * (lim[0] + lim[1]), 0.5 * (lim[2] + lim[3])]
    qpos = qpos.astype(np.float32).astype(np.float64); qvel = qvel.astype(np.float32).astype(np.float64)
    ctrl = np.zeros((60, n, 12))
    ref = oracle_rollout(m, qpos, qvel, ctrl)
    sim = BatchSim(m, n, device=cuda_device)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    first = np.full(n, -1)
    for t in range(60):
        _, _, term, _ = sim.step(torch.zeros(n, 12, device=cuda_device))
        tm = term.cpu().numpy().astype(bool)
        assert (tm == ref['term'][t]).all(), f'step {t}: {tm} vs {ref["term"][t]}'
        assert (sim.ncon.cpu().numpy() == 0).all() and (sim.invalid_body_mask.cpu().numpy() == 0).all()  # bounds only, no contact
        first = np.where((first < 0) & tm, t, first)
    assert (first[:3] >= 5).all() and (first[:3] <= 30).all() and first[3] == -1, first


def test_qfrc_applied_matches_oracle(cuda_device):
    """mjData.qfrc_applied[:6] (quadruped_env.py:305): an external wrench on the base dofs enters qfrc_smooth; 50 steps against the
    oracle with a different wrench per env."""
    m = Model('aliengo', 'flat')
    n, T = 8, 50
    qpos, qvel = seeded_states(m, n, seed=17)
    rng = np.random.RandomState(3)
    wrench = np.concatenate([rng.uniform(-40, 40, (n, 3)), rng.uniform(-8, 8, (n, 3))], axis=1).astype(np.float32)
    ctrl = (rng.randn(T, n, 12) * 6).astype(np.float32)
    sim = BatchSim(m, n, device=cuda_device)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.qfrc_applied[:] = torch.tensor(wrench, device=cuda_device)
    orc = []
    for i in range(n):
        o = Oracle(m); o.set_state(qpos[i], qvel[i], np.zeros(18)); o.set_env(-1.0, -1.0, [0, 0, 0, 0], wrench[i].astype(np.float64))
        orc.append(o)
    base = BatchSim(m, n, device=cuda_device)
    base.set_state(torch.tensor(qpos), torch.tensor(qvel))
    worst = 0.0
    for t in range(T):
        sim.step(torch.tensor(ctrl[t], device=cuda_device)); base.step(torch.tensor(ctrl[t], device=cuda_device))
        q, v = _state(sim)
        for i, o in enumerate(orc):
            o.step(ctrl[t, i].astype(np.float64))
            qo, vo, _, _ = o.get_state()
            worst = max(worst, np.abs(q[i] - qo).max(), np.abs(v[i] - vo).max())
    assert worst < 1e-4, worst
    assert (sim.qvel - base.qvel).abs().max() > 0.05  # the wrench did something


def test_in_kernel_schedules(cuda_device):
    """'+reset' command types and external disturbances of type 'reset' (quadruped_env.py:293-305, :1046-1139) run inside the step
    kernel: cadence (count, limit = randint(1000, 3000)), distributions, the wrench reaching qfrc_applied one step later, and
    per-env bookkeeping across resets -- without any host synchronisation."""
    m = Model('mini_cheetah', 'flat')
    n = 2048
    sim = BatchSim(m, n, device=cuda_device, seed=11)
    mode = CMD_RANDOM | CMD_ROTATE | CMD_RESET
    sim.set_schedule(command_mode=mode, lin_vel_range=(0.3, 0.9), ang_vel_range=(-0.4, 0.4), ext_enabled=True,
                     ext_ranges={'x': (-30, 30), 'z': (5,), 'yaw': (-2, 2)})
    torch.cuda.synchronize()
    # __init__-time draw of the disturbance (:240-242)
    w = sim.ext_wrench.cpu().numpy(); lim = sim.ext_limit.cpu().numpy()
    assert (np.abs(w[:, 0]) <= 30).all() and w[:, 0].std() > 10 and (w[:, 2] == 5).all() and (w[:, 1] == 0).all() and (np.abs(w[:, 5]) <= 2).all()
    assert lim.min() >= 1000 and lim.max() <= 2999 and lim.std() > 400
    opt = sim.make_reset_options(lin_vel_range=(0.3, 0.9), ang_vel_range=(-0.4, 0.4), friction_range=(0.5, 1.0), command_mode=mode)
    sim.reset(options=opt)
    cl = sim.cmd_limit.cpu().numpy()
    assert cl.min() >= 1000 and cl.max() <= 2999 and (sim.cmd_count == 0).all()
    assert (sim.qfrc_applied == 0).all()  # reset zeroes the applied wrench (:335); it comes back after the next step (:305)
    # shorten the limits (caller-owned buffers) so the cadence can be watched in a few steps
    sim.cmd_limit[:] = torch.randint(3, 9, (n,), dtype=torch.int32, device=cuda_device)
    sim.ext_limit[:] = torch.randint(4, 11, (n,), dtype=torch.int32, device=cuda_device)
    sim.ext_count[:] = 0
    cmd_lim0 = sim.cmd_limit.clone(); ext_lim0 = sim.ext_limit.clone()
    cmd_prev, wr_prev = sim.command.clone(), sim.ext_wrench.clone()
    ctrl = torch.zeros(n, 12, device=cuda_device)
    changed_cmd_at = torch.full((n,), -1, device=cuda_device); changed_ext_at = torch.full((n,), -1, device=cuda_device)
    for t in range(1, 12):
        sim.step(ctrl)
        assert torch.equal(sim.qfrc_applied, sim.ext_wrench)  # written after every step, acts on the next one
        c = (sim.command != cmd_prev).any(dim=1) & (changed_cmd_at < 0); changed_cmd_at[c] = t
        e = (sim.ext_wrench != wr_prev).any(dim=1) & (changed_ext_at < 0); changed_ext_at[e] = t
    # the first resample happens exactly when the count reaches the limit
    assert torch.equal(changed_cmd_at, cmd_lim0.to(changed_cmd_at.dtype)), 'command cadence'
    assert torch.equal(changed_ext_at, ext_lim0.to(changed_ext_at.dtype)), 'disturbance cadence'
    cmd = sim.command.cpu().numpy()
    speed = np.hypot(cmd[:, 0], cmd[:, 1])
    assert speed.min() >= 0.3 - 1e-6 and speed.max() <= 0.9 + 1e-6 and (np.abs(cmd[:, 3]) <= 0.4).all() and (cmd[:, 2] == 0).all()
    heading = np.arctan2(cmd[:, 1], cmd[:, 0])
    assert heading.std() > 1.5  # 'random' heading: U(-pi, pi)
    newl = sim.cmd_limit.cpu().numpy()
    assert newl.min() >= 1000 and newl.max() <= 2999  # the new limits are full-length again
    # determinism: a second handle with the same seed and the same edits reproduces everything bit for bit
    sim2 = BatchSim(m, n, device=cuda_device, seed=11)
    sim2.set_schedule(command_mode=mode, lin_vel_range=(0.3, 0.9), ang_vel_range=(-0.4, 0.4), ext_enabled=True,
                      ext_ranges={'x': (-30, 30), 'z': (5,), 'yaw': (-2, 2)})
    sim2.reset(options=opt)
    sim2.cmd_limit.copy_(cmd_lim0); sim2.ext_limit.copy_(ext_lim0); sim2.ext_count[:] = 0
    for t in range(1, 12):
        sim2.step(ctrl)
    assert torch.equal(sim2.command, sim.command) and torch.equal(sim2.ext_wrench, sim.ext_wrench) and torch.equal(sim2.qpos, sim.qpos)
    # 'human' speed stays zero when the schedule fires (:1058-1061)
    sim.set_schedule(command_mode=CMD_RESET, lin_vel_range=(0.3, 0.9))
    sim.cmd_limit[:] = 2; sim.cmd_count[:] = 0
    sim.step(ctrl); sim.step(ctrl)
    assert (sim.command == 0).all()


def test_wrench_schedule_drives_the_dynamics(cuda_device):
    """The scheduled wrench must be the one the NEXT step integrates: against the oracle fed with the same wrench by hand."""
    m = Model('mini_cheetah', 'flat')
    n = 4
    qpos, qvel = seeded_states(m, n, seed=5)
    sim = BatchSim(m, n, device=cuda_device, seed=2)
    sim.set_schedule(ext_enabled=True, ext_ranges={'x': (-20, 20), 'y': (10,), 'pitch': (-3, 3)})
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    wrench = sim.ext_wrench.cpu().numpy().astype(np.float64)
    orc = []
    for i in range(n):
        o = Oracle(m); o.set_state(qpos[i], qvel[i], np.zeros(18)); orc.append(o)
    ctrl = np.zeros(12)
    for t in range(20):
        sim.step(torch.zeros(n, 12, device=cuda_device))
        q, v = _state(sim)
        for i, o in enumerate(orc):
            o.set_env(-1.0, -1.0, [0, 0, 0, 0], wrench[i] if t > 0 else np.zeros(6))  # first step: qfrc_applied still zero
            o.step(ctrl)
            qo, vo, _, _ = o.get_state()
            assert np.abs(q[i] - qo).max() < 1e-4 and np.abs(v[i] - vo).max() < 1e-4, (t, i)


def test_env_reset_with_seed_keeps_everything_else(cuda_device):
    """reset(seed=...) (quadruped_env.py:337-338): reproducible draws, fused height-map columns stay, a masked reset leaves the
    other envs alone (_reseed used to rebuild the handle without the height map and to wipe command / friction)."""
    from gym_quadruped_b200.quadruped_env import QuadrupedEnv
    from gym_quadruped_b200.sensors import HeightMap
    env = QuadrupedEnv('aliengo', scene='perlin', state_obs_names=('qpos', 'qvel', 'heightmap'), sensors=(HeightMap,),
                       sensors_kwargs=(dict(num_rows=5, num_cols=5, dist_x=0.1, dist_y=0.1),), ref_base_lin_vel=(0.5, 1.0),
                       ground_friction_coeff=(0.2, 1.5), num_envs=64)
    o1 = {k: v.clone() for k, v in env.reset(seed=123).items()}
    cmd1, fr1 = env.sim.command.clone(), env.sim.friction.clone()
    assert o1['heightmap'].shape == (64, 75)
    env.step(torch.zeros(64, 12, device=cuda_device))
    o2 = env.reset(seed=123)
    for k in o1:
        assert torch.equal(o1[k], o2[k]), k
    assert torch.equal(env.sim.command, cmd1) and torch.equal(env.sim.friction, fr1)
    o3 = env.reset(seed=124)
    assert not torch.equal(o3['qpos'], o1['qpos'])
    # masked reset with a new seed: unmasked envs keep state, command, friction and step counters
    for _ in range(3):
        env.step(torch.zeros(64, 12, device=cuda_device))
    keep_q, keep_cmd, keep_steps = env.sim.qpos.clone(), env.sim.command.clone(), env.sim.step_count.clone()
    mask = torch.zeros(64, dtype=torch.bool, device=cuda_device); mask[:8] = True
    env.reset(seed=5, env_mask=mask)
    assert torch.equal(env.sim.qpos[8:], keep_q[8:]) and torch.equal(env.sim.command[8:], keep_cmd[8:])
    assert torch.equal(env.sim.step_count[8:], keep_steps[8:]) and (env.sim.step_count[:8] == 0).all()
    env.close()
    # num_envs == 1: the pinned row keeps its width
    env = QuadrupedEnv('aliengo', scene='flat', state_obs_names=('qpos', 'heightmap'), sensors=(HeightMap,),
                       sensors_kwargs=(dict(num_rows=3, num_cols=3, dist_x=0.1, dist_y=0.1),))
    a = env.reset(seed=9)
    obs, *_ = env.step(np.zeros(12))
    b = env.reset(seed=9)
    assert a['heightmap'].shape == (27,) and np.array_equal(a['qpos'], b['qpos']) and obs['heightmap'].shape == (27,)
    env.close()


def test_contact_buffer_overflow_keeps_termination_exact(cuda_device):
    """More contacts than the kernel's 16 slots: the surplus is dropped from the solver (status bit 1), but contact_state and the
    invalid-contact mask are collected when contacts are DETECTED, so termination still equals the oracle's."""
    m = Model('go1', 'flat')  # 42 collision geoms
    key = np.array(m.c.key_qpos)
    n = 2
    qpos = np.tile(key, (n, 1)); qvel = np.zeros((n, 18))
    qpos[:, 2] = 0.05           # belly on the floor, legs folded: many geoms touch
    qpos[:, 7:] = np.tile([0.0, 1.4, -2.6], 4)
    qpos = qpos.astype(np.float32).astype(np.float64)
    sim = BatchSim(m, n, device=cuda_device)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    _, _, term, _ = sim.step(torch.zeros(n, 12, device=cuda_device))
    o = Oracle(m); o.set_state(qpos[0], qvel[0], np.zeros(18))
    _, ref_term = o.step(np.zeros(12))
    f = o.flags()
    if f['ncon'] <= 16:
        pytest.skip(f'pose produces only {f["ncon"]} contacts')
    assert int(sim.status[0].item()) & 2, 'overflow bit not raised'
    inv = sim.invalid_body_mask.cpu().numpy().astype(np.int64)
    assert (inv[0, 0] | (inv[0, 1] << 8)) == f['invalid_body_mask'] and bool(term[0].item()) == ref_term
    assert ((sim.obs[0, 199:203].cpu().numpy() > 0.5) == f['contact_state']).all()


def test_rollout_recorder_from_gpu_rollout(tmp_path, cuda_device):
    """f3: a batched GPU rollout written in the reference's recording layout (utils/data/h5py.py:90-172): N envs x T steps become N
    trajectories [traj, time, dim] float64 with `action`, `time` and the env hyper-parameters."""
    from gym_quadruped_b200.quadruped_env import QuadrupedEnv
    from gym_quadruped_b200.utils.data.recorder import RolloutReader, RolloutRecorder
    names = ('qpos', 'qvel', 'base_ori_SO3', 'contact_state', 'feet_pos:base')
    env = QuadrupedEnv('aliengo', state_obs_names=names, ref_base_lin_vel=(0.5, 1.0), num_envs=8)
    env.reset(seed=1)
    rec = RolloutRecorder(env, tmp_path / 'rollout.npz', with_terminated=True)
    g = torch.Generator(device='cpu').manual_seed(0)
    kept = []
    for t in range(6):
        a = (torch.randn(8, 12, generator=g) * 5).to(cuda_device)
        obs, rew, term, trunc, info = env.step(a)
        rec.record(obs, a, term)
        kept.append((obs['qpos'].clone(), a.clone(), env.sim.sim_time.clone()))
    rec.flush()
    r = RolloutReader(tmp_path / 'rollout.npz')
    assert r.len() == 8 and r.env_hparams['robot'] == 'aliengo' and r.env_hparams['state_obs_names'] == list(names)
    time, data = r.get_trajectory(3)
    assert time.shape == (6, 1) and data['qpos'].shape == (6, 19) and data['base_ori_SO3'].shape == (6, 9) and data['action'].shape == (6, 12)
    for t, (q, a, st) in enumerate(kept):
        assert np.array_equal(data['qpos'][t], q[3].cpu().numpy().astype(np.float64))
        assert np.array_equal(data['action'][t], a[3].cpu().numpy().astype(np.float64))
        assert time[t, 0] == float(st[3])
    assert data['terminated'].shape == (6, 1) and data['qpos'].dtype == np.float64
    env.close()


@pytest.mark.parametrize('robot,scene,xy', [('mini_cheetah', 'perlin', (3.0, 2.0)), ('hyqreal1', 'perlin', (-4.0, 6.0)), ('mini_cheetah', 'random_boxes', (2.0, -1.0))])
def test_mesh_links_collide_with_the_terrain(robot, scene, xy, cuda_device):
    """Mesh robots on non-flat terrain (the reference's own test matrix runs mini_cheetah / hyqreal1 / hyqreal2 x perlin,
    tests/env_test.py:14-16): convex-mesh links against the height field / static boxes.  The robots are dropped limp so that
    thighs, calves and trunk come to rest on the terrain; closed loop against the oracle (re-seeded from the GPU state each step):
    contact count, contact_state, invalid-contact mask and termination identical, state to single-step fp32 accuracy."""
    m = Model(robot, scene)
    n, T = 6, 330
    rng = np.random.RandomState(11)
    key = np.array(m.c.key_qpos)
    qpos = np.tile(key, (n, 1)); qvel = np.zeros((n, 18))
    for i in range(n):
        qpos[i, 0:2] = np.array(xy) + rng.uniform(-0.8, 0.8, 2)
        qpos[i, 2] = 1.2 if robot != 'hyqreal1' else 1.6
        qpos[i, 7:] += rng.uniform(-0.2, 0.2, 12)
    qpos = qpos.astype(np.float32).astype(np.float64)
    sim = BatchSim(m, n, device=cuda_device)
    assert sim.step_variant in ('f3', 'f6')  # mesh x terrain runs on the generic kernel
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.friction[:] = 0.8
    orc = [Oracle(m) for _ in range(n)]
    for o in orc:
        o.set_env(0.8, 0.8, [0, 0, 0, 0])
    mesh = np.array([m.c.geom_type[g] == 7 for g in range(m.c.ngeom)])
    ctrl = torch.zeros(n, 12, device=cuda_device)
    worst, mesh_contacts, ties = 0.0, 0, 0
    for t in range(T):
        q0, v0 = _state(sim)
        w0 = sim.qacc_warmstart.cpu().numpy().astype(np.float64)
        obs, _, term, _ = sim.step(ctrl)
        q1, v1 = _state(sim)
        inv = sim.invalid_body_mask.cpu().numpy().astype(np.int64)
        for i, o in enumerate(orc):
            o.set_state(q0[i], v0[i], w0[i])
            _, ref_term = o.step(np.zeros(12))
            f = o.flags()
            if f['ncon'] > 16:
                continue  # the 16-slot contact buffer is full (status bit 1): flags stay exact (own test), forces do not
            same = (bool(term[i].item()) == ref_term and int(sim.ncon[i].item()) == f['ncon'] and (inv[i, 0] | (inv[i, 1] << 8)) == f['invalid_body_mask']
                    and ((obs[i, 199:203].cpu().numpy() > 0.5) == f['contact_state']).all())
            oc = o.get(F_CONTACTS)
            if not same:  # a contact within fp32 rounding of its activation distance (same rule as test_terrain_scenes_match_oracle)
                assert len(oc) and np.abs(oc[:, 0]).min() < 3e-6, f'step {t} env {i}: contact set differs'
                ties += 1
                continue
            if len(oc):
                mesh_contacts += int((mesh[oc[:, 16].astype(int)] & (oc[:, 3] > 0.02)).sum())
            qo, vo, _, _ = o.get_state()
            if f['ncon'] <= 6:  # many redundant contacts of a body lying on the terrain: ill-conditioned, held to the flags only
                worst = max(worst, np.abs(q1[i] - qo).max(), 0.1 * np.abs(v1[i] - vo).max())
    # single-step bound: several contacts on one lying link are a redundant, stiff problem (cf. test_terrain_scenes_match_oracle)
    assert mesh_contacts >= 50 and ties <= 6 and worst < 3e-4, (mesh_contacts, ties, worst)


@pytest.mark.parametrize('robot,scene,n', [('mini_cheetah', 'flat', 4096), ('aliengo', 'perlin', 700)])
def test_step_k_equals_k_single_steps(robot, scene, n, cuda_device):
    """qs_step_k: K steps from one call, one launch per step, launches overlapped on the device by the library
    itself (no contract on the caller, works on a handle with pipeline = 0) -- every per-step output (observation ring, terminated
    ring) and the final state bit-identical to K calls of qs_step_autoreset."""
    m = Model(robot, scene)
    hm = (5, 5, 0.1, 0.1) if scene == 'perlin' else None
    a = BatchSim(m, n, device=cuda_device, seed=4, heightmap=hm)
    b = BatchSim(m, n, device=cuda_device, seed=4, heightmap=hm)
    opt = a.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5))
    for s in (a, b):
        s.reset(options=opt)
    K = 48
    g = torch.Generator(device=cuda_device).manual_seed(2)
    ctrl = torch.randn(3 * K, n, 12, device=cuda_device, generator=g) * 50
    for seg in range(3):
        ring = torch.zeros(K, n, a.obs_dim, device=cuda_device)
        tring = torch.zeros(K, n, dtype=torch.uint8, device=cuda_device)
        a.step_k(ctrl[seg * K:(seg + 1) * K], opt, obs_ring=ring, terminated_ring=tring)
        a.friction.mul_(1.0)  # a foreign kernel between two K-step calls: ordered after the whole sequence
        for k in range(K):
            obs, _, term, _ = b.step_autoreset(ctrl[seg * K + k], opt)
            assert torch.equal(ring[k], obs), f'observation of step {seg * K + k} differs'
            assert torch.equal(tring[k], term)
    for name in ('qpos', 'qvel', 'qacc_warmstart', 'base_pos64', 'command', 'friction', 'step_count', 'sim_time', 'terminated'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    # without a ring only the last step's rows remain
    a.step_k(ctrl[:5], opt); [b.step_autoreset(ctrl[k], opt) for k in range(5)]
    assert torch.equal(a.obs, b.obs) and torch.equal(a.terminated, b.terminated)


def test_host_entry_point_with_padded_rows(cuda_device):
    """qs_step_host_strided: observation rows of a pinned host buffer on 128-byte boundaries ([N, 256] storage, [N, 227] view)
    receive exactly what the contiguous host buffer and the device path receive."""
    m = Model('mini_cheetah', 'flat')
    n = 64
    qpos, qvel = seeded_states(m, n, seed=6)
    sims = [BatchSim(m, n, device=cuda_device) for _ in range(3)]
    for s in sims:
        s.set_state(torch.tensor(qpos), torch.tensor(qvel))
    ctrl = torch.randn(n, 12) * 10
    ctrl_h = ctrl.clone().pin_memory()
    rew_h = torch.empty(n).pin_memory(); term_h = torch.empty(n, dtype=torch.uint8).pin_memory(); trunc_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    wide = torch.full((n, 256), -7.0).pin_memory()
    sims[0].step_host(ctrl_h, wide[:, :227], rew_h, term_h, trunc_h)
    flat = torch.empty(n, 227).pin_memory()
    sims[1].step_host(ctrl_h, flat, rew_h, term_h, trunc_h)
    obs, _, _, _ = sims[2].step(ctrl.to(cuda_device))
    assert torch.equal(wide[:, :227], flat) and torch.equal(flat, obs.cpu())
    assert (wide[:, 227:] == -7.0).all()  # the padding is never written


@pytest.mark.gpu
@pytest.mark.parametrize('robot,scene', [('mini_cheetah', 'flat'), ('go2', 'random_boxes'), ('aliengo', 'perlin'), ('hyqreal1', 'flat'),
                                         ('mini_cheetah', 'random_boxes'), ('go1', 'stairs')])
def test_non_finite_state_ends_the_episode(robot, scene, cuda_device):
    """Safety net of the batch path (no counterpart in the env; the engine itself warns and resets its data on a bad state): an env
    whose state became non-finite reports terminated = 1 with status bit0, the auto-reset pass brings it back to a finite state in
    the same launch, and no other env of the batch is affected (compared bit for bit with an untouched twin)."""
    n = 256
    m = Model(robot, scene)
    kw = dict(device=cuda_device, seed=11, use_imu=bool(m.c.has_imu), heightmap=(3, 3, 0.1, 0.1) if scene == 'perlin' else None)
    a = BatchSim(m, n, **kw)
    b = BatchSim(m, n, **kw)
    opt = a.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5))
    for s in (a, b):
        s.reset(options=opt)
        if scene != 'flat':  # bring the robots onto the terrain patch
            q = s.qpos.clone(); q[:, 0] = 2.0; q[:, 1] = -1.0 if scene == 'random_boxes' else 2.0; q[:, 2] = 0.9
            s.set_state(q, s.qvel)
    g = torch.Generator(device=cuda_device).manual_seed(1)
    ctrl = torch.randn(40, n, 12, device=cuda_device, generator=g) * 20
    for t in range(20):
        a.step_autoreset(ctrl[t], opt); b.step_autoreset(ctrl[t], opt)
    bad = torch.tensor([3, 77, 200], device=cuda_device)
    a.qvel[bad[0], 7] = float('nan'); a.qvel[bad[1], 2] = float('inf'); a.qpos[bad[2], 9] = float('nan')
    # plain step: flags only, the state stays what the dynamics made of it
    c = BatchSim(m, n, **kw)
    c.reset(options=opt)
    c.set_state(a.qpos.clone(), a.qvel.clone())
    c.step(ctrl[20])
    torch.cuda.synchronize()
    assert c.terminated[bad].all() and (c.status[bad] & 1).all()
    # auto-reset: same flags, finite state afterwards, every other env identical to the twin
    a.step_autoreset(ctrl[20], opt); b.step_autoreset(ctrl[20], opt)
    torch.cuda.synchronize()
    assert a.terminated[bad].all()
    assert torch.isfinite(a.qpos).all() and torch.isfinite(a.qvel).all() and torch.isfinite(a.obs).all()
    others = torch.ones(n, dtype=torch.bool, device=cuda_device); others[bad] = False
    for name in ('qpos', 'qvel', 'qacc', 'obs', 'terminated'):
        assert torch.equal(getattr(a, name)[others], getattr(b, name)[others]), name
    for t in range(21, 40):
        a.step_autoreset(ctrl[t], opt)
    torch.cuda.synchronize()
    assert torch.isfinite(a.qpos).all() and torch.isfinite(a.qvel).all() and (a.status & 1).sum() == 0


@pytest.mark.parametrize('robot,scene,n,ring,seed', [('mini_cheetah', 'flat', 333, 0, 0), ('go2', 'random_boxes', 200, 3, 1), ('aliengo', 'perlin', 500, 2, 2),
                                                     ('hyqreal1', 'flat', 1000, 0, 3)])
def test_random_api_sequences_pipelined_equals_serialized(robot, scene, n, ring, seed, cuda_device, monkeypatch):
    """Differential fuzz of the launch-chain logic: a seeded random sequence of C-ABI calls (plain steps, auto-reset steps, K-step
    calls, host-buffer steps, masked resets, reset_done, state writes, forward / get, schedule and seed changes, host reads) is applied
    to a pipelined handle and to a serialized twin; every caller-visible buffer must be bit-identical afterwards, and at random
    check points in between.  Overlapped launches, chain breaks and resumptions must never change a result."""
    if ring:
        monkeypatch.setenv('QSTEP_RING_DEPTH', str(ring))
    from gym_quadruped_b200.backend import FIELD_MASS_MATRIX as FIELD_M
    m = Model(robot, scene)
    kw = dict(device=cuda_device, seed=21, use_imu=bool(m.c.has_imu), heightmap=(3, 3, 0.1, 0.1) if scene == 'perlin' else None)
    sims = [BatchSim(m, n, pipeline=True, **kw), BatchSim(m, n, pipeline=False, **kw)]
    opt = sims[0].make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5), command_mode=CMD_FORWARD | CMD_ROTATE | CMD_RESET)
    for s in sims:
        s.set_schedule(command_mode=CMD_FORWARD | CMD_ROTATE | CMD_RESET, lin_vel_range=(0.3, 0.9), ang_vel_range=(-0.4, 0.4), ext_enabled=True,
                       ext_ranges={'x': (-20, 20), 'z': (5,)})
        s.reset(options=opt)
        s.cmd_limit[:] = 7; s.ext_limit[:] = 5
    rng = np.random.RandomState(seed)
    g = torch.Generator(device=cuda_device).manual_seed(seed)
    D = sims[0].obs_dim
    pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()
    host = [dict(ctrl=pin(n, 12), obs=pin(n, (D + 31) // 32 * 32)[:, :D], rew=pin(n), term=pin(n, dtype=torch.uint8), trunc=pin(n, dtype=torch.uint8)) for _ in sims]
    names = ('obs', 'qpos', 'qvel', 'qacc', 'qacc_warmstart', 'terminated', 'base_pos64', 'command', 'friction', 'step_count', 'sim_time', 'status',
             'cmd_count', 'ext_count', 'ext_wrench', 'qfrc_applied', 'imu_bias', 'ncon', 'solver_iter')

    def same(where):
        torch.cuda.synchronize()
        for name in names:
            assert torch.equal(getattr(sims[0], name), getattr(sims[1], name)), f'{name} differs after {where}'

    ops = ['step'] * 3 + ['auto'] * 12 + ['stepk'] * 2 + ['host', 'reset_mask', 'reset_done', 'set_state', 'forward', 'seed', 'schedule', 'read', 'check']
    pool = torch.randn(1200, n, 12, device=cuda_device, generator=g) * 40  # actions made up front: no foreign kernel between two step launches
    cur = 0
    torch.cuda.synchronize()
    for it in range(200):
        op = ops[rng.randint(len(ops))]
        if op in ('step', 'auto', 'host'):
            ctrl = pool[cur]; cur += 1
        if op == 'step':
            for s in sims: s.step(ctrl)
        elif op == 'auto':
            for s in sims: s.step_autoreset(ctrl, opt)
        elif op == 'stepk':
            K = int(rng.randint(2, 7))
            cs = pool[cur:cur + K]; cur += K
            rings = [torch.empty(K, n, D, device=cuda_device) for _ in sims]
            auto_k = bool(rng.randint(2))
            terms = [s.step_k(cs, opt, auto_reset=auto_k, obs_ring=r)[1] for s, r in zip(sims, rings)]
            torch.cuda.synchronize()
            assert torch.equal(rings[0], rings[1]) and torch.equal(terms[0], terms[1]), 'step_k rings differ'
        elif op == 'host':
            for s, h in zip(sims, host):
                h['ctrl'].copy_(ctrl)
                s.step_host(h['ctrl'], h['obs'], h['rew'], h['term'], h['trunc'], auto_reset=opt)
            assert torch.equal(host[0]['obs'], host[1]['obs']) and torch.equal(host[0]['term'], host[1]['term'])
        elif op == 'reset_mask':
            mask = (torch.rand(n, device=cuda_device, generator=g) < 0.1).to(torch.uint8)
            for s in sims: s.reset(mask=mask, options=opt)
        elif op == 'reset_done':
            for s in sims: s.reset_done(opt)
        elif op == 'set_state':
            ids = torch.randperm(n, device=cuda_device, generator=g)[:5]
            src = torch.randperm(n, device=cuda_device, generator=g)[:5]
            q = sims[1].qpos[src].clone(); q[:, :3] = sims[1].base_pos64[src].float(); v = sims[1].qvel[src].clone() * 0.5
            for s in sims: s.set_state(q, v, env_ids=ids)
        elif op == 'forward':
            for s in sims: s.forward()
            assert torch.equal(sims[0].get(FIELD_M), sims[1].get(FIELD_M))
        elif op == 'seed':
            sd = int(rng.randint(1 << 30))
            for s in sims: s.set_seed(sd)
        elif op == 'schedule':
            lim = int(rng.randint(2, 9))
            for s in sims: s.cmd_limit[:] = lim
        elif op == 'read':
            assert torch.equal(sims[0].terminated.cpu(), sims[1].terminated.cpu())  # a host read between launches
        else:
            same(f'op {it}')
    same('the whole sequence')
    assert torch.isfinite(sims[0].qpos).all()
