"""GPU parity tests proper: the fused sm_100a kernel (through the C-ABI) against the fp64 oracle on identical inputs.

Bar (BASELINE.json north_star): contact-set and termination flags bit-exact, fp32 state within 1e-4 over 100 steps.
The oracle itself is "parity unpinned" with respect to MuJoCo (see oracle/qstep_oracle.c).
"""
import numpy as np
import pytest
import torch

from gym_quadruped_b200.backend import (FIELD_CONTACTS, FIELD_FEET_JACP, FIELD_FEET_POS, FIELD_MASS_MATRIX, FIELD_QFRC_BIAS,
                                        FIELD_QFRC_SMOOTH, BatchSim)
from gym_quadruped_b200.model import Model
from oracle.oracle import F_BIAS, F_CONTACTS, F_FEET_JACP, F_FEET_POS, F_M, F_SMOOTH, Oracle
from tests.helpers import oracle_rollout, seeded_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model():
    return Model('mini_cheetah', 'flat')


def _sim(model, n, dev, precision=0):
    return BatchSim(model, n, device=dev, precision=precision)


@pytest.mark.parametrize('precision,tol', [(0, 2e-4), (1, 1e-5)])
def test_forward_tables_match_oracle(model, cuda_device, precision, tol):
    n = 16
    qpos, qvel = seeded_states(model, n, seed=3, lift=False)
    qpos[:, 2] -= 0.02  # push some feet into the ground so contacts exist
    qpos = qpos.astype(np.float32).astype(np.float64)
    sim = _sim(model, n, cuda_device, precision)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.forward()
    M = sim.get(FIELD_MASS_MATRIX).cpu().numpy(); bias = sim.get(FIELD_QFRC_BIAS).cpu().numpy()
    fsm = sim.get(FIELD_QFRC_SMOOTH).cpu().numpy(); jac = sim.get(FIELD_FEET_JACP).cpu().numpy(); fpos = sim.get(FIELD_FEET_POS).cpu().numpy()
    con = sim.get(FIELD_CONTACTS).cpu().numpy(); ncon = sim.ncon.cpu().numpy(); qacc = sim.qacc.cpu().numpy()
    for i in range(n):
        o = Oracle(model)
        o.set_state(qpos[i], qvel[i], np.zeros(18))
        o.forward(np.zeros(12))
        np.testing.assert_allclose(M[i], o.get(F_M), atol=tol)
        np.testing.assert_allclose(bias[i], o.get(F_BIAS), atol=tol * 50)
        np.testing.assert_allclose(fsm[i], o.get(F_SMOOTH), atol=tol * 50)
        np.testing.assert_allclose(jac[i], o.get(F_FEET_JACP), atol=tol)
        np.testing.assert_allclose(fpos[i], o.get(F_FEET_POS), atol=tol)
        oc = o.get(F_CONTACTS)
        assert ncon[i] == len(oc)
        gc = con[i, :ncon[i]]
        gc = gc[np.argsort(gc[:, 16])]; oc = oc[np.argsort(oc[:, 16])]
        assert (gc[:, 16] == oc[:, 16]).all() and (gc[:, 17] == oc[:, 17]).all()  # same geoms / bodies: bit-exact contact set
        np.testing.assert_allclose(gc[:, 0:13], oc[:, 0:13], atol=tol)       # dist, pos, frame
        scale = max(1.0, np.abs(oc[:, 13:16]).max()) if len(oc) else 1.0
        np.testing.assert_allclose(gc[:, 13:16], oc[:, 13:16], atol=2e-3 * scale if precision == 0 else 1e-4 * scale)
        np.testing.assert_allclose(qacc[i], o.get_state()[2], atol=(5e-3 if precision == 0 else 1e-4) * max(1.0, np.abs(o.get_state()[2]).max()))


# precision=1 runs the same kernel in fp64 arithmetic, but the state still round-trips through the fp32 buffers every step
@pytest.mark.parametrize('precision,tol', [(0, 1e-4), (1, 1e-4)])
@pytest.mark.parametrize('torque_scale', [4.0, 50.0])
def test_rollout_100_steps_matches_oracle(model, cuda_device, precision, tol, torque_scale):
    """Open-loop 100-step rollout: state within tol, contact / termination flags identical at every step."""
    n, T = 8, 100
    qpos, qvel = seeded_states(model, n, seed=11)
    rng = np.random.RandomState(5)
    ctrl = (rng.randn(T, n, 12) * torque_scale).astype(np.float32)
    ref = oracle_rollout(model, qpos, qvel, ctrl.astype(np.float64), mu=(0.8, 0.8), command=(0.6, 0.0, 0.0, 0.2))
    sim = _sim(model, n, cuda_device, precision)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    sim.friction[:] = 0.8
    sim.command[:] = torch.tensor([0.6, 0.0, 0.0, 0.2], device=cuda_device)
    ctrl_d = torch.tensor(ctrl, device=cuda_device)
    max_err = 0.0
    for t in range(T):
        obs, rew, term, trunc = sim.step(ctrl_d[t])
        q = sim.qpos.cpu().numpy().astype(np.float64); q[:, :3] = sim.base_pos64.cpu().numpy()
        v = sim.qvel.cpu().numpy()
        max_err = max(max_err, np.abs(q - ref['qpos'][t]).max(), np.abs(v - ref['qvel'][t]).max())
        cs = obs[:, 199:203].cpu().numpy() > 0.5
        assert (cs == ref['cstate'][t]).all(), f'contact_state differs at step {t}'
        assert (term.cpu().numpy().astype(bool) == ref['term'][t]).all(), f'termination differs at step {t}'
        inv = sim.invalid_body_mask.cpu().numpy().astype(np.int64)
        assert ((inv[:, 0] | (inv[:, 1] << 8)) == ref['invalid'][t]).all()
        assert (sim.ncon.cpu().numpy() == ref['ncon'][t]).all()
        assert (rew.cpu().numpy() == 0).all() and (trunc.cpu().numpy() == 0).all()
    assert max_err < tol, f'state error {max_err} over {T} steps'
    # observation pack at the last step (kinetic energy / work / forces scale with the state -> relative tolerance)
    o_gpu = sim.obs.cpu().numpy()[:, :227].astype(np.float64)
    o_ref = ref['obs'][-1]
    err = np.abs(o_gpu - o_ref) / np.maximum(1.0, np.abs(o_ref))
    assert err.max() < (2e-3 if precision == 0 else 1e-4), f'obs mismatch {err.max()} at column {np.argmax(err.max(axis=0))}'


def test_closed_loop_single_step_error(model, cuda_device):
    """Oracle re-seeded from the GPU state every step isolates single-step error from chaotic divergence."""
    n, T = 4, 60
    qpos, qvel = seeded_states(model, n, seed=21)
    sim = _sim(model, n, cuda_device, 0)
    sim.set_state(torch.tensor(qpos), torch.tensor(qvel))
    rng = np.random.RandomState(2)
    oracles = [Oracle(model) for _ in range(n)]
    worst = 0.0
    for t in range(T):
        ctrl = (rng.randn(n, 12) * 10).astype(np.float32)
        q0 = sim.qpos.cpu().numpy().astype(np.float64); q0[:, :3] = sim.base_pos64.cpu().numpy()
        v0 = sim.qvel.cpu().numpy().astype(np.float64); w0 = sim.qacc_warmstart.cpu().numpy().astype(np.float64)
        sim.step(torch.tensor(ctrl, device=cuda_device))
        q1 = sim.qpos.cpu().numpy().astype(np.float64); q1[:, :3] = sim.base_pos64.cpu().numpy()
        v1 = sim.qvel.cpu().numpy().astype(np.float64)
        for i, o in enumerate(oracles):
            o.set_state(q0[i], v0[i], w0[i])
            o.step(ctrl[i].astype(np.float64))
            qo, vo, _, _ = o.get_state()
            worst = max(worst, np.abs(q1[i] - qo).max(), np.abs(v1[i] - vo).max() * 0.1)
    assert worst < 1e-5, worst


def test_reset_given_state_and_random(model, cuda_device):
    n = 64
    sim = _sim(model, n, cuda_device, 0)
    qpos, qvel = seeded_states(model, n, seed=4)
    obs = sim.reset(qpos=torch.tensor(qpos), qvel=torch.tensor(qvel))
    torch.cuda.synchronize()
    # reset performs one full step with zero ctrl (quadruped_env.py:397): compare with the oracle
    for i in range(0, n, 16):
        o = Oracle(model)
        o.set_state(qpos[i], qvel[i], np.zeros(18))
        o.step(np.zeros(12))
        qo, vo, _, _ = o.get_state()
        np.testing.assert_allclose(sim.qpos[i].cpu().numpy(), qo, atol=2e-6)
        np.testing.assert_allclose(sim.qvel[i].cpu().numpy(), vo, atol=2e-4)
    assert (sim.step_count.cpu().numpy() == 0).all()
    np.testing.assert_allclose(sim.sim_time.cpu().numpy(), 0.002, rtol=1e-6)
    # random reset: distribution bounds and "no foot contact after lifting"
    opt = sim.make_reset_options(lin_vel_range=(0.5, 1.0), ang_vel_range=(-0.3, 0.3), friction_range=(0.2, 1.5), command_mode=1 | 4)
    sim.reset(options=opt)
    torch.cuda.synchronize()
    assert (sim.status.cpu().numpy() & 8 == 0).all(), 'lift loop failed'
    q = sim.qpos.cpu().numpy(); b = sim.base_pos64.cpu().numpy()
    assert np.abs(b[:, :2]).max() <= 1e4 and np.abs(b[:, :2]).max() > 1e3  # U(+-1e4) on flat (quadruped_env.py:352-356)
    key = np.array(model.c.key_qpos)
    assert np.abs(q[:, 7:] - key[7:]).max() <= 20 * np.pi / 180 + 0.5 * 0.002 + 1e-3
    cmd = sim.command.cpu().numpy(); fr = sim.friction.cpu().numpy()
    assert ((cmd[:, 0] >= 0.5) & (cmd[:, 0] <= 1.0)).all() and (cmd[:, 1] == 0).all() and (np.abs(cmd[:, 3]) <= 0.3).all()
    assert ((fr >= 0.2) & (fr <= 1.5)).all() and (fr[:, 0] == fr[:, 1]).all()
    yaw = sim.obs[:, 20].cpu().numpy()
    expect = np.arctan2(-b[:, 1], -b[:, 0])
    d = np.abs(((yaw - expect) + np.pi) % (2 * np.pi) - np.pi)
    assert d.max() < 5e-3  # facing the origin (quadruped_env.py:364); one free-fall step barely changes yaw
    # a second random reset must draw new numbers; same seed in a new sim reproduces the first
    b1 = sim.base_pos64.clone()
    sim.reset(options=opt)
    assert (sim.base_pos64 != b1).any()
    sim2 = _sim(model, n, cuda_device, 0)
    sim2.reset(qpos=torch.tensor(qpos), qvel=torch.tensor(qvel))  # same reset history -> same episode counters
    sim2.reset(options=opt)
    assert torch.equal(sim2.base_pos64, b1)


def test_full_size_batch_properties(model, cuda_device):
    """BASELINE configs[1] size (4096 envs): size-independent properties instead of an oracle run."""
    n = 4096
    sim = _sim(model, n, cuda_device, 0)
    opt = sim.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5))
    sim.reset(options=opt)
    g = torch.Generator(device='cpu').manual_seed(0)
    ctrl = (torch.randn(50, n, 12, generator=g) * 50).to(cuda_device)
    first = None
    for t in range(50):
        obs, _, term, _ = sim.step(ctrl[t])
        if t == 0:
            first = obs.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(sim.obs).all() and (sim.status & 1 == 0).all()
    q = sim.obs[:, 21:25]
    assert torch.allclose(q.norm(dim=1), torch.ones(n, device=cuda_device), atol=1e-5)   # unit quaternion
    R = sim.obs[:, 25:34].reshape(n, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, device=cuda_device).expand(n, 3, 3), atol=1e-5)
    assert (sim.obs[:, 89:101] == ctrl[-1]).all()                                          # tau_ctrl_setpoint is the raw ctrl
    assert torch.equal(sim.obs[:, 52:71][:, 2:], sim.qpos[:, 2:]) and torch.equal(sim.obs[:, 71:89], sim.qvel)
    assert (sim.step_count == 50).all()
    # determinism: same seed, same actions -> bitwise identical first step
    sim2 = _sim(model, n, cuda_device, 0)
    sim2.reset(options=opt)
    obs2, _, _, _ = sim2.step(ctrl[0])
    assert torch.equal(obs2, first)
    # translation invariance on flat ground: shifting the base by whole metres changes nothing but the positions
    sim3 = _sim(model, 8, cuda_device, 0)
    qpos, qvel = seeded_states(model, 8, seed=9)
    a = _sim(model, 8, cuda_device, 0)
    a.set_state(torch.tensor(qpos), torch.tensor(qvel))
    shifted = qpos.copy(); shifted[:, 0] += 5000.0; shifted[:, 1] -= 7000.0
    sim3.set_state(torch.tensor(shifted), torch.tensor(qvel))
    for t in range(20):
        a.step(ctrl[t, :8]); sim3.step(ctrl[t, :8])
    assert torch.equal(a.qvel, sim3.qvel) and torch.equal(a.qpos[:, 2:], sim3.qpos[:, 2:])
    assert torch.allclose(a.base_pos64[:, 0] + 5000.0, sim3.base_pos64[:, 0], atol=1e-9)


def test_host_buffer_entry_point(model, cuda_device):
    n = 32
    qpos, qvel = seeded_states(model, n, seed=6)
    a, b = _sim(model, n, cuda_device, 0), _sim(model, n, cuda_device, 0)
    for s in (a, b):
        s.set_state(torch.tensor(qpos), torch.tensor(qvel))
    ctrl = torch.randn(n, 12) * 10
    obs_h = torch.empty(n, a.obs_dim).pin_memory(); rew_h = torch.empty(n).pin_memory()
    term_h = torch.empty(n, dtype=torch.uint8).pin_memory(); trunc_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    ctrl_h = ctrl.clone().pin_memory()
    a.step_host(ctrl_h, obs_h, rew_h, term_h, trunc_h)
    obs, _, term, _ = b.step(ctrl.to(cuda_device))
    assert torch.equal(obs.cpu(), obs_h) and torch.equal(term.cpu(), term_h)


def test_fused_autoreset_equals_step_then_reset(model, cuda_device):
    """qs_step_autoreset == qs_step followed by qs_reset_done (same counter-based draws), in one launch."""
    n = 512
    a, b = _sim(model, n, cuda_device, 0), _sim(model, n, cuda_device, 0)
    opt = a.make_reset_options(lin_vel_range=(0.5, 1.0), friction_range=(0.2, 1.5))
    for s in (a, b):
        s.reset(options=opt)
    g = torch.Generator(device='cpu').manual_seed(3)
    n_term = 0
    for t in range(150):
        ctrl = (torch.randn(n, 12, generator=g) * 50).to(cuda_device)
        a.step_autoreset(ctrl, opt)
        b.step(ctrl)
        n_term += int(b.terminated.sum())
        term_b = b.terminated.clone()
        b.reset_done(opt)
        assert torch.equal(a.terminated, term_b)
        assert torch.equal(a.qpos, b.qpos) and torch.equal(a.qvel, b.qvel) and torch.equal(a.obs, b.obs)
        assert torch.equal(a.command, b.command) and torch.equal(a.friction, b.friction) and torch.equal(a.step_count, b.step_count)
    assert n_term > 0, 'workload never terminated: the test would be vacuous'
