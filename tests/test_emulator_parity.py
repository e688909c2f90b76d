"""CPU-only parity of the *kernel source*: gym_quadruped_b200/csrc/qs_env.cuh compiled for the host against a 32-thread warp
emulator (tests/emu) versus the fp64 oracle.  The same comparison runs on the real GPU in tests/test_gpu_parity.py; this one
lets `pytest -m "not gpu"` catch algorithmic regressions of the device code without a GPU."""
import numpy as np
import pytest

from gym_quadruped_b200.model import Model
from oracle.oracle import F_BIAS, F_CONTACTS, F_IMU, F_M, F_QACC_SMOOTH, Oracle
from tests.emu.emu import emu_step

ROBOTS = ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1']  # pyramidal/mesh, pyramidal/primitives+limits, elliptic condim 6, elliptic/mesh


def _start(model, rng):
    o = Oracle(model)
    q = np.array(model.c.key_qpos)
    q[7:] += rng.uniform(-0.25, 0.25, 12)
    o.set_state(q, np.zeros(18), np.zeros(18))
    assert o.lift() >= 0
    q = o.get_state()[0]
    v = np.zeros(18)
    v[6:] = rng.uniform(-0.5, 0.5, 12)
    return q.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize('robot', ROBOTS)
def test_forward_pass_fp64_matches_oracle_to_rounding(robot):
    m = Model(robot, 'flat')
    rng = np.random.RandomState(0)
    q, v = _start(m, rng)
    q[2] -= 0.015  # push the feet into the ground: contacts + friction cones active
    v[:6] = rng.uniform(-0.3, 0.3, 6)
    ctrl = rng.randn(12) * 15
    o = Oracle(m)
    o.set_state(q, v, np.zeros(18)); o.set_env(0.7, 0.7, [0.5, 0, 0, 0.1]); o.forward(ctrl)
    e = emu_step(m, q, v, np.zeros(18), ctrl, 0.7, 0.7, [0.5, 0, 0, 0.1], precision=1, mode=0)
    assert e['ncon'] == o.flags()['ncon'] > 0
    np.testing.assert_allclose(e['M'], o.get(F_M), atol=1e-12)
    np.testing.assert_allclose(e['bias'], o.get(F_BIAS), atol=1e-5)  # dumped through a float buffer
    np.testing.assert_allclose(e['qacc_smooth'], o.get(F_QACC_SMOOTH), rtol=1e-9, atol=1e-9)
    qa = o.get_state()[2]
    np.testing.assert_allclose(e['qacc'], qa, atol=1e-8 * max(1.0, np.abs(qa).max()))
    oc, ec = o.get(F_CONTACTS), e['contacts']
    oc, ec = oc[np.argsort(oc[:, 16], kind='stable')], ec[np.argsort(ec[:, 16], kind='stable')]
    assert (oc[:, 16:18] == ec[:, 16:18]).all()
    np.testing.assert_allclose(np.sort(ec[:, 0]), np.sort(oc[:, 0]), atol=1e-12)
    np.testing.assert_allclose(ec[:, 13].sum(), oc[:, 13].sum(), rtol=1e-7)
    if m.c.has_imu:
        np.testing.assert_allclose(e['imu'], o.get(F_IMU), atol=1e-9)


@pytest.mark.parametrize('robot', ROBOTS)
@pytest.mark.parametrize('precision,tol', [(1, 1e-10), (0, 1e-4)])
def test_rollout_matches_oracle(robot, precision, tol):
    """30 contact-rich steps: state within tol, contact / termination flags and Newton iteration counts (fp64) identical."""
    m = Model(robot, 'flat')
    rng = np.random.RandomState(7)
    q, v = _start(m, rng)
    o = Oracle(m)
    o.set_state(q, v, np.zeros(18)); o.set_env(0.9, 0.9, [0.6, 0, 0, 0.0])
    eq, ev, ew = q.copy(), v.copy(), np.zeros(18)
    scale = 0.08 * np.abs(np.array(m.c.act_ctrlrange)).max()
    worst = 0.0
    for k in range(30):
        ctrl = (rng.randn(12) * scale).astype(np.float32).astype(np.float64)
        obs, term = o.step(ctrl)
        e = emu_step(m, eq, ev, ew, ctrl, 0.9, 0.9, [0.6, 0, 0, 0.0], precision=precision, tol=1e-8 if precision else 1e-6, mode=1)
        eq, ev, ew = e['qpos'], e['qvel'], e['qacc']
        oq, ov, _, _ = o.get_state()
        f = o.flags()
        assert e['contact_mask'] == sum(int(b) << i for i, b in enumerate(f['contact_state']))
        assert e['invalid_mask'] == f['invalid_body_mask'] and e['ncon'] == f['ncon']
        if precision == 1:
            assert e['iters'] == f['solver_iter']
        worst = max(worst, np.abs(eq - oq).max(), np.abs(ev - ov).max())
        err = np.abs(e['obs'] - obs[:227]) / np.maximum(1.0, np.abs(obs[:227]))
        assert err.max() < (1e-8 if precision == 1 else 5e-3), f'obs column {np.argmax(err)} step {k}'
    assert worst < tol, worst
