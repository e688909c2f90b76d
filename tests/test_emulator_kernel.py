"""CPU-only checks of the KERNEL BODY (gym_quadruped_b200/csrc/qs_kernel.cuh): the step / reset kernel, compiled for the host and
run on the warp emulator (tests/emu/emu_kernel.cpp, one emulated CTA per env, fp64 arithmetic, fp32 state buffers as on the device),
against the fp64 oracle and against independent restatements of what the kernel adds around `mj_step`:

  * reset with a given state (quadruped_env.py:389-397), the random reset (:346-373: keyframe + noise, yaw towards the origin, lift
    loop :376-388, one step, command / friction sampling :400-404) reconstructed draw by draw from an independent Philox,
  * the same-launch auto-reset against step + masked reset,
  * the in-kernel command / disturbance schedules (:293-305) against the same Philox restatement,
  * the non-finite-state rule.

The GPU runs the same source (tests/test_gpu_*.py); this file lets `pytest -m "not gpu"` see regressions of the kernel-level logic."""
import math

import numpy as np
import pytest

from gym_quadruped_b200.model import Model, QsResetOptions
from oracle.oracle import Oracle
from tests.emu.emu import EmuSim, philox4x32, u32_to_unit

F32 = np.float32
CMD_FORWARD, CMD_RANDOM, CMD_ROTATE, CMD_RESET = 1, 2, 4, 8


def reset_options(model, randomize=True, lin=(0.5, 1.0), ang=(-0.3, 0.3), fric=(0.2, 1.5), mode=CMD_FORWARD | CMD_ROTATE):
    o = QsResetOptions()
    o.angle_sweep, o.vel_sweep, o.roll_sweep, o.pitch_sweep = 20 * math.pi / 180, 0.5, 10 * math.pi / 180, 10 * math.pi / 180
    o.hip_height = model.hip_height
    o.lin_vel_range[:] = lin; o.ang_vel_range[:] = ang; o.friction_range[:] = fric
    o.command_mode, o.randomize = mode, int(randomize)
    return o


def standing(model, n, seed):
    rng = np.random.RandomState(seed)
    q = np.tile(np.array(model.c.key_qpos), (n, 1))
    q[:, 7:] += rng.uniform(-0.2, 0.2, (n, 12))
    q[:, 2] = model.hip_height * rng.uniform(0.85, 1.0, n)
    v = rng.uniform(-0.5, 0.5, (n, 18))
    return q, v


@pytest.mark.parametrize('robot', ['mini_cheetah', 'go2', 'aliengo'])
def test_kernel_steps_match_oracle(robot):
    m = Model(robot, 'flat')
    n, T = 3, 25
    q, v = standing(m, n, 1)
    s = EmuSim(m, n, precision=1)
    s.set_state(q, v)
    orcs = []
    for i in range(n):
        o = Oracle(m); o.set_state(s.qpos[i].astype(float), s.qvel[i].astype(float), np.zeros(18)); o.set_env(-1.0, -1.0, [0, 0, 0, 0]); orcs.append(o)
    rng = np.random.RandomState(2)
    for t in range(T):
        ctrl = (rng.randn(n, 12) * 8).astype(F32)
        s.step(ctrl)
        for i, o in enumerate(orcs):
            obs, term = o.step(ctrl[i].astype(float))
            assert bool(s.terminated[i]) == term
            np.testing.assert_allclose(s.obs[i, :227], obs[:227], atol=2e-4, rtol=2e-4)
            # the device buffers are fp32: re-seed the oracle from them so that storage rounding does not accumulate
            qo, vo, _, wo = o.get_state()
            np.testing.assert_allclose(s.qpos[i], qo, atol=2e-6)
            np.testing.assert_allclose(s.qvel[i], vo, atol=2e-5)
            o.set_state(np.r_[s.base_pos64[i], s.qpos[i, 3:].astype(float)], s.qvel[i].astype(float), s.qacc_warmstart[i].astype(float))
    assert (s.step_count == T).all() and abs(s.sim_time[0] - T * 0.002) < 1e-6 and (s.ncon > 0).any()


def test_kernel_reset_with_given_state():
    m = Model('mini_cheetah', 'flat')
    n = 3
    q, v = standing(m, n, 3)
    s = EmuSim(m, n, precision=1)
    s.qacc_warmstart[:] = 5.0; s.qfrc_applied[:] = 3.0; s.step_count[:] = 17; s.sim_time[:] = 1.0   # must all be cleared (:332-335,:394-395)
    s.reset(reset_options(m, randomize=False), qpos=q.astype(F32), qvel=v.astype(F32))
    for i in range(n):
        o = Oracle(m)
        o.set_state(np.r_[q[i, :3], q[i, 3:].astype(F32).astype(float)], v[i].astype(F32).astype(float), np.zeros(18))
        o.set_env(-1.0, -1.0, [0, 0, 0, 0])
        obs, _ = o.step(np.zeros(12))                       # reset ends with one step at zero ctrl (:397)
        qo, vo, _, _ = o.get_state()
        np.testing.assert_allclose(s.qpos[i, 2:], qo[2:], atol=2e-6)
        np.testing.assert_allclose(s.base_pos64[i], qo[:3], atol=1e-5)
        np.testing.assert_allclose(s.qvel[i], vo, atol=2e-5)
    assert (s.step_count == 0).all() and np.allclose(s.sim_time, 0.002) and (s.qfrc_applied == 0).all() and (s.episode == 1).all()


def _reconstruct_random_reset(m, ro, env_g, ep, seed):
    """The random reset, draw by draw (qs_kernel.cuh reset pass; quadruped_env.py:346-373), from the independent Philox."""
    u = np.zeros(40, F32)
    r9 = None
    for lane in range(10):
        r = philox4x32(env_g, ep, lane, 0x5EED, seed & 0xffffffff, seed >> 32)
        for i in range(4):
            u[4 * lane + i] = u32_to_unit(r[i])
        if lane == 9:
            r9 = r
    d0 = (r9[0] * 4294967296.0 + r9[1]) / 18446744073709551616.0
    d1 = (r9[2] * 4294967296.0 + r9[3]) / 18446744073709551616.0
    q = np.array(m.c.key_qpos, dtype=float); v = np.zeros(18)
    for j in range(12):
        q[7 + j] += -ro.angle_sweep + 2 * ro.angle_sweep * float(u[j])
        v[6 + j] += -ro.vel_sweep + 2 * ro.vel_sweep * float(u[12 + j])
    roll = -ro.roll_sweep + 2 * ro.roll_sweep * float(u[24]); pitch = -ro.pitch_sweep + 2 * ro.pitch_sweep * float(u[25])
    tl = [float(x) for x in m.c.terrain_limits]
    bx = tl[0] + (tl[1] - tl[0]) * d0; by = tl[2] + (tl[3] - tl[2]) * d1
    yaw = math.atan2(-by, -bx)
    cr, sr, cp, sp, cy, sy = math.cos(roll / 2), math.sin(roll / 2), math.cos(pitch / 2), math.sin(pitch / 2), math.cos(yaw / 2), math.sin(yaw / 2)
    q[3:7] = [cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy]
    q[0], q[1], q[2] = bx, by, ro.hip_height
    return q, v, [float(x) for x in u[26:31]]


@pytest.mark.parametrize('robot,scene', [('mini_cheetah', 'flat'), ('aliengo', 'flat'), ('go2', 'random_boxes'), ('aliengo', 'perlin'),
                                         ('go2', 'perlin'), ('b2', 'perlin'), ('go1', 'random_pyramids'), ('spot', 'stairs')])
def test_kernel_random_reset_reconstructed(robot, scene):
    """Every number of a random reset is reproduced outside the kernel: the Philox draws and where they go, the yaw towards the origin,
    the lift loop (oracle: the reference's loop; kernel: closed-form recurrence on the flat floor, calf-only collision passes on
    terrain), the closing step with the OLD friction, and the command / friction sampled after it.  The perlin cases spawn robots
    deep inside the hills (hip height above z = 0): dozens of calf contacts overflow the 16-slot contact buffer, and the lift must
    still use the deepest DETECTED one (go2 / b2 / go1: a dropped contact used to shorten the first lift)."""
    m = Model(robot, scene)
    n, seed, off = 6, 0x1234ABCD5, 40
    ro = reset_options(m, mode=CMD_FORWARD | CMD_ROTATE | CMD_RESET)
    s = EmuSim(m, n, precision=1, seed=seed, env_id_offset=off)
    s.episode[:] = np.arange(n) + 2
    s.friction[:] = 0.7
    s.reset(ro)
    lifted = 0
    for i in range(n):
        q, v, ul = _reconstruct_random_reset(m, ro, i + off, i + 2, seed)
        o = Oracle(m)
        o.set_state(q, v, np.zeros(18))
        k = o.lift()
        lifted += k > 0
        assert (k >= 0) == (not (s.status[i] & 8))
        o.set_env(0.7, 0.7, [0, 0, 0, 0])
        o.step(np.zeros(12))
        qo, vo, _, _ = o.get_state()
        np.testing.assert_allclose(s.base_pos64[i], qo[:3], atol=2e-6, err_msg=f'env {i}')
        np.testing.assert_allclose(s.qpos[i, 3:], qo[3:], atol=2e-6)
        np.testing.assert_allclose(s.qvel[i], vo, atol=5e-5)
        vn = ro.lin_vel_range[0] + (ro.lin_vel_range[1] - ro.lin_vel_range[0]) * ul[0]
        yr = ro.ang_vel_range[0] + (ro.ang_vel_range[1] - ro.ang_vel_range[0]) * ul[2]
        mu = ro.friction_range[0] + (ro.friction_range[1] - ro.friction_range[0]) * ul[3]
        np.testing.assert_allclose(s.command[i], [vn, 0, 0, yr], atol=1e-6)
        np.testing.assert_allclose(s.friction[i], [mu, mu], atol=1e-6)
        assert s.cmd_limit[i] == 1000 + int(F32(ul[4]) * F32(2000.0)) and s.cmd_count[i] == 0
    assert (s.episode == np.arange(n) + 3).all() and (s.step_count == 0).all()
    assert lifted > 0 or scene == 'flat'


def _fallen(model, n, seed):
    """States that terminate at the next step (a non-foot geom on the ground), found with the oracle: the robot lying on its side."""
    q, v = standing(model, n, seed)
    o = Oracle(model)
    for i in range(n):
        for z in (0.3, 0.2, 0.15, 0.1, 0.07, 0.05, 0.03):
            for quat in ([math.sqrt(0.5), math.sqrt(0.5), 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]):
                q[i, 2] = z; q[i, 3:7] = quat
                o.set_state(q[i], v[i], np.zeros(18)); o.set_env(-1.0, -1.0, [0, 0, 0, 0])
                if o.step(np.zeros(12))[1]:
                    break
            else:
                continue
            break
        else:
            raise AssertionError('no terminating state found')
    return q, v


@pytest.mark.parametrize('robot,scene', [('mini_cheetah', 'flat'), ('go2', 'random_boxes')])
def test_kernel_autoreset_equals_step_then_masked_reset(robot, scene):
    m = Model(robot, scene)
    n = 4
    q, v = standing(m, n, 5)
    qf, vf = _fallen(m, n, 6)
    q[1], v[1], q[3], v[3] = qf[1], vf[1], qf[3], vf[3]
    ro = reset_options(m)
    a, b = EmuSim(m, n, precision=1, seed=9), EmuSim(m, n, precision=1, seed=9)
    for s in (a, b):
        s.set_state(q, v); s.friction[:] = 0.9; s.command[:, 0] = 0.4
    ctrl = (np.random.RandomState(7).randn(n, 12) * 5).astype(F32)
    a.step_autoreset(ctrl, ro)
    b.step(ctrl)
    term = b.terminated.copy()
    assert term.tolist() == [0, 1, 0, 1]
    b.reset(ro, mask=term)
    assert (a.terminated == term).all()
    for name in ('qpos', 'qvel', 'qacc', 'qacc_warmstart', 'base_pos64', 'command', 'friction', 'step_count', 'sim_time', 'episode', 'obs'):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_kernel_schedules_reconstructed():
    m = Model('mini_cheetah', 'flat')
    n, seed = 3, 77
    s = EmuSim(m, n, precision=1, seed=seed, env_id_offset=5)
    q, v = standing(m, n, 8)
    s.set_state(q, v)
    sc = s.sched
    sc.command_mode, sc.ext_enabled = CMD_FORWARD | CMD_RANDOM | CMD_ROTATE | CMD_RESET, 1
    sc.lin_vel_range[:] = (0.3, 0.9); sc.ang_vel_range[:] = (-0.4, 0.4)
    sc.ext_lo[:] = (-20, 0, 5, 0, 0, -1); sc.ext_hi[:] = (20, 0, 5, 0, 0, 1)
    s.cmd_limit[:] = [3, 2, 1000]; s.ext_limit[:] = [2, 1000, 4]
    s.cmd_epoch[:] = [0, 4, 0]; s.ext_epoch[:] = [1, 0, 2]
    s.ext_wrench[:] = 0.25
    cmd0 = s.command.copy()
    for t in range(1, 5):
        s.step(np.zeros((n, 12), F32))
        for i in range(n):
            g = i + 5
            if (i, t) in ((0, 3), (1, 2), (1, 4)):     # command resampled when the count reaches the limit (:293-296, :1046-1072)
                ce = {0: 0, 1: 4 + (t == 4)}[i]
                r = philox4x32(g, ce, 0, 0xC3D0, (seed & 0xffffffff) ^ 0x51ED270B, seed >> 32)
                vn = 0.3 + 0.6 * float(u32_to_unit(r[0])); ang = -math.pi + 2 * math.pi * float(u32_to_unit(r[1]))
                yr = -0.4 + 0.8 * float(u32_to_unit(r[2]))
                np.testing.assert_allclose(s.command[i], [vn * math.cos(ang), vn * math.sin(ang), 0, yr], atol=1e-6)
                assert s.cmd_limit[i] == 1000 + int(u32_to_unit(r[3]) * F32(2000.0)) and s.cmd_count[i] == 0
                if i == 1 and t == 2:
                    s.cmd_limit[1] = 2                 # keep the second env on a short cadence for another draw
        if t == 1:
            assert np.array_equal(s.command, cmd0) and (s.qfrc_applied == 0.25).all()   # wrench copied after every step (:305)
    # disturbance of env 0 was due at t = 2 (count 2 >= limit 2), of env 2 at t = 4
    for i, ee in ((0, 1), (2, 2)):
        g = i + 5
        w = [float(sc.ext_lo[k]) + (float(sc.ext_hi[k]) - float(sc.ext_lo[k])) * float(u32_to_unit(philox4x32(g, ee, k, 0xD157, (seed & 0xffffffff) ^ 0x7F4A7C15, seed >> 32)[0]))
             for k in range(6)]
        np.testing.assert_allclose(s.ext_wrench[i], w, atol=1e-5)
        np.testing.assert_allclose(s.qfrc_applied[i], w, atol=1e-5)
        assert s.ext_epoch[i] == ee + 1
    assert (s.ext_wrench[1] == 0.25).all() and s.ext_epoch[1] == 0 and s.cmd_epoch.tolist() == [1, 6, 0]


def test_kernel_non_finite_state_terminates_and_autoreset_recovers():
    m = Model('mini_cheetah', 'flat')
    n = 3
    q, v = standing(m, n, 9)
    ro = reset_options(m)
    a = EmuSim(m, n, precision=1, seed=3)
    a.set_state(q, v)
    a.qvel[1, 4] = np.nan
    a.step(np.zeros((n, 12), F32))
    assert a.terminated.tolist() == [0, 1, 0] and a.status[1] & 1 and not a.status[0] & 1
    b = EmuSim(m, n, precision=1, seed=3)
    b.set_state(q, v)
    b.qpos[2, 8] = np.inf
    b.step_autoreset(np.zeros((n, 12), F32), ro)
    assert b.terminated.tolist() == [0, 0, 1] and np.isfinite(b.qpos).all() and np.isfinite(b.qvel).all() and np.isfinite(b.obs).all()
    assert b.step_count.tolist() == [1, 1, 0]


def test_kernel_imu_columns_reconstructed():
    """sensors/imu.py:102-139 inside the kernel: reading = truth + bias + N(0, sigma), bias random walk; Box-Muller on Philox draws
    keyed by (global env id, per-env step counter).  Truth signals against the oracle, noise and bias against the restatement."""
    m = Model('hyqreal1', 'flat')
    assert m.c.has_imu
    n, seed, off = 2, 99, 7
    s = EmuSim(m, n, precision=1, seed=seed, use_imu=True, env_id_offset=off)
    s.imu_noise = (0.05, 0.01, 0.002, 0.0005)
    q, v = standing(m, n, 4)
    s.set_state(q, v)
    s.tick[:] = [3, 10]
    s.imu_bias[:] = 0.01
    orcs = []
    for i in range(n):
        o = Oracle(m); o.set_state(s.qpos[i].astype(float), s.qvel[i].astype(float), np.zeros(18)); o.set_env(-1.0, -1.0, [0, 0, 0, 0]); orcs.append(o)
    bias = s.imu_bias.copy()
    for t in range(3):
        ctrl = np.zeros((n, 12), F32)
        s.step(ctrl)
        for i, o in enumerate(orcs):
            obs, _ = o.step(np.zeros(12))
            truth = obs[227:233]
            io = s.obs[i, 227:245]
            for lane in range(3):
                r = philox4x32(i + off, int([3, 10][i]) + t, lane, 0x1A2B, (seed & 0xffffffff) ^ 0x9E3779B9, seed >> 32)
                u1, u2, u3, u4 = (max(float(u32_to_unit(r[0])), 5.9604645e-8), float(u32_to_unit(r[1])), max(float(u32_to_unit(r[2])), 5.9604645e-8), float(u32_to_unit(r[3])))
                ra, rb = math.sqrt(-2 * math.log(u1)), math.sqrt(-2 * math.log(u3))
                n_acc, n_ab = ra * math.cos(2 * math.pi * u2) * 0.05, ra * math.sin(2 * math.pi * u2) * 0.002
                n_gyr, n_gb = rb * math.cos(2 * math.pi * u4) * 0.01, rb * math.sin(2 * math.pi * u4) * 0.0005
                bias[i, lane] += n_ab; bias[i, 3 + lane] += n_gb
                np.testing.assert_allclose([io[3 + lane], io[6 + lane], io[12 + lane], io[15 + lane]], [n_acc, bias[i, lane], n_gyr, bias[i, 3 + lane]], atol=2e-6)
                np.testing.assert_allclose(io[lane], truth[lane] + bias[i, lane] + n_acc, atol=2e-3, rtol=1e-4)
                np.testing.assert_allclose(io[9 + lane], truth[3 + lane] + bias[i, 3 + lane] + n_gyr, atol=1e-4)
            qo, vo, _, _ = o.get_state()
            o.set_state(np.r_[s.base_pos64[i], s.qpos[i, 3:].astype(float)], s.qvel[i].astype(float), s.qacc_warmstart[i].astype(float))
    np.testing.assert_allclose(s.imu_bias, bias, atol=2e-6)
    assert s.tick.tolist() == [6, 13]


def test_kernel_heightmap_columns():
    """sensors/heightmap.py:106-169 as extra observation columns: the grid is cast around the post-step base position / heading."""
    m = Model('aliengo', 'perlin')
    n = 2
    s = EmuSim(m, n, precision=1, heightmap=(3, 4, 0.1, 0.15))
    q, v = standing(m, n, 12)
    q[:, 0] = [3.0, -2.5]; q[:, 1] = [2.0, 4.0]; q[:, 2] = 1.1
    s.set_state(q, v)
    s.step(np.zeros((n, 12), F32))
    for i in range(n):
        o = Oracle(m)
        center = np.r_[s.base_pos64[i, :2], float(s.qpos[i, 2])]
        hm = o.heightmap(center, float(s.obs[i, 20]), 3, 4, 0.1, 0.15)
        np.testing.assert_allclose(s.obs[i, 227:].reshape(3, 4, 3), hm, atol=2e-5)
    assert s.obs.shape[1] == 227 + 36 and np.ptp(s.obs[:, 227:].reshape(n, -1, 3)[:, :, 2]) > 1e-3


@pytest.mark.parametrize('robot', ['mini_cheetah', 'go2'])
def test_kernel_forward_tables(robot):
    """MODE_FORWARD (qs_forward + qs_get: mj_fullM, qfrc_bias, qfrc_passive, mj_jac, subtree_com users, quadruped_env.py:543-929): the
    accessor tables the kernel dumps, against the oracle's."""
    from oracle.oracle import F_BIAS, F_COM, F_FEET_JACP, F_FEET_POS, F_M, F_PASSIVE
    m = Model(robot, 'flat')
    n = 2
    q, v = standing(m, n, 21)
    s = EmuSim(m, n, precision=1)
    s.set_state(q, v)
    s.forward()
    for i in range(n):
        o = Oracle(m)
        o.set_state(np.r_[s.base_pos64[i], s.qpos[i, 3:].astype(float)], s.qvel[i].astype(float), np.zeros(18))
        o.set_env(-1.0, -1.0, [0, 0, 0, 0])
        o.forward(np.zeros(12))
        a = s.aux[i]
        np.testing.assert_allclose(a[0:324].reshape(18, 18), o.get(F_M), atol=2e-5, rtol=1e-5)
        np.testing.assert_allclose(a[324:342], o.get(F_BIAS)[:18], atol=2e-4, rtol=1e-5)
        np.testing.assert_allclose(a[342:360], o.get(F_PASSIVE)[:18], atol=1e-5)
        np.testing.assert_allclose(a[360:576].reshape(4, 3, 18), o.get(F_FEET_JACP), atol=1e-5)
        np.testing.assert_allclose(a[576:588].reshape(4, 3), o.get(F_FEET_POS), atol=1e-5)
        np.testing.assert_allclose(a[588:591], o.get(F_COM)[:3], atol=1e-5)


@pytest.mark.parametrize('stride', [256, 229])
def test_kernel_gather_rows_and_flags(stride):
    """The fused observation gather (qs_gather_*) on host stand-ins for the peer-mapped tensors: every env's row lands in
    this rank's block of EVERY rank's gathered tensor (also for a row stride that leaves the rows off 16-byte boundaries: scalar
    head / float4 body / scalar tail), an env that auto-resets sends its post-reset row, and the last env raises the flags."""
    m = Model('mini_cheetah', 'flat')
    n, world, rank = 4, 3, 1
    q, v = standing(m, n, 31)
    qf, vf = _fallen(m, n, 32)
    q[2], v[2] = qf[2], vf[2]
    ro = reset_options(m)
    ref = EmuSim(m, n, precision=1, seed=4)
    ref.set_state(q, v)
    ctrl = (np.random.RandomState(1).randn(n, 12) * 5).astype(F32)
    ref.step_autoreset(ctrl, ro)
    assert ref.terminated[2] == 1 and (ref.terminated == 0).any()
    s = EmuSim(m, n, precision=1, seed=4)
    s.set_state(q, v)
    peers = [np.full((world * n, stride), -7.0, F32) for _ in range(world)]
    flags = [np.zeros(8, np.uint32) for _ in range(world)]
    s.gather = dict(world=world, rank=rank, stride=stride, seq=5, peers=peers, flags=flags)
    s.step_autoreset(ctrl, ro)
    for k in range(world):
        block = peers[k][rank * n:(rank + 1) * n]
        assert np.array_equal(block[:, :227], ref.obs), f'rank {k}'
        assert (block[:, 227:] == -7.0).all()                                        # padding untouched
        others = np.delete(peers[k], np.s_[rank * n:(rank + 1) * n], axis=0)
        assert (others == -7.0).all()                                                # nobody else's rows touched
        assert flags[k].tolist() == [0, 5, 0, 0, 0, 0, 0, 0]
    assert np.array_equal(s.qpos, ref.qpos) and np.array_equal(s.terminated, ref.terminated)


def _random_states(m, scene, n, rng, low=0.05):
    q = np.tile(np.array(m.c.key_qpos), (n, 1))
    q[:, 7:] += rng.uniform(-0.6, 0.6, (n, 12))
    if scene != 'flat':
        q[:, 0:2] = rng.uniform(-4, 4, (n, 2))
    for i in range(n):
        ax = rng.randn(3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(0, 0.5) if rng.rand() < 0.6 else rng.uniform(0, np.pi)
        q[i, 3:7] = np.r_[np.cos(ang / 2), np.sin(ang / 2) * ax]
    q[:, 2] = rng.uniform(low, 1.2 * m.hip_height + (0.7 if scene != 'flat' else 0.0), n)
    return q, rng.uniform(-2, 2, (n, 18))


ROBOTS = ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1', 'hyqreal2', 'go1', 'b2', 'spot']


def test_kernel_autoreset_fuzz_all_robots():
    """Fused auto-reset == step + masked reset, bit for bit on every buffer, for random (often terminating) states of all eight robots
    on four scenes, IMU noise on where the robot has one.  (720 further cases of the generator were run once: no difference.)"""
    scenes = ['flat', 'random_boxes', 'perlin', 'stairs']
    rng = np.random.RandomState(5)
    nterm = 0
    for it in range(32):
        robot, scene = ROBOTS[it % 8], scenes[(it // 8) % 4]
        m = Model(robot, scene)
        n = 4
        q, v = _random_states(m, scene, n, rng)
        ctrl = (rng.randn(n, 12) * 30).astype(F32)
        ro = reset_options(m, mode=CMD_FORWARD | CMD_ROTATE | CMD_RESET)
        a, b = (EmuSim(m, n, precision=1, seed=it, use_imu=bool(m.c.has_imu)) for _ in range(2))
        for s in (a, b):
            s.set_state(q, v); s.friction[:] = 0.9; s.imu_noise = (0.05, 0.01, 0.002, 0.0005)
        a.step_autoreset(ctrl, ro)
        b.step(ctrl)
        term = b.terminated.copy()
        b.reset(ro, mask=term)
        nterm += int(term.sum())
        assert np.array_equal(a.terminated, term), (robot, scene)
        for name in ('qpos', 'qvel', 'qacc', 'qacc_warmstart', 'base_pos64', 'command', 'friction', 'step_count', 'sim_time', 'episode', 'status',
                     'cmd_count', 'cmd_limit', 'imu_bias', 'tick'):
            assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), (robot, scene, name)
        assert np.array_equal(a.obs[:, :227], b.obs[:, :227], equal_nan=True), (robot, scene)
    assert nterm >= 30


def test_kernel_step_fuzz_observation_rows_and_heightmap():
    """Random states of all robots on all seven scenes, one kernel step each: the packed observation row, the termination flag and the
    height-map columns against the oracle (fed with the fp32 state the kernel sees)."""
    scenes = ['flat', 'random_boxes', 'perlin', 'stairs', 'ramp', 'random_pyramids', 'slippery']
    rng = np.random.RandomState(0)
    compared = 0
    for it in range(36):
        robot, scene = ROBOTS[rng.randint(8)], scenes[rng.randint(7)]
        m = Model(robot, scene)
        n = 3
        q, v = _random_states(m, scene, n, rng, low=0.1)
        ctrl = (rng.randn(n, 12) * 30).astype(F32)
        hm = (3, 3, 0.1, 0.1) if scene != 'flat' else None
        s = EmuSim(m, n, precision=1, heightmap=hm)
        s.set_state(q, v); s.friction[:] = 0.8; s.command[:] = [0.4, 0.1, 0, 0.2]
        orcs = []
        for i in range(n):
            o = Oracle(m)
            o.set_state(np.r_[s.base_pos64[i], s.qpos[i, 3:].astype(float)], s.qvel[i].astype(float), np.zeros(18)); o.set_env(0.8, 0.8, [0.4, 0.1, 0, 0.2])
            orcs.append(o)
        s.step(ctrl)
        for i, o in enumerate(orcs):
            obs, term = o.step(ctrl[i].astype(float))
            if o.flags()['ncon'] > 16 or not np.isfinite(obs).all():
                continue
            assert bool(s.terminated[i]) == term, (robot, scene, it, i)
            rel = np.abs(s.obs[i, :227] - obs[:227]) / (1 + np.abs(obs[:227]))
            assert rel.max() < 3e-5, (robot, scene, it, i, int(np.argmax(rel)), float(rel.max()))
            if hm:
                want = Oracle(m).heightmap(np.r_[s.base_pos64[i, :2], float(s.qpos[i, 2])], float(s.obs[i, 20]), 3, 3, 0.1, 0.1)
                assert np.abs(s.obs[i, 227:].reshape(3, 3, 3) - want).max() < 5e-5, (robot, scene, it, i)
            compared += 1
    assert compared >= 60
