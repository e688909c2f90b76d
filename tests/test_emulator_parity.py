"""CPU-only parity of the *kernel source*: gym_quadruped_b200/csrc/qs_env.cuh compiled for the host against a 32-lane (fiber) warp
emulator (tests/emu) versus the fp64 oracle.  The same comparison runs on the real GPU in tests/test_gpu_parity.py; this one
lets `pytest -m "not gpu"` catch algorithmic regressions of the device code without a GPU."""
import numpy as np
import pytest

from gym_quadruped_b200.model import Model
from oracle.oracle import F_BIAS, F_CONTACTS, F_IMU, F_M, F_QACC_SMOOTH, Oracle
from tests.emu.emu import emu_step

ROBOTS = ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1', 'hyqreal2', 'b2', 'go1', 'spot']  # pyramidal/mesh, pyramidal/primitives+limits, elliptic condim 6, elliptic/mesh


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
            # with no active row the oracle's first line search returns alpha = 0 (0 iterations) while the kernel, which never
            # evaluates alpha = 0, accepts the unit step along a direction of rounding-level length (1 iteration): same qacc
            assert e['iters'] == f['solver_iter'] or (f['solver_iter'] == 0 and e['iters'] == 1)
        worst = max(worst, np.abs(eq - oq).max(), np.abs(ev - ov).max())
        err = np.abs(e['obs'] - obs[:227]) / np.maximum(1.0, np.abs(obs[:227]))
        assert err.max() < (1e-8 if precision == 1 else 5e-3), f'obs column {np.argmax(err)} step {k}'
    assert worst < tol, worst


def _keep_geoms(model, keep):
    """Compact the per-geom tables of a (fresh) Model to the geoms in `keep`: isolates one collider for a directed test."""
    c = model.c
    for name, ctype in c._fields_:
        arr = getattr(c, name)
        if name.startswith('geom_') and hasattr(arr, '__len__') and len(arr) == len(c.geom_type):
            vals = [arr[g] for g in keep]
            for k, v in enumerate(vals):
                if hasattr(v, '__len__'):
                    for i in range(len(v)):
                        arr[k][i] = v[i]
                elif hasattr(v, '_fields_'):
                    for fn, _ in v._fields_:
                        fv = getattr(v, fn)
                        if hasattr(fv, '__len__'):
                            for i in range(len(fv)):
                                getattr(arr[k], fn)[i] = fv[i]
                        else:
                            setattr(arr[k], fn, fv)
                else:
                    arr[k] = v
    for leg in range(4):
        c.foot_geom[leg] = keep.index(c.foot_geom[leg])
    c.ngeom = len(keep)
    return model


@pytest.mark.parametrize('roll', [1.45, 1.2, -1.5707963267948966, 0.9])
def test_cylinder_plane_collider_matches_oracle(roll):
    """b2 reduced to its feet and hip cylinders (b2.xml:96), lying on its side: the plane-cylinder routine (rim point of the lower
    cap, its twin, the two +-120 degree points; including the disk-parallel-to-plane branch at roll = -pi/2) is compared contact
    by contact between the kernel source and the oracle."""
    m = Model('b2', 'flat')
    cyl = [g for g in range(m.c.ngeom) if m.c.geom_type[g] == 5]
    assert len(cyl) == 4
    m = _keep_geoms(m, sorted(set(cyl) | set(m.c.foot_geom)))
    q = np.array(m.c.key_qpos)
    q[3:7] = [np.cos(roll / 2), np.sin(roll / 2), 0, 0]
    v = np.zeros(18)
    o = Oracle(m)
    for z in np.arange(0.6, 0.0, -0.004):  # lower the body until cylinders touch the floor
        q[2] = z
        q = q.astype(np.float32).astype(np.float64)
        o.set_state(q, v, np.zeros(18)); o.forward(np.zeros(12))
        oc = o.get(F_CONTACTS)
        if len(oc) and (oc[:, 16] < 99).sum() and np.isin(m.c.geom_type[:m.c.ngeom], 5)[oc[:, 16].astype(int)].sum() >= 3:
            break
    ncyl = np.isin(m.c.geom_type[:m.c.ngeom], 5)[oc[:, 16].astype(int)].sum()
    assert ncyl >= 3 and len(oc) <= 16, 'pose does not exercise the cylinder collider'
    e = emu_step(m, q, v, np.zeros(18), np.zeros(12), -1.0, -1.0, [0, 0, 0, 0], precision=1, mode=0)
    ec = e['contacts']
    assert e['ncon'] == len(oc)
    key = lambda c: np.lexsort((np.round(c[:, 2], 9), np.round(c[:, 1], 9), c[:, 16]))
    oc, ec = oc[key(oc)], ec[key(ec)]
    assert (oc[:, 16:18] == ec[:, 16:18]).all()
    np.testing.assert_allclose(ec[:, 0:13], oc[:, 0:13], atol=1e-9)   # dist, position, frame


@pytest.mark.parametrize('robot,scene,xy,z0', [('aliengo', 'perlin', (3.0, 2.0), 0.95), ('go2', 'random_boxes', (2.0, -1.0), 0.45),
                                               ('aliengo', 'stairs', (1.6, 0.0), 0.85), ('hyqreal2', 'random_pyramids', (3.0, 0.5), 1.32),
                                               ('aliengo', 'slippery', (12.0, 0.1), 0.55), ('b2', 'ramp', (1.0, 0.1), 0.8)])
def test_terrain_scenes_closed_loop(robot, scene, xy, z0):
    """Terrain colliders of the kernel source (height field, static boxes incl. the priority-2 strips of `slippery`) against the
    oracle, closed loop in fp64: the emulated step is re-seeded from the oracle's state every step while the robot drops onto the
    terrain under a PD hold; contact count / flags identical, state to rounding."""
    m = Model(robot, scene)
    rng = np.random.RandomState(4)
    key = np.array(m.c.key_qpos)
    q = key.copy()
    q[0:2] = np.array(xy) + rng.uniform(-0.3, 0.3, 2); q[2] = z0
    q[7:] += rng.uniform(-0.15, 0.15, 12)
    o = Oracle(m)
    o.set_state(q, np.zeros(18), np.zeros(18)); assert o.lift() >= 0
    o.set_env(0.8, 0.8, [0.5, 0, 0, 0])
    kp = 40.0 if m.c.body_mass[1] < 20 else 400.0
    max_ncon, worst = 0, 0.0
    for k in range(120):
        q0, v0, _, w0 = o.get_state()
        ctrl = (kp * (key[7:] - q0[7:]) - 0.05 * kp * v0[6:] + rng.randn(12) * 2).astype(np.float32).astype(np.float64)
        e = emu_step(m, q0, v0, w0, ctrl, 0.8, 0.8, [0.5, 0, 0, 0], precision=1, mode=1)
        o.step(ctrl)
        f = o.flags()
        assert e['ncon'] == f['ncon'] and e['invalid_mask'] == f['invalid_body_mask'], f'step {k}'
        assert e['contact_mask'] == sum(int(b) << i for i, b in enumerate(f['contact_state']))
        qo, vo, _, _ = o.get_state()
        worst = max(worst, np.abs(e['qpos'] - qo).max(), np.abs(e['qvel'] - vo).max())
        max_ncon = max(max_ncon, f['ncon'])
    assert max_ncon >= 2 and worst < 1e-8, (max_ncon, worst)


@pytest.mark.parametrize('robot,scene,xy', [('mini_cheetah', 'perlin', (3.0, 2.0)), ('hyqreal1', 'perlin', (-4.0, 6.0)), ('mini_cheetah', 'random_boxes', (2.0, -1.0)),
                                            ('hyqreal1', 'random_boxes', (1.0, 2.5)), ('spot', 'perlin', (5.0, -3.0))])
def test_mesh_links_collide_with_the_terrain(robot, scene, xy):
    """Convex-mesh links against the height field / static boxes (reference test matrix: mini_cheetah, hyqreal1, hyqreal2 x perlin,
    tests/env_test.py:14-16): the robot is dropped limp onto the terrain so that thighs, calves and the trunk end up lying on it;
    the kernel source (fp64, emulated warp) must produce the oracle's contact set -- including the mesh geoms -- at every step."""
    m = Model(robot, scene)
    rng = np.random.RandomState(11)
    key = np.array(m.c.key_qpos)
    q = key.copy()
    q[0:2] = np.array(xy) + rng.uniform(-0.3, 0.3, 2)
    q[2] = 1.2 if robot != 'hyqreal1' else 1.6
    o = Oracle(m)
    o.set_state(q, np.zeros(18), np.zeros(18)); assert o.lift() >= 0
    o.set_env(0.8, 0.8, [0, 0, 0, 0])
    mesh = np.array([m.c.geom_type[g] == 7 for g in range(m.c.ngeom)])
    assert mesh.any()
    mesh_terrain_contacts, worst, steps = 0, 0.0, 0
    for k in range(400):
        q0, v0, _, w0 = o.get_state()
        ctrl = np.zeros(12)
        o.step(ctrl)
        f = o.flags()
        if f['ncon'] == 0:
            continue  # free fall: nothing to compare yet
        e = emu_step(m, q0, v0, w0, ctrl, 0.8, 0.8, [0, 0, 0, 0], precision=1, mode=1)
        steps += 1
        if f['ncon'] > 16:
            assert e['overflow']
            continue
        assert e['ncon'] == f['ncon'] and e['invalid_mask'] == f['invalid_body_mask'], f'step {k}'
        assert e['contact_mask'] == sum(int(b) << i for i, b in enumerate(f['contact_state']))
        oc = o.get(F_CONTACTS) if False else None
        qo, vo, _, _ = o.get_state()
        worst = max(worst, np.abs(e['qpos'] - qo).max(), np.abs(e['qvel'] - vo).max())
        # contacts of the forward pass the step started from
        o2 = Oracle(m); o2.set_state(q0, v0, w0); o2.set_env(0.8, 0.8, [0, 0, 0, 0]); o2.forward(ctrl)
        oc = o2.get(F_CONTACTS); ec = e['contacts']
        key_ = lambda c: np.lexsort((np.round(c[:, 0], 9), c[:, 16]))
        oc, ec = oc[key_(oc)], ec[key_(ec)]
        assert (oc[:, 16:18] == ec[:, 16:18]).all()
        np.testing.assert_allclose(ec[:, 0:13], oc[:, 0:13], atol=1e-9)
        on_terrain = np.abs(oc[:, 3]) > 1e-3 if scene == 'random_boxes' else np.ones(len(oc), dtype=bool)  # contact point above the floor plane
        mesh_terrain_contacts += int((mesh[oc[:, 16].astype(int)] & on_terrain & (oc[:, 3] - 0.5 * oc[:, 0] > 0.01)).sum())
        if steps >= 100:
            break
    assert mesh_terrain_contacts >= 5 and worst < 1e-8, (mesh_terrain_contacts, worst)


@pytest.mark.parametrize('robot,scene,feat_bits', [('mini_cheetah', 'flat', 8101), ('aliengo', 'perlin', 1641), ('go2', 'random_boxes', 5746), ('hyqreal1', 'flat', 6022),
                                                   ('b2', 'flat', 5221), ('go1', 'flat', 4198), ('spot', 'flat', 6054)])
@pytest.mark.parametrize('precision', [0, 1])
def test_specialised_variants_are_bit_identical_to_generic(robot, scene, feat_bits, precision):
    """The four specialised step kernels (compile-time feature switches, csrc/qs_env.cuh FEAT_CFG2..5) against the generic
    variant on the host emulator: every number of a contact-rich rollout must be identical, in fp32 and fp64."""
    m = Model(robot, scene)
    rng = np.random.RandomState(5)
    key = np.array(m.c.key_qpos)
    q = key.copy()
    if scene != 'flat':
        q[0:2] = (3.0, 2.0) if scene == 'perlin' else (2.0, -1.0)
        q[2] = 0.95 if scene == 'perlin' else 0.45
    q[7:] += rng.uniform(-0.2, 0.2, 12)
    o = Oracle(m)
    o.set_state(q, np.zeros(18), np.zeros(18)); assert o.lift() >= 0
    q = o.get_state()[0]
    for _ in range(400):  # lower the robot until its feet press into the surface: contacts from the first step on
        o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
        if o.flags()['contact_state'].sum() >= 2:
            break
        q[2] -= 0.003
    q[2] -= 0.006
    q = q.astype(np.float32).astype(np.float64)
    qa, va, wa = q.copy(), np.zeros(18), np.zeros(18)
    qb, vb, wb = q.copy(), np.zeros(18), np.zeros(18)
    saw_contact = False
    for k in range(40):
        ctrl = (rng.randn(12) * 6).astype(np.float32).astype(np.float64)
        a = emu_step(m, qa, va, wa, ctrl, 0.8, 0.8, [0.5, 0, 0, 0.1], precision=precision, tol=1e-8 if precision else 1e-6, mode=1, specialised=True)
        b = emu_step(m, qb, vb, wb, ctrl, 0.8, 0.8, [0.5, 0, 0, 0.1], precision=precision, tol=1e-8 if precision else 1e-6, mode=1, specialised=False)
        assert a['feat'] == feat_bits and b['feat'] == 0
        for name in ('qpos', 'qvel', 'qacc', 'obs', 'fcon', 'M'):
            assert np.array_equal(a[name], b[name]), f'{name} differs at step {k}'
        assert a['iters'] == b['iters'] and a['ncon'] == b['ncon'] and a['contact_mask'] == b['contact_mask'] and a['invalid_mask'] == b['invalid_mask']
        saw_contact |= a['ncon'] > 0
        qa, va, wa = a['qpos'], a['qvel'], a['qacc']
        qb, vb, wb = b['qpos'], b['qvel'], b['qacc']
    assert saw_contact


def _capsule_segments(m, q):
    """World end points of every capsule geom from an independent numpy FK over the compiled tree (hinge angle = qpos - qpos0)."""
    c = m.c

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])

    def rot(qt, v):
        w, x, y, z = qt
        R = np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])
        return R @ np.asarray(v)

    pos = {1: q[:3]}; quat = {1: q[3:7] / np.linalg.norm(q[3:7])}
    for b in range(2, 14):
        pa = c.body_parent[b]
        pos[b] = pos[pa] + rot(quat[pa], c.body_pos[b])
        qb = qmul(quat[pa], np.array(c.body_quat[b]))
        ang = q[7 + b - 2] - c.qpos0[7 + b - 2]
        ax = np.array(c.jnt_axis[b - 2])
        quat[b] = qmul(qb, np.r_[np.cos(ang / 2), np.sin(ang / 2) * ax])
    out = {}
    for g in range(c.ngeom):
        if c.geom_type[g] != 3:
            continue
        b = c.geom_body[g]
        ctr = pos[b] + rot(quat[b], c.geom_pos[g])
        axis = rot(qmul(quat[b], np.array(c.geom_quat[g])), [0, 0, 1])
        out[g] = (ctr, axis, c.geom_size[g][1], c.geom_size[g][0])
    return out


def test_capsule_lying_across_a_box_edge_gets_a_mid_segment_contact():
    """A capsule link draped over a stair edge touches it between its two end spheres: the nearest point of the capsule axis to the
    box (bisection on the derivative of the squared segment-box distance) becomes a third sphere feature when it is strictly nearer
    than both ends.  Random poses of go1 on the stairs; the kernel source (fp64, emulated warp) must reproduce the oracle's
    contacts, and contacts whose foot point on the capsule axis lies well inside the segment must occur (checked with an independent
    numpy FK of the capsule end points)."""
    m = Model('go1', 'stairs')  # thin, long `fromto` capsules on thighs and calves (go1.xml:47-59)
    rng = np.random.RandomState(3)
    key = np.array(m.c.key_qpos)
    mids, checked = 0, 0
    for trial in range(600):
        q = key.copy()
        q[0] = rng.uniform(0.8, 3.0); q[1] = rng.uniform(-0.5, 0.5); q[2] = rng.uniform(0.25, 0.75)
        ang = rng.uniform(-0.6, 0.6, 3)
        cr, sr, cp, sp, cy, sy = np.cos(ang[0] / 2), np.sin(ang[0] / 2), np.cos(ang[1] / 2), np.sin(ang[1] / 2), np.cos(ang[2] / 2), np.sin(ang[2] / 2)
        q[3:7] = [cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy]
        q[7:] += rng.uniform(-0.6, 0.6, 12)
        q = q.astype(np.float32).astype(np.float64)
        o = Oracle(m)
        o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
        oc = o.get(F_CONTACTS)
        if len(oc) == 0 or len(oc) > 16:
            continue
        segs = _capsule_segments(m, q)
        n_mid = 0
        for row in oc:
            g = int(row[16])
            if g not in segs:
                continue
            ctr, axis, L, r = segs[g]
            t = float(np.dot(row[1:4] - ctr, axis))  # foot of the contact point on the capsule axis
            # an end-sphere contact projects to |t| >= L - r; anything well inside that can only be the mid-segment feature
            if abs(t) < L - r - 0.002 and row[3] > 0.02:
                n_mid += 1
        if n_mid == 0 and checked >= 12:
            continue  # enough poses without a mid contact have been compared already
        e = emu_step(m, q, np.zeros(18), np.zeros(18), np.zeros(12), -1.0, -1.0, [0, 0, 0, 0], precision=1, mode=0)
        assert e['ncon'] == len(oc), f'trial {trial}'
        ec = e['contacts']
        key_ = lambda cc: np.lexsort((np.round(cc[:, 2], 7), np.round(cc[:, 1], 7), np.round(cc[:, 0], 9), cc[:, 16]))
        oc2, ec2 = oc[key_(oc)], ec[key_(ec)]
        assert (oc2[:, 16:18] == ec2[:, 16:18]).all()
        np.testing.assert_allclose(ec2[:, 0:13], oc2[:, 0:13], atol=1e-7)
        checked += 1
        mids += n_mid
        if mids >= 6 and checked >= 20:
            break
    assert checked >= 12 and mids >= 3, (checked, mids)


def test_random_single_steps_all_robots_and_scenes():
    """Broad randomised comparison: 160 random (robot, scene, pose, velocity, ctrl, friction) cases over all eight robots and four
    scenes -- upright and tumbling orientations, bases from buried in the terrain to airborne -- one step each on the emulated
    kernel source (fp64) against the oracle: contact count, contact / invalid-contact body masks (also when the kernel's 16-slot
    contact buffer overflows: the masks are collected at detection), out-of-bounds flag, and the state to 1e-7.
    (2 300 further cases of the same generator, seeds 0-4, were run once without a mismatch; worst state difference 2.5e-8.)"""
    robots = ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1', 'hyqreal2', 'go1', 'b2', 'spot']
    scenes = ['flat', 'random_boxes', 'perlin', 'stairs']
    rng = np.random.RandomState(123)
    models = {}
    compared = overflowed = in_contact = 0
    for it in range(160):
        robot, scene = robots[rng.randint(len(robots))], scenes[rng.randint(len(scenes))]
        m = models.setdefault((robot, scene), Model(robot, scene))
        q = np.array(m.c.key_qpos)
        q[7:] += rng.uniform(-0.6, 0.6, 12)
        if scene != 'flat':
            q[0:2] = rng.uniform(-4, 4, 2)
        ax = rng.randn(3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(0, 0.5) if rng.rand() < 0.6 else rng.uniform(0, np.pi)
        q[3:7] = np.r_[np.cos(ang / 2), np.sin(ang / 2) * ax]
        q[2] = rng.uniform(0.02, 1.2 * m.hip_height + (0.7 if scene != 'flat' else 0.0))
        v = rng.uniform(-2, 2, 18)
        ctrl = rng.randn(12) * 30
        mu = rng.uniform(0.2, 1.5)
        o = Oracle(m)
        o.set_state(q, v, np.zeros(18)); o.set_env(mu, mu, [0.3, 0, 0, 0.1])
        o.step(ctrl)
        f = o.flags()
        qo, vo, _, _ = o.get_state()
        if not np.isfinite(qo).all():
            continue
        e = emu_step(m, q, v, np.zeros(18), ctrl, mu, mu, [0.3, 0, 0, 0.1], precision=1, mode=1)
        tag = f'case {it}: {robot} / {scene}'
        assert e['invalid_mask'] == f['invalid_body_mask'], tag
        assert e['contact_mask'] == sum(int(b) << i for i, b in enumerate(f['contact_state'])), tag
        assert bool(e['oob']) == bool(f['out_of_bounds']), tag
        if f['ncon'] > 16:
            assert e['overflow'], tag
            overflowed += 1
            continue
        assert e['ncon'] == f['ncon'], tag
        assert max(np.abs(e['qpos'] - qo).max(), np.abs(e['qvel'] - vo).max()) < 1e-7, tag
        compared += 1; in_contact += f['ncon'] > 0
    assert compared >= 100 and in_contact >= 40 and overflowed >= 20, (compared, in_contact, overflowed)
