"""Pins the fp64 oracle with analytic known-answers (the reference holds no golden vectors for this path and MuJoCo is
not installable here -- see oracle/qstep_oracle.c header: "parity unpinned")."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from gym_quadruped_b200.model import Model
from oracle.oracle import F_CONTACTS, F_EFC, F_FEET_JACP, F_FEET_POS, F_M, F_XPOS, Oracle

ROBOTS = ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1', 'hyqreal2', 'b2', 'go1', 'spot']
MASS = {'mini_cheetah': 12.473, 'aliengo': 24.638, 'go2': 15.206, 'hyqreal1': 107.573,
        'hyqreal2': 126.694, 'b2': 83.498, 'go1': 12.743, 'spot': 50.34}  # sum of the <inertial mass=...> attributes of hyqreal2.xml / b2.xml


def _airborne(model, rng, z=2.0):
    q = np.array(model.c.key_qpos)
    q[2] = z
    q[3:7] = Rotation.random(random_state=rng.randint(1 << 30)).as_quat()[[3, 0, 1, 2]]
    q[7:] += rng.uniform(-0.3, 0.3, 12)
    return q


@pytest.mark.parametrize('robot', ROBOTS)
def test_free_fall_acceleration(robot):
    m = Model(robot, 'flat')
    o = Oracle(m)
    rng = np.random.RandomState(0)
    o.set_state(_airborne(m, rng), np.zeros(18), np.zeros(18))
    o.forward(np.zeros(12))
    qacc = o.get_state()[2]
    np.testing.assert_allclose(qacc[:3], [0, 0, -9.81], atol=1e-9)
    np.testing.assert_allclose(qacc[3:], 0, atol=1e-8)  # every link falls alike: no relative acceleration
    assert o.flags()['ncon'] == 0


@pytest.mark.parametrize('robot', ROBOTS)
def test_mass_matrix_against_independent_formulation(robot):
    """Oracle CRB mass matrix at the compile pose vs the compiler's body-Jacobian sum (numpy, written independently)."""
    m = Model(robot, 'flat')
    o = Oracle(m)
    q = np.array(m.tables['qpos0_compile'])
    q[7:] += np.array(m.c.qpos0)[7:]  # hinge angle = qpos - qpos0 (mini_cheetah's override shifts the zero)
    o.set_state(q, np.zeros(18), np.zeros(18))
    o.forward(np.zeros(12))
    M = o.get(F_M)
    assert np.abs(M - M.T).max() == 0 and np.linalg.eigvalsh(M).min() > 0
    np.testing.assert_allclose(np.diag(M), m.tables['M0_diag'], rtol=1e-10)
    np.testing.assert_allclose(M[0, 0], MASS[robot], atol=2e-3)
    np.testing.assert_allclose(np.trace(M) / 18, m.c.meaninertia, rtol=1e-10)


@pytest.mark.parametrize('robot', ['mini_cheetah', 'hyqreal1'])
def test_kinetic_energy_matches_finite_difference_of_body_motion(robot):
    m = Model(robot, 'flat')
    rng = np.random.RandomState(3)
    q, v = _airborne(m, rng), rng.uniform(-1, 1, 18)
    o = Oracle(m)
    o.set_state(q, v, np.zeros(18)); o.forward(np.zeros(12))
    ke = 0.5 * v @ o.get(F_M) @ v
    # body COM velocities by central differences of the kinematics along the flow of v
    eps = 1e-6

    def poses(sign):
        qq = q.copy()
        qq[:3] += sign * eps * v[:3]
        dq = Rotation.from_rotvec(sign * eps * v[3:6])
        qq[3:7] = (Rotation.from_quat(q[[4, 5, 6, 3]]) * dq).as_quat()[[3, 0, 1, 2]]
        qq[7:] += sign * eps * v[6:]
        oo = Oracle(m)
        oo.set_state(qq, np.zeros(18), np.zeros(18)); oo.forward(np.zeros(12))
        return oo
    op, om = poses(+1), poses(-1)
    import ctypes
    from oracle.oracle import lib  # noqa: F401
    ke_fd = 0.0
    xp, xm = op.get(F_XPOS), om.get(F_XPOS)
    # rotation of each body from three points is overkill: use the feet Jacobian identity instead for the linear part and
    # compare total linear momentum, which is exact: p = sum m_b v_com_b = (M v)[0:3]
    mom = (o.get(F_M) @ v)[:3]
    masses = np.array(m.c.body_mass)[1:]
    ipos = np.array(m.c.body_ipos)[1:]
    # body COM = xpos + R * ipos; R from finite differences is unavailable through the C API, so restrict to bodies with
    # |ipos| small relative error: use all bodies but bound the error by |omega| * |ipos| * mass
    vcom = (xp - xm) / (2 * eps)
    approx = (masses[:, None] * vcom).sum(0)
    slack = (masses * np.linalg.norm(ipos, axis=1)).sum() * 8.0
    assert np.abs(mom - approx).max() < slack
    assert ke > 0
    # exact check on the base alone: with joints frozen and no rotation, KE = 1/2 m |v|^2
    v2 = np.zeros(18); v2[:3] = [0.3, -0.2, 0.5]
    np.testing.assert_allclose(0.5 * v2 @ o.get(F_M) @ v2, 0.5 * MASS[robot] * (v2[:3] @ v2[:3]), rtol=2e-4)


@pytest.mark.parametrize('robot', ['mini_cheetah', 'go2'])
def test_momentum_conserved_without_gravity_and_contacts(robot):
    """Internal torques, damping and joint friction cannot change the total linear momentum (M v)[0:3]; the semi-implicit Euler
    scheme conserves it up to O(h), so the drift must shrink ~linearly with the time step."""
    drift = {}
    for dt in (2e-3, 5e-4):
        m = Model(robot, 'flat', sim_dt=dt)
        m.c.gravity[2] = 0.0
        o = Oracle(m)
        rng = np.random.RandomState(1)
        q = _airborne(m, rng, z=5.0)
        v = np.zeros(18); v[:3] = [0.2, -0.1, 0.05]; v[6:] = rng.uniform(-1, 1, 12)
        o.set_state(q, v, np.zeros(18))
        o.forward(np.zeros(12))
        p0 = (o.get(F_M) @ v)[:3]
        ctrl = rng.uniform(-3, 3, 12)
        for _ in range(int(round(0.1 / dt))):
            o.step(ctrl)
        o.forward(np.zeros(12))
        p1 = (o.get(F_M) @ o.get_state()[1])[:3]
        drift[dt] = np.abs(p1 - p0).max() / np.abs(p0).max()
    assert drift[5e-4] < 2e-3 and drift[5e-4] < 0.45 * drift[2e-3], drift


@pytest.mark.parametrize('robot', ROBOTS)
def test_standing_reaction_force_equals_weight(robot):
    """PD-hold the home pose on flat ground; at rest the normal forces must add up to m*g."""
    m = Model(robot, 'flat')
    o = Oracle(m)
    key = np.array(m.c.key_qpos)
    o.set_state(key, np.zeros(18), np.zeros(18))
    assert o.lift() >= 0
    kp, kd = {'hyqreal1': (400.0, 20.0), 'hyqreal2': (3000.0, 60.0), 'b2': (1500.0, 40.0), 'spot': (800.0, 30.0)}.get(robot, (60.0, 3.0))  # heavy robots: stiffer hold
    for _ in range(5000):
        q, v, _, _ = o.get_state()
        o.step(kp * (key[7:] - q[7:]) - kd * v[6:])
    q, v, _, _ = o.get_state()
    assert np.abs(v).max() < 5e-3, 'robot did not come to rest'
    f = o.flags()
    assert f['contact_state'].all() and not f['invalid_contact']
    c = o.get(F_CONTACTS)
    np.testing.assert_allclose(c[:, 13].sum(), MASS[robot] * 9.81, rtol=2e-3)
    # complementarity / friction cone
    assert (c[:, 13] >= -1e-9).all()
    mu = c[:, 18]
    if m.c.cone == 0:
        assert (np.abs(c[:, 14]) + np.abs(c[:, 15]) <= mu * c[:, 13] + 1e-7).all()
    else:
        assert (np.hypot(c[:, 14], c[:, 15]) <= mu * c[:, 13] * 1.05 + 1e-6).all()


def test_feet_jacobian_matches_finite_differences():
    m = Model('aliengo', 'flat')
    rng = np.random.RandomState(5)
    q = _airborne(m, rng)
    o = Oracle(m)
    o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    J = o.get(F_FEET_JACP)
    eps = 1e-7
    for d in range(18):
        v = np.zeros(18); v[d] = 1.0
        qq = q.copy()
        qq[:3] += eps * v[:3]
        qq[3:7] = (Rotation.from_quat(q[[4, 5, 6, 3]]) * Rotation.from_rotvec(eps * v[3:6])).as_quat()[[3, 0, 1, 2]]
        qq[7:] += eps * v[6:]
        o2 = Oracle(m)
        o2.set_state(qq, np.zeros(18), np.zeros(18)); o2.forward(np.zeros(12))
        fd = (o2.get(F_FEET_POS) - o.get(F_FEET_POS)) / eps
        np.testing.assert_allclose(J[:, :, d], fd, atol=2e-6)


@pytest.mark.parametrize('robot', ['aliengo', 'mini_cheetah'])
def test_feet_jacobian_dot_matches_finite_difference_along_the_flow(robot):
    """mj_jacDot users (quadruped_env.py:742-797): Jdot(q, v) = d/dt J(q(t)) with q advanced along qvel (free-joint rotation in
    the body frame), for the translational and the rotational foot Jacobians."""
    from oracle.oracle import F_FEET_JACP_DOT, F_FEET_JACR, F_FEET_JACR_DOT
    m = Model(robot, 'flat')
    rng = np.random.RandomState(7)
    q = _airborne(m, rng)
    v = rng.uniform(-1.5, 1.5, 18)
    o = Oracle(m)
    o.set_state(q, v, np.zeros(18)); o.forward(np.zeros(12))
    Jp_dot, Jr_dot = o.get(F_FEET_JACP_DOT), o.get(F_FEET_JACR_DOT)
    J = {}
    eps = 1e-6
    for sgn in (-1, 1):
        qq = q.copy()
        qq[:3] += sgn * eps * v[:3]
        qq[3:7] = (Rotation.from_quat(q[[4, 5, 6, 3]]) * Rotation.from_rotvec(sgn * eps * v[3:6])).as_quat()[[3, 0, 1, 2]]
        qq[7:] += sgn * eps * v[6:]
        o2 = Oracle(m)
        o2.set_state(qq, v, np.zeros(18)); o2.forward(np.zeros(12))
        J[sgn] = (o2.get(F_FEET_JACP), o2.get(F_FEET_JACR))
    np.testing.assert_allclose(Jp_dot, (J[1][0] - J[-1][0]) / (2 * eps), atol=5e-6)
    np.testing.assert_allclose(Jr_dot, (J[1][1] - J[-1][1]) / (2 * eps), atol=5e-6)
    assert np.abs(Jp_dot).max() > 0.1 and np.abs(Jr_dot).max() > 0.1


def test_joint_limit_pushes_back():
    m = Model('aliengo', 'flat')
    o = Oracle(m)
    q = np.array(m.c.key_qpos); q[2] = 2.0
    q[7] = 1.22173 + 0.05  # FL hip beyond its upper limit (aliengo.xml:57)
    o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    efc = o.get(F_EFC)
    lim = efc[efc[:, 0] == 1]
    assert len(lim) == 1 and lim[0, 4] > 0
    assert o.get_state()[2][6] < -10.0  # accelerated back into the range


def test_observation_conventions_match_scipy():
    """Euler / SO3 / gravity-vector conventions of quadruped_env.py:962-1016."""
    m = Model('mini_cheetah', 'flat')
    rng = np.random.RandomState(7)
    for _ in range(5):
        o = Oracle(m)
        q = _airborne(m, rng); v = rng.uniform(-1, 1, 18)
        o.set_state(q, v, np.zeros(18)); o.set_env(-1, -1, [0.7, 0.0, 0.0, 0.3])
        obs, _ = o.step(rng.randn(12))
        qn, vn, qa, _ = o.get_state()
        R = Rotation.from_quat(qn[[4, 5, 6, 3]])
        np.testing.assert_allclose(obs[18:21], R.as_euler('xyz'), atol=1e-9)
        np.testing.assert_allclose(obs[25:34], R.as_matrix().ravel(), atol=1e-9)
        np.testing.assert_allclose(obs[34:37], R.as_matrix().T @ [0, 0, -1], atol=1e-9)
        Rh = Rotation.from_euler('xyz', R.as_euler('xyz') * [0, 0, 1]).as_matrix()
        np.testing.assert_allclose(obs[6:9], Rh @ [0.7, 0, 0] - vn[:3], atol=1e-9)
        np.testing.assert_allclose(obs[12:15], R.as_matrix() @ vn[3:6], atol=1e-9)
        np.testing.assert_allclose(obs[49:52], R.as_matrix().T @ [0, 0, 0.3] - vn[3:6], atol=1e-9)
        np.testing.assert_allclose(obs[52:71], qn, atol=0); np.testing.assert_allclose(obs[9:12], qa[:3], atol=0)


def test_fk_known_answers_at_home_keyframes():
    """SURVEY.md App. D.2 foot-sphere centres at the `home` keyframe."""
    expect = {'aliengo': ((0.2399, 0.134, 0.0392), (-0.2399, 0.134, 0.0392)),
              'go2': ((0.1922, 0.142, 0.0036), (-0.1946, 0.142, 0.0036)),
              'hyqreal1': ((0.4592, 0.256, 0.040), (-0.4278, 0.256, 0.040)),
              'mini_cheetah': ((0.1789, 0.111, -0.040), (-0.214, 0.111, -0.0392))}
    for robot, (fl, rl) in expect.items():
        m = Model(robot, 'flat')
        o = Oracle(m)
        o.set_state(np.array(m.c.key_qpos), np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
        fp = o.get(F_FEET_POS)
        np.testing.assert_allclose(fp[0], fl, atol=6e-4)
        np.testing.assert_allclose(fp[2], rl, atol=6e-4)
        np.testing.assert_allclose(fp[1], [fl[0], -fl[1], fl[2]], atol=6e-4)  # FR mirrors FL


@pytest.mark.parametrize('roll,expect', [(0.0, 'side'), (np.pi / 2, 'cap')])
def test_plane_cylinder_known_answers(roll, expect):
    """Geometry of the plane-cylinder contacts (b2 hip cylinders, b2.xml:96: radius 0.07, half-length 0.025, axis along the hip's y):
    lying on its side a cylinder touches along a line -> two contacts at the cap centres' projections, depth z_c - r; standing on a
    cap -> three rim contacts 120 degrees apart, depth z_c - h."""
    from oracle.oracle import F_XPOS
    m = Model('b2', 'flat')
    cyl = [g for g in range(m.c.ngeom) if m.c.geom_type[g] == 5]
    g0 = cyl[0]
    r, h = m.c.geom_size[g0][0], m.c.geom_size[g0][1]
    gpos = np.array(m.c.geom_pos[g0]); body = m.c.geom_body[g0]
    q = np.array(m.c.key_qpos); q[7:] = 0.0
    q[3:7] = [np.cos(roll / 2), np.sin(roll / 2), 0, 0]
    Rb = Rotation.from_quat(q[[4, 5, 6, 3]]).as_matrix()
    o = Oracle(m)
    # height at which the chosen cylinder penetrates the floor by 2 mm while (by symmetry) only hip cylinders can be lower
    q[2] = 5.0
    o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    centre = o.get(F_XPOS)[body - 1] + Rb @ gpos
    low = r if expect == 'side' else h
    q[2] -= centre[2] - low + 0.002
    o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    centre = o.get(F_XPOS)[body - 1] + Rb @ gpos
    c = o.get(F_CONTACTS)
    mine = c[c[:, 16] == g0]
    if expect == 'side':
        assert len(mine) == 2
        np.testing.assert_allclose(mine[:, 0], centre[2] - r, atol=1e-12)
        axis = Rb @ np.array([0.0, 1.0, 0.0])  # geom quat (1,1,0,0)/sqrt2 turns the cylinder's z onto the body's -y / +y line
        along = np.sort((mine[:, 1:4] - centre) @ axis)
        np.testing.assert_allclose(along, [-h, h], atol=1e-12)
    else:
        assert len(mine) == 3
        np.testing.assert_allclose(mine[:, 0], centre[2] - h, atol=1e-12)
        rel = mine[:, 1:3] - centre[:2]
        np.testing.assert_allclose(np.linalg.norm(rel, axis=1), r, atol=1e-12)
        ang = np.sort(np.mod(np.arctan2(rel[:, 1], rel[:, 0]), 2 * np.pi))
        np.testing.assert_allclose(np.diff(ang), [2 * np.pi / 3] * 2, atol=1e-9)
    np.testing.assert_allclose(mine[:, 4:7], [[0, 0, 1]] * len(mine), atol=1e-15)  # contact normal = plane normal


def _sliding_deceleration(x0, v0):
    m = Model('aliengo', 'slippery')
    o = Oracle(m)
    key = np.array(m.c.key_qpos)
    q = key.copy(); q[0] = x0; q[1] = 0.0
    for z in np.arange(0.6, 0.2, -0.001):  # lower the robot until its feet touch the strip
        q[2] = z
        o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
        if o.flags()['contact_state'].all():
            break
    o.set_env(0.5, 0.5, [0, 0, 0, 0])  # friction override of feet / floor: must not matter on the strip
    kp, kd = 400.0, 20.0
    for _ in range(600):  # come to rest under a PD hold
        qq, v, _, _ = o.get_state()
        o.step(kp * (key[7:] - qq[7:]) - kd * v[6:])
    qq, v, _, w = o.get_state()
    v = v.copy(); v[0] = v0
    o.set_state(qq, v, w)
    vs = []
    for _ in range(60):
        qq, v, _, _ = o.get_state()
        o.step(kp * (key[7:] - qq[7:]) - kd * v[6:])
        vs.append(o.get_state()[1][0])
    vs = np.array(vs)
    moving = vs > 0.05
    assert moving.sum() >= 15
    t = 0.002 * np.arange(1, len(vs) + 1)
    return -np.polyfit(t[moving], vs[moving], 1)[0]


def test_sliding_on_the_slippery_strips_follows_each_strip_s_friction():
    """scene_slippery.xml:39-40: a robot held by a PD controller and pushed along a priority-2 strip is decelerated by the strip's own
    friction coefficient, whatever the feet's / floor's coefficient is set to: mu * g on the 0.03 strip (rigid sliding), and more than
    ten times that on the 0.8 strip (where the legs give way before the feet slide, so only the order of magnitude is asserted)."""
    slow = _sliding_deceleration(12.0, 0.3)
    fast = _sliding_deceleration(2.0, 0.4)
    assert slow == pytest.approx(0.03 * 9.81, rel=0.2), slow
    assert fast > 10 * slow, (fast, slow)


def _fk_numpy(t, q):
    """Independent forward kinematics from the compiled tables: world rotation / position of bodies 1..13."""
    Rw = {0: np.eye(3)}; pw = {0: np.zeros(3)}
    for b in range(1, 14):
        if b == 1:
            Rw[1] = Rotation.from_quat(np.array(q[3:7])[[1, 2, 3, 0]]).as_matrix(); pw[1] = np.array(q[:3])
            continue
        p = t['body_parent'][b]
        R0 = Rw[p] @ Rotation.from_quat(np.array(t['body_quat'][b])[[1, 2, 3, 0]]).as_matrix()
        j = b - 2
        ang = q[7 + j] - t['qpos0'][7 + j]
        Rw[b] = R0 @ Rotation.from_rotvec(ang * np.array(t['jnt_axis'][j])).as_matrix()
        pw[b] = pw[p] + Rw[p] @ np.array(t['body_pos'][b])
    return Rw, pw


def test_plane_primitive_colliders_against_independent_geometry():
    """Every floor contact of aliengo (spheres, capsules, boxes) in a tilted, half-sunk pose must be one of the analytic candidates --
    sphere: z_c - r; capsule: end-sphere centres z - r; box: corner heights -- computed here with a separate numpy FK, and every
    candidate within the margin must appear (boxes: at most four per geom)."""
    m = Model('aliengo', 'flat')
    t = m.tables
    q = np.array(m.c.key_qpos); q[2] = 0.16
    q[3:7] = Rotation.from_euler('xyz', [0.5, 0.3, 0.7]).as_quat()[[3, 0, 1, 2]]
    q[7:] += np.random.RandomState(1).uniform(-0.3, 0.3, 12)
    o = Oracle(m)
    o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    c = o.get(F_CONTACTS)
    Rw, pw = _fk_numpy(t, q)
    expected = {}
    for g, geom in enumerate(t['geoms']):
        b = geom['body']
        Rg = Rw[b] @ Rotation.from_quat(np.array(geom['quat'])[[1, 2, 3, 0]]).as_matrix()
        cg = pw[b] + Rw[b] @ np.array(geom['pos'])
        sz, margin = geom['size'], geom['margin']
        if geom['type'] == 2:
            cand = [cg[2] - sz[0]]
        elif geom['type'] == 3:
            cand = [(cg + s * Rg[:, 2] * sz[1])[2] - sz[0] for s in (1, -1)]
        elif geom['type'] == 6:
            cand = [(cg + Rg @ (np.array(sz) * np.array([sx, sy, sz_])))[2] for sx in (-1, 1) for sy in (-1, 1) for sz_ in (-1, 1)]
            cand = [d for d in cand if d - cg[2] <= 0]   # only corners below the box centre are tested by the engine's routine
        else:
            continue
        expected[g] = sorted(d for d in cand if d <= margin)
    assert len(c) >= 6
    for g in set(c[:, 16].astype(int)) | {g for g, e in expected.items() if e}:
        got = np.sort(c[c[:, 16] == g][:, 0])
        exp = np.array(expected[g])
        if t['geoms'][g]['type'] == 6:
            assert len(got) == min(4, len(exp)) and all(np.abs(exp - d).min() < 1e-12 for d in got), g
        else:
            np.testing.assert_allclose(got, exp, atol=1e-12, err_msg=f'geom {g}')
