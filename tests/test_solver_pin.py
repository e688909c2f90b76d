"""Independent pins of the oracle (CPU): the Newton solver's optimum, and compiled model constants.

The soft-constraint problem of the engine is strictly convex, so its optimum is unique and any correct optimiser must find it:

* primal:  qacc* = argmin_a  1/2 (a - a0)' M (a - a0) + sum_i s_i(J_i a - aref_i)          (a0 = qacc_smooth)
  -- solved here by scipy L-BFGS on an independently written cost (this file's `_cost`), for every cone / condim the robots use;
* dual:    f* = argmin_{f in K}  1/2 f' (J M^-1 J' + R) f + f' (J a0 - aref),  qacc* = a0 + M^-1 J' f*
  -- the engine's DEFINING formulation (forces in the friction cone); solved by an active-set bounded least-squares method
  (scipy BVLS) for pyramidal cones / limits / friction loss, and by accelerated projected gradient onto the second-order cones
  for elliptic cones.  Agreement of the dual optimum with the oracle's primal Newton solution pins both the solver and the
  primal cone cost (the R scaling with impratio, the regularised middle zone).

Oracle = oracle/qstep_oracle.c:solve().  Nothing here needs MuJoCo; "parity unpinned" with respect to the engine itself remains.
"""
import numpy as np
import pytest
import scipy.linalg
import scipy.optimize

from gym_quadruped_b200.model import Model
from oracle.oracle import F_EFC_FULL, F_M, F_QACC_SMOOTH, Oracle

T_FRICTION, T_LIMIT, T_FRICTIONLESS, T_PYRAMIDAL, T_ELLIPTIC = range(5)


def _problem(robot, seed, sink=0.006, vel=0.4, torque=12.0, mu=0.7, tilt=0.04):
    """A contact-rich forward pass: feet pressed into the floor, joints near limits, random velocities / torques."""
    m = Model(robot, 'flat')
    rng = np.random.RandomState(seed)
    o = Oracle(m)
    q = np.array(m.c.key_qpos)
    q[7:] += rng.uniform(-0.12, 0.12, 12)
    r, p = rng.uniform(-tilt, tilt, 2)
    q[3:7] = [np.cos(r / 2) * np.cos(p / 2), np.sin(r / 2) * np.cos(p / 2), np.cos(r / 2) * np.sin(p / 2), -np.sin(r / 2) * np.sin(p / 2)]
    o.set_state(q, np.zeros(18), np.zeros(18)); assert o.lift() >= 0
    q = o.get_state()[0]
    for _ in range(60):  # lower the robot until at least three feet touch, then press them in
        o.set_state(q, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
        if o.flags()['contact_state'].sum() >= 3:
            break
        q[2] -= 0.002
    q[2] -= sink
    v = rng.uniform(-vel, vel, 18)
    o.set_state(q, v, np.zeros(18)); o.set_env(mu, mu, [0, 0, 0, 0])
    o.forward(rng.randn(12) * torque)
    efc = o.get(F_EFC_FULL)
    return dict(M=o.get(F_M), a0=o.get(F_QACC_SMOOTH), qacc=o.get_state()[2], efc=efc, J=efc[:, 16:34], cone=m.c.cone, flags=o.flags())


def _units(efc):
    """Group rows into units: scalar rows, or the `dim` rows of one elliptic contact."""
    units, r = [], 0
    while r < len(efc):
        n = int(efc[r, 14]) if int(efc[r, 0]) == T_ELLIPTIC else 1
        units.append((r, n)); r += n
    return units


def _cost(jar, efc):
    """s(jar) and its gradient: independent restatement of the engine's constraint cost (friction loss: Huber; limits, frictionless
    and pyramidal rows: one-sided quadratic; elliptic contacts: zero in the dual cone's polar... i.e. top zone, quadratic in the
    bottom zone, regularised distance to the cone in the middle zone)."""
    cost, g = 0.0, np.zeros_like(jar)
    for r, n in _units(efc):
        t, D, R = int(efc[r, 0]), efc[r, 2], efc[r, 3]
        x = jar[r]
        if t == T_FRICTION:
            f = efc[r, 7]
            if x <= -R * f: cost += -f * (0.5 * R * f + x); g[r] = -f
            elif x >= R * f: cost += -f * (0.5 * R * f - x); g[r] = f
            else: cost += 0.5 * D * x * x; g[r] = D * x
        elif t != T_ELLIPTIC:
            if x < 0: cost += 0.5 * D * x * x; g[r] = D * x
        else:
            mu, fri = efc[r, 8], efc[r, 9:14]
            Dk = efc[r:r + n, 2]
            xs = jar[r:r + n]
            N = xs[0] * mu
            U = xs[1:] * fri[:n - 1]
            T = np.linalg.norm(U)
            if N >= mu * T or (T <= 0 and N >= 0):
                pass
            elif mu * N + T <= 0 or (T <= 0 and N < 0):
                cost += 0.5 * float(Dk @ (xs * xs)); g[r:r + n] = Dk * xs
            else:
                Dm = D / max(1e-15, mu * mu * (1 + mu * mu))
                d = N - mu * T
                cost += 0.5 * Dm * d * d
                g[r] = Dm * d * mu
                g[r + 1:r + n] = -Dm * d * mu * U / T * fri[:n - 1]
    return cost, g


def _primal_scipy(P):
    M, a0, J, efc = P['M'], P['a0'], P['J'], P['efc']
    aref = efc[:, 4]
    L = np.linalg.cholesky(M)

    def f(y):  # a = a0 + L^-T y  ->  Gauss term = 1/2 |y|^2 (preconditioned)
        a = a0 + scipy.linalg.solve_triangular(L.T, y, lower=False)
        c, g = _cost(J @ a - aref, efc)
        ga = J.T @ g
        return 0.5 * y @ y + c, y + scipy.linalg.solve_triangular(L, ga, lower=True)

    y = np.zeros(18)
    for _ in range(6):  # restarts: L-BFGS history is rebuilt at the kinks of the piecewise-quadratic cost
        res = scipy.optimize.minimize(f, y, jac=True, method='L-BFGS-B', options=dict(maxiter=5000, ftol=1e-18, gtol=1e-13, maxcor=40))
        y = res.x
    return a0 + scipy.linalg.solve_triangular(L.T, y, lower=False), np.linalg.norm(f(y)[1])


def _dual_setup(P):
    M, a0, J, efc = P['M'], P['a0'], P['J'], P['efc']
    Minv = np.linalg.inv(M)
    Q = J @ Minv @ J.T + np.diag(efc[:, 3])
    b = J @ a0 - efc[:, 4]
    return Minv, Q, b


def _dual_bvls(P):
    """Box-constrained dual (no elliptic rows): friction loss |f| <= floss, everything else f >= 0."""
    efc = P['efc']
    Minv, Q, b = _dual_setup(P)
    lo = np.where(efc[:, 0] == T_FRICTION, -efc[:, 7], 0.0)
    hi = np.where(efc[:, 0] == T_FRICTION, efc[:, 7], np.inf)
    C = np.linalg.cholesky(Q).T  # Q = C' C
    rhs = -scipy.linalg.solve_triangular(C.T, b, lower=True)
    res = scipy.optimize.lsq_linear(C, rhs, bounds=(lo, hi), method='bvls', tol=1e-15, max_iter=2000)
    return P['a0'] + Minv @ P['J'].T @ res.x, res.x


def _dual_apg_elliptic(P, iters=400000):
    """Dual with elliptic cones by accelerated projected gradient in variables where every cone is a standard second-order cone:
    f_k = mu_k g_k for the tangential / torsional / rolling components."""
    efc = P['efc']
    Minv, Q, b = _dual_setup(P)
    n = len(efc)
    S = np.ones(n)  # f = S * g
    for r, k in _units(efc):
        if int(efc[r, 0]) == T_ELLIPTIC:
            S[r + 1:r + k] = efc[r, 9:9 + k - 1]
    Qs = (S[:, None] * Q) * S[None, :]
    for r, k in _units(efc):  # one positive scale per unit keeps boxes boxes and cones cones, and evens out the diagonal
        S[r:r + k] /= np.sqrt(np.diag(Qs)[r:r + k].mean())
    Qs = (S[:, None] * Q) * S[None, :]
    bs = S * b
    # Jacobi preconditioning keeps the cones (uniform scaling per contact would be needed to keep SOC shape -> scale per unit)
    Lmax = np.linalg.eigvalsh(Qs)[-1]
    units = _units(efc)

    def proj(g):
        out = g.copy()
        for r, k in units:
            t = int(efc[r, 0])
            if t == T_FRICTION:
                out[r] = np.clip(g[r], -efc[r, 7] / S[r], efc[r, 7] / S[r])
            elif t != T_ELLIPTIC:
                out[r] = max(g[r], 0.0)
            else:
                s, v = g[r], g[r + 1:r + k]
                nv = np.linalg.norm(v)
                if nv <= s: pass
                elif nv <= -s: out[r:r + k] = 0
                else:
                    a = 0.5 * (s + nv)
                    out[r] = a; out[r + 1:r + k] = a * v / nv
        return out

    g = np.zeros(n); y = g.copy(); tk = 1.0
    for it in range(iters):
        gn = proj(y - (Qs @ y + bs) / Lmax)
        tn = 0.5 * (1 + np.sqrt(1 + 4 * tk * tk))
        y = gn + (tk - 1) / tn * (gn - g)
        if (gn - g) @ (y - gn) > 0:  # adaptive restart
            y = gn.copy(); tn = 1.0
        if it % 500 == 0 and np.linalg.norm(gn - g) < 1e-14 * max(1.0, np.linalg.norm(gn)):
            g = gn
            break
        g, tk = gn, tn
    f = S * g
    return P['a0'] + Minv @ P['J'].T @ f, f


@pytest.mark.parametrize('robot', ['mini_cheetah', 'aliengo', 'hyqreal2', 'b2'])
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_pyramidal_optimum_primal_and_dual(robot, seed):
    """Pyramidal cones (condim 3 -> 4 rows per contact), joint limits, friction loss."""
    P = _problem(robot, seed)
    efc = P['efc']
    assert (efc[:, 0] == T_PYRAMIDAL).sum() >= 12 and (efc[:, 0] == T_FRICTION).sum() >= (12 if robot != 'b2' else 0), efc[:, 0]
    scale = max(1.0, np.abs(P['qacc']).max())
    a_primal, gnorm = _primal_scipy(P)
    assert np.abs(a_primal - P['qacc']).max() < 1e-5 * scale, (np.abs(a_primal - P['qacc']).max(), gnorm)
    a_dual, f = _dual_bvls(P)
    assert np.abs(a_dual - P['qacc']).max() < 1e-6 * scale, np.abs(a_dual - P['qacc']).max()
    # the dual forces are the oracle's efc_force
    assert np.abs(f - efc[:, 5]).max() < 1e-5 * max(1.0, np.abs(efc[:, 5]).max())


@pytest.mark.parametrize('robot', ['go2', 'hyqreal1', 'go1', 'spot'])
@pytest.mark.parametrize('seed', [0, 1])
def test_elliptic_optimum_primal_and_dual(robot, seed):
    """Elliptic cones with impratio 100: condim 3 (hyqreal1, go2 body geoms), condim 6 feet (go2, go1, spot), condim 1 (go1)."""
    P = _problem(robot, seed, sink=0.004)
    efc = P['efc']
    dims = {int(d) for t, d in zip(efc[:, 0], efc[:, 14]) if int(t) == T_ELLIPTIC}
    assert dims, 'no elliptic contact in this pose'
    if robot in ('go2', 'go1', 'spot'):
        assert 6 in dims
    scale = max(1.0, np.abs(P['qacc']).max())
    a_primal, gnorm = _primal_scipy(P)
    assert np.abs(a_primal - P['qacc']).max() < 2e-5 * scale, (np.abs(a_primal - P['qacc']).max(), gnorm)
    a_dual, f = _dual_apg_elliptic(P)
    assert np.abs(a_dual - P['qacc']).max() < 1e-7 * scale, np.abs(a_dual - P['qacc']).max()
    fscale = max(1.0, np.abs(efc[:, 5]).max())
    assert np.abs(f - efc[:, 5]).max() < 1e-6 * fscale, np.abs(f - efc[:, 5]).max()


@pytest.mark.parametrize('robot', ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1', 'hyqreal2', 'b2', 'go1', 'spot'])
def test_invweight0_against_dense_inverse(robot):
    """Compiled tables vs an independent computation (oracle and kernel share the compiled tables).
    dof_invweight0 / body_invweight0 [MJ: set0] at qpos0: hinge dofs get diag(M^-1); the 6 free-joint dofs get the mean of their
    translational / rotational diagonal entries; bodies get the mean translational / rotational diagonal of J M^-1 J' at the body
    centre of mass.  M is taken from the oracle's CRB pass at qpos0 and inverted densely here; the body Jacobians are rebuilt in
    this test from an independent numpy FK over the compiled kinematic tree."""
    m = Model(robot, 'flat')
    c = m.c
    o = Oracle(m)
    q0 = np.array(c.qpos0)
    o.set_state(q0, np.zeros(18), np.zeros(18)); o.forward(np.zeros(12))
    Minv = np.linalg.inv(o.get(F_M))
    d = np.diag(Minv)
    expect = np.concatenate([np.full(3, d[0:3].mean()), np.full(3, d[3:6].mean()), d[6:]])
    np.testing.assert_allclose(np.array(c.dof_invweight0), expect, rtol=1e-6)

    def quat2mat(q):
        w, x, y, z = q
        return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])

    # FK at qpos0 (hinge angle = qpos - qpos0 = 0): world pose of every body, joint axes / anchors
    pos = {0: np.zeros(3), 1: q0[:3]}; quat = {0: np.array([1.0, 0, 0, 0]), 1: q0[3:7] / np.linalg.norm(q0[3:7])}
    for b in range(2, 14):
        p = c.body_parent[b]
        pos[b] = pos[p] + quat2mat(quat[p]) @ np.array(c.body_pos[b])
        quat[b] = qmul(quat[p], np.array(c.body_quat[b]))
    for b in range(1, 14):
        R = quat2mat(quat[b])
        com = pos[b] + R @ np.array(c.body_ipos[b])
        Jp, Jr = np.zeros((3, 18)), np.zeros((3, 18))
        Rb = quat2mat(quat[1])
        Jp[:, 0:3] = np.eye(3)
        for k in range(3):
            ax = Rb[:, k]
            Jr[:, 3 + k] = ax; Jp[:, 3 + k] = np.cross(ax, com - pos[1])
        a = b
        while a >= 2:
            j = a - 2
            ax = quat2mat(quat[a]) @ np.array(c.jnt_axis[j])
            Jr[:, 6 + j] = ax; Jp[:, 6 + j] = np.cross(ax, com - pos[a])
            a = c.body_parent[a]
        tran = np.trace(Jp @ Minv @ Jp.T) / 3
        rot = np.trace(Jr @ Minv @ Jr.T) / 3
        np.testing.assert_allclose(np.array(c.body_invweight0[b]), [tran, rot], rtol=1e-6, err_msg=f'body {b}')
