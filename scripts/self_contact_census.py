"""How often would the engine generate robot-robot contacts in a benchmark workload?

libqstep (kernel and oracle alike) does not generate robot self-collisions (deliberate deviations).  This script
measures how much that matters: it rolls the fp64 oracle through the bench workload (ctrl = 50*N(0,1), auto-reset on termination,
steady state after a pre-roll) and, at every step, measures the distance between every geom pair the engine would test
(different bodies, not parent and child, spot's two <exclude> pairs honoured) with a bounding-sphere cull followed by Gilbert's
minimum-norm-point iteration on the convex hulls (spheres / capsules / boxes / cylinders / meshes all handled through their
support functions).  A pair counts as "in contact" when its distance is below the sum of the two geom margins (penetration
included).  CPU only; reads nothing outside the repository.

    python scripts/self_contact_census.py [workload=cfg2] [envs=24] [steps=400]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from gym_quadruped_b200.model import Model  # noqa: E402
from oracle.oracle import Oracle, build  # noqa: E402


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def qmat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


class Geoms:
    """World poses of the collision geoms from an independent numpy FK, and their support functions."""

    def __init__(self, model):
        self.m = model
        c = model.c
        self.n = c.ngeom
        self.vert = model._vert if c.nvert else np.zeros((0, 3))
        self.pairs = []
        names = list(model.tables['body_names'])
        # spot_arm.xml:253-255 excludes the trunk against the two front upper legs
        excl = {(names[1], names[3]), (names[1], names[6])} if model.robot == 'spot' else set()
        for a in range(self.n):
            for b in range(a + 1, self.n):
                ba, bb = c.geom_body[a], c.geom_body[b]
                if ba == bb or c.body_parent[ba] == bb or c.body_parent[bb] == ba:
                    continue
                if ((names[ba], names[bb]) in excl or (names[bb], names[ba]) in excl):
                    continue
                self.pairs.append((a, b))

    def place(self, q):
        c = self.m.c
        pos = {1: q[:3]}
        quat = {1: q[3:7] / np.linalg.norm(q[3:7])}
        for b in range(2, 14):
            pa = c.body_parent[b]
            pos[b] = pos[pa] + qmat(quat[pa]) @ np.array(c.body_pos[b])
            qb = qmul(quat[pa], np.array(c.body_quat[b]))
            ang = q[7 + b - 2] - c.qpos0[7 + b - 2]
            quat[b] = qmul(qb, np.r_[np.cos(ang / 2), np.sin(ang / 2) * np.array(c.jnt_axis[b - 2])])
        self.R, self.p, self.W = [], [], []
        for g in range(self.n):
            b = c.geom_body[g]
            R = qmat(qmul(quat[b], np.array(c.geom_quat[g])))
            p = pos[b] + qmat(quat[b]) @ np.array(c.geom_pos[g])
            self.R.append(R); self.p.append(p)
            if c.geom_type[g] == 7:
                v = self.vert[c.geom_vertadr[g]:c.geom_vertadr[g] + c.geom_vertnum[g]]
                self.W.append(v @ R.T + p)
            else:
                self.W.append(None)

    def center_radius(self, g):
        c = self.m.c
        return self.p[g] + self.R[g] @ np.array(c.geom_bcenter[g]), c.geom_rbound[g]

    def support(self, g, d):
        """Farthest point of geom g in world direction d."""
        c = self.m.c
        t, s, R, p = c.geom_type[g], c.geom_size[g], self.R[g], self.p[g]
        if t == 7:
            W = self.W[g]
            return W[np.argmax(W @ d)]
        dl = R.T @ d
        n = np.linalg.norm(dl) + 1e-300
        if t == 2:
            loc = s[0] * dl / n
        elif t == 3:
            loc = s[0] * dl / n + np.array([0, 0, np.sign(dl[2]) * s[1]])
        elif t == 5:
            h = np.hypot(dl[0], dl[1]) + 1e-300
            loc = np.array([s[0] * dl[0] / h, s[0] * dl[1] / h, np.sign(dl[2]) * s[1]])
        elif t == 6:
            loc = np.sign(dl) * np.array(s[:3])
        else:
            raise ValueError(t)
        return p + R @ loc

    def distance(self, a, b, iters=60, tol=1e-5):
        """Gilbert's algorithm on the Minkowski difference A - B: distance between the two convex sets (0 when they overlap)."""
        v = self.p[a] - self.p[b]
        if not np.any(v):
            return 0.0
        for _ in range(iters):
            w = self.support(a, -v) - self.support(b, v)
            vv = v @ v
            if vv - v @ w <= tol * max(vv, 1e-12) or vv < 1e-14:
                break
            d = w - v
            t = min(1.0, max(0.0, -(v @ d) / (d @ d)))
            v = v + t * d
        return float(np.sqrt(v @ v))


def main():
    args = sys.argv[1:]
    wl = args[0] if args else 'cfg2'
    n_envs = int(args[1]) if len(args) > 1 else 24
    steps = int(args[2]) if len(args) > 2 else 400
    bench.select_workload(wl)
    build()
    bench._cpu_init(100, n_envs, 1)
    W = bench._W
    model = Model(bench.ROBOT, bench.SCENE)
    G = Geoms(model)
    c = model.c
    margin = [float(e['margin']) for e in model.tables['geoms']]
    bench._cpu_block(bench.PREROLL)
    rng, table = W['rng'], W['table']
    hit_steps = near_steps = total = 0
    pair_hits = {}
    cur = W['cursor']
    for s in range(steps):
        ctrl = rng.randn(1, 12) * bench.TORQUE_SCALE
        for e in W['envs']:
            _, cur = e.rollout_autoreset(ctrl, table, cur)
            q = e.get_state()[0]
            G.place(q)
            hit = near = False
            for a, b in G.pairs:
                ca, ra = G.center_radius(a)
                cb, rb = G.center_radius(b)
                mg = margin[a] + margin[b]
                if np.linalg.norm(ca - cb) > ra + rb + mg + 0.01:
                    continue
                d = G.distance(a, b)
                if d <= mg + 1e-4:
                    hit = True
                    pair_hits[(a, b)] = pair_hits.get((a, b), 0) + 1
                elif d <= mg + 0.01:
                    near = True
            hit_steps += hit; near_steps += near and not hit; total += 1
    print(f'workload {wl}: {bench.ROBOT}/{bench.SCENE}, {n_envs} envs x {steps} steps after a {bench.PREROLL}-step pre-roll, '
          f'{len(G.pairs)} candidate geom pairs')
    print(f'env-steps with at least one robot-robot geom pair in contact: {hit_steps} of {total} = {100 * hit_steps / total:.2f} %')
    print(f'env-steps with a pair within 1 cm but not touching:            {near_steps} of {total} = {100 * near_steps / total:.2f} %')
    bn = lambda g: model.tables['body_names'][c.geom_body[g]]
    for (a, b), k in sorted(pair_hits.items(), key=lambda kv: -kv[1])[:12]:
        print(f'   geom {a} ({bn(a)}) - geom {b} ({bn(b)}): {k} env-steps')


if __name__ == '__main__':
    main()
