"""Model compiler: reference MJCF -> committed JSON constant tables (`gym_quadruped_b200/assets/<robot>.json`).

    python -m gym_quadruped_b200.compiler.compile --reference /root/reference --out gym_quadruped_b200/assets

Besides re-shaping the XML attributes into the `QsModel` layout (include/qstep.h) it derives the compile-time
constants the engine needs: `body_invweight0`, `dof_invweight0`, `meaninertia`, all evaluated at
the XML reference pose `qpos0` *before* the env's joint-zero override (quadruped_env.py:171-173), bounding spheres for
the broad phase, and the convex hulls of collision meshes expressed in the owning body frame.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np

from .mjcf import GEOM_TYPES, convex_hull_vertices, load_mesh_vertices, parse_robot, quat_mul, quat_to_mat

LEGS = ('FL', 'FR', 'RL', 'RR')
EXPECTED_BODIES = ['base'] + [f'{leg}_{part}' for leg in LEGS for part in ('hip', 'thigh', 'calf')]

# robot name -> (mjcf path under gym_quadruped/robot_model, hip_height, qpos0_js override) ; robot_cfgs.py:31-60
ROBOTS = {
    'mini_cheetah': ('mini_cheetah/mini_cheetah.xml', 0.225, [0, -np.pi / 2, 0] * 2 + [0, np.pi / 2, 0] * 2),
    'aliengo': ('aliengo/aliengo.xml', 0.35, None),
    'go2': ('go2/go2.xml', 0.28, None),
    'hyqreal1': ('hyqreal1/hyqreal1.xml', 0.498, None),
    'hyqreal2': ('hyqreal2/hyqreal2.xml', 0.498, None),   # joint-level actuatorfrcrange clamps
    'go1': ('go1/go1.xml', 0.3, None),                    # fromto capsules, cylinders, condim-6 priority feet
    'b2': ('b2/b2.xml', 0.485, None),                     # cylinders
    'spot': ('spot/spot.xml', 0.46, None),                # implicitfast (== Euler + implicit damping here), OBJ hulls, contact excludes
}


def _fk(bodies, joints_by_body, qpos):
    """World poses of every body (index 0 = world) for generalized position qpos (nq=19)."""
    n = len(bodies) + 1
    xpos = np.zeros((n, 3))
    xquat = np.zeros((n, 4))
    xquat[0] = [1, 0, 0, 0]
    xaxis = {}
    for i, b in enumerate(bodies, start=1):
        p = b['parent']
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ b['pos']
        quat = quat_mul(xquat[p], b['quat'])
        for j in joints_by_body.get(i, []):
            if j['type'] == 'free':
                pos = qpos[0:3].copy()
                quat = qpos[3:7] / np.linalg.norm(qpos[3:7])
            else:
                R = quat_to_mat(quat)
                anchor = pos + R @ j['pos']
                axis = R @ j['axis']
                ang = qpos[j['qadr']]
                qloc = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * j['axis']])
                quat = quat_mul(quat, qloc)
                pos = anchor - quat_to_mat(quat) @ j['pos']
                xaxis[j['dof']] = (axis, anchor)
        xpos[i] = pos
        xquat[i] = quat / np.linalg.norm(quat)
    return xpos, xquat, xaxis


def _mass_matrix_and_invweights(bodies, joints_by_body, qpos0, armature):
    """Dense M(q0) from body-COM Jacobians (an independent formulation from the CRB recursion used at run time),
    then invweight0 = per-body / per-dof averages of diag(J M^-1 J^T)  (SURVEY.md App. A.6)."""
    nb = len(bodies) + 1
    nv = 18
    xpos, xquat, xaxis = _fk(bodies, joints_by_body, qpos0)
    Rb = quat_to_mat(xquat[1])

    def chain(i):
        c = []
        while i != 0:
            c.append(i)
            i = bodies[i - 1]['parent']
        return c

    def jac_at(i, point):
        jp = np.zeros((3, nv))
        jr = np.zeros((3, nv))
        for b in chain(i):
            for j in joints_by_body.get(b, []):
                if j['type'] == 'free':
                    jp[:, 0:3] = np.eye(3)
                    for k in range(3):
                        ax = Rb[:, k]
                        jr[:, 3 + k] = ax
                        jp[:, 3 + k] = np.cross(ax, point - xpos[1])
                else:
                    ax, anchor = xaxis[j['dof']]
                    jr[:, j['dof']] = ax
                    jp[:, j['dof']] = np.cross(ax, point - anchor)
        return jp, jr

    M = np.diag(np.asarray(armature, dtype=np.float64))
    J_com = {}
    for i, b in enumerate(bodies, start=1):
        R = quat_to_mat(xquat[i])
        xipos = xpos[i] + R @ b['ipos']
        Ri = R @ quat_to_mat(b['iquat'])
        Iw = Ri @ np.diag(b['inertia']) @ Ri.T
        jp, jr = jac_at(i, xipos)
        M += b['mass'] * jp.T @ jp + jr.T @ Iw @ jr
        J_com[i] = np.vstack([jp, jr])
    Minv = np.linalg.inv(M)
    body_invweight0 = np.zeros((nb, 2))
    for i in range(1, nb):
        A = J_com[i] @ Minv @ J_com[i].T
        body_invweight0[i, 0] = np.trace(A[0:3, 0:3]) / 3
        body_invweight0[i, 1] = np.trace(A[3:6, 3:6]) / 3
    d = np.diag(Minv).copy()
    dof_invweight0 = d.copy()
    dof_invweight0[0:3] = d[0:3].mean()
    dof_invweight0[3:6] = d[3:6].mean()
    meaninertia = float(np.trace(M) / nv)
    return M, body_invweight0, dof_invweight0, meaninertia


def compile_robot(reference_root: Path, robot: str) -> dict:
    rel, hip_height, qpos0_js = ROBOTS[robot]
    xml_path = reference_root / 'gym_quadruped' / 'robot_model' / rel
    r = parse_robot(xml_path)
    bodies, joints, geoms = r['bodies'], r['joints'], r['geoms']

    names = [b['name'] for b in bodies]
    assert names[1:] == EXPECTED_BODIES[1:], f'unexpected body tree {names}'  # the root body may be called base / trunk
    assert joints[0]['type'] == 'free' and joints[0]['body'] == 1
    hinges = joints[1:]
    assert len(hinges) == 12
    for k, j in enumerate(hinges):
        assert j['type'] == 'hinge' and j['body'] == k + 2, 'hinge k must live on body k+2'
        j['qadr'], j['dof'] = 7 + k, 6 + k
    for b in range(2, 14):
        expect_parent = 1 if (b - 2) % 3 == 0 else b - 1
        assert bodies[b - 1]['parent'] == expect_parent
    assert len(r['actuators']) == 12
    for k, a in enumerate(r['actuators']):
        assert a['joint'] == hinges[k]['name'], 'actuator k must drive hinge k'

    joints_by_body: dict[int, list] = {}
    for j in joints:
        joints_by_body.setdefault(j['body'], []).append(j)

    # compile-time reference pose: base at its XML pose, all hinge refs 0
    qpos0_xml = np.zeros(19)
    qpos0_xml[0:3] = bodies[0]['pos']
    qpos0_xml[3:7] = bodies[0]['quat']
    armature = np.zeros(18)
    damping = np.zeros(18)
    frictionloss = np.zeros(18)
    for j in hinges:
        armature[j['dof']] = j['armature']
        damping[j['dof']] = j['damping']
        frictionloss[j['dof']] = j['frictionloss']
    M0, body_invweight0, dof_invweight0, meaninertia = _mass_matrix_and_invweights(
        bodies, joints_by_body, qpos0_xml, armature)

    qpos0 = qpos0_xml.copy()
    if qpos0_js is not None:  # quadruped_env.py:171-173 -- applied AFTER compilation
        qpos0[7:] = qpos0_js
    key_qpos = r['key_qpos'] if r['key_qpos'] is not None else qpos0_xml
    assert key_qpos.shape == (19,)

    # geoms: keep those that can collide with the world (floor contype=conaffinity=1)
    verts_all = []
    gout = []
    foot_geom = {}
    for g in geoms:
        if not ((g['contype'] & 1) or (g['conaffinity'] & 1)):
            continue
        gtype = GEOM_TYPES[g['type']]
        R = quat_to_mat(g['quat'])
        entry = {
            'name': g['name'], 'type': gtype, 'body': g['body'], 'pos': g['pos'], 'quat': g['quat'],
            'size': g['size'], 'friction': g['friction'], 'condim': g['condim'], 'priority': g['priority'],
            'solref': g['solref'], 'solimp': g['solimp'], 'solmix': g['solmix'], 'margin': g['margin'],
            'gap': g['gap'], 'vertadr': 0, 'vertnum': 0, 'foot_leg': -1,
        }
        if g['type'] == 'mesh':
            path, scale = r['meshes'][g['mesh']]
            v = load_mesh_vertices(path) * scale
            hull = convex_hull_vertices(v)
            vb = hull @ R.T + g['pos']  # hull vertices in the BODY frame
            entry['vertadr'] = sum(len(x) for x in verts_all)
            entry['vertnum'] = len(vb)
            verts_all.append(vb)
            c = vb.mean(axis=0)
            entry['bcenter'] = c
            entry['rbound'] = float(np.linalg.norm(vb - c, axis=1).max())
        else:
            entry['bcenter'] = g['pos']
            s = g['size']
            if g['type'] == 'sphere':
                entry['rbound'] = float(s[0])
            elif g['type'] == 'capsule':
                entry['rbound'] = float(s[0] + s[1])
            elif g['type'] == 'box':
                entry['rbound'] = float(np.linalg.norm(s))
            elif g['type'] == 'cylinder':
                entry['rbound'] = float(np.hypot(s[0], s[1]))
            else:
                raise ValueError(f'unsupported collision geom type {g["type"]}')
        if g['name'] in LEGS:
            entry['foot_leg'] = LEGS.index(g['name'])
            foot_geom[g['name']] = len(gout)
            assert g['type'] == 'sphere' and g['body'] == 4 + 3 * LEGS.index(g['name']), 'foot must be a calf sphere'
        gout.append(entry)
    assert set(foot_geom) == set(LEGS), 'feet geoms FL/FR/RL/RR not found'

    imu = None
    acc = [s for s in r['sensors'] if s['type'] == 'accelerometer']
    gyr = [s for s in r['sensors'] if s['type'] == 'gyro']
    if acc and gyr:
        site = next(s for s in r['sites'] if s['name'] == acc[0]['site'])
        assert site['body'] == 1 and gyr[0]['site'] == acc[0]['site']
        imu = {'accel_name': acc[0]['name'], 'gyro_name': gyr[0]['name'], 'site': site['name'],
               'pos': site['pos'], 'quat': site['quat']}
    # sensordata address of each sensor (imu.py:231-238 sums dims of preceding sensors)
    dims = {'accelerometer': 3, 'gyro': 3, 'framepos': 3, 'framequat': 4, 'jointpos': 1, 'jointvel': 1,
            'velocimeter': 3, 'framelinvel': 3, 'frameangvel': 3}
    adr, sensor_adr = 0, {}
    for s in r['sensors']:
        sensor_adr[s['name']] = adr
        adr += dims[s['type']]

    # [MJ] actuatorfrcrange clamps the summed actuator force of a joint after the per-actuator forcerange; with one unit-gear
    # motor per hinge the two clamps nest into one interval
    for k, a in enumerate(r['actuators']):
        fr = hinges[k].get('actuatorfrcrange')
        if fr is not None:
            if a['forcelimited']:
                a['forcerange'] = np.array([max(a['forcerange'][0], fr[0]), min(a['forcerange'][1], fr[1])])
            else:
                a['forcerange'], a['forcelimited'] = np.asarray(fr, dtype=np.float64), True

    model = {
        'robot': robot, 'mjcf': rel, 'hip_height': hip_height,
        'cone': r['cone'], 'impratio': r['impratio'], 'meaninertia': meaninertia,
        'total_mass': float(sum(b['mass'] for b in bodies)),
        'body_names': ['world'] + names,
        'body_parent': [0] + [b['parent'] for b in bodies],
        'body_pos': [[0, 0, 0]] + [b['pos'] for b in bodies],
        'body_quat': [[1, 0, 0, 0]] + [b['quat'] for b in bodies],
        'body_ipos': [[0, 0, 0]] + [b['ipos'] for b in bodies],
        'body_iquat': [[1, 0, 0, 0]] + [b['iquat'] for b in bodies],
        'body_mass': [0.0] + [b['mass'] for b in bodies],
        'body_inertia': [[0, 0, 0]] + [b['inertia'] for b in bodies],
        'body_invweight0': body_invweight0,
        'joint_names': [j['name'] for j in hinges],
        'jnt_pos': [j['pos'] for j in hinges], 'jnt_axis': [j['axis'] for j in hinges],
        'jnt_range': [j['range'] for j in hinges], 'jnt_limited': [int(j['limited']) for j in hinges],
        'jnt_margin': [j['margin'] for j in hinges],
        'jnt_solref': [j['solreflimit'] for j in hinges], 'jnt_solimp': [j['solimplimit'] for j in hinges],
        'qpos0': qpos0, 'qpos0_compile': qpos0_xml, 'key_qpos': key_qpos,
        'dof_damping': damping, 'dof_armature': armature, 'dof_frictionloss': frictionloss,
        'dof_invweight0': dof_invweight0,
        'dof_solref': [[0.02, 1.0]] * 6 + [j['solreffriction'] for j in hinges],
        'dof_solimp': [[0.9, 0.95, 0.001, 0.5, 2.0]] * 6 + [j['solimpfriction'] for j in hinges],
        'actuator_names': [a['name'] for a in r['actuators']],
        'act_ctrlrange': [a['ctrlrange'] for a in r['actuators']],
        'act_ctrllimited': [int(a['ctrllimited']) for a in r['actuators']],
        'act_forcerange': [a['forcerange'] for a in r['actuators']],
        'act_forcelimited': [int(a['forcelimited']) for a in r['actuators']],
        'geoms': gout, 'foot_geom': [foot_geom[leg] for leg in LEGS],
        'vert': np.concatenate(verts_all) if verts_all else np.zeros((0, 3)),
        'imu': imu, 'sensor_adr': sensor_adr, 'M0_diag': np.diag(M0),
    }
    return model


def _jsonable(x):
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--out', default=str(Path(__file__).resolve().parents[1] / 'assets'))
    ap.add_argument('--robots', nargs='*', default=list(ROBOTS))
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    for robot in args.robots:
        model = compile_robot(Path(args.reference), robot)
        path = out / f'{robot}.json'
        path.write_text(json.dumps(_jsonable(model)))
        print(f'{robot}: mass={model["total_mass"]:.4f} ngeom={len(model["geoms"])} nvert={len(model["vert"])} '
              f'meaninertia={model["meaninertia"]:.5f} -> {path}')


if __name__ == '__main__':
    main()
