"""MJCF-subset reader for the quadruped models of the reference.

Reads `gym_quadruped/robot_model/<robot>/<robot>.xml` (e.g. mini_cheetah.xml:1-211, aliengo.xml:1-248,
go2.xml:1-270, hyqreal1.xml:1-243) and produces the flat constant tables of `include/qstep.h::QsModel`.
Only the MJCF features those files use are implemented: nested <default> classes with `childclass` /
`class` inheritance, <inertial pos quat mass diaginertia>, free + hinge joints, sphere / capsule / box /
mesh geoms, <motor> actuators, <site>, accelerometer / gyro sensors, <keyframe>, <option cone impratio>,
<compiler autolimits>.  Engine-side defaults (solref, solimp, friction, ...) are the documented MJCF defaults
(SURVEY.md App. A).

This module runs on the build host only (it needs the reference checkout); its JSON output is committed under
`gym_quadruped_b200/assets/` and is what the runtime loads.
"""
from __future__ import annotations

import copy
import struct
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

GEOM_TYPES = {'plane': 0, 'hfield': 1, 'sphere': 2, 'capsule': 3, 'ellipsoid': 4, 'cylinder': 5, 'box': 6, 'mesh': 7}

GEOM_DEFAULTS = {
    'type': 'sphere', 'size': '0 0 0', 'pos': '0 0 0', 'quat': '1 0 0 0', 'friction': '1 0.005 0.0001',
    'condim': '3', 'contype': '1', 'conaffinity': '1', 'priority': '0', 'solref': '0.02 1',
    'solimp': '0.9 0.95 0.001 0.5 2', 'solmix': '1', 'margin': '0', 'gap': '0', 'group': '0',
}
JOINT_DEFAULTS = {
    'type': 'hinge', 'pos': '0 0 0', 'axis': '0 0 1', 'damping': '0', 'armature': '0', 'frictionloss': '0',
    'stiffness': '0', 'ref': '0', 'margin': '0', 'solreflimit': '0.02 1', 'solimplimit': '0.9 0.95 0.001 0.5 2',
    'solreffriction': '0.02 1', 'solimpfriction': '0.9 0.95 0.001 0.5 2',
}
MOTOR_DEFAULTS = {'gear': '1'}


def _vec(s, n=None, fill=None):
    v = [float(x) for x in str(s).split()]
    if n is not None and len(v) < n:
        assert fill is not None, f'expected {n} numbers, got {s!r}'
        v = v + list(fill[len(v):n])
    return np.asarray(v, dtype=np.float64)


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


class Defaults:
    """Nested <default> classes -> {class_name: {tag: {attr: value}}} with parent inheritance resolved."""

    def __init__(self, root: ET.Element):
        self.classes: dict[str, dict[str, dict[str, str]]] = {'main': {}}
        for top in root.findall('default'):
            self._walk(top, 'main', is_top=True)

    def _walk(self, node: ET.Element, parent: str, is_top: bool):
        name = node.attrib.get('class', 'main' if is_top else None)
        assert name is not None, 'nested <default> needs a class name'
        if name != 'main' or 'main' not in self.classes:
            self.classes[name] = copy.deepcopy(self.classes[parent])
        elif name == 'main' and not is_top:
            raise ValueError('class "main" re-declared')
        tbl = self.classes[name]
        for child in node:
            if child.tag == 'default':
                continue
            tbl.setdefault(child.tag, {}).update(child.attrib)
        for child in node.findall('default'):
            self._walk(child, name, is_top=False)

    def resolve(self, tag: str, elem: ET.Element, childclass: str | None, builtin: dict[str, str]) -> dict[str, str]:
        cls = elem.attrib.get('class', childclass) or 'main'
        assert cls in self.classes, f'unknown default class {cls!r}'
        out = dict(builtin)
        out.update(self.classes[cls].get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != 'class'})
        return out


def load_mesh_vertices(path: Path) -> np.ndarray:
    """OBJ ('v x y z' lines) or STL (binary / ascii) -> (n,3) float64 vertex array."""
    suffix = path.suffix.lower()
    if suffix == '.obj':
        verts = []
        with open(path, 'r', errors='ignore') as f:
            for line in f:
                if line.startswith('v '):
                    p = line.split()
                    verts.append((float(p[1]), float(p[2]), float(p[3])))
        return np.asarray(verts, dtype=np.float64)
    if suffix == '.stl':
        data = path.read_bytes()
        ntri = struct.unpack('<I', data[80:84])[0]
        if 84 + 50 * ntri == len(data):  # binary
            arr = np.frombuffer(data, dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]),
                                count=ntri, offset=84)
            return arr['v'].reshape(-1, 3).astype(np.float64)
        verts = []
        for line in data.decode('ascii', errors='ignore').splitlines():
            p = line.split()
            if len(p) == 4 and p[0] == 'vertex':
                verts.append((float(p[1]), float(p[2]), float(p[3])))
        return np.asarray(verts, dtype=np.float64)
    raise ValueError(f'unsupported mesh format: {path}')


def convex_hull_vertices(v: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull (qhull, the same library the engine uses for mesh collision)."""
    from scipy.spatial import ConvexHull

    v = np.unique(np.round(v, 12), axis=0)
    hull = ConvexHull(v)
    return v[np.sort(hull.vertices)]


def parse_robot(xml_path: str | Path) -> dict:
    """Parse one robot MJCF into an intermediate dict of numpy arrays (tree order = document order)."""
    xml_path = Path(xml_path)
    root = ET.parse(xml_path).getroot()
    defaults = Defaults(root)

    compiler = root.find('compiler')
    angle = (compiler.attrib.get('angle', 'degree') if compiler is not None else 'degree')
    assert angle == 'radian', 'only angle="radian" models are supported'
    autolimits = (compiler.attrib.get('autolimits', 'true') if compiler is not None else 'true') == 'true'
    assert autolimits, 'autolimits="false" not supported'

    option = root.find('option')
    opt = dict(option.attrib) if option is not None else {}
    out: dict = {
        'name': root.attrib.get('model', xml_path.stem),
        'cone': opt.get('cone', 'pyramidal'),
        'impratio': float(opt.get('impratio', 1.0)),
        'integrator': opt.get('integrator', 'Euler'),
    }
    # [MJ] implicitfast (spot.xml:4) solves (M - h dF/dv) qacc = f with dF/dv restricted to passive and actuator forces; for these
    # models (joint damping only, velocity-independent motors) that is the Euler step with implicit joint damping, (M + h D).
    assert out['integrator'] in ('Euler', 'implicitfast'), f"integrator {out['integrator']} is not implemented"

    meshes = {}
    asset = root.find('asset')
    if asset is not None:
        for m in asset.findall('mesh'):
            mattr = defaults.resolve('mesh', m, None, {'scale': '1 1 1'})
            name = mattr.get('name', Path(mattr['file']).stem)
            meshes[name] = (xml_path.parent / mattr['file'], _vec(mattr['scale']))

    bodies, joints, geoms, sites = [], [], [], []

    def walk(body: ET.Element, parent: int, childclass: str | None):
        childclass = body.attrib.get('childclass', childclass)
        bid = len(bodies) + 1
        inertial = body.find('inertial')
        assert inertial is not None, f'body {body.attrib.get("name")} needs an explicit <inertial>'
        assert 'diaginertia' in inertial.attrib, 'fullinertia not supported'
        bodies.append({
            'name': body.attrib.get('name', f'body{bid}'),
            'parent': parent,
            'pos': _vec(body.attrib.get('pos', '0 0 0')),
            'quat': quat_normalize(_vec(body.attrib.get('quat', '1 0 0 0'))),
            'ipos': _vec(inertial.attrib.get('pos', '0 0 0')),
            'iquat': quat_normalize(_vec(inertial.attrib.get('quat', '1 0 0 0'))),
            'mass': float(inertial.attrib['mass']),
            'inertia': _vec(inertial.attrib['diaginertia']),
        })
        for j in list(body.findall('joint')) + list(body.findall('freejoint')):
            if j.tag == 'freejoint':
                joints.append({'name': j.attrib.get('name', 'root'), 'type': 'free', 'body': bid})
                continue
            a = defaults.resolve('joint', j, childclass, JOINT_DEFAULTS)
            if a['type'] == 'free':
                joints.append({'name': a.get('name', 'root'), 'type': 'free', 'body': bid})
                continue
            assert a['type'] == 'hinge', f'unsupported joint type {a["type"]}'
            assert float(a['stiffness']) == 0.0 and float(a['ref']) == 0.0
            has_range = 'range' in a
            lim = a.get('limited', 'auto')
            limited = has_range if lim == 'auto' else (lim == 'true')
            # joint-level clamp of the total actuator force (hyqreal2.xml:36); `actuatorfrclimited` follows autolimits
            frc_lim = a.get('actuatorfrclimited', 'auto')
            has_frc = ('actuatorfrcrange' in a) if frc_lim == 'auto' else (frc_lim == 'true')
            joints.append({
                'actuatorfrcrange': _vec(a['actuatorfrcrange']) if has_frc else None,
                'name': a['name'], 'type': 'hinge', 'body': bid,
                'pos': _vec(a['pos']), 'axis': quat_normalize(_vec(a['axis'])),
                'range': _vec(a['range']) if has_range else np.zeros(2), 'limited': bool(limited),
                'damping': float(a['damping']), 'armature': float(a['armature']),
                'frictionloss': float(a['frictionloss']), 'margin': float(a['margin']),
                'solreflimit': _vec(a['solreflimit']), 'solimplimit': _vec(a['solimplimit'], 5, (0.9, 0.95, 0.001, 0.5, 2)),
                'solreffriction': _vec(a['solreffriction']),
                'solimpfriction': _vec(a['solimpfriction'], 5, (0.9, 0.95, 0.001, 0.5, 2)),
            })
        for g in body.findall('geom'):
            a = defaults.resolve('geom', g, childclass, GEOM_DEFAULTS)
            contype, conaff = int(a['contype']), int(a['conaffinity'])
            if contype == 0 and conaff == 0:
                continue  # visual only; bodies carry explicit inertials so these never matter
            gtype = a['type']
            if 'mesh' in a and 'type' not in g.attrib and gtype == 'sphere':
                gtype = 'mesh'
            if 'fromto' in a:
                # [MJ] fromto (go1.xml:47-59): centre at the midpoint, half-length from the segment, z axis along from - to
                # (shortest-arc rotation of +z, as the engine's z2quat)
                assert gtype in ('capsule', 'cylinder'), 'fromto only for capsules / cylinders'
                ft = _vec(a['fromto'])
                p1, p2 = ft[:3], ft[3:]
                vec = p1 - p2
                length = float(np.linalg.norm(vec))
                z = vec / length
                axis = np.cross([0.0, 0.0, 1.0], z)
                sn = float(np.linalg.norm(axis))
                ang = float(np.arctan2(sn, z[2]))
                if sn < 1e-10:
                    quat = np.array([1.0, 0, 0, 0]) if z[2] > 0 else np.array([0.0, 1.0, 0, 0])
                else:
                    quat = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis / sn])
                a = dict(a)
                a['pos'] = ' '.join(repr(float(x)) for x in 0.5 * (p1 + p2))
                a['quat'] = ' '.join(repr(float(x)) for x in quat)
                a['size'] = f"{float(_vec(a['size'], 3, (0, 0, 0))[0])!r} {0.5 * length!r}"
            geoms.append({
                'name': a.get('name', ''), 'body': bid, 'type': gtype, 'mesh': a.get('mesh'),
                'size': _vec(a['size'], 3, (0, 0, 0)), 'pos': _vec(a['pos']), 'quat': quat_normalize(_vec(a['quat'])),
                'friction': _vec(a['friction'], 3, (1, 0.005, 0.0001)), 'condim': int(a['condim']),
                'contype': contype, 'conaffinity': conaff, 'priority': int(a['priority']),
                'solref': _vec(a['solref']), 'solimp': _vec(a['solimp'], 5, (0.9, 0.95, 0.001, 0.5, 2)),
                'solmix': float(a['solmix']), 'margin': float(a['margin']), 'gap': float(a['gap']),
            })
        for s in body.findall('site'):
            a = defaults.resolve('site', s, childclass, {'pos': '0 0 0', 'quat': '1 0 0 0'})
            sites.append({'name': a.get('name', ''), 'body': bid, 'pos': _vec(a['pos']),
                          'quat': quat_normalize(_vec(a['quat']))})
        for child in body.findall('body'):
            walk(child, bid, childclass)

    worldbody = root.find('worldbody')
    tops = worldbody.findall('body')
    assert len(tops) == 1, 'expected a single root body'
    walk(tops[0], 0, None)

    actuators = []
    act = root.find('actuator')
    for m in (act.findall('motor') if act is not None else []):
        a = defaults.resolve('motor', m, None, MOTOR_DEFAULTS)
        assert float(str(a['gear']).split()[0]) == 1.0
        actuators.append({
            'name': a.get('name', ''), 'joint': a['joint'],
            'ctrlrange': _vec(a['ctrlrange']) if 'ctrlrange' in a else np.zeros(2),
            'ctrllimited': ('ctrlrange' in a) if a.get('ctrllimited', 'auto') == 'auto' else a['ctrllimited'] == 'true',
            'forcerange': _vec(a['forcerange']) if 'forcerange' in a else np.zeros(2),
            'forcelimited': ('forcerange' in a) if a.get('forcelimited', 'auto') == 'auto' else a['forcelimited'] == 'true',
        })

    sensors = []
    sens = root.find('sensor')
    for s in (list(sens) if sens is not None else []):
        sensors.append({'type': s.tag, 'name': s.attrib.get('name', ''), 'site': s.attrib.get('site'),
                        'objname': s.attrib.get('objname'), 'joint': s.attrib.get('joint')})

    key = root.find('keyframe')
    key_qpos = None
    if key is not None and key.find('key') is not None and 'qpos' in key.find('key').attrib:
        key_qpos = _vec(key.find('key').attrib['qpos'])

    out.update({'bodies': bodies, 'joints': joints, 'geoms': geoms, 'sites': sites, 'actuators': actuators,
                'sensors': sensors, 'key_qpos': key_qpos, 'meshes': meshes})
    return out
