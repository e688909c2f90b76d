"""Generates tests/golden/terrain_boxes_<robot>.json by running the REFERENCE's own terrain generator
(/root/reference/gym_quadruped/utils/mujoco/terrain.py) in this container. `noise` (absent) is stubbed: the
random_boxes scene never calls it.  Run once on the build host; the JSON is committed, the reference is not needed
at test time.

    python tests/golden/make_terrain_golden.py
"""
import importlib.util
import json
import sys
import types
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

REF = Path('/root/reference/gym_quadruped/utils/mujoco')
sys.modules.setdefault('noise', types.ModuleType('noise'))
spec = importlib.util.spec_from_file_location('ref_terrain', REF / 'terrain.py')
ref_terrain = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_terrain)

out_dir = Path(__file__).resolve().parent
for robot, hip in (('mini_cheetah', 0.225), ('aliengo', 0.35), ('go2', 0.28), ('hyqreal1', 0.498)):
    np.random.seed(1234)
    before = np.random.get_state()[1].copy()
    scene, limits = ref_terrain.generate_terrain(Path('/nonexistent/scene_random_boxes.xml'), REF / 'assets', hip,
                                                 'random_boxes', seed=10)
    assert (np.random.get_state()[1] == before).all(), 'global RNG must be restored'
    geoms = scene.getroot().find('worldbody').findall('geom')
    boxes = [g for g in geoms if g.attrib.get('type') == 'box']
    rec = {'robot': robot, 'hip_height': hip, 'terrain_limits': [float(x) for x in limits],
           'n_world_geoms': len(geoms),
           'pos': [[float(x) for x in g.attrib['pos'].split()] for g in boxes],
           'half': [[float(x) for x in g.attrib['size'].split()] for g in boxes],
           'quat': [[float(x) for x in g.attrib['quat'].split()] for g in boxes]}
    (out_dir / f'terrain_boxes_{robot}.json').write_text(json.dumps(rec))
    print(robot, len(boxes), limits)


# static XML scenes (stairs, ramp) and the procedural pyramid stack -> tests/golden/terrain_static.json
def _boxes_of(scene):
    geoms = scene.getroot().find('worldbody').findall('geom')
    boxes = [g for g in geoms if g.attrib.get('type') == 'box']
    return {'pos': [[float(x) for x in g.attrib['pos'].split()] for g in boxes],
            'half': [[float(x) for x in g.attrib['size'].split()] for g in boxes],
            'quat': [[float(x) for x in g.attrib.get('quat', '1 0 0 0').split()] for g in boxes]}


static = {}
for name in ('stairs', 'ramp', 'slippery'):
    scene, limits = ref_terrain.generate_terrain(Path(f'/root/reference/gym_quadruped/robot_model/scene_{name}.xml'), REF / 'assets',
                                                 0.35, name, seed=10)
    static[name] = dict(_boxes_of(scene), terrain_limits=[float(x) for x in limits])
    if name == 'slippery':
        boxes = [g for g in scene.getroot().find('worldbody').findall('geom') if g.attrib.get('type') == 'box']
        static[name]['friction'] = [[float(x) for x in g.attrib['friction'].split()] for g in boxes]
        static[name]['priority'] = [int(g.attrib['priority']) for g in boxes]
for robot, hip in (('mini_cheetah', 0.225), ('aliengo', 0.35), ('go2', 0.28), ('hyqreal1', 0.498)):
    scene, limits = ref_terrain.generate_terrain(Path('/nonexistent/scene_random_pyramids.xml'), REF / 'assets', hip,
                                                 'random_pyramids', seed=10)
    static[f'random_pyramids_{robot}'] = dict(_boxes_of(scene), terrain_limits=[float(x) for x in limits], hip_height=hip)
(out_dir / 'terrain_static.json').write_text(json.dumps(static))
print({k: len(v['pos']) for k, v in static.items()})
