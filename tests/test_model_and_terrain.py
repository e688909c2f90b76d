"""Compiled model tables and procedural terrain against the reference's own numbers."""
import json
from pathlib import Path

import numpy as np
import pytest

from gym_quadruped_b200.model import Model, load_robot_tables
from gym_quadruped_b200.terrain import FLAT_LIMITS, generate_terrain, perlin_image, world_of_boxes

GOLDEN = Path(__file__).resolve().parent / 'golden'


@pytest.mark.parametrize('robot', ['mini_cheetah', 'aliengo', 'go2', 'hyqreal1'])
def test_random_boxes_bit_exact_against_reference_generator(robot):
    g = json.loads((GOLDEN / f'terrain_boxes_{robot}.json').read_text())  # dumped from the reference's terrain.py
    t = world_of_boxes(g['hip_height'])
    assert len(t['box_pos']) == 100 and g['n_world_geoms'] == 101
    assert np.array_equal(t['box_pos'], np.array(g['pos']))
    assert np.array_equal(t['box_half'], np.array(g['half']))
    np.testing.assert_allclose(t['box_quat'], np.array(g['quat']), atol=1e-15)
    assert tuple(t['terrain_limits']) == tuple(g['terrain_limits'])


def test_flat_and_perlin_scene_tables():
    assert generate_terrain('flat', 0.3)['terrain_limits'] == FLAT_LIMITS == (10000, -10000, 10000, -10000)
    t = generate_terrain('perlin', 0.35)
    assert t['data'].shape == (128, 128) and t['data'].min() == 0.0 and t['data'].max() == 1.0
    np.testing.assert_allclose(t['size'], (17.5, 17.5, 0.70, 0.005))
    np.testing.assert_allclose(t['terrain_limits'], (14.0, -14.0, 14.0, -14.0))
    img = perlin_image()  # regression values of SURVEY.md App. D.5 (best-effort restatement of the absent `noise` package)
    assert (img.min(), img.max()) == (64, 195) and list(img[0, :6]) == [127, 136, 129, 135, 143, 143]
    assert list(img[64, 60:66]) == [158, 152, 152, 159, 164, 177]
    with pytest.raises(ValueError):
        generate_terrain('lava', 0.3)


@pytest.mark.parametrize('key,scene,hip', [('stairs', 'stairs', 0.35), ('ramp', 'ramp', 0.35), ('slippery', 'slippery', 0.35), ('random_pyramids_mini_cheetah', 'random_pyramids', 0.225),
                                           ('random_pyramids_aliengo', 'random_pyramids', 0.35), ('random_pyramids_go2', 'random_pyramids', 0.28),
                                           ('random_pyramids_hyqreal1', 'random_pyramids', 0.498)])
def test_static_box_scenes_against_reference_generator(key, scene, hip):
    """stairs / ramp (the reference's XML scenes) and random_pyramids (its generator run here) -> tests/golden/terrain_static.json."""
    g = json.loads((GOLDEN / 'terrain_static.json').read_text())[key]
    t = generate_terrain(scene, hip)
    assert t['type'] == 'boxes' and len(t['box_pos']) == len(g['pos']) > 0
    assert np.array_equal(t['box_pos'], np.array(g['pos'])) and np.array_equal(t['box_half'], np.array(g['half']))
    q = np.array(g['quat'])
    np.testing.assert_allclose(t['box_quat'], q / np.linalg.norm(q, axis=1, keepdims=True), atol=1e-15)
    assert tuple(float(x) for x in t['terrain_limits']) == tuple(g['terrain_limits'])
    if 'friction' in g:  # scene_slippery.xml: per-surface friction triples and priority 2
        assert np.array_equal(t['box_friction'], np.array(g['friction'])) and set(g['priority']) == {t['box_priority']} == {2}


def test_model_constants_from_the_mjcf():
    """SURVEY.md App. C, values read from the XML files."""
    mc = load_robot_tables('mini_cheetah')
    assert abs(mc['total_mass'] - 12.473) < 1e-3 and mc['cone'] == 'pyramidal' and mc['hip_height'] == 0.225
    assert sum(mc['jnt_limited']) == 0                                   # ranges live in unused default classes (App. B.14)
    assert len(mc['geoms']) == 15 and sum(g['type'] == 7 for g in mc['geoms']) == 11
    assert all(g['condim'] == 1 and g['margin'] == 0.001 and g['friction'][0] == 0.6 for g in mc['geoms'])
    np.testing.assert_allclose(mc['act_ctrlrange'][:3], [[-23.7, 23.7], [-23.7, 23.7], [-45.43, 45.43]])
    np.testing.assert_allclose(mc['qpos0'][7:13], [0, -np.pi / 2, 0, 0, -np.pi / 2, 0])
    np.testing.assert_allclose(mc['dof_damping'][6:], 0.2); np.testing.assert_allclose(mc['dof_frictionloss'][6:], 0.2)
    al = load_robot_tables('aliengo')
    assert abs(al['total_mass'] - 24.638) < 1e-3 and al['jnt_limited'] == [1, 0, 1] * 4
    assert sorted(g['type'] for g in al['geoms']).count(6) == 9 and sum(g['type'] == 3 for g in al['geoms']) == 4
    go2 = load_robot_tables('go2')
    assert go2['cone'] == 'elliptic' and go2['impratio'] == 100 and len(go2['geoms']) == 31
    foot = go2['geoms'][go2['foot_geom'][0]]
    assert foot['condim'] == 6 and foot['priority'] == 1 and foot['friction'] == [0.8, 0.02, 0.01]
    hy = load_robot_tables('hyqreal1')
    assert abs(hy['total_mass'] - 107.573) < 1e-2 and hy['sensor_adr']['Body_Acc'] == 24  # 24 joint sensors come first
    radii = [hy['geoms'][i]['size'][0] for i in hy['foot_geom']]
    assert radii == [0.036, 0.032, 0.032, 0.032]                          # App. B.15
    assert hy['imu'] is not None and hy['imu']['accel_name'] == 'Body_Acc'


def test_qsmodel_struct_roundtrip():
    m = Model('go2', 'random_boxes')
    assert m.c.nbox == 100 and m.c.terrain_type == 2 and m.c.cone == 1 and m.c.ngeom == 31
    np.testing.assert_allclose(m.terrain_limits[0], 8.051757569573601)
    assert list(m.c.body_parent) == [0, 0, 1, 2, 3, 1, 5, 6, 1, 8, 9, 1, 11, 12]
    with pytest.raises(ValueError):
        Model('hyqreal', 'flat')  # the reference rejects this name too (robot_cfgs.py:49-58)


def test_slippery_scene_and_new_robots_fill_the_model_struct():
    m = Model('aliengo', 'slippery')
    assert m.c.terrain_type == 2 and m.c.nbox == 2 and m.c.box_par.priority == 2 and m.c.box_par.condim == 3
    assert list(m.c.box_friction[0]) == [0.03, 0.05, 0.07] and list(m.c.box_friction[1]) == [0.8, 0.2, 0.3]
    assert tuple(m.c.terrain_limits) == FLAT_LIMITS
    m = Model('aliengo', 'random_boxes')
    assert m.c.box_par.priority == 0 and list(m.c.box_friction[5]) == [1.0, 0.005, 0.0001]
    go1 = Model('go1', 'flat')
    assert go1.c.ngeom == 42 and go1.c.cone == 1 and sorted(set(go1.c.geom_type[:42])) == [2, 3, 5, 6]  # sphere, capsule, cylinder, box
    b2, spot, h2 = Model('b2', 'flat'), Model('spot', 'flat'), Model('hyqreal2', 'flat')
    assert list(b2.c.geom_type[:b2.c.ngeom]).count(5) == 4 and spot.c.nvert > 1000 and spot.c.cone == 1
    # hyqreal2.xml:36-51: joint-level actuatorfrcrange (150 / 250 / 350 N m) folded into the actuator force clamp
    assert [h2.c.act_forcerange[a][1] for a in range(3)] == [150.0, 250.0, 350.0] and all(h2.c.act_forcelimited[a] for a in range(12))


def test_fromto_capsules_of_go1_reproduce_their_end_points():
    """go1.xml:47-59 gives its thigh / calf capsules as `fromto` segments; the compiler turns them into centre + half-length + a
    frame whose z axis runs along from - to.  Rebuild the end points from the compiled tables."""
    t = load_robot_tables('go1')
    segs = {(-0.02, 0, 0, -0.02, 0, -0.16): 0.015, (0, 0, 0, -0.02, 0, -0.1): 0.015, (-0.02, 0, -0.16, 0, 0, -0.2): 0.015,
            (0, 0, 0, 0.02, 0, -0.13): 0.01, (0.02, 0, -0.13, 0, 0, -0.2): 0.01}
    found = 0
    for g in t['geoms']:
        if g['type'] != 3:
            continue
        q = np.array(g['quat']); w, x, y, z = q
        zaxis = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), w * w - x * x - y * y + z * z])
        p1, p2 = np.array(g['pos']) + zaxis * g['size'][1], np.array(g['pos']) - zaxis * g['size'][1]
        for seg, rad in segs.items():
            if np.allclose(np.r_[p1, p2], seg, atol=1e-12) and abs(g['size'][0] - rad) < 1e-15:
                found += 1
    assert found == 20  # five segments on each of the four legs


def test_self_contact_census_distance_routine():
    """scripts/self_contact_census.py quantifies the robot-robot contacts the engine would generate (deviations): its
    convex-distance routine (Gilbert's iteration on support functions) against closed forms, on geoms of a compiled robot."""
    import importlib.util
    import sys
    from pathlib import Path

    spec = importlib.util.spec_from_file_location('census', Path(__file__).resolve().parents[1] / 'scripts' / 'self_contact_census.py')
    census = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['census']
    try:
        spec.loader.exec_module(census)
    finally:
        sys.argv = argv
    from gym_quadruped_b200.model import Model

    m = Model('go2', 'flat')
    G = census.Geoms(m)
    q = np.array(m.c.key_qpos, dtype=float)
    G.place(q)
    c = m.c
    spheres = [g for g in range(c.ngeom) if c.geom_type[g] == 2]
    boxes = [g for g in range(c.ngeom) if c.geom_type[g] == 6]
    assert len(spheres) >= 4 and boxes
    # sphere - sphere: |c1 - c2| - r1 - r2
    a, b = spheres[1], spheres[2]
    want = np.linalg.norm(G.p[a] - G.p[b]) - c.geom_size[a][0] - c.geom_size[b][0]
    assert abs(G.distance(a, b, iters=200, tol=1e-9) - want) < 1e-4
    # sphere - box: closed form in the box frame
    a, b = spheres[1], boxes[0]
    loc = G.R[b].T @ (G.p[a] - G.p[b])
    half = np.array(c.geom_size[b][:3])
    want = np.linalg.norm(np.maximum(np.abs(loc) - half, 0.0)) - c.geom_size[a][0]
    assert want > 0 and abs(G.distance(a, b, iters=400, tol=1e-10) - want) < 1e-4
    # overlapping geoms report (almost) zero; the same geom pair list as the engine's filter: no parent-child, no same-body pairs
    assert G.distance(boxes[0], boxes[0]) == 0.0
    for a, b in G.pairs:
        ba, bb = c.geom_body[a], c.geom_body[b]
        assert ba != bb and c.body_parent[ba] != bb and c.body_parent[bb] != ba
