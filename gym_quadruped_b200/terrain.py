"""Procedural scenes as numeric tables (no XML round trip).

Restates what `gym_quadruped/utils/mujoco/terrain.py:309-365` builds for the scenes BASELINE.json names:

* ``flat``          -- the `floor` plane of scene_flat.xml:32, limits (1e4,-1e4,1e4,-1e4)            terrain.py:357-359
* ``random_boxes``  -- floor + 10x10 randomly sized / tilted static boxes                             terrain.py:145-238,325-335
* ``perlin``        -- floor + a 128x128 height field of 5-octave Perlin noise                        terrain.py:25-119,345-356
* ``stairs``        -- floor + 50 static steps (robot_model/scene_stairs.xml:38-89), flat limits           terrain.py:319-321
* ``ramp``          -- floor + one tilted slab (robot_model/scene_ramp.xml:37)                             terrain.py:319-321
* ``random_pyramids`` -- floor + a stack of shrinking slabs                                            terrain.py:240-292,336-344
* ``slippery``      -- floor + two priority-2 strips with their own friction (robot_model/scene_slippery.xml:39-40)

All scenes are generated under the reference's fixed seed 10 (quadruped_env.py:155, terrain.py:299-306) with NumPy's
legacy MT19937 stream, so every env instance of a robot sees the same terrain.  `random_boxes` is pinned bit-for-bit by
`tests/golden/terrain_boxes_*.json` (dumped from the reference's own terrain.py).  `perlin` depends on the third-party
`noise` C extension (absent here); `pnoise2` below restates its published algorithm and is flagged
best-effort -- parity tests always feed the same explicit height array to oracle and kernel.
"""
from __future__ import annotations

import numpy as np

FLAT_LIMITS = (10000, -10000, 10000, -10000)


def _euler_xyz_to_quat_wxyz(e):
    """scipy `Rotation.from_euler('xyz', e).as_quat(canonical=True, scalar_first=True)` (extrinsic x-y-z)."""
    cr, sr = np.cos(e[0] / 2), np.sin(e[0] / 2)
    cp, sp = np.cos(e[1] / 2), np.sin(e[1] / 2)
    cy, sy = np.cos(e[2] / 2), np.sin(e[2] / 2)
    q = np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                  cr * cp * sy - sr * sp * cy])
    if q[0] < 0 or (q[0] == 0 and (q[1] < 0 or (q[1] == 0 and (q[2] < 0 or (q[2] == 0 and q[3] < 0))))):
        q = -q
    return q


def world_of_boxes(hip_height: float, seed: int = 10):
    """terrain.py:325-335 -> add_world_of_boxes(:145-238). Draw order per box: size_xy(2), size_z(1), euler(3), sep_x(1),
    sep_y(1), after one initial separation draw(2)."""
    rng = np.random.RandomState(seed)
    origin = np.array([0.5, -3.0, 0.02])
    grid = (10, 10)
    sep0 = np.array([2 * hip_height, 2 * hip_height])
    sep_rand = np.array([0.0, 1.0])
    size0 = np.array([2 * hip_height, 2 * hip_height, hip_height / 2.0])
    size_rand = np.array([0.5 * hip_height, 0.5 * hip_height, hip_height / 2])
    euler_rand = np.array([0.1, 0.1, 2 * np.pi])

    pos_l, quat_l, half_l = [], [], []
    sep = sep0 + sep_rand * rng.uniform(-1.0, 1.0, 2)
    x = 0.0
    far_x = far_y = 0.0
    sgn_x = sgn_y = 0
    for _ in range(grid[0]):
        x += sep[0]
        y = 0.0
        for _ in range(grid[1]):
            sxy = size0[:2] + size_rand[:2] * rng.uniform(-0.2, 0.2, 2)
            sz = size0[2] + size_rand[2] * rng.uniform(-0.1, 0.15, 1)
            euler = euler_rand * rng.uniform(-1.0, 1.0, 3)
            sep = np.array([sep0[0] + sep_rand[0] * rng.uniform(0, 0.5, 1)[0],
                            sep0[1] + sep_rand[1] * rng.uniform(-0.5, 0.5, 1)[0]])
            y += sep[1]
            pos_l.append(np.array([x, y, 0.0]) + origin)
            quat_l.append(_euler_xyz_to_quat_wxyz(euler))
            half_l.append(0.5 * np.array([sxy[0], sxy[1], sz[0]]))
            ax, ay = abs(x + origin[0]), abs(y + origin[1])
            if ax >= far_x:
                far_x, sgn_x = ax, (1 if ax > 0 else -1)
            if ay >= far_y:
                far_y, sgn_y = ay, (1 if ay > 0 else -1)
    mx, my = far_x * sgn_x, far_y * sgn_y
    cx, cy = (mx + origin[0]) / 2, (my + origin[1]) / 2
    rad = 1.2 * np.sqrt(2 * (mx - cx) ** 2) if far_x >= far_y else 1.2 * np.sqrt(2 * (my - cy) ** 2)
    return {'type': 'boxes', 'box_pos': np.array(pos_l), 'box_quat': np.array(quat_l), 'box_half': np.array(half_l),
            'terrain_limits': (cx + rad, cx - rad, cy + rad, cy - rad)}


# ---- caseman/noise `pnoise2` (improved Perlin, fp32) -- [from memory; unverifiable offline] ---------
_PERM = np.array([
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10,
    23, 190, 6, 148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87,
    174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211,
    133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208,
    89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250, 124, 123, 5,
    202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119,
    248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232,
    178, 185, 112, 104, 218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249,
    14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205,
    93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180], dtype=np.int64)
_PERM = np.concatenate([_PERM, _PERM])
_GRAD3 = np.array([(1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0), (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1),
                   (0, 1, 1), (0, -1, 1), (0, 1, -1), (0, -1, -1), (1, 0, -1), (-1, 0, -1), (0, -1, 1), (0, 1, 1)],
                  dtype=np.float32)


def _noise2(x, y, rx, ry, base=0):
    f32 = np.float32
    x, y = f32(x), f32(y)
    i = int(np.floor(np.fmod(x, f32(rx))))
    j = int(np.floor(np.fmod(y, f32(ry))))
    ii = int(np.fmod(f32(i + 1), f32(rx)))
    jj = int(np.fmod(f32(j + 1), f32(ry)))
    i, j, ii, jj = (i & 255) + base, (j & 255) + base, (ii & 255) + base, (jj & 255) + base
    x = f32(x - np.floor(x))
    y = f32(y - np.floor(y))
    fx = f32(x * x * x * (x * (x * f32(6) - f32(15)) + f32(10)))
    fy = f32(y * y * y * (y * (y * f32(6) - f32(15)) + f32(10)))
    A, B = _PERM[i], _PERM[ii]
    AA, AB, BA, BB = _PERM[A + j], _PERM[A + jj], _PERM[B + j], _PERM[B + jj]

    def grad(h, gx, gy):
        g = _GRAD3[h & 15]
        return f32(gx * g[0] + gy * g[1])

    def lerp(t, a, b):
        return f32(a + t * (b - a))

    return lerp(fy, lerp(fx, grad(_PERM[AA], x, y), grad(_PERM[BA], f32(x - 1), y)),
                lerp(fx, grad(_PERM[AB], x, f32(y - 1)), grad(_PERM[BB], f32(x - 1), f32(y - 1))))


def pnoise2(x, y, octaves=1, persistence=0.5, lacunarity=2.0, repeatx=1024, repeaty=1024, base=0):
    f32 = np.float32
    freq, amp, mx, total = f32(1), f32(1), f32(0), f32(0)
    for _ in range(octaves):
        total = f32(total + _noise2(f32(x) * freq, f32(y) * freq, f32(repeatx) * freq, f32(repeaty) * freq, base) * amp)
        mx = f32(mx + amp)
        freq = f32(freq * f32(lacunarity))
        amp = f32(amp * f32(persistence))
    return float(total / mx)


def perlin_image(width=128, height=128, smooth=50.0, octaves=5, persistence=0.5, lacunarity=4.0):
    """8-bit terrain image exactly as terrain.py:75-86 fills it."""
    img = np.zeros((height, width), dtype=np.uint8)
    for y in range(width):
        for x in range(width):
            img[y, x] = int((pnoise2(x / smooth, y / smooth, octaves=octaves, persistence=persistence,
                                     lacunarity=lacunarity) + 1) / 2 * 255)
    return img


_PERLIN_CACHE: dict = {}


def hfield_from_image(img: np.ndarray) -> np.ndarray:
    """Engine-side PNG -> height data: rows flipped vertically, values normalised to [0,1] [MJ]."""
    data = np.asarray(img, dtype=np.float64)[::-1].copy()
    lo, hi = data.min(), data.max()
    data = (data - lo) / (hi - lo) if hi > lo else np.zeros_like(data)
    return data.astype(np.float32)


def perlin_heightfield(hip_height: float):
    """terrain.py:345-356 -> add_perlin_heightfield(:25-119)."""
    if 'img' not in _PERLIN_CACHE:
        _PERLIN_CACHE['img'] = perlin_image()
    size = hip_height * 100
    radius = 0.8 * (size / 2.0)
    return {'type': 'hfield', 'data': hfield_from_image(_PERLIN_CACHE['img']),
            'size': (size / 2.0, size / 2.0, 2 * hip_height, 0.005), 'pos': (0.0, 0.0, 0.0),
            'terrain_limits': (radius, -radius, radius, -radius)}


def stairs_scene():
    """robot_model/scene_stairs.xml: 50 steps of half size (0.05, 1.25, 0.025); the XML's positions are running sums (0.1 k + 1 and
    0.05 k - 0.025, accumulated in double precision), reproduced here the same way and pinned by tests/golden/terrain_static.json."""
    pos, lx, z = [], 0.0, -0.025
    for _ in range(50):  # local x and z accumulate step by step; the stair case starts 1 m in front of the origin
        lx += 0.1
        z += 0.05
        pos.append([lx + 1.0, 0.0, z])
    n = len(pos)
    return {'type': 'boxes', 'box_pos': np.array(pos), 'box_quat': np.tile([1.0, 0, 0, 0], (n, 1)),
            'box_half': np.tile([0.05, 1.25, 0.025], (n, 1)), 'terrain_limits': FLAT_LIMITS}


def ramp_scene():
    """robot_model/scene_ramp.xml:37: one slab pitched by quat (1, 0, -0.2, 0) (normalised by the engine's compiler)."""
    q = np.array([1.0, 0.0, -0.20, 0.0])
    return {'type': 'boxes', 'box_pos': np.array([[0.5, 0.0, 0.025]]), 'box_quat': (q / np.linalg.norm(q))[None],
            'box_half': np.array([[4.05, 1.25, 0.025]]), 'terrain_limits': FLAT_LIMITS}


def slippery_scene():
    """robot_model/scene_slippery.xml:39-40: two 1 m wide strips, 1 cm above the floor, whose priority 2 makes their own friction
    triple win over every robot geom (the feet carry priority 0 or 1)."""
    return {'type': 'boxes', 'box_pos': np.array([[18.0, 0.0, -0.19], [2.0, 0.0, -0.19]]), 'box_quat': np.tile([1.0, 0, 0, 0], (2, 1)),
            'box_half': np.array([[13.0, 0.5, 0.2], [3.0, 0.5, 0.2]]), 'box_friction': np.array([[0.03, 0.05, 0.07], [0.8, 0.2, 0.3]]),
            'box_priority': 2, 'terrain_limits': FLAT_LIMITS}


def world_of_pyramid(hip_height: float, seed: int = 10):
    """terrain.py:336-344 -> add_world_of_pyramid(:240-292).  Draw order: stair_nums (an argument, drawn by the caller), then
    height_rand, stride_rand."""
    rng = np.random.RandomState(seed)
    stair_nums = rng.uniform(2, 8, 1)
    init_pos = [3.0, 0.0, 0.02]
    width, max_height, length = 10 * hip_height, 5 * hip_height, 10 * hip_height
    local_z = -0.05
    height_rand = rng.uniform(0.08, max_height, 1)
    stride_rand = rng.uniform(0.5, 1.0, 1)
    pos, half = [], []
    max_abs_x = max_abs_y = 0.0
    center = (0.0, 0.0)
    for i in range(int(stair_nums[0])):
        local_z += height_rand[0]
        new_width, new_length = width - stride_rand[0] * i, length - stride_rand[0] * i
        if new_width < 0.3 or new_length < 0.3:
            break
        pos.append([0.0 + init_pos[0], 0.0 + init_pos[1], local_z])
        half.append(0.5 * np.array([new_width, new_length, height_rand[0]]))
        if i == 0:
            max_abs_x = abs(init_pos[0] + new_width / 2.0)
            max_abs_y = abs(init_pos[1] + new_length / 2.0)
            center = (init_pos[0], init_pos[1])
    if max_abs_x >= max_abs_y:
        radius = 1.5 * np.sqrt(2 * (max_abs_x - center[0]) * (max_abs_x - center[0]))
    else:
        radius = 1.5 * np.sqrt(2 * (max_abs_y - center[1]) * (max_abs_y - center[1]))
    n = len(pos)
    return {'type': 'boxes', 'box_pos': np.array(pos), 'box_quat': np.tile([1.0, 0, 0, 0], (n, 1)), 'box_half': np.array(half),
            'terrain_limits': (center[0] + radius, center[0] - radius, center[1] + radius, center[1] - radius)}


def generate_terrain(scene: str, hip_height: float, seed: int = 10) -> dict:
    if scene == 'flat':
        return {'type': 'flat', 'terrain_limits': FLAT_LIMITS}
    if scene == 'random_boxes':
        return world_of_boxes(hip_height, seed)
    if scene == 'perlin':
        return perlin_heightfield(hip_height)
    if scene == 'stairs':
        return stairs_scene()
    if scene == 'ramp':
        return ramp_scene()
    if scene == 'random_pyramids':
        return world_of_pyramid(hip_height, seed)
    if scene == 'slippery':
        return slippery_scene()
    raise ValueError(f'Invalid scene name: {scene}, available are: flat, random_boxes, random_pyramids, perlin, stairs, ramp, slippery')
