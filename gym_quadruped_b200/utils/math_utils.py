"""Small math helpers with the reference's names (gym_quadruped/utils/math_utils.py:7-60)."""
from __future__ import annotations

import numpy as np


def skew(x):
    x = np.asarray(x, dtype=float).reshape(3)
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])


def homogenous_transform(vec, X):
    vec = np.asarray(vec, dtype=float).reshape(-1)
    assert vec.shape == (3,), f'Expected 3D vector, got shape {vec.shape}'
    X = np.asarray(X, dtype=float)
    assert X.shape == (4, 4) and X[3, 3] == 1, 'Expected a homogeneous transformation matrix'
    return X[:3, :3] @ vec + X[:3, 3]


def angle_between_vectors(vector1, vector2) -> float:
    """Heading of (vector2 - vector1) in the xy plane (math_utils.py:50-51)."""
    d = np.asarray(vector2, dtype=float) - np.asarray(vector1, dtype=float)
    return float(np.arctan2(d[1], d[0]))


def _process_range(values):
    if isinstance(values, (int, float, np.number)):
        return (values, values)
    if isinstance(values, (tuple, list, np.ndarray)):
        assert len(values) == 2, f'Invalid range values, expected (min, max) got: {values}'
        return values
    return None


def quat_wxyz_to_matrix(q):
    w, x, y, z = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
