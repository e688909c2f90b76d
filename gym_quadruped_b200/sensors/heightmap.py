"""Height-map sensor (mirror of gym_quadruped/sensors/heightmap.py:17-221).

A rows x cols grid of downward rays around the robot, laid out in the heading frame: ray origin
`(center_xy + R_z(yaw) @ offset, center_z + 0.6 - 0.07)`, row index grows toward -x, column index toward -y, even grid
sizes are re-centred by half a cell (:106-136); a miss returns the point 1 m above the ray origin (:90-104 with
`mj_ray -> -1`).  Only static terrain is hit (geom groups {0,4,5}, `flg_static=1`: floor, height field, boxes).

The reference casts rows*cols rays through pybind one by one; here one kernel launch (qs_raycast_heightmap) casts them for
every env.  Like the reference class it is used stand-alone, after `env.step` (examples/aliengo_with_heightmap.py:39-41).
"""
from __future__ import annotations

import ctypes as C

import torch


class HeightMap:
    def __init__(self, num_rows, num_cols, dist_x, dist_y, mj_model=None, mj_data=None, env=None):
        self.env = env if env is not None else mj_data  # the batched QuadrupedEnv plays the role of (mj_model, mj_data)
        if self.env is None or not hasattr(self.env, 'sim'):
            raise ValueError('HeightMap needs the QuadrupedEnv it belongs to: HeightMap(rows, cols, dx, dy, env=env)')
        self.num_rows, self.num_cols, self.dist_x, self.dist_y = int(num_rows), int(num_cols), float(dist_x), float(dist_y)
        sim = self.env.sim
        self.data = torch.zeros(sim.N, self.num_rows, self.num_cols, 1, 3, dtype=torch.float32, device=sim.device)
        self.last_time = 0.0

    @property
    def last_sim_time(self) -> float:
        return self.last_time

    @last_sim_time.setter
    def last_sim_time(self, t) -> None:
        self.last_time = t

    def update_height_map(self, center=None, yaw=None):
        """Cast the grid for every env around its current base position / heading (arguments kept for API parity; the
        kernel reads base position and yaw from the env state, which is what the reference's callers pass)."""
        sim = self.env.sim
        sim._check(sim.L.qs_raycast_heightmap(sim.h, self.num_rows, self.num_cols, C.c_double(self.dist_x), C.c_double(self.dist_y),
                                              self.data.data_ptr(), sim._stream()))
        if sim.N == 1:
            return self.data[0].detach().cpu().numpy().astype('float64')
        return self.data

    create_sensor_matrix = update_height_map

    def get_height(self, target):
        """Height of the grid point nearest to `target` (+0.02 as in heightmap.py:209-221)."""
        pts = self.data.reshape(self.data.shape[0], -1, 3)
        t = torch.as_tensor(target, dtype=torch.float32, device=pts.device).reshape(-1, 3)[:, :2]
        d = (pts[:, :, :2] - t[:, None, :]).norm(dim=2)
        idx = d.argmin(dim=1)
        z = pts[torch.arange(pts.shape[0], device=pts.device), idx, 2] + 0.02
        return float(z[0].item()) if pts.shape[0] == 1 else z
