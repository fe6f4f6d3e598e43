"""ORACLE — test infrastructure only.

CPU restatement of ``MazeNAMO`` (reference benchpush/environments/maze_NAMO/maze_NAMO_env.py)
without gymnasium / pymunk / matplotlib / skimage: same reset/step flow, same constants, same
info keys.  Physics through the Chipmunk restatement (oracle/chipmunk_oracle.cpp: static
``Segment`` walls, compound kinematic robot, (1,3) wall handler) — PARITY UNPINNED there (see that
file's header).  Rasters: ``skimage.draw.polygon`` restated (raster_oracle.py);
``scipy.ndimage.rotate(order=1)`` restated here and PINNED against the scipy in this image by
tests/test_oracle_maze_cpu.py (bit-exact on random images and angles).

The obstacle layouts are INPUTS (the reference draws them from the unseeded global ``random``
module, maze_NAMO_env.py:275-301): ``layouts`` is a list of ``[[x, y], ...]`` obstacle centres,
episode k uses ``layouts[k % len(layouts)]``.
"""
import math
import numpy as np

from . import cpo
from .raster_oracle import OccupancyGrid, draw_polygon, poly_area, poly_centroid
from .ship_ice_oracle import merge_cfg, total_work_done

BOUNDARY_PENALTY = -50      # maze_NAMO_env.py:29-31
TERMINAL_REWARD = 200

# maze_NAMO/config.yaml restated (reference defaults)
DEFAULT_CFG = {
    "dt": 0.8, "target_speed": 0.15,
    "robot": {
        "mass": 1, "padding": 0.25, "min_r": 0.86,
        "vertices": [[0.7, -0.5], [0.55, -0.6], [-0.55, -0.6], [-0.7, -0.5], [-0.7, 0.5], [-0.55, 0.6], [0.55, 0.6], [0.7, 0.5]],
        "wheel_vertices": [
            [[0.2, 0.6], [0.2, 0.7], [0.5, 0.7], [0.55, 0.6]],
            [[0.2, -0.7], [0.2, -0.6], [0.55, -0.6], [0.55, -0.7]],
            [[-0.2, 0.6], [-0.2, 0.7], [-0.55, 0.7], [-0.55, 0.6]],
            [[-0.2, -0.6], [-0.2, -0.7], [-0.55, -0.7], [-0.55, -0.6]],
        ],
        "min_obstacle_dist": 2.5,
    },
    "low_dim_state": False, "fixed_trial_idx": 1, "random_start": False, "start_x_range": 20, "start_y_range": 20,
    "sim": {"steps": 400, "gravity": (0, 0), "iterations": 10, "damping": 0, "obstacle_density": 0.001},
    "occ": {"m_to_pix_scale": 16, "local_width": 12, "local_height": 12},
    "log_obs": False, "render_snapshot": False,
    "randomize_obstacles": True, "min_obs_dist": 0.62, "goal_radius": 1,
    "env1": {"length": 15, "width": 15, "goal_x": 3.75, "goal_y": 3.75},
    "env2": {"length": 20, "width": 20, "goal_x": 3.33, "goal_y": 3.33},
    "output_dir": "logs/", "num_obstacles": 5, "obstacle_size": 0.5, "render_scale": 80, "maze_version": 1,
}


# ---------------------------------------------------------------- scipy.special.sindg / cosdg (cephes sindg.c), restated
_SINCOF = [1.58962301572218447952E-10, -2.50507477628503540135E-8, 2.75573136213856773549E-6,
           -1.98412698295895384658E-4, 8.33333333332211858862E-3, -1.66666666666666307295E-1]
_COSCOF = [1.13678171382044553091E-11, -2.08758833757683644217E-9, 2.75573155429816611547E-7,
           -2.48015872936186303776E-5, 1.38888888888806666760E-3, -4.16666666666666348141E-2, 4.99999999999999999798E-1]
_PI180 = 1.74532925199432957692E-2


def _polevl(x, c):
    a = c[0]
    for k in c[1:]:
        a = a * x + k
    return a


def _octant(x):
    y = math.floor(x / 45.0)
    z = math.ldexp(y, -4)
    z = math.floor(z)
    z = y - math.ldexp(z, 4)
    j = int(z)
    if j & 1:
        j += 1
        y += 1.0
    return j & 7, y


def sindg(x):
    sign = 1
    if x < 0:
        x = -x
        sign = -1
    j, y = _octant(x)
    if j > 3:
        sign = -sign
        j -= 4
    z = (x - y * 45.0) * _PI180
    zz = z * z
    if j == 1 or j == 2:
        y = 1.0 - zz * _polevl(zz, _COSCOF)
    else:
        y = z + z * (zz * _polevl(zz, _SINCOF))
    return -y if sign < 0 else y


def cosdg(x):
    sign = 1
    if x < 0:
        x = -x
    j, y = _octant(x)
    if j > 3:
        j -= 4
        sign = -sign
    if j > 1:
        sign = -sign
    z = (x - y * 45.0) * _PI180
    zz = z * z
    if j == 1 or j == 2:
        y = z + z * (zz * _polevl(zz, _SINCOF))
    else:
        y = 1.0 - zz * _polevl(zz, _COSCOF)
    return -y if sign < 0 else y


def rotate_bilinear(inp, angle_deg, cval):
    """scipy.ndimage.rotate(inp, angle, reshape=False, order=1, mode='constant', cval=cval), restated.

    Arithmetic order of scipy 1.1x (ndimage/_interpolation.py::rotate -> affine_transform ->
    ni_interpolation.c::NI_GeometricTransform), established by bit-exact comparison in this image:
      c, s = cosdg(angle), sindg(angle);  M = [[c, s], [-s, c]];  offset = centre - M @ centre
      (numpy evaluates the 2x2 @ 2 product as fma(M[h,0], v0, M[h,1] * v1));
      coordinate h of output pixel (i, j) = (offset[h] + i * M[h,0]) + j * M[h,1];
      outside [0, len - 1] on either axis -> cval;
      weights per axis: w0 = 1 - (x - floor(x)), w1 = 1 - w0;
      value = sum over (a, b) in (0,0),(0,1),(1,0),(1,1) of (v[a,b] * w0[a]) * w1[b], accumulated from 0.
    """
    inp = np.asarray(inp, dtype=np.float64)
    H, W = inp.shape
    c, s = cosdg(float(angle_deg)), sindg(float(angle_deg))
    R = np.array([[c, s], [-s, c]])
    cen = (np.array([H, W]) - 1) / 2
    off = cen - R @ cen
    I = np.arange(H, dtype=np.float64)[:, None]
    J = np.arange(W, dtype=np.float64)[None, :]
    c0 = (off[0] + I * R[0, 0]) + J * R[0, 1]
    c1 = (off[1] + I * R[1, 0]) + J * R[1, 1]
    outside = (c0 < 0) | (c0 > H - 1) | (c1 < 0) | (c1 > W - 1)
    c0 = np.where(outside, 0.0, c0)
    c1 = np.where(outside, 0.0, c1)
    f0 = np.floor(c0)
    f1 = np.floor(c1)
    a0 = 1.0 - (c0 - f0)
    b0 = 1.0 - (c1 - f1)
    a1 = 1.0 - a0
    b1 = 1.0 - b0
    i0 = f0.astype(np.int64)
    j0 = f1.astype(np.int64)
    i1 = i0 + 1
    j1 = j0 + 1
    i1 = np.where(i1 >= H, 2 * H - 2 - i1, i1)      # mirror border of the spline support (weight is 0 there)
    j1 = np.where(j1 >= W, 2 * W - 2 - j1, j1)
    t = 0.0 + (inp[i0, j0] * a0) * b0
    t = t + (inp[i0, j1] * a0) * b1
    t = t + (inp[i1, j0] * a1) * b0
    t = t + (inp[i1, j1] * a1) * b1
    return np.where(outside, float(cval), t)


def _trig(heading):
    """(cos, sin) of the heading: libm (what the reference's math.cos/sin call) or the portable routine."""
    if cpo.portable_trig():
        s, c = cpo.sincos(heading)
        return c, s
    return math.cos(heading), math.sin(heading)


class MazeOccupancyGrid(OccupancyGrid):
    """The maze methods of occupancy_map.py (:67-94, :142-202, :338-376, :435-485)."""

    # occupancy_map.py:67-94
    def compute_occ_img_walls(self, walls, width, height, wall_radius=0.5):
        scale = height / self.map_height
        raw = np.zeros((height, width))
        for wall in walls:
            d = np.array(wall[1]) - np.array(wall[0])
            length = np.linalg.norm(d)
            u = d / length
            p = np.array([-u[1], u[0]])
            vertices = [wall[0] + wall_radius * p - wall_radius * u,
                        wall[0] - wall_radius * p - wall_radius * u,
                        wall[1] - wall_radius * p + wall_radius * u,
                        wall[1] + wall_radius * p + wall_radius * u]
            vertices = np.array(vertices) * scale
            rr, cc = draw_polygon(vertices[:, 1], vertices[:, 0], shape=raw.shape)
            if len(rr) == 0 or len(cc) == 0:
                continue
            raw[rr, cc] = 1.0
        return raw

    # occupancy_map.py:338-376
    def _compute_global_footprint_maze(self, ship_state, ship_vertices):
        g = np.zeros((self.occ_map_height, self.occ_map_width))
        sx = self.occ_map_width / self.map_width
        sy = self.occ_map_height / self.map_height
        position = ship_state[:2]
        c, s = _trig(ship_state[2])
        R = np.asarray([[c, -s], [s, c]])
        vertices = np.asarray(ship_vertices) @ R.T + np.asarray(position)
        r, cc_ = [], []
        for x, y in vertices:
            gx = x * sx
            gy = y * sy
            if gy < 0 or gy >= self.occ_map_height or gx < 0 or gx >= self.occ_map_width:
                continue
            r.append(gy)
            cc_.append(gx)
        if len(r) == 0:
            return None
        rr, cc = draw_polygon(r, cc_)
        # a vertex kept inside the grid can still produce pixels on row/col == size via ceil(): numpy would raise
        g[rr, cc] = 1.0
        return g

    # occupancy_map.py:435-485 (FIFO wavefront; the hop count does not depend on the neighbour order)
    def global_goal_point_dist_transform(self, goal, walls, wall_radius=0.5):
        g2m = self.map_width / self.occ_map_width
        gx = int(goal[0] / g2m)
        gy = int(goal[1] / g2m)
        H, W = self.occ_map_height, self.occ_map_width
        edt = np.zeros((H, W))
        edt[gy, gx] = 1.0
        wall = self.compute_occ_img_walls(walls, W, H, wall_radius=wall_radius)
        visited = np.zeros((H, W), dtype=bool)
        visited[gy, gx] = True
        frontier = [(gy, gx)]
        neighbours = [(0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (1, -1), (-1, 1), (-1, -1)]
        while frontier:
            nxt = []
            for y, x in frontier:
                for dy, dx in neighbours:
                    ny, nx = y + dy, x + dx
                    if ny < 0 or ny >= H or nx < 0 or nx >= W:
                        continue
                    if visited[ny, nx]:
                        continue
                    if wall[ny, nx] == 1.0:
                        continue
                    visited[ny, nx] = True
                    edt[ny, nx] = edt[y, x] + 1
                    nxt.append((ny, nx))
            frontier = nxt
        norm = edt / edt.max()
        norm[wall == 1.0] = 1.0
        return norm, edt

    # occupancy_map.py:142-202
    def ego_view_map_maze(self, agent_state, agent_vertices, obstacles, maze_walls, global_dist_map, wall_map=None):
        s = self.meter_to_pixel_scale
        gobs = self.compute_occ_img(obstacles, ice_binary_w=int(self.map_width * s), ice_binary_h=int(self.map_height * s))
        gwall = wall_map if wall_map is not None else self.compute_occ_img_walls(maze_walls, self.occ_map_width, self.occ_map_height)
        gfoot = self._compute_global_footprint_maze(agent_state, agent_vertices)
        robot_x, robot_y = agent_state[:2]
        wx, wy = int(robot_x * s), int(robot_y * s)
        inflation = max(self.local_window_width, self.local_window_height) // 2
        Hh = self.local_window_height + inflation
        Ww = self.local_window_width + inflation
        gi = (np.arange(Hh) + wy - (Hh / 2)).astype(np.int64)
        gj = (np.arange(Ww) + wx - (Ww / 2)).astype(np.int64)
        vi = (gi >= 0) & (gi < gobs.shape[0])
        vj = (gj >= 0) & (gj < gobs.shape[1])

        def crop(m, fill):
            out = np.full((Hh, Ww), fill, dtype=np.float64)
            out[np.ix_(vi, vj)] = m[np.ix_(gi[vi], gj[vj])]
            return out
        lf, lo, lw, ld = crop(gfoot, 0.0), crop(gobs, 0.0), crop(gwall, 0.0), crop(global_dist_map, 1.0)
        corrected = agent_state[2] - np.pi / 2
        deg = corrected * (180 / np.pi)
        lf = rotate_bilinear(lf, deg, 0)
        lo = rotate_bilinear(lo, deg, 0)
        lw = rotate_bilinear(lw, deg, 0)
        ld = rotate_bilinear(ld, deg, 1)
        h = inflation // 2
        sl = (slice(h, h + self.local_window_height), slice(h, h + self.local_window_width))
        return lf[sl], lo[sl], lw[sl], ld[sl]


def R(theta):
    return np.asarray([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])


class MazeNamoOracleEnv:
    def __init__(self, cfg=None, layouts=None, order_mode=0):
        self.cfg = merge_cfg(DEFAULT_CFG, cfg)
        if self.cfg["maze_version"] == 1:
            self.cfg["env"] = self.cfg["env1"]
        elif self.cfg["maze_version"] == 2:
            self.cfg["env"] = self.cfg["env2"]
        else:
            raise Exception("Invalid Maze Version!")
        occ = self.cfg["occ"]
        grid = 1 / occ["m_to_pix_scale"]
        self.occupancy = MazeOccupancyGrid(grid_width=grid, grid_height=grid, map_width=self.cfg["env"]["width"],
                                           map_height=self.cfg["env"]["length"], local_width=occ["local_width"],
                                           local_height=occ["local_height"], meter_to_pixel_scale=occ["m_to_pix_scale"])
        self.beta = 1.5
        self.k_increment = 150
        self.episode_idx = None
        self.prev_dist_value = None
        self.max_yaw_rate_step = (np.pi / 2) / 15
        assert layouts is not None, "obstacle layouts are inputs to the oracle"
        self.layouts = layouts
        self.order_mode = order_mode
        import random as _random
        self.rng = _random
        self.start_override = None
        self._maps = None

    def construct_maze_walls(self):
        L = self.cfg["env"]["length"]
        W = self.cfg["env"]["width"]
        if self.cfg["maze_version"] == 1:      # maze_NAMO_env.py:365-370
            self.maze_walls = [[(0, 0), (W, 0)], [(0, 0), (0, L)], [(W, 0), (W, L)], [(0, L), (W, L)],
                               [(2 * W / 2, L), (2 * W / 2, 5)], [(W / 2, 0), (W / 2, L - L / 3)]]
        else:                                   # :371-375
            self.maze_walls = [[(0, 0), (W, 0)], [(0, 0), (0, L)], [(W, 0), (W, L)], [(0, L), (W, L)],
                               [(W / 3, 0), (W / 3, 2 * L / 3)], [(2 * W / 3, L), (2 * W / 3, L / 3)]]

    def init_sim(self):
        self.steps = self.cfg["sim"]["steps"]
        self.dt = self.cfg["dt"]
        self.target_speed = self.cfg["target_speed"]
        self.space = cpo.Space(iterations=self.cfg["sim"]["iterations"], damping=self.cfg["sim"]["damping"],
                               gravity=self.cfg["sim"]["gravity"])
        self.space.set_order(self.order_mode)
        self.total_work = [0, []]

    def init_env(self, settle=True):
        self.construct_maze_walls()
        # sim_utils.py:174-181 generate_sim_maze: one static body + Segment(r=.5) per wall, type 3, e=.5, mu=.5
        for a, b in self.maze_walls:
            body = self.space.add_body(cpo.STATIC, 0.0, 0.0, 0.0)
            seg = self.space.add_segment(body, a, b, 0.5)
            self.space.set_material(seg, 0.5, 0.5, 3)
        if self.cfg["random_start"]:
            # maze_NAMO_env.py:229-238: redraw until no wall is closer than min_obstacle_dist (point_query, strict <); the
            # reference draws from the global `random` module (self.rng); tests replay the start the CUDA path drew
            while True:
                x_start = 1 + self.rng.random() * (self.cfg["start_x_range"] - 1)
                y_start = 1 + self.rng.random() * (self.cfg["start_y_range"] - 1)
                if not self.space.point_query_any((x_start, y_start), self.cfg["robot"]["min_obstacle_dist"]):
                    break
            self.start = (x_start, y_start, np.pi * 3 / 2)
            if self.start_override is not None:
                self.start = tuple(self.start_override)
        elif self.cfg["maze_version"] == 1:
            self.start = (11.25, 3.75, np.pi / 2)
        else:
            self.start = (16.66, 16.66, 3 * np.pi / 2)
        centres = self.layouts[self.episode_idx % len(self.layouts)]
        size = self.cfg["obstacle_size"]
        obs_dicts = []
        for ox, oy in centres:
            obs_dicts.append({"centre": np.array([ox, oy]),
                              "vertices": np.array([[ox + size, oy + size], [ox - size, oy + size],
                                                    [ox - size, oy - size], [ox + size, oy - size]])})
        obs_dicts = [ob for ob in obs_dicts if poly_area(ob["vertices"]) != 0]
        self.obs_dicts = obs_dicts
        self.obstacles = [ob["vertices"] for ob in obs_dicts]
        self.goal = (self.cfg["env"]["goal_x"], self.cfg["env"]["goal_y"])
        # generate_sim_obs -> create_polygon (sim_utils.py:136-163)
        self.poly_shapes, self.poly_bodies = [], []
        for ob in obs_dicts:
            centre = np.array(ob["centre"], dtype=np.float64)
            verts = np.asarray(ob["vertices"], dtype=np.float64) - centre
            b = self.space.add_body(cpo.DYNAMIC, centre[0], centre[1], 0.0)
            cog = cpo.poly_centroid_hull(verts)
            s = self.space.add_poly(b, verts - cog, 0.02)
            self.space.set_density(s, self.cfg["sim"]["obstacle_density"])
            self.space.set_material(s, 0.01, 1.0, 2)
            self.poly_shapes.append(s)
            self.poly_bodies.append(b)
        # Robot.sim (robot.py:77-118): kinematic body, chassis Poly(r=.02, e=.01, friction default 0) type 1,
        # four wheel Polys (type 0)
        x, y, th = self.start
        self.robot_body = self.space.add_body(cpo.KINEMATIC, x, y, th)
        self.robot_shape = self.space.add_poly(self.robot_body, np.asarray(self.cfg["robot"]["vertices"], dtype=np.float64), 0.02)
        self.space.set_mass(self.robot_shape, 1.0)
        self.space.set_material(self.robot_shape, 0.01, 0.0, 1)
        self.wheels = []
        for wv in self.cfg["robot"]["wheel_vertices"]:
            w = self.space.add_poly(self.robot_body, np.asarray(wv, dtype=np.float64), 0.02)
            self.space.set_mass(w, 1.0)
            self.space.set_material(w, 0.01, 0.0, 0)
            self.wheels.append(w)
        self.local_verts = [self.space.get_vertices(s) for s in self.poly_shapes]
        if self._maps is None:
            self._maps = self.occupancy.global_goal_point_dist_transform(self.goal, self.maze_walls)
            self._wall_map = self.occupancy.compute_occ_img_walls(self.maze_walls, self.occupancy.occ_map_width, self.occupancy.occ_map_height)
        self.global_distance_map, self.unnormalized_dist_map = self._maps
        if settle:
            for _ in range(1000):
                self.space.step(self.dt / self.steps)
        self.prev_obs = self.get_obs_from_poly()

    def get_obs_from_poly(self):
        st = self.space.state()
        out = []
        for k, b in enumerate(self.poly_bodies):
            x, y, a = st[b, 0], st[b, 1], st[b, 2]
            if cpo.portable_trig():
                sn, cs = cpo.sincos(a)
                Rm = np.asarray([[cs, -sn], [sn, cs]])
            else:
                Rm = R(a)
            out.append(self.local_verts[k] @ Rm.T + np.array([x, y]))
        return out

    def robot_pose(self):
        b = self.space.body(self.robot_body)
        return (b[0], b[1], b[2])

    def reset(self, seed=None, options=None, settle=True):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        self.init_sim()
        self.init_env(settle=settle)
        self.t = 0
        self.prev_dist_value = None
        # NB: wall_collision is cleared AFTER the settle substeps (maze_NAMO_env.py:340-341), so a chassis that
        # touches a wall while settling does not count; hits during later steps do
        self.wall_hits_at_reset = self.space.hits_13()
        updated = self.get_obs_from_poly()
        px, py, pa = self.robot_pose()
        info = {"state": (round(px, 2), round(py, 2), round(pa, 2)), "total_work": self.total_work[0], "obs": updated,
                "box_count": 0, "goal_dt": self.unnormalized_dist_map, "m_to_pix_scale": self.cfg["occ"]["m_to_pix_scale"]}
        if self.cfg["low_dim_state"]:           # maze_NAMO_env.py:353-354
            return self.generate_observation_low_dim(updated), info
        return self.generate_observation(), info

    @property
    def wall_collision(self):
        return self.space.hits_13() > self.wall_hits_at_reset

    def get_distance_value(self):
        px, py, _ = self.robot_pose()
        s = self.cfg["occ"]["m_to_pix_scale"]
        return self.global_distance_map[int(py * s), int(px * s)]

    def goal_is_reached(self):
        px, py, _ = self.robot_pose()
        d = ((px - self.goal[0]) ** 2 + (py - self.goal[1]) ** 2) ** 0.5
        return d <= self.cfg["goal_radius"] + self.cfg["robot"]["min_r"]

    def step(self, action, want_obs=True):
        self.t += 1
        px, py, pa = self.robot_pose()
        if cpo.portable_trig():
            sn, cs = cpo.sincos(pa)
            gv = [cs * self.target_speed + (-sn) * 0.0, sn * self.target_speed + cs * 0.0]
        else:
            gv = R(pa) @ [self.target_speed, 0]
        action = action * self.max_yaw_rate_step
        self.space.set_angular_velocity(self.robot_body, action)
        self.space.set_velocity(self.robot_body, gv[0], gv[1])
        flags = self.space.maze_substeps(self.robot_body, self.steps, self.dt / self.steps, self.cfg["env"]["width"])
        boundary_violated = bool(flags & 2)
        updated = self.get_obs_from_poly()
        work = total_work_done(self.prev_obs, updated)
        self.total_work[0] += work
        self.total_work[1].append(work)
        self.prev_obs = updated
        self.obstacles = updated
        wall = self.wall_collision
        terminated = bool(self.goal_is_reached() or wall)
        px, py, pa = self.robot_pose()
        if px != self.goal[0] or py != self.goal[1]:
            dist_value = self.get_distance_value()
            if self.prev_dist_value is None:
                dist_increment_reward = 0
            else:
                dist_increment_reward = (self.prev_dist_value - dist_value) * self.k_increment
            self.prev_dist_value = dist_value
        else:
            dist_increment_reward = 0
        collision_reward = -work
        reward = self.beta * collision_reward + dist_increment_reward
        if boundary_violated or wall:
            reward += BOUNDARY_PENALTY
        trial_success = False
        if terminated and not wall:
            reward += TERMINAL_REWARD
            trial_success = True
        info = {"state": (round(px, 2), round(py, 2), round(pa, 2)), "total_work": self.total_work[0],
                "collision reward": collision_reward, "scaled collision reward": collision_reward * self.beta,
                "dist increment reward": dist_increment_reward, "trial_success": trial_success, "obs": updated}
        if self.cfg["low_dim_state"]:           # maze_NAMO_env.py:477-478
            observation = self.generate_observation_low_dim(updated)
        else:
            observation = self.generate_observation() if want_obs else None
        return observation, reward, terminated, False, info

    # maze_NAMO_env.py:488-504.  The loop starts at 1 (as in the reference): obstacle 0 is never written and the
    # last pair stays zero.
    def generate_observation_low_dim(self, updated_obstacles):
        observation = np.zeros(((len(updated_obstacles) + 1) * 2))
        px, py, _ = self.robot_pose()
        observation[0] = px
        observation[1] = py
        for i in range(1, len(updated_obstacles)):
            center = np.abs(poly_centroid(updated_obstacles[i]))
            observation[i * 2] = center[0]
            observation[i * 2 + 1] = center[1]
        return observation

    # maze_NAMO_env.py:514-525
    def generate_observation(self):
        pose = self.robot_pose()
        lf, lo, lw, ld = self.occupancy.ego_view_map_maze(pose, self.cfg["robot"]["vertices"], self.obstacles, self.maze_walls,
                                                           self.global_distance_map, wall_map=self._wall_map)
        observation = np.concatenate((np.array([lf]), np.array([lo]), np.array([lw]), np.array([ld])))
        return (observation * 255).astype(np.uint8)


def generate_obstacle_centres(rng, space, num_obstacles, length, min_dist):
    """maze_NAMO_env.py:275-301 with ``rng`` (a ``random.Random``) in place of the global ``random`` module
    and ``space`` = the Space holding only the maze walls.  Quirk preserved: the wall test sits inside the
    loop over previous obstacles, so the FIRST obstacle is never tested against the walls."""
    obstacles = []
    while len(obstacles) < num_obstacles:
        cx = rng.random() * (length - 0) + 0
        cy = rng.random() * (length - 0) + 0
        overlapped = False
        for px, py in obstacles:
            if ((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 <= min_dist:
                overlapped = True
                break
            if space.point_query_any((cx, cy), min_dist):
                overlapped = True
                break
        if not overlapped:
            obstacles.append([cx, cy])
    return obstacles


class MazeNamoMetricOracle:
    """common/metrics/maze_namo_metric.py:5-75 (+ the list fields of base_metric.py)."""

    def __init__(self, alg_name, robot_mass):
        self.rewards, self.efficiency_scores, self.effort_scores, self.success_rates = [], [], [], []
        self.alg_name = alg_name
        self.eps_reward = 0
        self.total_mass_dist = 0
        self.robot_mass = robot_mass
        self.total_robot_dist = 0

    def compute_efficiency_score(self):
        if not self.trial_success:
            return 0
        return self.L / self.total_robot_dist

    def compute_effort_score(self):
        return (self.robot_mass * self.total_robot_dist) / (self.robot_mass * self.total_robot_dist + self.total_mass_dist)

    def update(self, info, reward, eps_complete=False):
        self.eps_reward += reward
        self.total_mass_dist = info["total_work"]
        self.trial_success = info["trial_success"]
        robot_state = info["state"]
        self.total_robot_dist += np.linalg.norm(np.array(self.robot_state[:2]) - np.array(robot_state[:2]))
        self.robot_state = robot_state
        if eps_complete:
            self.rewards.append(self.eps_reward)
            self.efficiency_scores.append(self.compute_efficiency_score())
            self.effort_scores.append(self.compute_effort_score())

    def reset(self, info):
        self.eps_reward = 0
        self.total_mass_dist = 0
        self.total_robot_dist = 0
        self.trial_success = False
        self.robot_state = info["state"]
        goal_dt = info["goal_dt"]
        s = info["m_to_pix_scale"]
        self.L = goal_dt[int(self.robot_state[1] * s), int(self.robot_state[0] * s)] / s
