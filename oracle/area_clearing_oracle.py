"""ORACLE — test infrastructure only.

CPU restatement of ``AreaClearingEnv`` (reference benchpush/environments/area_clearing/area_clearing.py) for the four
environments of ``area_clearing/envs/*.yaml`` (``clear_env``, ``clear_env_small``, ``walled_env``,
``walled_env_with_columns``: inner walls and static obstacles are fixed per environment), on top of the
box-delivery restatement (same position-control chain, same collision handlers, same local-map machinery):

* physics: oracle/chipmunk_oracle.cpp, handler_mode 1 — PARITY UNPINNED.  Shapes are created in the order robot,
  boxes, walls (the reference adds walls, boxes, robot: area_clearing.py:361-401); the pair order is the oracle's
  definition anyway (chipmunk_oracle.cpp header), and this order keeps ``shapes[0]`` of the (1,3) / (2,3) handlers the
  first shape of the pair.
* shapely is absent: ``Polygon.intersects`` between the (convex) clearance polygon and a box is restated as a
  separating-axis test with closed-set semantics (touching counts); ``Polygon.contains(Point)`` as a strict
  point-in-convex-polygon test; the goal points of ``_compute_boundary_goals`` (:225-263) for a boundary without walls
  as ``LineString.interpolate`` on each edge, p0 + (d / length) (p1 - p0).  Unpinned.
* layouts are INPUTS (the reference draws them from the unseeded global ``random``, :363-365, :515-541):
  ``sample_layout(rng)`` restates the sampler for a ``random.Random``.
"""
import numpy as np
import cv2
from scipy.ndimage import distance_transform_edt, binary_dilation

from . import cpo
from .ship_ice_oracle import merge_cfg, total_work_done
from .raster_oracle import poly_centroid
from .box_delivery_oracle import (BoxDeliveryOracleEnv, disk, distance, restrict_heading_range, heading_difference,
                                  NONMOVEMENT_DIST_THRESHOLD, NONMOVEMENT_TURN_THRESHOLD)

BOUNDARY_PENALTY = -0.25            # area_clearing.py:41-52
BOX_PUTBACK_PENALTY = -10
TRUNCATION_PENALTY = 0
TERMINAL_REWARD = 50
BOX_CLEARED_REWARD = 10
BOX_PUSHING_REWARD_MULTIPLIER = 0.2
NONMOVEMENT_PENALTY = 0
TIME_PENALTY = 0
DISTANCE_SCALE_MAX = 0.5
FLOOR_SEG_INDEX, GOAL_AREA_SEG_INDEX, COMPLETED_CUBE_SEG_INDEX, CUBE_SEG_INDEX, ROBOT_SEG_INDEX, MAX_SEG_INDEX = 1, 3, 7, 4, 5, 8
STEP_LIMIT = 5000

def polygon_from_line(line, width=0.2):
    """``create_polygon_from_line`` (geometry/polygon.py:201-207): shapely ``LineString(line).buffer(width / 2, cap_style=2,
    join_style=2)`` of a two-point line = the rectangle of that width around it, flat caps (restated; shapely is absent)."""
    p0, p1 = np.asarray(line[0], dtype=np.float64), np.asarray(line[1], dtype=np.float64)
    d = p1 - p0
    n = np.array([-d[1], d[0]]) / np.hypot(d[0], d[1]) * (width / 2)
    return np.array([p0 + n, p1 + n, p1 - n, p0 - n])


ENVS = {   # area_clearing/envs/*.yaml
    "clear_env": {"outer_boundary": [[-8.0, -8.0], [-8.0, 8.0], [8.0, 8.0], [8.0, -8.0]],
                  "boundary": [[-5.0, -5.0], [-5.0, 5.0], [5.0, 5.0], [5.0, -5.0]],
                  "local_map_pixel_width": 224, "sam_local_map_pixel_width": 96, "local_map_width": 24},
    "clear_env_small": {"outer_boundary": [[-6.0, -6.0], [-6.0, 6.0], [6.0, 6.0], [6.0, -6.0]],
                        "boundary": [[-4.0, -4.0], [-4.0, 4.0], [4.0, 4.0], [4.0, -4.0]],
                        "local_map_pixel_width": 224, "sam_local_map_pixel_width": 96, "local_map_width": 12},
    "walled_env": {"outer_boundary": [[-8.0, -8.0], [-8.0, 8.0], [8.0, 8.0], [8.0, -8.0]],
                   "boundary": [[-5.0, -5.0], [-5.0, 5.0], [5.0, 5.0], [5.0, -5.0]],
                   "walls": [[[-6.0, -6.0], [-6.0, 6.0]], [[6.0, 6.0], [6.0, -6.0]]],
                   "local_map_pixel_width": 224, "sam_local_map_pixel_width": 96, "local_map_width": 24},
    "walled_env_with_columns": {"outer_boundary": [[-12.0, -12.0], [-12.0, 12.0], [12.0, 12.0], [12.0, -12.0]],
                                "boundary": [[-7.5, -8.0], [-7.5, 8.0], [7.5, 8.0], [7.5, -8.0]],
                                "walls": [[[-8.0, -8.0], [-8.0, 8.0]], [[8.0, 8.0], [8.0, -8.0]]],
                                "static_obstacles": [[[-1.0, -6.0], [-1.0, -4.0], [1.0, -4.0], [1.0, -6.0]], [[-1.0, -1.0], [-1.0, 1.0], [1.0, 1.0], [1.0, -1.0]],
                                                     [[-1.0, 4.0], [-1.0, 6.0], [1.0, 6.0], [1.0, 4.0]]],
                                "local_map_pixel_width": 224, "sam_local_map_pixel_width": 96, "local_map_width": 24},
}
MAX_YAW_RATE_STEP = (np.pi / 2) / 15          # area_clearing.py:199

DEFAULT_CFG = {   # area_clearing/config.yaml (keys that reach the hot path)
    "env": "clear_env", "num_obstacles": 10, "obstacle_size": 0.5, "min_obs_dist": 1.5, "low_dim_state": False, "random_start": True,
    "agent": {"action_type": "heading", "mass": 1, "padding": 0.25, "length": 1.38, "width": 2.4, "movement_step_size": 1.5,
              "footprint_vertices": [[0.69, -1.2], [-0.69, -0.89], [-0.69, 0.89], [0.69, 1.2]],
              "vertices": [[0.6, -0.8], [-0.7, -0.8], [-0.7, 0.8], [0.6, 0.8]],
              "front_bumper_vertices": [[0.7, -1.2], [0.45, -0.9], [0.45, 0.9], [0.7, 1.2]],
              "wheel_vertices": [[[0.2, 0.8], [0.2, 0.9], [0.55, 0.9], [0.55, 0.8]], [[0.2, -0.9], [0.2, -0.8], [0.55, -0.8], [0.55, -0.9]],
                                 [[-0.2, 0.8], [-0.2, 0.9], [-0.55, 0.9], [-0.55, 0.8]], [[-0.2, -0.8], [-0.2, -0.9], [-0.55, -0.9], [-0.55, -0.8]]]},
    "sim": {"t_max": 200, "steps": 100, "gravity": (0, 0), "iterations": 10, "damping": 0, "obstacle_density": 0.001},
    "controller": {"dt": 0.2, "Lfc": 0.5, "target_speed": 0.2},
}


def sample_layout(rng, cfg=None):
    """Robot start + obstacle centres (area_clearing.py:363-368, 515-541); rng = random.Random(seed)."""
    cfg = merge_cfg(DEFAULT_CFG, cfg or {})
    b = ENVS[cfg["env"]]["boundary"]
    min_x, max_x = min(p[0] for p in b), max(p[0] for p in b)
    min_y, max_y = min(p[1] for p in b), max(p[1] for p in b)
    if cfg["random_start"]:
        start = ((min_x + 1) + rng.random() * ((max_x - min_x) - 2), min_y + 1.0, np.pi / 2)
    else:
        start = ((min_x + max_x) / 2, min_y + 1.0, np.pi / 2)
    lo_x, hi_x, lo_y, hi_y = min_x + 1, max_x - 1, min_y + 1, max_y - 1
    obs = []
    while len(obs) < cfg["num_obstacles"]:
        cx = rng.random() * (hi_x - lo_x) + lo_x
        cy = rng.random() * (hi_y - lo_y) + lo_y
        if all(((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 > cfg["min_obs_dist"] for px, py in obs):
            obs.append([cx, cy])
    return {"start": tuple(float(v) for v in start), "boxes": [[float(a), float(b_)] for a, b_ in obs]}


def convex_intersects(P, Q):
    """shapely Polygon(P).intersects(Polygon(Q)) for convex polygons: closed sets, touching counts (SAT)."""
    for A, B in ((P, Q), (Q, P)):
        n = len(A)
        for i in range(n):
            ex, ey = A[(i + 1) % n][0] - A[i][0], A[(i + 1) % n][1] - A[i][1]
            nx, ny = ey, -ex
            a = [nx * p[0] + ny * p[1] for p in A]
            b = [nx * p[0] + ny * p[1] for p in B]
            if max(a) < min(b) or max(b) < min(a):
                return False
    return True


class AreaClearingOracleEnv(BoxDeliveryOracleEnv):
    def __init__(self, cfg=None, layouts=None, order_mode=1):
        self.cfg = merge_cfg(DEFAULT_CFG, cfg or {})
        c = self.cfg
        self.env_cfg = ENVS[c["env"]]
        self.order_mode = order_mode
        self.layouts = layouts
        self.episode_idx = None
        self.path = None
        self.dp = None
        self.local_map_pixel_width = self.env_cfg["local_map_pixel_width"]
        self._configure_state_and_observation_channels()
        self.target_speed = c["controller"]["target_speed"]
        self.outer = self.env_cfg["outer_boundary"]
        self.boundary_vertices = self.env_cfg["boundary"]
        xs, ys = [p[0] for p in self.boundary_vertices], [p[1] for p in self.boundary_vertices]
        self.min_x_boundary, self.max_x_boundary, self.min_y_boundary, self.max_y_boundary = min(xs), max(xs), min(ys), max(ys)
        xo, yo = [p[0] for p in self.outer], [p[1] for p in self.outer]
        self.map_width, self.map_height = max(xo) - min(xo), max(yo) - min(yo)
        # the attributes the shared control chain reads
        self.room_width, self.room_length = self.map_width, self.map_height      # create_padded_room uses map_width first (area_clearing.py:956-960)
        self.bound_x_half, self.bound_y_half = self.map_height / 2, self.map_width / 2
        self.lfc = c["controller"]["Lfc"]
        self.omega_gain, self.v_gain = 0.5, 5                                    # apply_controller (:903-906): omega / 2, v * 5
        self.step_limit = STEP_LIMIT
        self.num_box = c["num_obstacles"]
        # _compute_boundary_goals without walls: 10 points per boundary edge at ((i + 1/2) / 10) of its length
        self.goal_points = []
        nb = len(self.boundary_vertices)
        for k in range(nb):
            p0, p1 = np.array(self.boundary_vertices[k], dtype=float), np.array(self.boundary_vertices[(k + 1) % nb], dtype=float)
            length = float(np.hypot(*(p1 - p0)))
            for i in range(10):
                d = ((i + 1 / 2) / 10) * length
                self.goal_points.append(tuple(p0 + (d / length) * (p1 - p0)))
        self.goal_point_global_map = None
        self.configuration_space = None

    def _configure_state_and_observation_channels(self):
        """area_clearing.py:168-190: everything that depends on ``local_map_pixel_width``."""
        c = self.cfg
        self.local_map_width = self.env_cfg["local_map_width"]
        self.local_map_pixels_per_meter = self.local_map_pixel_width / self.local_map_width
        ag = c["agent"]
        self.robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
        W = self.local_map_pixel_width
        rpw = int(2 * self.robot_radius * self.local_map_pixels_per_meter)
        self.robot_state_channel = np.zeros((W, W), dtype=np.float32)
        start = int(np.floor(W / 2 - rpw / 2))
        for i in range(start, start + rpw):
            for j in range(start, start + rpw):
                if (((i + 0.5) - W / 2) ** 2 + ((j + 0.5) - W / 2) ** 2) ** 0.5 < rpw / 2:
                    self.robot_state_channel[i, j] = 1
        self.goal_point_global_map = None
        self.configuration_space = None

    def configure_env_for_SAM(self):
        """area_clearing.py:221-223."""
        self.local_map_pixel_width = self.env_cfg["sam_local_map_pixel_width"]
        self._configure_state_and_observation_channels()

    def _build_space(self, layout):
        c = self.cfg
        self.steps = c["sim"]["steps"]
        self.t_max = c["sim"]["t_max"] if c["sim"]["t_max"] else np.inf
        self.dt = c["controller"]["dt"]
        sp = cpo.Space(iterations=c["sim"]["iterations"], damping=c["sim"]["damping"], gravity=c["sim"]["gravity"])
        sp.set_order(self.order_mode)
        sp.set_handler_mode(1)
        self.space = sp
        self.total_work = [0, []]
        self.start = tuple(layout["start"])
        ag = c["agent"]
        x, y, h = self.start
        rb = sp.add_body(cpo.KINEMATIC, x, y, h)
        self.robot_body = rb
        verts = np.asarray(ag["vertices"], dtype=np.float64)
        cg = cpo.poly_centroid_hull(verts)
        self.robot_shape = sp.add_poly(rb, np.array([[vx - cg[0], vy - cg[1]] for vx, vy in verts]), 0.0)
        sp.set_mass(self.robot_shape, 10.0)
        sp.set_material(self.robot_shape, 0.01, 1.0, 1)
        for wv in ag["wheel_vertices"]:
            s = sp.add_poly(rb, np.asarray(wv, dtype=np.float64), 0.02)
            sp.set_mass(s, 1)
            sp.set_material(s, 0.01, 0.0, 0)
        s = sp.add_poly(rb, np.asarray(ag["front_bumper_vertices"], dtype=np.float64), 0.02)
        sp.set_mass(s, 1)
        sp.set_material(s, 0.01, 0.0, 0)
        # boxes: generate_sim_obs -> create_polygon(vertices - centre, *centre, density): heading 0, r .02, e .01, mu 1
        sz = c["obstacle_size"]
        self.box_bodies, self.box_shapes, self.box_idx = [], [], []
        for i, (bx, by) in enumerate(layout["boxes"]):
            vertices = np.array([[bx + sz, by + sz], [bx - sz, by + sz], [bx - sz, by - sz], [bx + sz, by - sz]])
            local = (vertices - np.array([bx, by])).tolist()
            b = sp.add_body(cpo.DYNAMIC, bx, by, 0.0)
            cg = cpo.poly_centroid_hull(np.asarray(local))
            s = sp.add_poly(b, np.array([[vx - cg[0], vy - cg[1]] for vx, vy in local]), 0.02)
            sp.set_density(s, c["sim"]["obstacle_density"])
            sp.set_material(s, 0.01, 1.0, 2)
            self.box_bodies.append(b); self.box_shapes.append(s); self.box_idx.append(i)
        # static obstacles (generate_static_obstacles :437-453), inner walls as polygons around their centre lines and the
        # four outer walls (generate_walls :455-513): static body, radius .1, friction .99, default elasticity 0; all of them
        # end up with collision type 3 (:384-385)
        room_length, room_width, wt = abs(self.outer[0][0]) * 2, abs(self.outer[0][1]) * 2, 24
        self.boundary_shapes = []
        rect = lambda x, y, l, w: np.array([[x - l / 2, y - w / 2], [x + l / 2, y - w / 2], [x + l / 2, y + w / 2], [x - l / 2, y + w / 2]])
        for kind, polys in (("static", [np.asarray(o, dtype=np.float64) for o in self.env_cfg.get("static_obstacles", [])]),
                            ("wall", [polygon_from_line(w) for w in self.env_cfg.get("walls", [])])):
            for poly in polys:
                b = sp.add_body(cpo.STATIC, 0, 0, 0)
                s = sp.add_poly(b, poly, 0.1)
                sp.set_material(s, 0.0, 0.99, 3)
                self.boundary_shapes.append((b, s, kind))
        for (x, y, length, width) in [(-room_length / 2 - wt / 2, 0, wt, room_width), (room_length / 2 + wt / 2, 0, wt, room_width),
                                      (0, -room_width / 2 - wt / 2, room_length + 2 * wt, wt), (0, room_width / 2 + wt / 2, room_length + 2 * wt, wt)]:
            b = sp.add_body(cpo.STATIC, 0, 0, 0)
            s = sp.add_poly(b, rect(x, y, length, width), 0.1)
            sp.set_material(s, 0.0, 0.99, 3)
            self.boundary_shapes.append((b, s, "wall"))
        self.box_local = {s: sp.get_vertices(s) for s in self.box_shapes}
        self.robot_local = sp.get_vertices(self.robot_shape)

    @staticmethod
    def _R(theta):
        from .box_delivery_oracle import R
        return R(theta)

    # ---- maps (area_clearing.py:956-1120)
    def update_configuration_space(self):
        W, ppm = self.local_map_pixel_width, self.local_map_pixels_per_meter
        obstacle_map = self.create_padded_room(0.0)
        small = np.zeros((W + 20, W + 20), dtype=np.float32)
        for (b, s, label) in self.boundary_shapes:
            vertices = self._world_verts(b, self.space.get_vertices(s))
            cv2.fillPoly(small, [self._verts_px(vertices, small.shape[0])], color=1)
        si, sj = int(obstacle_map.shape[0] / 2 - small.shape[0] / 2), int(obstacle_map.shape[1] / 2 - small.shape[1] / 2)
        obstacle_map[si:si + small.shape[0], sj:sj + small.shape[1]] = small
        rpw = int(2 * self.robot_radius * ppm)
        selem = disk(np.floor(rpw / 4))
        self.configuration_space = 1 - binary_dilation(obstacle_map, selem).astype(np.float32)
        self.configuration_space_thin = 1 - binary_dilation(obstacle_map, selem).astype(np.float32)
        self.closest_cspace_indices = distance_transform_edt(1 - self.configuration_space, return_distances=False, return_indices=True)
        self.small_obstacle_map = 1 - small
        self.small_start = (si, sj)

    def update_global_overhead_map(self):
        ppm = self.local_map_pixels_per_meter
        small = self.small_obstacle_map.copy()
        small[small == 1] = FLOOR_SEG_INDEX / MAX_SEG_INDEX
        H = small.shape[0]
        il, iw = abs(self.boundary_vertices[0][0]) * 2, abs(self.boundary_vertices[0][1]) * 2
        rt = abs(self.outer[0][0]) - abs(self.boundary_vertices[0][0])
        for x, y, length, width in [(-il / 2 - rt / 2, 0, rt, iw), (il / 2 + rt / 2, 0, rt, iw),
                                    (0, -iw / 2 - rt / 2, il + 2 * rt, rt), (0, iw / 2 + rt / 2, il + 2 * rt, rt)]:
            poly = np.array([[x - length / 2, y - width / 2], [x + length / 2, y - width / 2], [x + length / 2, y + width / 2], [x - length / 2, y + width / 2]])
            cv2.fillPoly(small, [self._verts_px(poly, H)], color=GOAL_AREA_SEG_INDEX / MAX_SEG_INDEX)
        for i, (b, s) in enumerate(zip(self.box_bodies, self.box_shapes)):
            px = self._verts_px(self._world_verts(b, self.box_local[s]), H)
            cv2.fillPoly(small, [px], color=(COMPLETED_CUBE_SEG_INDEX if self.box_clearance_statuses[i] else CUBE_SEG_INDEX) / MAX_SEG_INDEX)
        foot = np.asarray(self.cfg["agent"]["footprint_vertices"], dtype=np.float64)
        cv2.fillPoly(small, [self._verts_px(self._world_verts(self.robot_body, foot), H)], color=ROBOT_SEG_INDEX / MAX_SEG_INDEX)
        si, sj = self.small_start
        self.global_overhead_map[si:si + H, sj:sj + small.shape[1]] = small

    def create_global_shortest_path_map(self, robot_position):
        i, j = self.position_to_pixel_indices(robot_position[0], robot_position[1], self.configuration_space.shape)
        i, j = self.closest_valid_cspace_indices(i, j)
        m, _ = cpo.spfa(self.configuration_space, (i, j))
        ppm, W = self.local_map_pixels_per_meter, self.local_map_pixel_width
        m /= ppm
        m /= (np.sqrt(2) * W) / ppm
        return m

    def create_global_shortest_path_to_goal_points(self):
        ppm, W = self.local_map_pixels_per_meter, self.local_map_pixel_width
        global_map = self.create_padded_room(0.0) + np.inf
        for (rx, ry) in self.goal_points:
            i, j = self.position_to_pixel_indices(rx, ry, self.configuration_space.shape)
            i, j = self.closest_valid_cspace_indices(i, j)
            img, _ = cpo.spfa(self.configuration_space, (i, j))
            img /= ppm
            global_map = np.minimum(global_map, img)
        global_map /= (np.sqrt(2) * W) / ppm
        max_value, min_value = np.max(global_map), np.min(global_map)
        global_map = (global_map - min_value) / (max_value - min_value) * DISTANCE_SCALE_MAX
        # points outside the boundary polygon -> 0, outside the outer boundary -> 1 (strict containment, axis-aligned squares)
        H, Wd = global_map.shape
        I, J = np.meshgrid(np.arange(H), np.arange(Wd), indexing="ij")
        X, Y = self.pixel_indices_to_position(I, J, self.configuration_space.shape)
        inside = lambda poly: ((X > min(p[0] for p in poly)) & (X < max(p[0] for p in poly)) & (Y > min(p[1] for p in poly)) & (Y < max(p[1] for p in poly)))
        global_map[~inside(self.boundary_vertices)] = 0
        global_map[~inside(self.outer)] = 1
        global_map += 1 - self.configuration_space
        return global_map

    def generate_observation(self, done=False):
        self.update_global_overhead_map()
        pos, ang = self.robot_pose()
        scale = lambda a: (a * 255).astype(np.uint8)
        ch = [scale(self.get_local_map(self.global_overhead_map, pos, ang)), scale(self.robot_state_channel.copy()),
              scale(self.get_local_distance_map(self.create_global_shortest_path_map(pos), pos, ang)),
              scale(self.get_local_distance_map(self.goal_point_global_map.copy(), pos, ang))]
        return np.stack(ch, axis=2).astype(np.uint8)

    # ---- completion / rewards (area_clearing.py:1122-1140; evaluation/metrics.py:73-94)
    def boxes_completed(self, updated):
        count = 0
        for i, ob in enumerate(updated):
            out = not convex_intersects(self.boundary_vertices, ob)
            count += out
            self.box_clearance_statuses[i] = bool(out)
        return count, count == self.num_box

    def obs_to_goal_difference(self, prev, updated):
        total = 0
        for a, b in zip(prev, updated):
            if not convex_intersects(self.boundary_vertices, a):
                continue
            ca, cb = poly_centroid(a), poly_centroid(b)
            da = min(((ca[0] - g[0]) ** 2 + (ca[1] - g[1]) ** 2) ** 0.5 for g in self.goal_points)
            db = min(((cb[0] - g[0]) ** 2 + (cb[1] - g[1]) ** 2) ** 0.5 for g in self.goal_points)
            total += da - db
        return total

    # ---- gym API
    def reset(self, seed=None, options=None):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        layout = self.layouts[self.episode_idx % len(self.layouts)]
        self._build_space(layout)
        self.box_clearance_statuses = [False for _ in self.box_bodies]
        if self.configuration_space is None:
            self.update_configuration_space()
        self.n_substeps = 0
        for _ in range(1000):
            self.space.step(self.dt / self.steps)
        self.prev_obs = self.get_obs_from_poly()
        self.dp = None
        self.global_overhead_map = self.create_padded_room(0.0)
        self.update_global_overhead_map()
        if self.goal_point_global_map is None:
            self.goal_point_global_map = self.create_global_shortest_path_to_goal_points()
        self.t = 0
        self.cleared_box_count = 0
        self.space.set_robot_hit(0)
        updated = self.get_obs_from_poly()
        b = self.space.body(self.robot_body)
        info = {"state": (round(float(b[0]), 2), round(float(b[1]), 2), round(float(b[2]), 2)), "total_work": self.total_work[0], "obs": updated,
                "box_count": 0, "boundary": self.boundary_vertices, "goal_positions": self.goal_points}
        return self.generate_observation(), info

    def step(self, action, want_obs=True):
        self.t += 1
        p_init, ang = self.robot_pose()
        h_init = restrict_heading_range(ang)
        self.dp = None
        at = self.cfg["agent"]["action_type"]
        assert at in ("position", "heading", "velocity")
        robot_distance, robot_turn_angle = 0, 0
        if at == "velocity":
            # area_clearing.py:656-663: the velocity is set ONCE (not refreshed while the heading changes); the policy's action
            # is a float32 array (action_space dtype), so the two products are float32 (NumPy >= 2 promotion)
            action = np.asarray(action, dtype=np.float32)
            self.space.set_angular_velocity(self.robot_body, float(MAX_YAW_RATE_STEP * action[1] / 2))
            scaled_vel = self.target_speed * action[0]
            gv = self._R(ang) @ np.array([scaled_vel, 0])
            self.space.set_velocity(self.robot_body, float(gv[0]), float(gv[1]))
            self.last_path = []
        else:
            if at == "heading":
                W, ppm = self.local_map_pixel_width, self.local_map_pixels_per_meter
                angle = (float(np.asarray(action).reshape(-1)[0]) + 1) * np.pi + np.pi / 2
                step_size = self.cfg["agent"]["movement_step_size"]
                x_pixel = int(W / 2 + step_size * np.cos(angle) * ppm)
                y_pixel = int(W / 2 - step_size * np.sin(angle) * ppm)
                action = y_pixel * W + x_pixel
            self.path, move_sign = self.get_waypoints_to_spatial_action(p_init, h_init, int(action))
            self.last_path = [list(p) for p in self.path]
            robot_distance, robot_turn_angle = self.execute_robot_path(p_init, h_init, move_sign)
        for _ in range(self.steps):
            self.space.step(self.dt / self.steps)
            self.n_substeps += 1
        collision_penalty = BOUNDARY_PENALTY if self.space.robot_hit() == 1 else 0
        updated = self.get_obs_from_poly()
        num_completed, all_completed = self.boxes_completed(updated)
        diff_reward = self.obs_to_goal_difference(self.prev_obs, updated)
        pushing_reward = diff_reward * BOX_PUSHING_REWARD_MULTIPLIER
        if num_completed > self.cleared_box_count:
            box_completion_reward = abs(num_completed - self.cleared_box_count) * BOX_CLEARED_REWARD
            self.t = 0
        else:
            box_completion_reward = abs(num_completed - self.cleared_box_count) * BOX_PUTBACK_PENALTY
        self.cleared_box_count = num_completed
        nonmovement_penalty = 0
        if robot_distance < NONMOVEMENT_DIST_THRESHOLD and abs(robot_turn_angle) < NONMOVEMENT_TURN_THRESHOLD:
            nonmovement_penalty = NONMOVEMENT_PENALTY
        work = total_work_done(self.prev_obs, updated)
        self.total_work[0] += work
        self.total_work[1].append(work)
        self.prev_obs = updated
        terminated = bool(all_completed)
        reward = box_completion_reward + collision_penalty + pushing_reward + nonmovement_penalty
        truncated = self.t >= self.t_max
        if truncated:
            reward += TRUNCATION_PENALTY
        elif terminated:
            reward += TERMINAL_REWARD
        b = self.space.body(self.robot_body)
        info = {"state": (round(float(b[0]), 2), round(float(b[1]), 2), round(float(b[2]), 2)), "total_work": self.total_work[0],
                "collision reward": -work, "diff_reward": diff_reward, "box_completed_reward": box_completion_reward, "obs": updated,
                "box_completed_statuses": list(self.box_clearance_statuses), "box_count": num_completed, "ministeps": robot_distance / 2.5}
        obs = self.generate_observation() if want_obs else None
        self.update_global_overhead_map()
        self.space.set_robot_hit(0)
        return obs, reward, terminated, truncated, info
