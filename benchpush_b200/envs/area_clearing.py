"""``area-clearing-v0`` as a batched environment on one B200.

Keeps the reference's gym-style API (``AreaClearingEnv``, area_clearing.py:79-778): cfg dict with the same keys and
merge, ``reset() -> (obs, info)``, ``step(action) -> (obs, reward, terminated, truncated, info)`` with the reference's
info keys (``state``, ``total_work``, ``collision reward``, ``diff_reward``, ``box_completed_reward``, ``obs``,
``box_completed_statuses``, ``box_count``, ``ministeps``) batched on the GPU.  Runs on the box-delivery kernels
(csrc/boxdel_*.cuh, ``env_kind 1``): same position-control chain and wall handlers, area-clearing constants, the fixed
``cfg.sim.steps`` substeps after the path, clearance-polygon completion, push / completion / put-back rewards and the
truncation clock that restarts whenever a box is cleared (D13).

Host work done here once per handle: configuration space, nearest-free-cell map and the goal-point distance map
(``create_global_shortest_path_to_goal_points``, area_clearing.py:1055-1082: one spfa solve per goal point — run on the
GPU through ``bp_spfa_batch`` — min-combined, normalised and masked by the two boundary squares).

Implemented: the four environments of ``area_clearing/envs/*.yaml`` (``clear_env``, ``clear_env_small``, ``walled_env``,
``walled_env_with_columns``: inner walls and static obstacles are fixed per environment, so they are static shapes and
map pixels of the handle), ``action_type`` heading / position / velocity.
Layouts are inputs or drawn by ``layouts.sample_area_clearing_layout`` (the reference uses the unseeded global ``random``).
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Sequence

import numpy as np

from .. import _lib
from ..layouts import sample_area_clearing_layout
from ..utils import DotDict, merge_cfg
from .box_delivery import LAYOUT_EPISODE_STRIDE, BoxDelConfig, BoxDeliveryVecEnv, bind, default_pool_episodes, disk

def polygon_from_line(line, width=0.2):
    """``create_polygon_from_line`` (geometry/polygon.py:201-207: shapely ``LineString.buffer(width / 2, cap_style=2, join_style=2)``)
    of a two-point wall: the rectangle of that width around the line, flat caps."""
    p0, p1 = np.asarray(line[0], dtype=np.float64), np.asarray(line[1], dtype=np.float64)
    d = p1 - p0
    n = np.array([-d[1], d[0]]) / np.hypot(d[0], d[1]) * (width / 2)
    return [list(p0 + n), list(p1 + n), list(p1 - n), list(p0 - n)]


ENVS = {   # area_clearing/envs/*.yaml restated
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

DEFAULT_CFG = {   # area_clearing/config.yaml restated (keys that reach the hot path)
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

# area_clearing.py:41-60
BOUNDARY_PENALTY, BOX_PUTBACK_PENALTY, TRUNCATION_PENALTY, TERMINAL_REWARD = -0.25, -10, 0, 50
BOX_CLEARED_REWARD, BOX_PUSHING_REWARD_MULTIPLIER, NONMOVEMENT_PENALTY, DISTANCE_SCALE_MAX = 10, 0.2, 0, 0.5
GOAL_AREA_SEG_INDEX, COMPLETED_CUBE_SEG_INDEX, CUBE_SEG_INDEX, ROBOT_SEG_INDEX = 3, 7, 4, 5
STEP_LIMIT = 5000


class ClearingMaps:
    """Static shapes and maps of one clearing environment (area_clearing.py:455-513, 956-966, 1087-1120)."""

    def __init__(self, cfg, env_cfg):
        import cv2
        from scipy.ndimage import binary_dilation, distance_transform_edt
        outer, bnd = env_cfg["outer_boundary"], env_cfg["boundary"]
        self.LW = LW = env_cfg["local_map_pixel_width"]
        self.lmw = lmw = env_cfg["local_map_width"]
        self.ppm = ppm = LW / lmw
        ag = cfg["agent"]
        self.robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
        xo, yo = [p[0] for p in outer], [p[1] for p in outer]
        self.map_width, self.map_height = max(xo) - min(xo), max(yo) - min(yo)
        # outer walls: 4 rectangles of thickness 24 on the static body, radius .1 (generate_walls :488-513)
        room_length, room_width, wt = abs(outer[0][0]) * 2, abs(outer[0][1]) * 2, 24
        rect = lambda x, y, l, w: [[x - l / 2, y - w / 2], [x + l / 2, y - w / 2], [x + l / 2, y + w / 2], [x - l / 2, y + w / 2]]
        # static obstacles (generate_static_obstacles :437-453) and inner walls (polygons around their centre lines, :455-480)
        # come first, in the reference's creation order; all statics are radius .1, friction .99, collision type 3 (:384-385)
        st = [([list(map(float, p)) for p in o], (0.0, 0.0, 0.0), 0.1, 3) for o in env_cfg.get("static_obstacles", [])]
        st += [(polygon_from_line(w), (0.0, 0.0, 0.0), 0.1, 3) for w in env_cfg.get("walls", [])]
        st += [(rect(x, y, l, w), (0.0, 0.0, 0.0), 0.1, 3) for (x, y, l, w) in
               [(-room_length / 2 - wt / 2, 0, wt, room_width), (room_length / 2 + wt / 2, 0, wt, room_width),
                (0, -room_width / 2 - wt / 2, room_length + 2 * wt, wt), (0, room_width / 2 + wt / 2, room_length + 2 * wt, wt)]]
        self.statics = st
        self.st_nv = np.array([4] * len(st), dtype=np.int32)
        self.st_verts = np.array([s[0] for s in st], dtype=np.float64)
        self.st_pose = np.array([s[1] for s in st], dtype=np.float64)
        self.st_rad = np.array([s[2] for s in st], dtype=np.float64)
        self.st_ctype = np.array([s[3] for s in st], dtype=np.int32)
        shape = (int(2 * np.ceil((self.map_width * ppm + LW * np.sqrt(2)) / 2)), int(2 * np.ceil((self.map_height * ppm + LW * np.sqrt(2)) / 2)))
        self.shape = shape
        obstacle_map = np.zeros(shape, dtype=np.float32)
        small = np.zeros((LW + 20, LW + 20), dtype=np.float32)
        off = int(lmw * ppm / 2) + 10

        def to_px(world):
            px = (np.asarray(world, dtype=np.float64) * ppm).astype(np.int32)
            px[:, 0] += off
            px[:, 1] += off
            px[:, 1] = small.shape[0] - px[:, 1]
            return px
        for verts, _, _, _ in st:
            cv2.fillPoly(small, [to_px(verts)], color=1)
        si, sj = int(shape[0] / 2 - small.shape[0] / 2), int(shape[1] / 2 - small.shape[1] / 2)
        obstacle_map[si:si + small.shape[0], sj:sj + small.shape[1]] = small
        rpw = int(2 * self.robot_radius * ppm)
        selem = disk(np.floor(rpw / 4))
        self.configuration_space = 1 - binary_dilation(obstacle_map, selem).astype(np.float32)
        self.configuration_space_thin = self.configuration_space.copy()          # both use disk(floor(rpw / 4)) (D20)
        self.closest_cspace_indices = distance_transform_edt(1 - self.configuration_space, return_distances=False, return_indices=True)
        # overhead-map codes of the static part: 0 obstacle, 1 floor, 3 goal area (update_global_overhead_map :968-996)
        codes = (1 - small).astype(np.float32)
        il, iw = abs(bnd[0][0]) * 2, abs(bnd[0][1]) * 2
        rt = abs(outer[0][0]) - abs(bnd[0][0])
        for x, y, l, w in [(-il / 2 - rt / 2, 0, rt, iw), (il / 2 + rt / 2, 0, rt, iw), (0, -iw / 2 - rt / 2, il + 2 * rt, rt), (0, iw / 2 + rt / 2, il + 2 * rt, rt)]:
            cv2.fillPoly(codes, [to_px(rect(x, y, l, w))], color=GOAL_AREA_SEG_INDEX)
        self.small_obstacle_map = codes
        W = LW
        chan = np.zeros((W, W), dtype=np.float32)
        start = int(np.floor(W / 2 - rpw / 2))
        ii = np.arange(start, start + rpw)
        I, J = np.meshgrid(ii, ii, indexing="ij")
        inside = (((I + 0.5) - W / 2) ** 2 + ((J + 0.5) - W / 2) ** 2) ** 0.5 < rpw / 2
        chan[I[inside], J[inside]] = 1
        self.robot_state_channel = chan


class AreaClearingVecEnv(BoxDeliveryVecEnv):
    def __init__(self, num_envs: int = 1, cfg=None, device: Optional[int] = None, layouts: Optional[Sequence[dict]] = None,
                 seed: int = 1, env_offset: int = 0, layout_pool_episodes: Optional[int] = None,
                 exact_spfa_parents: bool = True, sam: bool = False):
        # ``sam``: the observation of the spatial-action-map baselines — ``sam_local_map_pixel_width`` (96) instead of
        # ``local_map_pixel_width`` (224) pixels for the same local map (what ``configure_env_for_SAM`` switches to,
        # area_clearing.py:221-223)
        self._ctor = dict(num_envs=num_envs, cfg=cfg, device=device, layouts=layouts, seed=seed, env_offset=env_offset,
                          layout_pool_episodes=layout_pool_episodes, exact_spfa_parents=exact_spfa_parents)
        self.sam = bool(sam)
        self.path = None
        self.cfg = merge_cfg(DotDict(DEFAULT_CFG), cfg)
        c = self.cfg
        if c["env"] not in ENVS:
            raise KeyError("unknown area-clearing environment %r (have: %s)" % (c["env"], sorted(ENVS)))
        if c["agent"]["action_type"] not in ("heading", "position", "velocity"):
            raise NotImplementedError("unknown action_type %r" % (c["agent"]["action_type"],))
        self.env_cfg = ENVS[c["env"]]
        if self.sam:
            self.env_cfg = dict(self.env_cfg, local_map_pixel_width=self.env_cfg["sam_local_map_pixel_width"])
        self.num_envs = E = int(num_envs)
        self._open(device)
        self.num_boxes = NB = int(c["num_obstacles"])
        self.num_channels = 4
        self.room = r = ClearingMaps(c, self.env_cfg)
        self.local_map_pixel_width = LW = r.LW
        bnd = self.env_cfg["boundary"]
        # _compute_boundary_goals (:225-263) for a boundary without walls: 10 points per edge at ((i + 1/2) / 10) of its length
        self.goal_points = []
        for k in range(len(bnd)):
            p0, p1 = np.array(bnd[k], dtype=float), np.array(bnd[(k + 1) % len(bnd)], dtype=float)
            length = float(np.hypot(*(p1 - p0)))
            for i in range(10):
                self.goal_points.append(tuple(p0 + ((((i + 1 / 2) / 10) * length) / length) * (p1 - p0)))
        if layouts is None:
            # pool of F episodes per env keyed by (GLOBAL env id, episode), see BoxDeliveryVecEnv (env_offset = first env of this rank's shard)
            F = default_pool_episodes(E) if layout_pool_episodes is None else max(1, int(layout_pool_episodes))
            layouts = [sample_area_clearing_layout(random.Random(seed + int(env_offset) + e + f * LAYOUT_EPISODE_STRIDE), c, self.env_cfg)
                       for f in range(F) for e in range(E)]
        self.layouts = [{"start": l["start"], "boxes": [[b[0], b[1], 0.0] for b in l["boxes"]]} for l in layouts]
        L = self._L
        k = BoxDelConfig()
        L.bp_boxdel_default_config(C.byref(k))
        k.n_envs = E; k.num_boxes = NB; k.device = self.device_index
        k.action_type = {"position": 0, "heading": 1, "velocity": 2}[c["agent"]["action_type"]]
        self._action_dim = 2 if k.action_type == 2 else 1
        k.steps = int(c["sim"]["steps"]); k.iterations = int(c["sim"]["iterations"]); k.settle_steps = 1000
        k.local_map_pixel_width = LW
        k.ctrl_dt = float(c["controller"]["dt"]); k.target_speed = float(c["controller"]["target_speed"]); k.damping = float(c["sim"]["damping"])
        k.box_size = 2.0 * float(c["obstacle_size"]); k.box_density = float(c["sim"]["obstacle_density"])
        k.room_length = float(r.map_height); k.room_width = float(r.map_width); k.receptacle_width = 0.0
        k.local_map_width = float(r.lmw); k.step_size = float(c["agent"]["movement_step_size"])
        k.robot_length = float(c["agent"]["length"]); k.robot_width = float(c["agent"]["width"])
        k.shortest_path_channel_scale = 1.0                       # create_global_shortest_path_map (:1046-1053) has no channel scale
        for name, key in (("chassis_verts", "vertices"), ("bumper_verts", "front_bumper_vertices"), ("wheel_verts", "wheel_vertices"), ("footprint_verts", "footprint_vertices")):
            for i, v in enumerate(np.asarray(c["agent"][key], dtype=np.float64).reshape(-1)):
                getattr(k, name)[i] = float(v)
        k.n_footprint = len(c["agent"]["footprint_vertices"])
        k.env_kind = 1; k.step_limit = STEP_LIMIT; k.post_steps = int(c["sim"]["steps"]); k.t_max = int(c["sim"]["t_max"])
        k.lfc = float(c["controller"].get("Lfc", 0.0) or 0.0); k.omega_mul = 0.5; k.v_mul = 5.0
        k.bound_x = float(r.map_height); k.bound_y = float(r.map_width)          # PositionController(cfg, radius, map_width, map_height, ...): D17
        k.static_e = 0.0; k.static_u = 0.99
        k.robot_code = ROBOT_SEG_INDEX; k.box_code = CUBE_SEG_INDEX; k.box_done_code = COMPLETED_CUBE_SEG_INDEX
        k.n_boundary = len(bnd)
        for i, v in enumerate(np.asarray(bnd, dtype=np.float64).reshape(-1)):
            k.boundary[i] = float(v)
        k.n_goals = len(self.goal_points)
        for i, v in enumerate(np.asarray(self.goal_points, dtype=np.float64).reshape(-1)):
            k.goals[i] = float(v)
        k.r_boundary = BOUNDARY_PENALTY; k.r_cleared = BOX_CLEARED_REWARD; k.r_putback = BOX_PUTBACK_PENALTY
        k.r_push_mult = BOX_PUSHING_REWARD_MULTIPLIER; k.r_terminal = TERMINAL_REWARD; k.r_truncation = TRUNCATION_PENALTY; k.r_nonmove = NONMOVEMENT_PENALTY
        self.goal_point_global_map = self._goal_point_map()
        k.exact_spfa_parents = 1 if exact_spfa_parents else 0      # see BoxDeliveryVecEnv
        self._finish_init(k, r, ch3_map=self.goal_point_global_map)

    def configure_env_for_SAM(self):
        """``AreaClearingEnv.configure_env_for_SAM`` (area_clearing.py:221-223): observations, maps and the position-action
        space switch to ``sam_local_map_pixel_width``.  The reference calls it right after construction; here the native handle,
        whose rasters and room maps are sized at creation, is rebuilt with the same arguments (episode state is discarded, call
        ``reset()`` next)."""
        if self.sam:
            return
        kw = self._ctor
        self.close()
        type(self).__init__(self, sam=True, **kw)

    def update_path(self, new_path):
        """area_clearing.py:922-925: remembers the path for a renderer (no renderer here)."""
        self.path = new_path

    def _spfa_many(self, cspace_u8, sources):
        """One shortest-path field per source on the shared configuration space (bp_spfa_batch_host of the bound library)."""
        n = len(sources)
        maps = np.ascontiguousarray(np.broadcast_to(cspace_u8, (n,) + cspace_u8.shape))
        src = np.ascontiguousarray(np.asarray(sources, dtype=np.int32))
        d = np.zeros((n,) + cspace_u8.shape, dtype=np.float32)
        p = np.zeros((n,) + cspace_u8.shape, dtype=np.int32)
        self._L.bp_spfa_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        self._check(self._L.bp_spfa_batch_host(maps.ctypes.data, src.ctypes.data, n, cspace_u8.shape[0], cspace_u8.shape[1], d.ctypes.data, p.ctypes.data))
        return d

    def _goal_point_map(self):
        """create_global_shortest_path_to_goal_points (area_clearing.py:1055-1082)."""
        r = self.room
        ppm, W = r.ppm, r.LW
        cs = r.configuration_space
        H, Wd = cs.shape
        srcs = []
        for (gx, gy) in self.goal_points:
            i = int(np.clip(np.floor(H / 2 - gy * ppm).astype(np.int32), 0, H - 1))
            j = int(np.clip(np.floor(Wd / 2 + gx * ppm).astype(np.int32), 0, Wd - 1))
            srcs.append(r.closest_cspace_indices[:, i, j])
        fields = self._spfa_many(np.ascontiguousarray(cs.astype(np.uint8)), srcs)
        global_map = np.zeros(cs.shape, dtype=np.float32) + np.inf
        for img in fields:
            img /= ppm
            global_map = np.minimum(global_map, img)
        global_map /= (np.sqrt(2) * W) / ppm
        max_value, min_value = np.max(global_map), np.min(global_map)
        global_map = (global_map - min_value) / (max_value - min_value) * DISTANCE_SCALE_MAX
        I, J = np.meshgrid(np.arange(H), np.arange(Wd), indexing="ij")
        X, Y = (J - Wd / 2) / ppm, (H / 2 - I) / ppm
        strict = lambda poly: ((X > min(p[0] for p in poly)) & (X < max(p[0] for p in poly)) & (Y > min(p[1] for p in poly)) & (Y < max(p[1] for p in poly)))
        global_map[~strict(self.env_cfg["boundary"])] = 0        # shapely Polygon.contains(Point): strict interior of the axis-aligned squares
        global_map[~strict(self.env_cfg["outer_boundary"])] = 1
        global_map += 1 - cs
        return global_map

    def _info(self):
        b = self._buf
        info, ii = b["info"], b["info_i"]
        return {"state": info[:, 0:3], "total_work": info[:, 5], "collision reward": -info[:, 7], "diff_reward": info[:, 3],
                "box_completed_reward": info[:, 4], "obs": b["box_verts"], "n_boxes_left": ii[:, 2], "box_completed_statuses": b["completed"],
                "box_count": ii[:, 0], "ministeps": info[:, 6], "t": ii[:, 1], "status": ii[:, 3], "goal_positions": self.goal_points,
                "boundary": self.env_cfg["boundary"], "walls": self.env_cfg.get("walls", []), "static_obstacles": self.env_cfg.get("static_obstacles", [])}
