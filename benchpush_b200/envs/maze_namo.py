"""``maze-NAMO-v0`` as a batched environment on one B200.

Keeps the reference's gym-style API (``MazeNAMO``, maze_NAMO_env.py:42-536): ``reset() -> (obs, info)``,
``step(action) -> (obs, reward, terminated, truncated, info)``, the cfg dict with the same keys and the
same one-level merge, and the same ``info`` keys (``state``, ``total_work``, ``collision reward``,
``scaled collision reward``, ``dist increment reward``, ``trial_success``, ``obs``; at reset also
``box_count``, ``goal_dt``, ``m_to_pix_scale``) — every value with a leading batch dimension, as torch
tensors on the GPU.  All simulation work happens in ``libbenchpush_b200.so`` (CUDA, sm_100a).

Obstacle layouts are inputs (the reference samples them with the unseeded global ``random`` module,
maze_NAMO_env.py:275-301): pass ``layouts`` (lists of [x, y] centres) or let the env draw
``n_synthetic_layouts`` seeded ones with the same rejection sampler (benchpush_b200.layouts).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..layouts import generate_maze_layouts, maze_walls, pack_maze_layouts
from ..utils import DotDict, merge_cfg
from .ship_ice import ShipIceVecEnv, _LazyObstacles

_CFG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", "maze_namo.yaml")


class MazeNamoVecEnv(ShipIceVecEnv):
    _prefix = "maze"
    _obs_shape = _lib.MAZE_OBS_SHAPE

    def __init__(self, num_envs: int = 1, cfg=None, device: Optional[int] = None, layouts: Optional[Sequence] = None,
                 auto_reset: bool = False, first_layout: Optional[Sequence[int]] = None, synthetic_seed: int = 0,
                 n_synthetic_layouts: int = 256, start_seed: int = 0, precision: str = "fp64"):
        if not torch.cuda.is_available():
            raise _lib.BenchPushError("MazeNamoVecEnv needs a CUDA device (there is no CPU fallback)")
        self.cfg = merge_cfg(DotDict.load_from_file(_CFG_PATH), cfg)
        if self.cfg.maze_version == 1:          # maze_NAMO_env.py:67-73
            self.cfg.env = self.cfg.env1
        elif self.cfg.maze_version == 2:
            self.cfg.env = self.cfg.env2
        else:
            raise Exception("Invalid Maze Version!")
        self.low_dim_state = bool(self.cfg.low_dim_state)
        self.auto_reset = bool(auto_reset)
        if int(self.cfg.occ.local_width) * int(self.cfg.occ.m_to_pix_scale) != 192 or int(self.cfg.occ.local_height) * int(self.cfg.occ.m_to_pix_scale) != 192:
            raise NotImplementedError("the observation kernel is built for a 192 x 192 local window (occ.local_* x occ.m_to_pix_scale)")
        if tuple(self.cfg.sim.gravity) != (0, 0):
            raise NotImplementedError("non-zero gravity is not supported by the awake-set solver")
        self.num_envs = int(num_envs)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.goal = (self.cfg.env.goal_x, self.cfg.env.goal_y)
        self.max_yaw_rate_step = (np.pi / 2) / 15
        self.beta = 1.5
        self.k_increment = 150
        self.num_channels = 4
        self.maze_walls = maze_walls(self.cfg.maze_version, self.cfg.env.width, self.cfg.env.length)
        if layouts is None:
            if not self.cfg.randomize_obstacles:      # maze_NAMO_env.py:303-310
                layouts = [[[8.5, 11], [10, 9], [11.25, 11.5], [6, 10], [3.5, 8.5]]]
            else:
                layouts = generate_maze_layouts(n_synthetic_layouts, seed=synthetic_seed, maze_version=self.cfg.maze_version,
                                                num_obstacles=int(self.cfg.num_obstacles), width=float(self.cfg.env.width),
                                                length=float(self.cfg.env.length), min_dist=float(self.cfg.min_obs_dist))
        self.layouts = [list(map(list, l)) for l in layouts]
        self.env_max_trial = len(self.layouts)
        nb, cen = pack_maze_layouts(self.layouts)
        self.max_boxes = int(cen.shape[1])
        # ---- native handle
        self.precision = precision
        L = _lib.lib(precision)
        c = _lib.MazeConfig()
        L.bp_maze_default_config(C.byref(c), int(self.cfg.maze_version))
        c.n_envs = self.num_envs; c.max_boxes = self.max_boxes; c.device = self.device_index; c.auto_reset = 1 if auto_reset else 0
        c.dt = float(self.cfg.dt); c.steps = int(self.cfg.sim.steps); c.iterations = int(self.cfg.sim.iterations)
        c.damping = float(self.cfg.sim.damping); c.obstacle_density = float(self.cfg.sim.obstacle_density)
        c.obstacle_size = float(self.cfg.obstacle_size); c.target_speed = float(self.cfg.target_speed)
        c.width = float(self.cfg.env.width); c.length = float(self.cfg.env.length)
        c.goal_x = float(self.cfg.env.goal_x); c.goal_y = float(self.cfg.env.goal_y)
        c.goal_radius = float(self.cfg.goal_radius); c.min_r = float(self.cfg.robot.min_r)
        c.m_to_pix_scale = int(self.cfg.occ.m_to_pix_scale)
        c.n_walls = len(self.maze_walls)
        for i, (a, b) in enumerate(self.maze_walls):
            c.walls[4 * i], c.walls[4 * i + 1], c.walls[4 * i + 2], c.walls[4 * i + 3] = float(a[0]), float(a[1]), float(b[0]), float(b[1])
        rv = np.asarray(self.cfg.robot.vertices, dtype=np.float64)
        c.n_robot_verts = rv.shape[0]
        for i, v in enumerate(rv.reshape(-1)):
            c.robot_verts[i] = float(v)
        wheels = [np.asarray(w, dtype=np.float64) for w in self.cfg.robot.wheel_vertices]
        c.n_wheels = len(wheels)
        for w, wv in enumerate(wheels):
            c.wheel_nv[w] = wv.shape[0]
            for i, v in enumerate(wv.reshape(-1)):
                c.wheel_verts[16 * w + i] = float(v)
        self._h = C.c_void_p()
        _lib.check(L.bp_maze_create(C.byref(c), C.byref(self._h)))
        self._L = L
        nb = np.ascontiguousarray(nb, dtype=np.int32)
        cen = np.ascontiguousarray(cen, dtype=np.float64)
        _lib.check(L.bp_maze_load_layouts(self._h, len(self.layouts), nb.ctypes.data, cen.ctypes.data))
        hw = np.zeros(2, dtype=np.int32)
        _lib.check(L.bp_maze_get_maps(self._h, None, None, None, hw.ctypes.data))
        H, W = int(hw[0]), int(hw[1])
        self.wall_map = np.zeros((H, W), dtype=np.uint8)
        self.global_distance_map = np.zeros((H, W), dtype=np.float64)
        self.unnormalized_dist_map = np.zeros((H, W), dtype=np.float64)
        _lib.check(L.bp_maze_get_maps(self._h, self.wall_map.ctypes.data, self.global_distance_map.ctypes.data,
                                      self.unnormalized_dist_map.ctypes.data, None))
        # ---- buffers
        E, dev = self.num_envs, self.device
        self.obs = torch.zeros((E,) + _lib.MAZE_OBS_SHAPE, dtype=torch.uint8, device=dev)
        self._reward = torch.zeros(E, dtype=torch.float64, device=dev)
        self._terminated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._truncated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._state = torch.zeros((E, 3), dtype=torch.float64, device=dev)
        self._pose = torch.zeros((E, 3), dtype=torch.float64, device=dev)
        self._total_work = torch.zeros(E, dtype=torch.float64, device=dev)
        self._coll = torch.zeros(E, dtype=torch.float64, device=dev)
        self._scoll = torch.zeros(E, dtype=torch.float64, device=dev)
        self._dist = torch.zeros(E, dtype=torch.float64, device=dev)
        self._success = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._status = torch.zeros(E, dtype=torch.int32, device=dev)
        o = _lib.ShipIceOutputs()
        o.obs = self.obs.data_ptr(); o.reward = self._reward.data_ptr(); o.terminated = self._terminated.data_ptr()
        o.truncated = self._truncated.data_ptr(); o.state = self._state.data_ptr(); o.pose = self._pose.data_ptr()
        o.total_work = self._total_work.data_ptr(); o.collision_reward = self._coll.data_ptr()
        o.scaled_collision_reward = self._scoll.data_ptr(); o.dist_reward = self._dist.data_ptr()
        o.trial_success = self._success.data_ptr(); o.status = self._status.data_ptr()
        self._out = o
        self._init_low_dim(2 * (self.max_boxes + 1))
        self._start_poses = None
        self.random_start, self.redraw_starts = bool(self.cfg.random_start), True
        if first_layout is None:
            first_layout = np.arange(E) % self.env_max_trial
        self._next_trial = torch.as_tensor(np.asarray(first_layout, dtype=np.int32), device=dev)
        self._actions = torch.zeros(E, dtype=torch.float32, device=dev)
        if self.random_start:
            self._start_rng = np.random.default_rng(start_seed)
            self.set_start_poses(self._sample_starts())
        if not self.low_dim_state:
            self.observation_shape = _lib.MAZE_OBS_SHAPE
        self.action_low, self.action_high = -1.0, 1.0
        self.check_status = True

    def _sample_starts(self) -> np.ndarray:
        """cfg.random_start (maze_NAMO_env.py:229-238): x ~ 1 + U[0,1) (start_x_range - 1), y likewise, redrawn until no wall
        segment (radius 0.5) is closer than cfg.robot.min_obstacle_dist (`space.point_query(p, min_dist)`: strict <), heading
        3 pi / 2.  The obstacles are sampled independently of the start (:275-301).  The reference draws from the unseeded global
        `random`; here one seeded generator serves all envs (drawn before each reset and after each step, like ship-ice)."""
        E = self.num_envs
        sx, sy = float(self.cfg.start_x_range), float(self.cfg.start_y_range)
        md, wr = float(self.cfg.robot.min_obstacle_dist), 0.5
        A = np.array([w[0] for w in self.maze_walls], dtype=np.float64); B = np.array([w[1] for w in self.maze_walls], dtype=np.float64)
        out = np.zeros((E, 3)); out[:, 2] = np.pi * 3 / 2
        todo = np.arange(E)
        while todo.size:
            p = np.stack([1 + self._start_rng.random(todo.size) * (sx - 1), 1 + self._start_rng.random(todo.size) * (sy - 1)], axis=1)
            d = B - A                                                          # distance to every wall segment
            t = np.clip(((p[:, None, :] - A[None]) * d[None]).sum(2) / (d * d).sum(1)[None], 0.0, 1.0)
            dist = np.linalg.norm(p[:, None, :] - (A[None] + t[..., None] * d[None]), axis=2) - wr
            ok = ~(dist < md).any(axis=1)
            out[todo[ok], :2] = p[ok]
            todo = todo[~ok]
        return out

    def _info(self, reset: bool):
        info = {"state": self._state, "total_work": self._total_work, "obs": _LazyObstacles(self), "pose": self._pose}
        if reset:      # maze_NAMO_env.py:345-352
            info.update({"box_count": 0, "goal_dt": self.unnormalized_dist_map, "m_to_pix_scale": self.cfg.occ.m_to_pix_scale})
        else:          # :465-474
            info.update({"collision reward": self._coll, "scaled collision reward": self._scoll,
                         "dist increment reward": self._dist, "trial_success": self._success.bool()})
        return info

    def get_obstacles(self, e: int):
        verts = np.zeros((max(self.max_boxes, 1), _lib.BP_VMAX, 2), dtype=np.float64)
        nv = np.zeros(max(self.max_boxes, 1), dtype=np.int32)
        n = _lib.check(self._fn("get_obstacles")(self._h, e, verts.ctypes.data, nv.ctypes.data))
        return [verts[i, : nv[i]].copy() for i in range(n)]

    def get_bodies(self, e: int) -> np.ndarray:
        out = np.zeros((len(self.maze_walls) + self.max_boxes + 1 + len(self.cfg.robot.wheel_vertices), 6), dtype=np.float64)
        n = _lib.check(self._fn("get_bodies")(self._h, e, out.ctypes.data))
        return out[:n]
