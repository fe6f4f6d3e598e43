"""``box-delivery-v0`` as a batched environment on one B200.

Keeps the reference's gym-style API (``BoxDeliveryEnv``, box_delivery_env.py:68-830): ``reset() -> (obs, info)``,
``step(action) -> (obs, reward, terminated, truncated, info)``, the cfg dict with the same keys and the same
one-level merge, and the same ``info`` keys — every value with a leading batch dimension, on the GPU as
torch tensors.  The physics, the position-control chain, the 2 N + 2 shortest-path solves per step and the
224 x 224 x 4 observation run in ``libbenchpush_b200.so`` (CUDA, sm_100a; csrc/boxdel_*.cuh).

Host work done here, once per handle (not on the hot path): the room's configuration spaces and the
nearest-free-cell map (``update_configuration_space``, box_delivery_env.py:1140-1175), built with the same
library calls as the reference (cv2.fillPoly, scipy binary_dilation with a disk footprint, scipy
distance_transform_edt).  They depend on the room only, so all envs of a handle share them.

Implemented: every ``obstacle_config`` (``small_empty``, ``small_columns``, ``large_columns``, ``large_divider``; a
``*_empty`` room of any size), ``action_type`` ``heading``, ``position`` and ``velocity`` (actions ``float32[B, 2]``),
``low_dim_state``.  In the column / divider configurations the reference redraws the obstacles at every reset and rebuilds
its configuration space; here every layout of the pool carries its own room (static shapes + maps, built with the same
library calls), the pool is capped at ``max_rooms`` layouts and the kernels index the room tables by layout.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from .. import _lib
from ..layouts import sample_box_delivery_layout
from ..utils import DotDict, merge_cfg

# box_delivery/config.yaml restated (reference defaults that reach the hot path)
DEFAULT_CFG = {
    "teleop_mode": False, "low_dim_state": False, "fixed_trial_idx": 1, "output_dir": "logs/",
    "render": {"log_obs": False, "show": False, "show_obs": False, "frequency": 20},
    "sim": {"t_max": 2000, "steps": 100, "gravity": (0, 0), "iterations": 10, "damping": 0},
    "controller": {"dt": 0.2, "target_speed": 0.3},
    "boxes": {"num_boxes_small": 10, "num_boxes_large": 20, "box_size": 0.44, "min_box_dist": 0.62, "box_density": 0.001},
    "agent": {
        "action_type": "position", "step_size": 1.75, "random_start": True, "mass": 1, "padding": 0.25,
        "length": 0.6, "width": 0.92,
        "vertices": [[0.26, -0.35], [-0.31, -0.35], [-0.31, 0.35], [0.26, 0.35]],
        "front_bumper_vertices": [[0.31, -0.53], [0.20, -0.39], [0.20, 0.39], [0.31, 0.53]],
        "wheel_vertices": [
            [[0.09, 0.35], [0.09, 0.39], [0.24, 0.39], [0.24, 0.35]],
            [[0.09, -0.39], [0.09, -0.35], [0.24, -0.35], [0.24, -0.39]],
            [[-0.09, 0.35], [-0.09, 0.39], [-0.24, 0.39], [-0.24, 0.35]],
            [[-0.09, -0.35], [-0.09, -0.39], [-0.24, -0.39], [-0.24, -0.35]],
        ],
    },
    "env": {"obstacle_config": "small_empty", "room_length": 10, "room_width_small": 5, "room_width_large": 10,
            "receptacle_width": 1.5, "shortest_path_channel_scale": 0.25, "local_map_pixel_width": 224,
            "local_map_pixel_width_sam": 96, "local_map_width": 10, "wall_thickness": 14, "invert_receptacle_map": False},
    "misc": {"ministep_size": 2.5, "inactivity_cutoff_sam": 100, "inactivity_cutoff": 200, "random_seed": 42},
    "rewards": {"partial_rewards_scale": 0.2, "goal_reward": 10.0, "collision_penalty": 0.25, "non_movement_penalty": 0,
                "correct_direction_reward_scale": 3},
    "rewards_sam": {"partial_rewards_scale": 0.2, "goal_reward": 1.0, "collision_penalty": 0.25, "non_movement_penalty": 0.25,
                    "correct_direction_reward_scale": 1},
    "train": {"job_type": "ppo", "use_correct_direction_reward": False},
}


class BoxDelConfig(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("num_boxes", C.c_int32), ("device", C.c_int32), ("action_type", C.c_int32),
        ("steps", C.c_int32), ("iterations", C.c_int32), ("settle_steps", C.c_int32), ("inactivity_cutoff", C.c_int32),
        ("local_map_pixel_width", C.c_int32), ("use_correct_direction_reward", C.c_int32),
        ("force_box_spfa", C.c_int32), ("phase_kernels", C.c_int32),
        ("ctrl_dt", C.c_double), ("target_speed", C.c_double), ("damping", C.c_double),
        ("box_size", C.c_double), ("box_density", C.c_double),
        ("room_length", C.c_double), ("room_width", C.c_double), ("receptacle_width", C.c_double),
        ("local_map_width", C.c_double), ("step_size", C.c_double), ("robot_length", C.c_double), ("robot_width", C.c_double),
        ("partial_rewards_scale", C.c_double), ("goal_reward", C.c_double), ("collision_penalty", C.c_double),
        ("non_movement_penalty", C.c_double), ("correct_direction_reward_scale", C.c_double),
        ("ministep_size", C.c_double), ("shortest_path_channel_scale", C.c_double),
        ("chassis_verts", C.c_double * 8), ("bumper_verts", C.c_double * 8), ("wheel_verts", C.c_double * 32),
        ("env_kind", C.c_int32), ("step_limit", C.c_int32), ("post_steps", C.c_int32), ("t_max", C.c_int32),
        ("n_boundary", C.c_int32), ("n_goals", C.c_int32),
        ("lfc", C.c_double), ("omega_mul", C.c_double), ("v_mul", C.c_double), ("bound_x", C.c_double), ("bound_y", C.c_double),
        ("static_e", C.c_double), ("static_u", C.c_double),
        ("robot_code", C.c_double), ("box_code", C.c_double), ("box_done_code", C.c_double),
        ("footprint_verts", C.c_double * 8), ("n_footprint", C.c_int32), ("exact_spfa_parents", C.c_int32),
        ("boundary", C.c_double * 16), ("goals", C.c_double * 128),
        ("r_boundary", C.c_double), ("r_cleared", C.c_double), ("r_putback", C.c_double), ("r_push_mult", C.c_double),
        ("r_terminal", C.c_double), ("r_truncation", C.c_double), ("r_nonmove", C.c_double),
    ]


class BoxDelOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "reward", "terminated", "truncated", "info", "info_i", "completed", "box_verts")]


def bind(L):
    """Declare the bp_boxdel_* signatures on a loaded library (the product library, or the test-only emulation)."""
    L.bp_boxdel_default_config.argtypes = [C.POINTER(BoxDelConfig)]
    L.bp_boxdel_default_config.restype = None
    L.bp_boxdel_create.argtypes = [C.POINTER(BoxDelConfig), C.POINTER(C.c_void_p)]
    L.bp_boxdel_destroy.argtypes = [C.c_void_p]
    L.bp_boxdel_load.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + \
                                [C.c_int32, C.c_int32] + [C.c_void_p] * 6
    L.bp_boxdel_load_rooms.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + \
                                      [C.c_int32, C.c_int32] + [C.c_void_p] * 6
    L.bp_boxdel_load_rooms_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 7 + \
                                             [C.c_int32, C.c_int32, C.c_void_p]
    L.bp_boxdel_get_room_maps.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    L.bp_boxdel_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BoxDelOutputs), C.c_void_p]
    L.bp_boxdel_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BoxDelOutputs), C.c_void_p]
    L.bp_boxdel_step_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(BoxDelOutputs)]
    L.bp_boxdel_step_host_obs_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(BoxDelOutputs), C.c_void_p]
    L.bp_boxdel_get_bodies.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.bp_boxdel_get_step_debug.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bp_boxdel_launch_count.argtypes = [C.c_void_p]
    L.bp_boxdel_launch_count.restype = C.c_int64
    L.bp_boxdel_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    L.bp_boxdel_get_kernel_ms.argtypes = [C.c_void_p, C.c_void_p]
    L.bp_boxdel_env_bytes.argtypes = [C.c_void_p]
    L.bp_boxdel_env_bytes.restype = C.c_size_t
    return L


def disk(radius: int) -> np.ndarray:
    """skimage.morphology.disk"""
    r = int(radius)
    L = np.arange(-r, r + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= r ** 2, dtype=np.uint8)


class RoomMaps:
    """Static shapes and maps of one room (box_delivery_env.py:322-335, 342-358, 472-489, 1098-1175; sim_utils.py:78-131)."""

    # static-shape slots of the obstacle configurations (the same slot list for every room of a handle; a room that draws
    # fewer columns leaves slots empty): receptacle, 4 walls, [columns | divider], 3 room corners x 3 triangles, [2 divider corners x 3]
    MAX_COLUMNS = {"small_columns": 2, "large_columns": 7}      # randint(1, 3) / randint(1, 8): box_delivery_env.py:466-469

    def __init__(self, cfg, layout=None, build_maps=True):
        """build_maps=False: static shapes and their pixel polygons only (the maps are then built on the device at every reset,
        csrc/boxdel_room.cuh)."""
        env = cfg["env"]
        small = env["obstacle_config"].split("_")[0] == "small"
        self.room_length = rl = env["room_length"]
        self.room_width = rw = env["room_width_small"] if small else env["room_width_large"]
        wt = env["wall_thickness"]
        size = env["receptacle_width"]
        self.LW = LW = env["local_map_pixel_width"] if cfg["train"]["job_type"] != "sam" else env["local_map_pixel_width_sam"]
        self.lmw = lmw = env["local_map_width"]
        self.ppm = ppm = LW / lmw
        ag = cfg["agent"]
        self.robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
        self.robot_half_width = max(ag["length"], ag["width"]) / 2
        self.receptacle_position = (rl / 2 - size / 2, rw / 2 - size / 2)
        rect = lambda x, y, l, w: [[x - l / 2, y - w / 2], [x + l / 2, y - w / 2], [x + l / 2, y + w / 2], [x - l / 2, y + w / 2]]
        # shape creation order: receptacle, walls, corners (generate_sim_bounds)
        st = [(rect(*self.receptacle_position, size, size), (0.0, 0.0, 0.0), 0.0, 4)]
        for (x, y, length, width) in [(-rl / 2 - wt / 2, 0, wt, rw), (rl / 2 + wt / 2, 0, wt, rw),
                                      (0, -rw / 2 - wt / 2, rl + 2 * wt, wt), (0, rw / 2 + wt / 2, rl + 2 * wt, wt)]:
            st.append((rect(x, y, length, width), (0.0, 0.0, 0.0), 0.0, 3))
        # random obstacles of the layout (generate_boundary :358-471; create_static: world vertices, no radius)
        oc = env["obstacle_config"]
        columns = list((layout or {}).get("columns") or [])
        divider = (layout or {}).get("divider")
        empty = ([], (0.0, 0.0, 0.0), 0.0, 3)
        if oc in self.MAX_COLUMNS:
            assert len(columns) <= self.MAX_COLUMNS[oc]
            for k in range(self.MAX_COLUMNS[oc]):
                st.append((rect(columns[k][0], columns[k][1], 1, 1), (0.0, 0.0, 0.0), 0.0, 3) if k < len(columns) else empty)
        elif oc == "large_divider":
            st.append((rect(divider[0], divider[1], 8, 0.5), (0.0, 0.0, 0.0), 0.0, 3) if divider is not None else empty)
        s225, c225 = np.sin(22.5 * np.pi / 180), np.cos(22.5 * np.pi / 180)
        tri = [[(0, 0), (0, -1), (0.5 * s225, -1 + 0.5 * c225)],
               [(0, 0), (0.5 * s225, -1 + 0.5 * c225), (1 - 0.5 * c225, -0.5 * s225)],
               [(0, 0), (1, 0), (1 - 0.5 * c225, -0.5 * s225)]]
        corners = []
        for i, (x, y) in enumerate([(-rl / 2, rw / 2), (rl / 2, rw / 2), (rl / 2, -rw / 2), (-rl / 2, -rw / 2)]):
            if i == 1:
                continue                      # the receptacle corner
            corners.append((x, y, -np.radians(i * 90)))
        if oc == "large_divider":             # corners where the divider meets the right wall (:491-505)
            corners += [(rl / 2, divider[1] - 0.5 / 2, np.radians(-90)), (rl / 2, divider[1] + 0.5 / 2, np.radians(180))] if divider is not None else [None, None]
        for cr in corners:
            for vs in tri:
                st.append(([list(map(float, v)) for v in vs], (float(cr[0]), float(cr[1]), float(cr[2])), 0.02, 3) if cr is not None else empty)
        self.statics = st
        ns = len(st)
        self.st_nv = np.array([len(s[0]) for s in st], dtype=np.int32)
        self.st_verts = np.zeros((ns, 4, 2), dtype=np.float64)
        for k, s in enumerate(st):
            if len(s[0]):
                self.st_verts[k, :len(s[0])] = np.asarray(s[0], dtype=np.float64)
        self.st_pose = np.array([s[1] for s in st], dtype=np.float64)
        self.st_rad = np.array([s[2] for s in st], dtype=np.float64)
        self.st_ctype = np.array([s[3] for s in st], dtype=np.int32)
        # ---- maps
        shape = (int(2 * np.ceil((rw * ppm + LW * np.sqrt(2)) / 2)), int(2 * np.ceil((rl * ppm + LW * np.sqrt(2)) / 2)))
        self.shape = shape
        small_shape = (LW + 20, LW + 20)
        off = int(lmw * ppm / 2) + 10
        # pixel polygon of every collision_type-3 shape as update_configuration_space passes it to cv2.fillPoly (:1151-1162)
        self.st_npx = np.zeros(ns, dtype=np.int32)
        self.st_pxy = np.zeros((ns, 4, 2), dtype=np.int32)
        polys = []
        for k, (verts, pose, _, ctype) in enumerate(st):
            if ctype != 3 or len(verts) == 0:
                continue
            cs, sn = np.cos(pose[2]), np.sin(pose[2])
            v = np.asarray(verts, dtype=np.float64)
            world = np.stack([cs * v[:, 0] - sn * v[:, 1] + pose[0], sn * v[:, 0] + cs * v[:, 1] + pose[1]], axis=1)
            px = (world * ppm).astype(np.int32)
            px[:, 0] += off
            px[:, 1] += off
            px[:, 1] = small_shape[0] - px[:, 1]
            polys.append(px)
            self.st_npx[k] = len(px)
            self.st_pxy[k, :len(px)] = px
        W = LW
        robot_pixel_width = int(2 * self.robot_radius * ppm)
        chan = np.zeros((W, W), dtype=np.float32)
        start = int(np.floor(W / 2 - robot_pixel_width / 2))
        ii = np.arange(start, start + robot_pixel_width)
        I, J = np.meshgrid(ii, ii, indexing="ij")
        inside = (((I + 0.5) - W / 2) ** 2 + ((J + 0.5) - W / 2) ** 2) ** 0.5 < robot_pixel_width / 2
        chan[I[inside], J[inside]] = 1
        self.robot_state_channel = chan
        if not build_maps:
            return
        import cv2
        from scipy.ndimage import binary_dilation, distance_transform_edt
        obstacle_map = np.ones(shape, dtype=np.float32)
        small_map = np.zeros(small_shape, dtype=np.float32)
        for px in polys:
            cv2.fillPoly(small_map, [px], color=1)
        si, sj = int(shape[0] / 2 - small_map.shape[0] / 2), int(shape[1] / 2 - small_map.shape[1] / 2)
        obstacle_map[si:si + small_map.shape[0], sj:sj + small_map.shape[1]] = small_map
        self.configuration_space = 1 - binary_dilation(obstacle_map, disk(np.floor(self.robot_radius * ppm))).astype(np.float32)
        self.configuration_space_thin = 1 - binary_dilation(obstacle_map, disk(np.floor(self.robot_half_width * ppm))).astype(np.float32)
        self.closest_cspace_indices = distance_transform_edt(1 - self.configuration_space, return_distances=False, return_indices=True)
        self.small_obstacle_map = 1 - small_map


LAYOUT_EPISODE_STRIDE = 1000003


def default_pool_episodes(num_envs: int) -> int:
    """Episodes per env in the default layout pool: 4, fewer for very large batches (the pool is built on the host)."""
    return max(1, min(4, 65536 // max(int(num_envs), 1)))


class BoxDeliveryVecEnv:
    metadata = {"render_modes": []}

    def __init__(self, num_envs: int = 1, cfg=None, device: Optional[int] = None, layouts: Optional[Sequence[dict]] = None,
                 seed: Optional[int] = None, force_box_spfa: bool = False, env_offset: int = 0,
                 layout_pool_episodes: Optional[int] = None, max_rooms: Optional[int] = 64,
                 exact_spfa_parents: bool = True, device_rooms: Optional[bool] = None):
        self.cfg = merge_cfg(DotDict(DEFAULT_CFG), cfg)
        c = self.cfg
        if c["env"]["obstacle_config"] not in ("small_empty", "large_empty", "small_columns", "large_columns", "large_divider"):
            raise ValueError("Invalid obstacle config: %r" % (c["env"]["obstacle_config"],))
        if c["agent"]["action_type"] not in ("heading", "position", "velocity"):
            raise NotImplementedError("unknown action_type %r" % (c["agent"]["action_type"],))
        if c["teleop_mode"]:
            raise NotImplementedError("teleop_mode (keyboard tele-operation) is out of scope")
        if tuple(c["sim"]["gravity"]) != (0, 0):
            raise NotImplementedError("non-zero gravity is not supported")
        self.num_envs = E = int(num_envs)
        self._open(device)
        small = c["env"]["obstacle_config"].split("_")[0] == "small"
        self.num_boxes = NB = int(c["boxes"]["num_boxes_small"] if small else c["boxes"]["num_boxes_large"])
        self.num_channels = 4
        per_layout_rooms = not c["env"]["obstacle_config"].endswith("_empty")
        # device_rooms (default in the column / divider configurations): every layout carries its static shapes only and the maps of
        # an env are rebuilt ON THE DEVICE at each of its resets (csrc/boxdel_room.cuh), as the reference rebuilds them at every
        # reset; False: the maps of every layout are built here with cv2 / scipy and uploaded (a pool of at most max_rooms)
        self.device_rooms = per_layout_rooms if device_rooms is None else bool(device_rooms) and per_layout_rooms
        self.room = RoomMaps(c, {"columns": [], "divider": [0.0, 0.0]} if per_layout_rooms else None,
                             build_maps=not per_layout_rooms)      # sizes and constants (and the maps of an obstacle-free room)
        self.local_map_pixel_width = LW = self.room.LW
        self.goal_points = [self.room.receptacle_position]
        if layouts is None:
            base = int(c["misc"]["random_seed"]) if seed is None else int(seed)
            # The reference draws a fresh layout at every reset (box_delivery_env.py:509-546 from RandomState(42)).  Here the
            # layouts are inputs: a pool of F episodes per env, entry f * E + e drawn from RandomState(base + global env id +
            # f * LAYOUT_EPISODE_STRIDE) — keyed by (GLOBAL env id, episode), so results depend neither on the number of ranks
            # (env_offset = first env of this rank's shard) nor on when other envs finish; an env's episode k runs on entry k mod F.
            F = default_pool_episodes(E) if layout_pool_episodes is None else max(1, int(layout_pool_episodes))
            keys = [(f, e) for f in range(F) for e in range(E)]
            if per_layout_rooms and not self.device_rooms and max_rooms is not None:
                # every layout of a column / divider configuration carries its own room maps (3 - 4 MB each on the device,
                # built on the host): the pool is capped; env e's episode k then runs on entry (e + k * E) mod len(pool)
                keys = keys[:max(1, int(max_rooms))]
            layouts = [sample_box_delivery_layout(np.random.RandomState(base + int(env_offset) + e + f * LAYOUT_EPISODE_STRIDE), c)
                       for f, e in keys]
        self.layouts = list(layouts)
        self.rooms = [RoomMaps(c, l, build_maps=not self.device_rooms) for l in self.layouts] if per_layout_rooms else [self.room]
        # ---- native handle
        L = self._L
        k = BoxDelConfig()
        L.bp_boxdel_default_config(C.byref(k))
        rw = c["rewards_sam"] if c["train"]["job_type"] == "sam" else c["rewards"]
        # True (default): robot paths / box distances follow the reference solver's own parent tie-breaks (its queue restated by one
        # thread per solve); False: parallel relaxation + deterministic parents — same distances, possibly another equal-cost path
        k.exact_spfa_parents = 1 if exact_spfa_parents else 0
        k.force_box_spfa = 1 if force_box_spfa else 0      # testing aid: solve spfa per box even when the exact shortcut applies
        k.n_envs = E; k.num_boxes = NB; k.device = self.device_index
        k.action_type = {"position": 0, "heading": 1, "velocity": 2}[c["agent"]["action_type"]]
        self._action_dim = 2 if k.action_type == 2 else 1
        k.steps = int(c["sim"]["steps"]); k.iterations = int(c["sim"]["iterations"]); k.settle_steps = 1000
        k.inactivity_cutoff = int(c["misc"]["inactivity_cutoff"] if c["train"]["job_type"] != "sam" else c["misc"]["inactivity_cutoff_sam"])
        k.local_map_pixel_width = LW; k.use_correct_direction_reward = 1 if c["train"]["use_correct_direction_reward"] else 0
        k.ctrl_dt = float(c["controller"]["dt"]); k.target_speed = float(c["controller"]["target_speed"]); k.damping = float(c["sim"]["damping"])
        k.box_size = float(c["boxes"]["box_size"]); k.box_density = float(c["boxes"]["box_density"])
        k.room_length = float(self.room.room_length); k.room_width = float(self.room.room_width)
        k.receptacle_width = float(c["env"]["receptacle_width"])
        k.local_map_width = float(c["env"]["local_map_width"]); k.step_size = float(c["agent"]["step_size"])
        k.robot_length = float(c["agent"]["length"]); k.robot_width = float(c["agent"]["width"])
        k.partial_rewards_scale = float(rw["partial_rewards_scale"]); k.goal_reward = float(rw["goal_reward"])
        k.collision_penalty = float(rw["collision_penalty"]); k.non_movement_penalty = float(rw["non_movement_penalty"])
        k.correct_direction_reward_scale = float(rw["correct_direction_reward_scale"])
        k.ministep_size = float(c["misc"]["ministep_size"]); k.shortest_path_channel_scale = float(c["env"]["shortest_path_channel_scale"])
        for i, v in enumerate(np.asarray(c["agent"]["vertices"], dtype=np.float64).reshape(-1)):
            k.chassis_verts[i] = float(v)
        for i, v in enumerate(np.asarray(c["agent"]["front_bumper_vertices"], dtype=np.float64).reshape(-1)):
            k.bumper_verts[i] = float(v)
        for i, v in enumerate(np.asarray(c["agent"]["wheel_vertices"], dtype=np.float64).reshape(-1)):
            k.wheel_verts[i] = float(v)
        k.bound_x = float(self.room.room_length); k.bound_y = float(self.room.room_width)     # D17: x clamps with map_height = room_length
        self._finish_init(k, self.rooms, ch3_map=None)

    def _open(self, device):
        """Bind the CUDA library and pick the device.  There is no CPU fallback: no device, no env."""
        import torch
        if not torch.cuda.is_available():
            raise _lib.BenchPushError("%s needs a CUDA device (there is no CPU fallback)" % type(self).__name__)
        self._L = bind(_lib.lib())
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)

    def _init_spaces(self):
        """The reference's spaces (box_delivery_env.py:157-173, area_clearing.py:200-221) as ``single_*``; ``action_space`` /
        ``observation_space`` carry a leading axis of num_envs.  Area-clearing declares its position action as int32 (1,)."""
        from .. import spaces
        at = self.cfg["agent"]["action_type"]
        LW = int(self.local_map_pixel_width)
        area = any(k.__name__ == "AreaClearingVecEnv" for k in type(self).__mro__)
        if at == "velocity":
            a = spaces.box(-1, 1, (2,), np.float32)
        elif at == "heading":
            a = spaces.box(-1, 1, (1,), np.float32)
        else:
            a = spaces.box(0, LW * LW, (1,), np.int32) if area else spaces.box(0, LW * LW, (), np.float32)
        self.single_action_space = a
        self.observation_shape = (LW, LW, self.num_channels)
        self.single_observation_space = spaces.box(0, 255, self.observation_shape, np.uint8)
        self.action_space = spaces.batched(a, self.num_envs)
        self.observation_space = spaces.batched(self.single_observation_space, self.num_envs)

    def _finish_init(self, k, r, ch3_map):
        """Create the native handle from the filled config, upload layouts / room maps, allocate the output buffers."""
        self._init_spaces()
        import os
        if os.environ.get("BENCHPUSH_BD_PHASED") is not None:          # development aid: A/B of the substep-loop variants
            k.phase_kernels = int(os.environ["BENCHPUSH_BD_PHASED"])
        L, E, NB, LW = self._L, self.num_envs, self.num_boxes, self.local_map_pixel_width
        self._h = C.c_void_p()
        self._check(L.bp_boxdel_create(C.byref(k), C.byref(self._h)))
        nl = len(self.layouts)
        starts = np.array([l["start"] for l in self.layouts], dtype=np.float64)
        nbx = np.array([len(l["boxes"]) for l in self.layouts], dtype=np.int32)
        boxes = np.zeros((nl, NB, 3), dtype=np.float64)
        for i, l in enumerate(self.layouts):
            boxes[i, :len(l["boxes"])] = np.asarray(l["boxes"], dtype=np.float64).reshape(-1, 3)
        # rooms: one for the handle, or one per layout (column / divider configurations); every room array gains a leading axis
        rooms = list(r) if isinstance(r, (list, tuple)) else [r]
        r = rooms[0]
        if getattr(self, "device_rooms", False):
            stack = lambda f, dt: np.ascontiguousarray(np.stack([f(q) for q in rooms]).astype(dt))
            st_nv = stack(lambda q: q.st_nv, np.int32); st_verts = stack(lambda q: q.st_verts, np.float64)
            st_pose = stack(lambda q: q.st_pose, np.float64); st_rad = stack(lambda q: q.st_rad, np.float64); st_ctype = stack(lambda q: q.st_ctype, np.int32)
            st_npx = stack(lambda q: q.st_npx, np.int32); st_pxy = stack(lambda q: q.st_pxy, np.int32)
            rc = np.ascontiguousarray((r.robot_state_channel * 255).astype(np.uint8))
            self._check(L.bp_boxdel_load_rooms_device(self._h, nl, starts.ctypes.data, nbx.ctypes.data, boxes.ctypes.data, len(r.statics),
                                                      st_nv.ctypes.data, st_verts.ctypes.data, st_pose.ctypes.data, st_rad.ctypes.data, st_ctype.ctypes.data,
                                                      st_npx.ctypes.data, st_pxy.ctypes.data, r.shape[0], r.shape[1], rc.ctypes.data))
            self._alloc_buffers()
            return
        if len(rooms) > 1:
            assert len(rooms) == nl, "one room per layout"
        layout_room = np.arange(nl, dtype=np.int32) if len(rooms) > 1 else np.zeros(nl, dtype=np.int32)
        stack = lambda f, dt: np.ascontiguousarray(np.stack([f(q) for q in rooms]).astype(dt))
        cs = stack(lambda q: q.configuration_space, np.uint8)
        th = stack(lambda q: q.configuration_space_thin, np.uint8)
        cl = stack(lambda q: q.closest_cspace_indices, np.int32)
        sm = stack(lambda q: q.small_obstacle_map, np.uint8)
        rc = np.ascontiguousarray((r.robot_state_channel * 255).astype(np.uint8))
        c3 = None if ch3_map is None else np.ascontiguousarray(ch3_map, dtype=np.float32)
        st_nv = stack(lambda q: q.st_nv, np.int32); st_verts = stack(lambda q: q.st_verts, np.float64)
        st_pose = stack(lambda q: q.st_pose, np.float64); st_rad = stack(lambda q: q.st_rad, np.float64); st_ctype = stack(lambda q: q.st_ctype, np.int32)
        self._check(L.bp_boxdel_load_rooms(self._h, nl, starts.ctypes.data, nbx.ctypes.data, boxes.ctypes.data, len(rooms), layout_room.ctypes.data,
                                           len(r.statics), st_nv.ctypes.data, st_verts.ctypes.data, st_pose.ctypes.data, st_rad.ctypes.data, st_ctype.ctypes.data,
                                           r.shape[0], r.shape[1], cs.ctypes.data, th.ctypes.data, cl.ctypes.data, sm.ctypes.data, rc.ctypes.data,
                                           None if c3 is None else c3.ctypes.data))
        self._alloc_buffers()

    def _alloc_buffers(self):
        L, E, NB, LW = self._L, self.num_envs, self.num_boxes, self.local_map_pixel_width
        # ---- buffers
        self._buf = {}
        self._alloc("obs", (E, LW, LW, 4), "uint8"); self._alloc("reward", (E,), "float64")
        self._alloc("terminated", (E,), "uint8"); self._alloc("truncated", (E,), "uint8")
        self._alloc("info", (E, 8), "float64"); self._alloc("info_i", (E, 4), "int32")
        self._alloc("completed", (E, NB), "uint8"); self._alloc("box_verts", (E, NB, 4, 2), "float64")
        self._alloc("actions", (E,) if getattr(self, "_action_dim", 1) == 1 else (E, 2), "float32"); self._alloc("layout_ids", (E,), "int32"); self._alloc("mask", (E,), "uint8")
        o = BoxDelOutputs()
        for n in ("obs", "reward", "terminated", "truncated", "info", "info_i", "completed", "box_verts"):
            setattr(o, n, self._ptr(n))
        self._out = o
        self.episode_idx = None
        # cfg.low_dim_state (box_delivery_env.py:166-169, 606-607; area_clearing.py:212-215, 599-600): the observation is
        # generate_observation_low_dim and the image kernel is not launched (NULL observation pointer)
        self.low_dim_state = bool(self.cfg["low_dim_state"])
        if self.low_dim_state:
            from .. import spaces
            self._out.obs = None
            self.observation_shape = (2 * NB,)
            self.single_observation_space = spaces.box(-10, 30, self.observation_shape, np.float32)
            self.observation_space = spaces.batched(self.single_observation_space, E)

    def _observation(self):
        return self.low_dim_observation() if self.low_dim_state else self._buf["obs"]

    # ---- buffers: torch CUDA tensors
    def _alloc(self, name, shape, dtype):
        import torch
        self._buf[name] = torch.zeros(shape, dtype=getattr(torch, dtype), device=self.device)

    def _ptr(self, name):
        return self._buf[name].data_ptr()

    def _put(self, name, values):
        import torch
        t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values))
        self._buf[name].copy_(t.to(self.device).reshape(self._buf[name].shape).to(self._buf[name].dtype))

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _as_bool(x):
        return x.bool()

    def _check(self, rc):
        if rc < 0:
            raise _lib.BenchPushError("benchpush_b200 error %d: %s" % (rc, self._L.bp_last_error().decode("utf-8", "replace")
                                                                       if hasattr(self._L, "bp_last_error") and self._L.bp_last_error.restype is not None
                                                                       else rc))
        return rc

    @property
    def obs(self):
        return self._buf["obs"]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.bp_boxdel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- gym API
    def reset(self, seed=None, options=None, mask=None):
        """Every env (or those in ``mask``) starts its next episode: env e keeps its own episode counter k_e (advanced only
        when e is reset) and runs on layout (e + k_e * num_envs) % n_layouts."""
        if mask is not None and hasattr(mask, "cpu"):
            mask = mask.to("cpu").numpy()
        if self.episode_idx is None:
            self.episode_idx = np.full(self.num_envs, -1, dtype=np.int64)
        sel = np.ones(self.num_envs, dtype=bool) if mask is None else np.asarray(mask).astype(bool).reshape(self.num_envs)
        self.episode_idx[sel] += 1
        ids = (np.arange(self.num_envs) + np.maximum(self.episode_idx, 0) * self.num_envs) % len(self.layouts)
        self._put("layout_ids", ids.astype(np.int32))
        mptr = None
        if mask is not None:
            self._put("mask", mask)
            mptr = self._ptr("mask")
        self._check(self._L.bp_boxdel_reset(self._h, self._ptr("layout_ids"), mptr, C.byref(self._out), self._stream()))
        return self._observation(), self._info()

    def step(self, action, mask=None):
        self._put("actions", action)
        mptr = None
        if mask is not None:
            self._put("mask", mask)
            mptr = self._ptr("mask")
        self._check(self._L.bp_boxdel_step(self._h, self._ptr("actions"), mptr, C.byref(self._out), self._stream()))
        b = self._buf
        term, trunc = self._as_bool(b["terminated"]), self._as_bool(b["truncated"])
        return self._observation(), b["reward"], term, trunc, self._info()

    def _info(self):
        b = self._buf
        info, ii = b["info"], b["info_i"]
        return {"state": info[:, 0:3], "cumulative_distance": info[:, 3], "cumulative_boxes": ii[:, 0], "cumulative_reward": info[:, 4],
                "total_work": info[:, 5], "obs": b["box_verts"], "n_boxes_left": ii[:, 2], "box_completed_statuses": b["completed"],
                "goal_positions": self.goal_points, "ministeps": info[:, 6], "inactivity": ii[:, 1], "status": ii[:, 3], "work": info[:, 7]}

    def step_host(self, actions: np.ndarray, out: Optional["BoxDelOutputs"] = None, want_obs=False):
        """bp_boxdel_step_host: actions and results in HOST buffers (copies inside the call, synchronises).
        Returns a dict of numpy arrays (reward, terminated, truncated, info, info_i).  want_obs="device"
        (bp_boxdel_step_host_obs_device): the observation kernel runs too and renders into ``self.obs`` on the device — the
        call a GPU-resident policy makes; False: no observation is produced."""
        if not hasattr(self, "_hb"):
            E = self.num_envs
            self._hb = {"reward": np.zeros(E), "terminated": np.zeros(E, dtype=np.uint8), "truncated": np.zeros(E, dtype=np.uint8),
                        "info": np.zeros((E, 8)), "info_i": np.zeros((E, 4), dtype=np.int32), "completed": np.zeros((E, self.num_boxes), dtype=np.uint8)}
            o = BoxDelOutputs()
            for n, a in self._hb.items():
                setattr(o, n, a.ctypes.data)
            self._ho = o
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if want_obs == "device" and not self.low_dim_state:
            self._check(self._L.bp_boxdel_step_host_obs_device(self._h, a.ctypes.data, C.byref(self._ho), self._ptr("obs")))
        else:
            self._check(self._L.bp_boxdel_step_host(self._h, a.ctypes.data, C.byref(self._ho)))
        return self._hb

    def set_profiling(self, on: bool):
        self._check(self._L.bp_boxdel_set_profiling(self._h, 1 if on else 0))

    def kernel_ms(self):
        o = np.zeros(3, dtype=np.float64)
        self._check(self._L.bp_boxdel_get_kernel_ms(self._h, o.ctypes.data))
        return {"move_ms": float(o[0]), "obs_ms": float(o[1]), "steps": int(o[2])}

    def launch_count(self):
        return int(self._L.bp_boxdel_launch_count(self._h))

    def low_dim_observation(self):
        """``generate_observation_low_dim`` (box_delivery_env.py:1025-1037, area_clearing.py:908-920), batched: for every box left,
        ``np.abs(poly_centroid(vertices))`` (x, y) in list order; zero padded to ``2 * num_boxes``.  Computed from ``info['obs']``."""
        v = self._buf["box_verts"]
        n = self._buf["info_i"][:, 2]
        import torch
        x, y = v[..., 0], v[..., 1]
        xp, yp = torch.roll(x, 1, dims=2), torch.roll(y, 1, dims=2)
        A = 0.5 * ((x * yp).sum(2) - (y * xp).sum(2)).abs()
        u = x * yp - xp * y
        cx = (((x + xp) * u).sum(2) / (6 * A)).abs(); cy = (((y + yp) * u).sum(2) / (6 * A)).abs()
        valid = torch.arange(self.num_boxes, device=v.device)[None, :] < n[:, None]
        zero = torch.zeros((), dtype=v.dtype, device=v.device)
        return torch.stack([torch.where(valid, cx, zero), torch.where(valid, cy, zero)], dim=2).reshape(self.num_envs, 2 * self.num_boxes)

    # ---- parity / debugging aids
    def get_room_maps(self, e: int) -> dict:
        """The maps env ``e`` uses at the moment (bp_boxdel_get_room_maps): configuration_space, configuration_space_thin,
        closest_cspace_indices, small_obstacle_map (1 = free floor), receptacle_map (fourth channel before the window shift)."""
        H, W = self.room.shape
        S = self.local_map_pixel_width + 20
        cs = np.zeros((H, W), np.uint8); th = np.zeros((H, W), np.uint8); cl = np.zeros((2, H, W), np.int32)
        sm = np.zeros((S, S), np.uint8); c3 = np.zeros((H, W), np.float32)
        self._check(self._L.bp_boxdel_get_room_maps(self._h, int(e), cs.ctypes.data, th.ctypes.data, cl.ctypes.data, sm.ctypes.data, c3.ctypes.data))
        return {"configuration_space": cs, "configuration_space_thin": th, "closest_cspace_indices": cl, "small_obstacle_map": sm, "receptacle_map": c3}

    def get_bodies(self, e: int) -> np.ndarray:
        out = np.zeros((1 + self.num_boxes, 7), dtype=np.float64)
        self._check(self._L.bp_boxdel_get_bodies(self._h, int(e), out.ctypes.data))
        return out

    def get_step_debug(self, e: int):
        o = np.zeros(16, dtype=np.float64)
        d = np.zeros((2, self.num_boxes), dtype=np.float64)
        p = np.zeros((24, 3), dtype=np.float64)
        self._check(self._L.bp_boxdel_get_step_debug(self._h, int(e), o.ctypes.data, d.ctypes.data, p.ctypes.data))
        return {"n_path": int(o[0]), "move_sign": int(o[1]), "need_spfa": int(o[2]), "status": int(o[3]), "robot_hit": int(o[4]),
                "substeps": int(o[5]), "candidates": int(o[6]), "arbiters": int(o[7]), "robot_distance": float(o[8]),
                "init_dist": d[0], "final_dist": d[1], "path": p[:int(o[0])]}
