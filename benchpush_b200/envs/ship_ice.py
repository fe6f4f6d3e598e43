"""``ship-ice-v0`` as a batched environment on one B200.

Keeps the reference's gym-style API (``ShipIceEnv``, ship_ice_env.py:37-355): ``reset() -> (obs, info)``,
``step(action) -> (obs, reward, terminated, truncated, info)``, the cfg dict with the same keys and
the same one-level merge, and the same ``info`` keys — but every return value carries a leading
batch dimension and lives on the GPU as a torch tensor.  All simulation work happens in
``libbenchpush_b200.so`` (CUDA, sm_100a); torch only owns the buffers and the stream.

Ownership differs from the reference in one way: the reference allocates fresh numpy arrays per call, whereas the tensors
returned by ``reset()`` / ``step()`` (``obs``, ``reward``, ``terminated``, ``truncated`` and the tensors inside ``info``) are
the env's persistent device buffers, overwritten in place by the next ``step()`` / ``reset()`` (zero-copy hand-off to a
policy on the GPU).  A caller that keeps a value across steps — a replay buffer, a per-episode trace — must ``.clone()`` it.
The same holds for ``MazeNamoVecEnv`` and the room envs.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..layouts import PackedLayouts, generate_experiment, load_reference_pickle, pack_trials
from ..utils import DotDict, merge_cfg

_CFG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", "ship_ice.yaml")


class _LazyObstacles:
    """``info['obs']`` of the reference is the list of world-space vertex arrays of every floe
    (cost_map.py:274-281).  Materialising it for a whole batch every step would move gigabytes, so
    it is fetched per env on demand: ``info['obs'][e]`` -> list of (V_i, 2) float64 arrays."""

    def __init__(self, env):
        self._env = env

    def __len__(self):
        return self._env.num_envs

    def __getitem__(self, e):
        return self._env.get_obstacles(int(e))


class ShipIceVecEnv:
    metadata = {"render_modes": []}
    _prefix = "shipice"          # C ABI family of this env: bp_<prefix>_reset / step / ... (include/benchpush_b200.h)
    _obs_shape = _lib.OBS_SHAPE

    def _fn(self, name):
        return getattr(self._L, "bp_%s_%s" % (self._prefix, name))

    def __init__(self, num_envs: int = 1, cfg=None, device: Optional[int] = None, layouts: Optional[PackedLayouts] = None,
                 trials: Optional[Sequence[dict]] = None, ice_file: Optional[str] = None, auto_reset: bool = False,
                 first_trial: Optional[Sequence[int]] = None, synthetic_seed: int = 99, n_synthetic_trials: int = 100,
                 start_seed: int = 0, precision: str = "fp64"):
        if not torch.cuda.is_available():
            raise _lib.BenchPushError("ShipIceVecEnv needs a CUDA device (there is no CPU fallback)")
        self.cfg = merge_cfg(DotDict.load_from_file(_CFG_PATH), cfg)
        self.low_dim_state = bool(self.cfg.low_dim_state)
        self.egocentric_obs = bool(self.cfg.egocentric_obs)
        self.random_start = bool(self.cfg.random_start)
        assert self.cfg.concentration in [0.1, 0.2, 0.3, 0.4, 0.5], "concentration must be one of 0.1 .. 0.5 (ship_ice_env.py:75)"
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.goal = (0, self.cfg.goal_y)
        self.max_yaw_rate_step = (np.pi / 2) / 7
        self.beta = 30
        self.num_channels = 4
        # ---- layouts (the reference loads experiments_XX_100_r06_d40x12.pk, ship_ice_env.py:76-80)
        if layouts is None:
            if trials is None:
                if ice_file is not None:
                    trials = load_reference_pickle(ice_file, self.cfg.concentration)
                else:
                    trials = generate_experiment(n_synthetic_trials, self.cfg.concentration, seed=synthetic_seed)
            layouts = pack_trials(trials)
        if self.low_dim_state:
            # the reference runs every episode on trial cfg.fixed_trial_idx when low_dim_state (ship_ice_env.py:84-87, 189-190),
            # which also fixes the observation length at 2 * n_floes of that trial
            k = int(self.cfg.fixed_trial_idx) % layouts.n_layouts
            layouts = PackedLayouts(layouts.n_floes[k:k + 1], layouts.n_verts[k:k + 1], layouts.verts[k:k + 1],
                                    layouts.centres[k:k + 1], layouts.ship_state[k:k + 1])
        self.layouts = layouts
        self.env_max_trial = layouts.n_layouts
        # ---- native handle
        self.precision = precision          # "fp64": bit-comparable with the oracle; "fp32": fp32 body state / contacts / solver (tolerance parity)
        L = _lib.lib(precision)
        c = _lib.ShipIceConfig()
        L.bp_shipice_default_config(C.byref(c))
        c.n_envs = self.num_envs
        c.max_floes = layouts.max_floes
        c.device = self.device_index
        c.auto_reset = 1 if auto_reset else 0
        c.dt = float(self.cfg.dt); c.steps = int(self.cfg.sim.steps); c.iterations = int(self.cfg.sim.iterations)
        c.damping = float(self.cfg.sim.damping); c.obstacle_density = float(self.cfg.sim.obstacle_density)
        c.target_speed = float(self.cfg.target_speed); c.goal_y = float(self.cfg.goal_y)
        c.map_width = float(self.cfg.occ.map_width); c.map_height = float(self.cfg.occ.map_height)
        c.m_to_pix_scale = int(self.cfg.occ.m_to_pix_scale)
        if tuple(self.cfg.sim.gravity) != (0, 0):
            raise NotImplementedError("non-zero gravity is not supported by the awake-set solver")
        sv = np.asarray(self.cfg.ship.vertices, dtype=np.float64)
        c.n_ship_verts = sv.shape[0]
        for i, v in enumerate(sv.reshape(-1)):
            c.ship_verts[i] = float(v)
        c.ship_head[0], c.ship_head[1] = map(float, self.cfg.ship.head)
        c.ship_tail[0], c.ship_tail[1] = map(float, self.cfg.ship.tail)
        self._h = C.c_void_p()
        _lib.check(L.bp_shipice_create(C.byref(c), C.byref(self._h)))
        self._L = L
        nf = np.ascontiguousarray(layouts.n_floes, dtype=np.int32)
        nv = np.ascontiguousarray(layouts.n_verts, dtype=np.int32)
        vs = np.ascontiguousarray(layouts.verts, dtype=np.float64)
        ce = np.ascontiguousarray(layouts.centres, dtype=np.float64)
        ss = np.ascontiguousarray(layouts.ship_state, dtype=np.float64)
        _lib.check(L.bp_shipice_load_layouts(self._h, layouts.n_layouts, nf.ctypes.data, nv.ctypes.data, vs.ctypes.data,
                                             ce.ctypes.data, ss.ctypes.data))
        # ---- buffers
        E, dev = self.num_envs, self.device
        self._global_obs = not self.low_dim_state and not self.egocentric_obs
        if self._global_obs:
            # generate_observation's second branch (ship_ice_env.py:96, 397-407): concentration + footprint on 0.2 m cells
            hw = np.zeros(2, dtype=np.int32)
            _lib.check(L.bp_shipice_obs_global_shape(self._h, hw.ctypes.data))
            self._obs_shape = (2, int(hw[0]), int(hw[1]))
            self.obs = torch.zeros((E,) + self._obs_shape, dtype=torch.uint8, device=dev)
        else:
            self.obs = torch.zeros((E,) + _lib.OBS_SHAPE, dtype=torch.uint8, device=dev)
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
        self._init_low_dim(2 * layouts.max_floes)
        self._start_poses = None
        self.redraw_starts = True          # False: keep the poses given to set_start_poses() (replaying recorded starts)
        if self.random_start:
            # ship_ice_env.py:201-203: x ~ 1 + U[0,1) * (start_x_range - 1), y = 1, theta = pi / 2, drawn per episode.  The
            # reference uses the unseeded global `random`; here one seeded generator draws a fresh start for every env before
            # each reset and after each step (an env that auto-resets inside the next step consumes its current one).
            self._start_rng = np.random.default_rng(start_seed)
            self.set_start_poses(self._sample_starts())
        if self._global_obs:
            self._out.obs = None                      # the egocentric kernel is not launched; _observation() renders
        if first_trial is None:
            first_trial = np.arange(E) % self.env_max_trial       # env e starts at trial e % n
        self._next_trial = torch.as_tensor(np.asarray(first_trial, dtype=np.int32), device=dev)
        self._actions = torch.zeros(E, dtype=torch.float32, device=dev)
        if not self.low_dim_state:
            self.observation_shape = self._obs_shape
        self.action_low, self.action_high = -1.0, 1.0
        self.check_status = True

    # ------------------------------------------------------------------ low-dimensional observation
    def _init_spaces(self, action_dtype=np.float32):
        """``single_action_space`` / ``single_observation_space`` = the reference's spaces (ship_ice_env.py:72, 84-99;
        maze_NAMO_env.py:102-124); ``action_space`` / ``observation_space`` = the same with a leading axis of num_envs."""
        from .. import spaces
        self.single_action_space = spaces.box(-1, 1, (), action_dtype)
        if self.low_dim_state:
            self.single_observation_space = spaces.box(-10, 30, (self._low_dim_width,), np.float64)
        else:
            self.single_observation_space = spaces.box(0, 255, tuple(self._obs_shape), np.uint8)
        self.action_space = spaces.batched(self.single_action_space, self.num_envs)
        self.observation_space = spaces.batched(self.single_observation_space, self.num_envs)

    def _init_low_dim(self, width: int):
        """cfg.low_dim_state: the observation is ``generate_observation_low_dim`` (ship_ice_env.py:358-370,
        maze_NAMO_env.py:488-504) — float64 [E, width], zero padded — and the image kernel is not launched."""
        self._low_dim_width = int(width)
        self.low_dim_obs = None
        if self.low_dim_state:
            self.low_dim_obs = torch.zeros((self.num_envs, self._low_dim_width), dtype=torch.float64, device=self.device)
            self._low_dim_counts = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            self._out.obs = None                      # NULL: bp_*_step / reset skip the observation kernel
            self.observation_shape = (self._low_dim_width,)
        self._init_spaces(np.float64 if self._prefix == "maze" else np.float32)

    def low_dim_observation(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``generate_observation_low_dim`` of the current state for every env (available in both observation modes)."""
        if out is None:
            out = torch.empty((self.num_envs, self._low_dim_width), dtype=torch.float64, device=self.device)
        cnt = getattr(self, "_low_dim_counts", None)
        with torch.cuda.device(self.device):
            _lib.check(self._fn("low_dim")(self._h, C.c_void_p(out.data_ptr()),
                                           C.c_void_p(cnt.data_ptr()) if cnt is not None else None, self._stream()))
        return out

    def _observation(self, mask=None):
        if self.low_dim_state:
            return self.low_dim_observation(self.low_dim_obs)
        if getattr(self, "_global_obs", False):
            with torch.cuda.device(self.device):
                mptr = C.c_void_p(mask.data_ptr()) if mask is not None else None
                _lib.check(self._L.bp_shipice_obs_global(self._h, C.c_void_p(self.obs.data_ptr()), mptr, self._stream()))
        return self.obs

    # ------------------------------------------------------------------ cfg.random_start
    def _sample_starts(self) -> np.ndarray:
        E = self.num_envs
        x = 1 + self._start_rng.random(E) * (float(self.cfg.start_x_range) - 1)
        return np.stack([x, np.full(E, 1.0), np.full(E, np.pi / 2)], axis=1)

    def set_start_poses(self, poses: Optional[np.ndarray]):
        """Start pose (x, y, theta) of every env for its following resets (``bp_*_set_start_poses``); ``None`` restores
        the layouts' poses.  ``start_poses`` keeps the host copy."""
        if poses is None:
            self._start_poses, self.start_poses = None, None
            _lib.check(self._fn("set_start_poses")(self._h, None))
            return
        p = np.ascontiguousarray(poses, dtype=np.float64).reshape(self.num_envs, 3)
        self.start_poses = p.copy()
        t = torch.as_tensor(p, device=self.device)
        if self._start_poses is None:
            self._start_poses = t.clone()
            _lib.check(self._fn("set_start_poses")(self._h, C.c_void_p(self._start_poses.data_ptr())))
        else:
            self._start_poses.copy_(t)           # same buffer: ordered after the kernels already enqueued

    # ------------------------------------------------------------------ gym-style API
    @property
    def unwrapped(self):
        return self

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _info(self, reset: bool):
        info = {"state": self._state, "total_work": self._total_work, "obs": _LazyObstacles(self), "pose": self._pose}
        if not reset:
            info.update({"collision reward": self._coll, "scaled collision reward": self._scoll, "dist reward": self._dist,
                         "trial_success": self._success.bool()})
        return info

    def _raise_on_status(self):
        if not self.check_status:
            return
        bad = int((self._status & 7).max().item())
        if bad:
            raise _lib.BenchPushError(f"per-env capacity overflow (status bits {bad}): raise pair_capacity / live_capacity")

    def reset(self, seed=None, options=None, mask: Optional[torch.Tensor] = None):
        """Reset all envs (or those with ``mask[e]`` set) to their next trial.  ``seed`` is accepted and
        ignored, as in the reference (ship_ice_env.py:223): layouts are deterministic inputs."""
        with torch.cuda.device(self.device):
            mptr = None
            if mask is not None:
                mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
                mptr = C.c_void_p(mask.data_ptr())
            if getattr(self, "random_start", False) and self.redraw_starts:
                self.set_start_poses(self._sample_starts())
            _lib.check(self._fn("reset")(self._h, C.c_void_p(self._next_trial.data_ptr()), mptr, C.byref(self._out), self._stream()))
            if mask is None:
                self._next_trial = (self._next_trial + 1) % self.env_max_trial
            else:
                self._next_trial = torch.where(mask.bool(), (self._next_trial + 1) % self.env_max_trial, self._next_trial)
        self._raise_on_status()
        return self._observation(mask), self._info(reset=True)

    def step(self, action, mask: Optional[torch.Tensor] = None):
        """``mask`` (bool/uint8 [E], optional): step only the envs with mask set; the others keep their
        state and their previous outputs (no counterpart in the single-env reference)."""
        with torch.cuda.device(self.device):
            a = torch.as_tensor(action, dtype=torch.float32, device=self.device).reshape(self.num_envs)
            self._actions.copy_(a)
            self.step_async_nocheck(self._actions, mask)
            if getattr(self, "random_start", False) and self.auto_reset and self.redraw_starts:
                self.set_start_poses(self._sample_starts())
        self._raise_on_status()
        return (self._observation(getattr(self, "_mask", None) if mask is not None else None), self._reward, self._terminated.bool(),
                self._truncated.bool(), self._info(reset=False))

    def step_async_nocheck(self, action_dev: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """Enqueue one step without any host synchronisation (actions already on the device)."""
        mptr = None
        if mask is not None:
            self._mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()      # kept alive until the next call
            mptr = C.c_void_p(self._mask.data_ptr())
        _lib.check(self._fn("step_masked")(self._h, C.c_void_p(action_dev.data_ptr()), mptr, C.byref(self._out), self._stream()))

    def step_host(self, actions: np.ndarray, want_obs=False):
        """The same step through host buffers (numpy in, numpy out; copies inside).  ``want_obs``: False = the observation
        kernel is not launched; True = the image is rendered and copied to the host; "device" = the image is rendered into the
        env's device tensor (``self.obs``, for a policy that runs on the GPU) and not copied (ship-ice egocentric image only)."""
        E = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(E)
        if not hasattr(self, "_host"):
            self._host = {"reward": np.zeros(E), "terminated": np.zeros(E, np.uint8), "truncated": np.zeros(E, np.uint8),
                          "state": np.zeros((E, 3)), "total_work": np.zeros(E), "trial_success": np.zeros(E, np.uint8),
                          "status": np.zeros(E, np.int32), "obs": None}
        hb = self._host
        if want_obs and (self.low_dim_state or getattr(self, "_global_obs", False)):
            raise NotImplementedError("step_host returns the egocentric image only; use step() for the low-dimensional / global observation")
        if want_obs is True and hb["obs"] is None:
            hb["obs"] = np.zeros((E,) + self._obs_shape, np.uint8)
        o = _lib.ShipIceOutputs()
        o.reward = hb["reward"].ctypes.data; o.terminated = hb["terminated"].ctypes.data; o.truncated = hb["truncated"].ctypes.data
        o.state = hb["state"].ctypes.data; o.total_work = hb["total_work"].ctypes.data
        o.trial_success = hb["trial_success"].ctypes.data; o.status = hb["status"].ctypes.data
        if want_obs is True:
            o.obs = hb["obs"].ctypes.data
        if want_obs == "device":
            _lib.check(self._fn("step_host_obs_device")(self._h, a.ctypes.data, C.byref(o), C.c_void_p(self.obs.data_ptr())))
        else:
            _lib.check(self._fn("step_host")(self._h, a.ctypes.data, C.byref(o)))
        return hb

    def update_path(self, new_path):
        """ship_ice_env.py:373-375: a planner hands its current path to the env for the renderer; kept in ``self.path``
        (there is no renderer here)."""
        self.path = new_path

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self):
        raise NotImplementedError("rendering is out of scope for the batched simulator")

    # ------------------------------------------------------------------ inspection helpers
    def get_obstacles(self, e: int):
        F = self.layouts.max_floes
        verts = np.zeros((F, _lib.BP_VMAX, 2), dtype=np.float64)
        nv = np.zeros(F, dtype=np.int32)
        n = _lib.check(self._fn("get_obstacles")(self._h, e, verts.ctypes.data, nv.ctypes.data))
        return [verts[i, : nv[i]].copy() for i in range(n)]

    def get_bodies(self, e: int) -> np.ndarray:
        out = np.zeros((self.layouts.max_floes + 1, 6), dtype=np.float64)
        n = _lib.check(self._fn("get_bodies")(self._h, e, out.ctypes.data))
        return out[:n]

    def get_state(self, e: int) -> np.ndarray:
        """Checkpoint of env ``e`` (``bp_*_get_state``): its whole state block as bytes; feed it to :meth:`set_state` of a
        handle with the same configuration and layouts to resume exactly (bit-identical continuation)."""
        buf = np.zeros(int(self._fn("state_bytes")(self._h)), dtype=np.uint8)
        _lib.check(self._fn("get_state")(self._h, e, buf.ctypes.data))
        return buf

    def set_state(self, e: int, state: np.ndarray):
        s = np.ascontiguousarray(state, dtype=np.uint8)
        if s.size != int(self._fn("state_bytes")(self._h)):
            raise ValueError("state block of %d bytes does not belong to this handle (%d)" % (s.size, int(self._fn("state_bytes")(self._h))))
        _lib.check(self._fn("set_state")(self._h, e, s.ctypes.data))

    def set_bodies(self, e: int, bodies: np.ndarray):
        """Put env ``e`` into an externally given body state ([n_bodies, 6] = x, y, angle, vx, vy, w in :meth:`get_bodies`
        order), e.g. one taken from the oracle in the middle of an episode; the arbiter cache starts empty."""
        b = np.ascontiguousarray(bodies, dtype=np.float64)
        _lib.check(self._fn("set_bodies")(self._h, e, b.ctypes.data, b.shape[0]))

    def counters(self, e: int) -> dict:
        o = np.zeros(32, dtype=np.int64)
        _lib.check(self._L.bp_shipice_get_counters(self._h, e, o.ctypes.data))
        keys = ["stamp", "n_awake", "n_pairs", "n_free", "n_order", "n_colours", "status", "n_apairs", "n_cached", "gjk",
                "pairs_eval", "live_sum", "awake_sum", "recolour", "rebuild", "substeps", "epa", "skipped"]
        out = dict(zip(keys, o[:18].tolist()))
        out["phase_clocks"] = o[18:26].tolist()
        return out

    def set_profiling(self, on: bool):
        _lib.check(self._L.bp_shipice_set_profiling(self._h, 1 if on else 0))

    def kernel_ms(self) -> dict:
        o = np.zeros(4, dtype=np.float64)
        _lib.check(self._L.bp_shipice_get_kernel_ms(self._h, o.ctypes.data))
        return {"step_ms": o[0], "step_launches": int(o[1]), "obs_ms": o[2], "obs_launches": int(o[3])}

    def env_bytes(self) -> int:
        return int(self._L.bp_shipice_env_bytes(self._h))

    def launch_count(self) -> int:
        return int(self._L.bp_shipice_launch_count(self._h))
