"""Single-environment view with the reference's exact return types, for callers written against
``ShipIceEnv`` (ship_ice_env.py:37-355): ``reset() -> (np.uint8[4,150,150], info)``,
``step(a) -> (obs, float, bool, bool, info)`` with ``info`` holding python / numpy values
(ship_ice_env.py:241-247, 337-345).  It owns a ``ShipIceVecEnv`` with one env; nothing is simulated
on the host.  gymnasium is not imported (absent from this image); the class is duck-typed and can be
registered by a gymnasium user."""
from __future__ import annotations

import numpy as np

from .ship_ice import ShipIceVecEnv


def _obs0(v, obs):
    """Observation of env 0 as numpy; a low-dimensional observation is cut to its valid length (the batched
    tensor is zero padded to the largest layout)."""
    if getattr(v, "low_dim_state", False):
        return obs[0, : int(v._low_dim_counts[0])].cpu().numpy()
    return obs[0].cpu().numpy()


class ShipIceEnvB200:
    metadata = {"render_modes": []}

    def __init__(self, cfg=None, **kw):
        self._v = ShipIceVecEnv(num_envs=1, cfg=cfg, **kw)
        self.cfg = self._v.cfg
        self.goal = self._v.goal
        self.max_yaw_rate_step = self._v.max_yaw_rate_step
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.beta = self._v.beta
        self.num_channels = self._v.num_channels
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.episode_idx = None
        self.t = 0
        self.total_work = [0, []]

    @property
    def unwrapped(self):
        return self

    def _info(self, info, reset):
        out = {"state": tuple(float(x) for x in info["state"][0].tolist()),
               "total_work": float(info["total_work"][0]) if not reset else 0,
               "obs": info["obs"][0]}
        if not reset:
            out.update({"collision reward": float(info["collision reward"][0]),
                        "scaled collision reward": float(info["scaled collision reward"][0]),
                        "dist reward": float(info["dist reward"][0]),
                        "trial_success": bool(info["trial_success"][0])})
        return out

    def reset(self, seed=None, options=None):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        obs, info = self._v.reset()
        self.t = 0
        self.total_work = [0, []]
        return _obs0(self._v, obs), self._info(info, True)

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1)
        obs, r, term, trunc, info = self._v.step(a)
        self.t += 1
        out = self._info(info, False)
        self.total_work[0] = out["total_work"]
        return _obs0(self._v, obs), float(r[0]), bool(term[0]), bool(trunc[0]), out

    def render(self):
        raise NotImplementedError("rendering is out of scope for the batched simulator")

    def update_path(self, new_path):
        """ship_ice_env.py:373-375 (the lattice-planner baseline calls it every step)."""
        self._v.update_path(new_path)
        self.path = new_path

    def close(self):
        self._v.close()


class MazeNamoEnvB200:
    """Single-environment view of :class:`MazeNamoVecEnv` with the reference's return types
    (``MazeNAMO``, maze_NAMO_env.py:325-359, 402-485): obs ``np.uint8[4,192,192]``, python scalars in ``info``."""
    metadata = {"render_modes": []}

    def __init__(self, cfg=None, **kw):
        from .maze_namo import MazeNamoVecEnv
        self._v = MazeNamoVecEnv(num_envs=1, cfg=cfg, **kw)
        self.cfg = self._v.cfg
        self.goal = self._v.goal
        self.max_yaw_rate_step = self._v.max_yaw_rate_step
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.beta = self._v.beta
        self.maze_walls = self._v.maze_walls
        self.episode_idx = None
        self.t = 0
        self.total_work = [0, []]

    @property
    def unwrapped(self):
        return self

    def _info(self, info, reset):
        out = {"state": tuple(float(x) for x in info["state"][0].tolist()),
               "total_work": float(info["total_work"][0]) if not reset else 0, "obs": info["obs"][0]}
        if reset:
            out.update({"box_count": 0, "goal_dt": info["goal_dt"], "m_to_pix_scale": info["m_to_pix_scale"]})
        else:
            out.update({"collision reward": float(info["collision reward"][0]),
                        "scaled collision reward": float(info["scaled collision reward"][0]),
                        "dist increment reward": float(info["dist increment reward"][0]),
                        "trial_success": bool(info["trial_success"][0])})
        return out

    def reset(self, seed=None, options=None):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        obs, info = self._v.reset()
        self.t = 0
        self.total_work = [0, []]
        return _obs0(self._v, obs), self._info(info, True)

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1)
        obs, r, term, trunc, info = self._v.step(a)
        self.t += 1
        out = self._info(info, False)
        self.total_work[0] = out["total_work"]
        return _obs0(self._v, obs), float(r[0]), bool(term[0]), bool(trunc[0]), out

    def render(self):
        raise NotImplementedError("rendering is out of scope for the batched simulator")

    def close(self):
        self._v.close()


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


class BoxDeliveryEnvB200:
    """Single-environment view of :class:`BoxDeliveryVecEnv` with the reference's return types (``BoxDeliveryEnv``,
    box_delivery_env.py:571-632, 634-830): obs ``np.uint8[224,224,4]``, python scalars / lists in ``info``
    (``state``, ``cumulative_distance``, ``cumulative_boxes``, ``cumulative_reward``, ``total_work``, ``obs`` = list of
    (4, 2) vertex arrays of the boxes left, ``box_completed_statuses``, ``goal_positions``, ``ministeps``, ``inactivity``).
    ``TaskDrivenMetric`` of the reference reads ``goal.x / goal.y`` of ``goal_positions``; a small point type provides them."""
    metadata = {"render_modes": []}

    class _Point(tuple):
        x = property(lambda self: self[0])
        y = property(lambda self: self[1])

    @staticmethod
    def _vec_class():
        from .box_delivery import BoxDeliveryVecEnv
        return BoxDeliveryVecEnv

    def __init__(self, cfg=None, **kw):
        self._v = self._vec_class()(num_envs=1, cfg=cfg, **kw)
        self.cfg = self._v.cfg
        self.num_channels = self._v.num_channels
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.local_map_pixel_width = self._v.local_map_pixel_width
        self.episode_idx = None

    @property
    def unwrapped(self):
        return self

    def _info(self, info):
        n = int(_np(info["n_boxes_left"])[0])
        st = _np(info["state"])[0]
        return {"state": (float(st[0]), float(st[1]), float(st[2])),
                "cumulative_distance": float(_np(info["cumulative_distance"])[0]), "cumulative_boxes": int(_np(info["cumulative_boxes"])[0]),
                "cumulative_reward": float(_np(info["cumulative_reward"])[0]), "total_work": float(_np(info["total_work"])[0]),
                "obs": [v.copy() for v in _np(info["obs"])[0][:n]],
                "box_completed_statuses": [bool(b) for b in _np(info["box_completed_statuses"])[0]],
                "goal_positions": [self._Point(g) for g in info["goal_positions"]],
                "ministeps": float(_np(info["ministeps"])[0]), "inactivity": int(_np(info["inactivity"])[0])}

    def reset(self, seed=None, options=None):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        obs, info = self._v.reset()
        return _np(obs)[0].copy(), self._info(info)

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1)
        obs, r, term, trunc, info = self._v.step(a)
        return _np(obs)[0].copy(), float(_np(r)[0]), bool(_np(term)[0]), bool(_np(trunc)[0]), self._info(info)

    def render(self):
        raise NotImplementedError("rendering is out of scope for the batched simulator")

    def close(self):
        self._v.close()


class AreaClearingEnvB200(BoxDeliveryEnvB200):
    """Single-environment view of :class:`AreaClearingVecEnv` (``AreaClearingEnv``, area_clearing.py:563-778): info keys
    ``state``, ``total_work``, ``collision reward``, ``diff_reward``, ``box_completed_reward``, ``obs``,
    ``box_completed_statuses``, ``box_count``, ``ministeps`` (+ ``goal_positions`` / ``boundary`` as at reset)."""

    @staticmethod
    def _vec_class():
        from .area_clearing import AreaClearingVecEnv
        return AreaClearingVecEnv

    def __init__(self, cfg=None, **kw):
        self._v = self._vec_class()(num_envs=1, cfg=cfg, **kw)
        self.cfg = self._v.cfg
        self.num_channels = self._v.num_channels
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.local_map_pixel_width = self._v.local_map_pixel_width
        self.episode_idx = None

    def configure_env_for_SAM(self):
        """area_clearing.py:221-223: switch to the 96-pixel observation of the spatial-action-map baselines (call before reset)."""
        self._v.configure_env_for_SAM()
        self.action_space, self.observation_space = self._v.single_action_space, self._v.single_observation_space
        self.local_map_pixel_width = self._v.local_map_pixel_width

    def update_path(self, new_path):
        """area_clearing.py:922-925."""
        self._v.update_path(new_path)
        self.path = new_path

    def activate_demo_mode(self):
        """area_clearing.py:266-272 prepares keyboard tele-operation of the rendered env: interactive, out of scope."""
        raise NotImplementedError("demo / tele-operation mode needs the interactive renderer, which the batched simulator does not have")

    def _info(self, info):
        n = int(_np(info["n_boxes_left"])[0])
        st = _np(info["state"])[0]
        return {"state": (float(st[0]), float(st[1]), float(st[2])), "total_work": float(_np(info["total_work"])[0]),
                "collision reward": float(_np(info["collision reward"])[0]), "diff_reward": float(_np(info["diff_reward"])[0]),
                "box_completed_reward": float(_np(info["box_completed_reward"])[0]),
                "obs": [v.copy() for v in _np(info["obs"])[0][:n]],
                "box_completed_statuses": [bool(b) for b in _np(info["box_completed_statuses"])[0][:n]],
                "box_count": int(_np(info["box_count"])[0]), "ministeps": float(_np(info["ministeps"])[0]),
                "goal_positions": [self._Point(g) for g in info["goal_positions"]], "boundary": info["boundary"]}
