"""ORACLE — test infrastructure only.

CPU restatement of ``ShipIceEnv`` (reference benchpush/environments/ship_ice_nav/ship_ice_env.py)
without gymnasium / pymunk / matplotlib / skimage: same reset/step flow, same constants, same
info keys; physics through the Chipmunk restatement (oracle/chipmunk_oracle.cpp), rasters through
oracle/raster_oracle.py.  PARITY UNPINNED for the physics (see chipmunk_oracle.cpp header).

The ice fields are INPUTS: ``experiment`` is the reference's ``ddict['exp'][concentration]``
structure (dict/list of trials ``{goal, ship_state, obstacles:[{vertices, centre, ...}]}``,
ship_ice_env.py:79-80) — loaded from the reference pickles where present, or from the committed
fixtures / the synthetic generator elsewhere.
"""
import math
import numpy as np

from . import cpo
from .raster_oracle import OccupancyGrid, poly_area, poly_centroid

YAW_CONSTRAINT_PENALTY = 0      # ship_ice_env.py:29-31
BOUNDARY_PENALTY = -50
TERMINAL_REWARD = 200

# ship_ice_nav/config.yaml (values are the reference's defaults; restated, not loaded from the tree)
DEFAULT_CFG = {
    "dt": 0.8, "target_speed": 0.3,
    "low_dim_state": False, "fixed_trial_idx": 4, "random_start": False, "start_x_range": 7,
    "ship": {
        "mass": 1, "padding": 0.25,
        "vertices": [[1., -0.], [0.9, 0.10], [0.5, 0.25], [0.25, 0.25], [0, 0.25], [-0.25, 0.25],
                     [-0.5, 0.25], [-0.75, 0.25], [-1., 0.25], [-1., -0.25], [-0.75, -0.25],
                     [-0.5, -0.25], [-0.25, -0.25], [0, -0.25], [0.25, -0.25], [0.5, -0.25], [0.9, -0.10]],
        "head": [1., 0.], "tail": [-1., 0.],
    },
    "sim": {"steps": 400, "gravity": (0, 0), "iterations": 10, "damping": 0, "obstacle_density": 0.001},
    "log_obs": False, "render_snapshot": False,
    "occ": {"m_to_pix_scale": 25, "map_height": 40, "map_width": 12},
    "output_dir": "logs/", "egocentric_obs": True, "concentration": 0.1, "goal_y": 9, "render_scale": 40,
}


def R(theta):
    return np.asarray([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])


def merge_cfg(base, cfg):
    """One-level-deep merge (ship_ice_env.py:47-56)."""
    out = {k: (dict(v) if isinstance(v, dict) else v) for k, v in base.items()}
    if cfg is not None:
        for k in cfg:
            if isinstance(cfg[k], dict):
                if k not in out:
                    out[k] = {}
                for p in cfg[k]:
                    out[k][p] = cfg[k][p]
            else:
                out[k] = cfg[k]
    return out


def total_work_done(obs_initial, obs_final):
    """evaluation/metrics.py:96-126 (mass=None branch)."""
    work = 0
    for ob_a, ob_b in zip(obs_initial, obs_final):
        area = poly_area(ob_a)
        centre_a = np.abs(poly_centroid(ob_a))
        centre_b = np.abs(poly_centroid(ob_b))
        work += ((centre_a[0] - centre_b[0]) ** 2 + (centre_a[1] - centre_b[1]) ** 2) ** 0.5 * area
    return work


class ShipIceOracleEnv:
    def __init__(self, cfg=None, experiment=None, order_mode=0):
        self.cfg = merge_cfg(DEFAULT_CFG, cfg)
        self.local_window_v_shift = 2
        self.beta = 30
        self.directional_reward_scale = 1.0
        self.episode_idx = None
        self.goal = (0, self.cfg["goal_y"])
        self.max_yaw_rate_step = (np.pi / 2) / 7
        assert experiment is not None, "ice fields are inputs to the oracle"
        self.experiment = experiment
        self.env_max_trial = len(self.experiment)
        occ = self.cfg["occ"]
        # ship_ice_env.py:91-97: cells of 1 / scale for the egocentric image, 0.2 m cells for the global observation
        grid_size = 1 / occ["m_to_pix_scale"] if self.cfg["egocentric_obs"] else 0.2
        self.occupancy = OccupancyGrid(grid_width=grid_size, grid_height=grid_size, map_width=occ["map_width"],
                                       map_height=occ["map_height"], local_width=6, local_height=6,
                                       meter_to_pixel_scale=occ["m_to_pix_scale"])
        self.yaw_lim = (0, np.pi)
        self.boundary_violation_limit = 0.0
        self.order_mode = order_mode
        import random as _random
        self.rng = _random
        self.start_override = None

    def clone(self, order_mode=None):
        """Independent copy of the env in its current (mid-episode) state, optionally with another solver order."""
        import copy
        c = copy.copy(self)
        c.space = self.space.clone()
        c.total_work = [self.total_work[0], list(self.total_work[1])]
        if order_mode is not None:
            c.order_mode = order_mode
            c.space.set_order(order_mode)
        return c

    # ship_ice_env.py:109-183 (the KE / impulse / contact bookkeeping of the handlers has no reader)
    def init_ship_ice_sim(self):
        self.steps = self.cfg["sim"]["steps"]
        self.dt = self.cfg["dt"]
        self.target_speed = self.cfg["target_speed"]
        self.space = cpo.Space(iterations=self.cfg["sim"]["iterations"], damping=self.cfg["sim"]["damping"],
                               gravity=self.cfg["sim"]["gravity"])
        self.space.set_order(self.order_mode)
        self.total_work = [0, []]

    # ship_ice_env.py:186-220
    def init_ship_ice_env(self, settle=True):
        trial_idx = self.episode_idx % self.env_max_trial
        if self.cfg["low_dim_state"]:           # ship_ice_env.py:189-190: always the fixed trial
            trial_idx = self.cfg["fixed_trial_idx"]
        trial = self.experiment[trial_idx]
        self.start = trial["ship_state"]
        if self.cfg["random_start"]:            # ship_ice_env.py:201-203 (the reference draws from the global `random` module)
            x_start = 1 + self.rng.random() * (self.cfg["start_x_range"] - 1)
            self.start = (x_start, 1.0, np.pi / 2)
            if self.start_override is not None:  # tests: replay the start the CUDA path drew
                self.start = tuple(self.start_override)
        obs_dicts = [ob for ob in trial["obstacles"] if poly_area(np.asarray(ob["vertices"])) != 0]
        self.obs_dicts = obs_dicts
        self.obstacles = [np.asarray(ob["vertices"]) for ob in obs_dicts]
        # sim_utils.py:156-163 generate_sim_obs -> :136-153 create_polygon
        self.poly_shapes = []
        self.poly_bodies = []
        density = self.cfg["sim"]["obstacle_density"]
        for ob in obs_dicts:
            centre = np.array(ob["centre"], dtype=np.float64)
            verts = (np.asarray(ob["vertices"], dtype=np.float64) - centre)
            b = self.space.add_body(cpo.DYNAMIC, centre[0], centre[1], 0.0)
            cog = cpo.poly_centroid_hull(verts)
            vs = verts - cog
            s = self.space.add_poly(b, vs, 0.02)
            self.space.set_density(s, density)
            self.space.set_material(s, 0.01, 1.0, 2)
            self.poly_shapes.append(s)
            self.poly_bodies.append(b)
        # ship.py:77-98 Ship.sim: kinematic, Poly(radius=0.02), mass 10 (ignored: kinematic), e .01, mu 1
        x, y, theta = self.start
        self.ship_body = self.space.add_body(cpo.KINEMATIC, x, y, theta)
        self.ship_shape = self.space.add_poly(self.ship_body, np.asarray(self.cfg["ship"]["vertices"], dtype=np.float64), 0.02)
        self.space.set_mass(self.ship_shape, 10.0)
        self.space.set_material(self.ship_shape, 0.01, 1.0, 1)
        self.local_verts = [self.space.get_vertices(s) for s in self.poly_shapes]
        if settle:
            for _ in range(1000):
                self.space.step(self.dt / self.steps)
        self.prev_obs = self.get_obs_from_poly()

    # cost_map.py:274-281
    def get_obs_from_poly(self):
        st = self.space.state()
        out = []
        for k, b in enumerate(self.poly_bodies):
            x, y, a = st[b, 0], st[b, 1], st[b, 2]
            out.append(self.local_verts[k] @ R(a).T + np.array([x, y]))
        return out

    def ship_pose(self):
        b = self.space.body(self.ship_body)
        return (b[0], b[1], b[2])

    def reset(self, seed=None, options=None, settle=True):
        if self.episode_idx is None:
            self.episode_idx = 0
        else:
            self.episode_idx += 1
        self.init_ship_ice_sim()
        self.init_ship_ice_env(settle=settle)
        self.t = 0
        updated = self.get_obs_from_poly()
        # NB: the reset observation rasterises the RAW trial vertices (self.obstacles, ship_ice_env.py:207,248)
        px, py, pa = self.ship_pose()
        info = {"state": (round(px, 2), round(py, 2), round(pa, 2)), "total_work": self.total_work[0], "obs": updated}
        if self.cfg["low_dim_state"]:           # ship_ice_env.py:244-245
            return self.generate_observation_low_dim(updated), info
        return self.generate_observation(), info

    def step(self, action, want_obs=True):
        self.t += 1
        action = action * self.max_yaw_rate_step
        px, py, pa = self.ship_pose()
        if cpo.portable_trig():
            sn, cs = cpo.sincos(pa)          # same R(pa) @ [speed, 0], with the portable cos / sin
            gv = [cs * self.target_speed + (-sn) * 0.0, sn * self.target_speed + cs * 0.0]
        else:
            gv = R(pa) @ [self.target_speed, 0]
        self.space.set_angular_velocity(self.ship_body, action)
        self.space.set_velocity(self.ship_body, gv[0], gv[1])
        flags = self.space.shipice_substeps(self.ship_body, self.steps, self.dt / self.steps,
                                            self.yaw_lim[0], self.yaw_lim[1], self.cfg["occ"]["map_width"])
        yaw_violated = bool(flags & 1)
        boundary_violated = bool(flags & 2)
        px, py, pa = self.ship_pose()
        W = self.cfg["occ"]["map_width"]
        boundary_terminal = False
        if px < 0 and abs(px - 0) >= self.boundary_violation_limit:
            boundary_terminal = True
        if px > W and abs(px - W) >= self.boundary_violation_limit:
            boundary_terminal = True
        updated = self.get_obs_from_poly()
        work = total_work_done(self.prev_obs, updated)
        self.total_work[0] += work
        self.total_work[1].append(work)
        self.prev_obs = updated
        self.obstacles = updated
        if py >= self.goal[1]:
            terminated = True
        elif boundary_terminal:
            terminated = True
        else:
            terminated = False
        if py < self.goal[1]:
            dist_reward = self.directional_reward_scale * float(np.dot(np.array([np.cos(pa), np.sin(pa)]), np.array([0, 1])))
        else:
            dist_reward = 0
        collision_reward = -work
        reward = self.beta * collision_reward + dist_reward
        if yaw_violated:
            reward += YAW_CONSTRAINT_PENALTY
        if boundary_violated:
            reward += BOUNDARY_PENALTY
        trial_success = False
        if terminated and not boundary_terminal:
            reward += TERMINAL_REWARD
            trial_success = True
        info = {"state": (round(px, 2), round(py, 2), round(pa, 2)), "total_work": self.total_work[0],
                "collision reward": collision_reward, "scaled collision reward": collision_reward * self.beta,
                "dist reward": dist_reward, "trial_success": trial_success, "obs": updated}
        if self.cfg["low_dim_state"]:           # ship_ice_env.py:347-348
            observation = self.generate_observation_low_dim(updated)
        else:
            observation = self.generate_observation() if want_obs else None
        return observation, reward, terminated, False, info

    # ship_ice_env.py:358-370
    def generate_observation_low_dim(self, updated_obstacles):
        observation = np.zeros((len(updated_obstacles) * 2))
        for i in range(len(updated_obstacles)):
            center = np.abs(poly_centroid(updated_obstacles[i]))
            observation[i * 2] = center[0]
            observation[i * 2 + 1] = center[1]
        return observation

    # ship_ice_env.py:378-409
    def generate_observation(self):
        occ = self.cfg["occ"]
        pose = self.ship_pose()
        if not self.cfg["egocentric_obs"]:      # :397-407: 5 x 5 block-mean concentration of all floes + planner footprint, 200 x 60 cells
            raw = self.occupancy.compute_occ_img(self.obstacles, ice_binary_w=int(self.occupancy.map_width * occ["m_to_pix_scale"]),
                                                 ice_binary_h=int(self.occupancy.map_height * occ["m_to_pix_scale"]))
            occupancy = np.copy(self.occupancy.compute_con_gridmap(raw))
            self.occupancy.compute_ship_footprint_planner(pose, self.cfg["ship"]["vertices"])
            footprint = np.copy(self.occupancy.footprint)
            observation = np.concatenate((np.array([occupancy]), np.array([footprint])))
            return (observation * 255).astype(np.uint8)
        raw = self.occupancy.compute_occ_img(self.obstacles, ice_binary_w=int(self.occupancy.map_width * occ["m_to_pix_scale"]),
                                             ice_binary_h=int(self.occupancy.map_height * occ["m_to_pix_scale"]),
                                             local_range=[12, 12], ship_state=pose)
        occupancy = self.occupancy.ego_view_obstacle_map(raw, pose, self.local_window_v_shift)
        footprint = self.occupancy.ego_view_footprint(pose, self.cfg["ship"]["vertices"], self.local_window_v_shift)
        edt = self.occupancy.ego_view_goal_dist_transform(self.goal[1], pose, self.local_window_v_shift)
        orient = self.occupancy.ego_view_orientation_map(pose, self.cfg["ship"]["head"], self.cfg["ship"]["tail"],
                                                         self.local_window_v_shift)
        observation = np.concatenate((np.array([footprint]), np.array([edt]), np.array([orient]), np.array([occupancy])))
        return (observation * 255).astype(np.uint8)


class ShipIceMetricOracle:
    """common/metrics/ship_ice_metric.py:5-72 (+ the list fields of base_metric.py:12-19)."""

    def __init__(self, alg_name, ship_mass, goal):
        self.rewards, self.efficiency_scores, self.effort_scores, self.success_rates = [], [], [], []
        self.alg_name = alg_name
        self.eps_reward = 0
        self.total_mass_dist = 0
        self.ship_mass = ship_mass
        self.total_ship_dist = 0
        self.goal_line = goal[1]

    def compute_efficiency_score(self):
        if not self.trial_success:
            return 0
        return self.L / self.total_ship_dist

    def compute_effort_score(self):
        return (self.ship_mass * self.total_ship_dist) / (self.ship_mass * self.total_ship_dist + self.total_mass_dist)

    def update(self, info, reward, eps_complete=False):
        self.eps_reward += reward
        self.total_mass_dist = info["total_work"]
        self.trial_success = info["trial_success"]
        ship_state = info["state"]
        self.total_ship_dist += np.linalg.norm(np.array(self.ship_state[:2]) - np.array(ship_state[:2]))
        self.ship_state = ship_state
        if eps_complete:
            self.rewards.append(self.eps_reward)
            self.efficiency_scores.append(self.compute_efficiency_score())
            self.effort_scores.append(self.compute_effort_score())

    def reset(self, info):
        self.eps_reward = 0
        self.total_mass_dist = 0
        self.total_ship_dist = 0
        self.trial_success = False
        self.ship_state = info["state"]
        self.L = self.goal_line - self.ship_state[1]
