"""The env layer of the oracle (and the product's host-side metric / cfg code) against vectors produced by the REFERENCE'S OWN
modules, imported from /root/reference by tools/make_ref_fixtures.py (tests/golden/ref_envlayer.npz).  Third-party primitives
that are absent in the build image (skimage.draw.polygon / line, shapely) were stubbed with the oracle's restatements there, so
those primitives stay unpinned; everything around them is the reference's code run as is.  Physics (Chipmunk2D) is not covered
by these pins."""
import os
import types

import numpy as np
import pytest
import torch

from tests.conftest import GOLD, SHIP_VERTS
from oracle import raster_oracle as ro
from oracle.ship_ice_oracle import ShipIceMetricOracle, total_work_done, R
from oracle.maze_oracle import MazeNamoMetricOracle, MazeOccupancyGrid


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_envlayer.npz"))


def _split(flat, nv):
    out, o = [], 0
    for n in nv:
        out.append(flat[o:o + int(n)]); o += int(n)
    return out


def test_polygon_area_centroid_work(ref):
    a, b = _split(ref["poly_a"], ref["poly_nv"]), _split(ref["poly_b"], ref["poly_nv"])
    np.testing.assert_allclose([ro.poly_area(p) for p in a], ref["poly_area"], rtol=1e-13)          # polygon.py:25-29
    np.testing.assert_allclose([ro.poly_centroid(p) for p in a], ref["poly_centroid"], rtol=1e-12)  # polygon.py:32-41 (abs-folded)
    np.testing.assert_allclose(total_work_done(a, b), ref["total_work"][0], rtol=1e-12)             # evaluation/metrics.py:96-126
    np.testing.assert_allclose(total_work_done(a[:7], b[:7]), ref["total_work"][1], rtol=1e-12)


def test_push_reward_obs_to_goal_difference(ref):
    from oracle.area_clearing_oracle import AreaClearingOracleEnv
    a, b = _split(ref["poly_a"], ref["poly_nv"]), _split(ref["poly_b"], ref["poly_nv"])
    env = types.SimpleNamespace(boundary_vertices=ref["square"], goal_points=[tuple(g) for g in ref["goal_pts"]])
    f = AreaClearingOracleEnv.obs_to_goal_difference                                                # evaluation/metrics.py:73-94
    np.testing.assert_allclose(f(env, a, b), ref["push_diff"][0], rtol=1e-12)
    env.goal_points = env.goal_points[:3]
    np.testing.assert_allclose(f(env, a[5:20], b[5:20]), ref["push_diff"][1], rtol=1e-12)


def test_get_obs_from_poly(ref):
    loc = _split(ref["gop_local"], ref["gop_nv"])
    got = np.concatenate([l @ R(p[2]).T + p[:2] for l, p in zip(loc, ref["gop_pose"])])             # cost_map.py:274-281
    np.testing.assert_allclose(got, ref["gop_world"], rtol=0, atol=1e-14)


def test_cfg_merge_one_level(ref):
    from benchpush_b200.utils import DotDict, merge_cfg
    base = DotDict.to_dot_dict({"a": 1, "sim": {"steps": 400, "gravity": (0, 0)}, "ship": {"mass": 1}})
    c = merge_cfg(base, {"sim": {"steps": 10}, "b": 2, "new": {"k": 7}})                           # ship_ice_env.py:47-56
    assert [c.a, c.sim.steps, c.sim.gravity[0], c.ship.mass, c.b, c.new.k] == ref["dotdict"].tolist()


def test_ship_ice_metric(ref):
    from benchpush_b200.metrics import ShipIceMetric
    st, work, rew = ref["sm_state"], ref["sm_work"], ref["sm_rew"]
    T = len(work)
    o = ShipIceMetricOracle("pin", ship_mass=1.0, goal=(0, 9.0))
    p = ShipIceMetric("pin", ship_mass=1.0, goal=(0, 9.0))
    for succ in (True, False):
        o.reset({"state": tuple(st[0])})
        p.reset({"state": torch.tensor(st[:1])})
        for t in range(T):
            s = succ and t == T - 1
            o.update({"state": tuple(st[t + 1]), "total_work": float(work[t]), "trial_success": s}, float(rew[t]), eps_complete=(t == T - 1))
            p.update({"state": torch.tensor(st[t + 1:t + 2]), "total_work": torch.tensor([work[t]]), "trial_success": torch.tensor([s])},
                     torch.tensor([rew[t]]), eps_complete=torch.tensor([t == T - 1]))
    for m in (o, p):                                                                                # ship_ice_metric.py:26-72
        np.testing.assert_allclose([m.efficiency_scores, m.effort_scores, m.rewards], ref["sm_scores"], rtol=1e-12)


def test_maze_namo_metric(ref):
    from benchpush_b200.metrics import MazeNamoMetric
    st, work, rew, gdt = ref["mm_state"], ref["sm_work"], ref["sm_rew"], ref["mm_goal_dt"]
    T = len(work)
    o = MazeNamoMetricOracle("pin", robot_mass=2.5)
    p = MazeNamoMetric("pin", robot_mass=2.5)
    for succ in (True, False):
        o.reset({"state": tuple(st[0]), "goal_dt": gdt, "m_to_pix_scale": 16})
        p.reset({"state": torch.tensor(st[:1]), "goal_dt": gdt, "m_to_pix_scale": 16})
        for t in range(T):
            s = succ and t == T - 1
            o.update({"state": tuple(st[t + 1]), "total_work": float(work[t]), "trial_success": s}, float(rew[t]), eps_complete=(t == T - 1))
            p.update({"state": torch.tensor(st[t + 1:t + 2]), "total_work": torch.tensor([work[t]]), "trial_success": torch.tensor([s])},
                     torch.tensor([rew[t]]), eps_complete=torch.tensor([t == T - 1]))
    for m in (o, p):                                                                                # maze_namo_metric.py:26-75
        np.testing.assert_allclose([m.efficiency_scores, m.effort_scores, m.rewards], ref["mm_scores"], rtol=1e-12)


def test_task_driven_metric(ref):
    from benchpush_b200.metrics import TaskDrivenMetric
    boxes, goals, st, status = ref["tm_boxes"], ref["tm_goals"], ref["tm_state"], ref["tm_status"]
    work, rew = ref["sm_work"], ref["sm_rew"]
    T = len(work)
    m = TaskDrivenMetric("pin", robot_mass=1.2)
    m.reset({"state": st[:1], "obs": boxes[None], "n_boxes_left": np.array([len(boxes)]), "goal_positions": [tuple(g) for g in goals]})
    for t in range(T):
        m.update({"state": st[t + 1:t + 2], "total_work": np.array([work[t]]), "box_completed_statuses": status[t][None]},
                 np.array([rew[t]]), eps_complete=np.array([t == T - 1]))
    got = [m.success_rates[0], m.efficiency_scores[0], m.effort_scores[0], m.rewards[0]]            # task_driven_metric.py:29-157
    np.testing.assert_allclose(got, ref["tm_scores"], rtol=1e-10)


def test_ideal_controller_and_target_course(ref):
    """BoxDeliveryEnv.controller (box_delivery_env.py:867-885) = DP.ideal_control + TargetCourse.advance (dp.py:63-116, 217-248)."""
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv
    paths = [np.array([[0.0, 0.0, 0.3], [1.2, 0.7, 0.5]]), np.array([[0.0, 0.0, 0.3], [0.3, 0.2, 0.5]]), np.array([[1.0, 1.0, -1.0], [-0.8, 2.5, 2.0]])]
    lfc = [0.0, 0.5, 0.5]
    rows = ref["dp_rows"]
    for case in range(3):
        env = types.SimpleNamespace(dp=None, path=[tuple(p) for p in paths[case]], lfc=lfc[case], dt=0.2 / 100, target_speed=0.2)
        for r in rows[rows[:, 0] == case]:
            om, v = BoxDeliveryOracleEnv.controller(env, (r[1], r[2]), r[3])
            np.testing.assert_allclose([om, v[0], v[1]], r[4:7], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(env.dp["setpoint"][:2], r[7:9], rtol=0, atol=0)


def _obstacles():
    from benchpush_b200.layouts import load_packed, unpack_trials
    trial = unpack_trials(load_packed(os.path.join(GOLD, "ice_c30_t7.npz")))[0]
    return [np.asarray(o["vertices"], dtype=np.float64) for o in trial["obstacles"]]


def test_ship_ice_observation_channels_bit_exact(ref):
    """The four egocentric channels through the reference's OccupancyGrid methods (occupancy_map.py:37-65, 112-140, 300-337,
    379-433, 492-587) incl. poses outside the map and near the yaw limits: byte-identical."""
    obstacles = _obstacles()
    occ = ro.OccupancyGrid(grid_width=1 / 25, grid_height=1 / 25, map_width=12, map_height=40, local_width=6, local_height=6, meter_to_pixel_scale=25)
    for pose, want in zip(ref["occ_poses"], ref["occ_obs"]):
        pose = tuple(pose)
        raw = occ.compute_occ_img(obstacles, ice_binary_w=300, ice_binary_h=1000, local_range=[12, 12], ship_state=pose)
        ch = [occ.ego_view_footprint(pose, SHIP_VERTS, 2), occ.ego_view_goal_dist_transform(9.0, pose, 2),
              occ.ego_view_orientation_map(pose, [1., 0.], [-1., 0.], 2), occ.ego_view_obstacle_map(raw, pose, 2)]
        got = (np.stack([np.array(c) for c in ch]) * 255).astype(np.uint8)
        assert np.array_equal(got, want), (pose, int((got != want).sum()))


def test_ship_ice_global_observation_bit_exact(ref):
    obstacles = _obstacles()
    occ = ro.OccupancyGrid(grid_width=0.2, grid_height=0.2, map_width=12, map_height=40, local_width=6, local_height=6, meter_to_pixel_scale=25)
    for pose, want in zip(ref["occ_poses"][:3], ref["occ_global"]):
        raw = occ.compute_occ_img(obstacles, ice_binary_w=300, ice_binary_h=1000)
        o = np.copy(occ.compute_con_gridmap(raw))                                                    # occupancy_map.py:97-109
        occ.compute_ship_footprint_planner(tuple(pose), np.array(SHIP_VERTS))                       # :253-296
        got = (np.stack([o, np.copy(occ.footprint)]) * 255).astype(np.uint8)
        assert np.array_equal(got, want)


def test_maze_maps_and_observation_bit_exact(ref):
    walls = [[(0, 0), (15, 0)], [(0, 0), (0, 15)], [(15, 0), (15, 15)], [(0, 15), (15, 15)], [(15, 15), (15, 5)], [(7.5, 0), (7.5, 10)]]
    mo = MazeOccupancyGrid(grid_width=1 / 16, grid_height=1 / 16, map_width=15, map_height=15, local_width=12, local_height=12, meter_to_pixel_scale=16)
    assert np.array_equal(mo.compute_occ_img_walls(walls, 240, 240, wall_radius=0.5), ref["maze_wall_img"])      # :67-94
    norm, raw = mo.global_goal_point_dist_transform((3.75, 3.75), walls, wall_radius=0.5)                         # :435-485
    assert np.array_equal(raw, ref["maze_dist_raw"]) and np.array_equal(norm, ref["maze_dist_map"])
    boxes = [b for b in ref["maze_boxes"]]
    for pose, want in zip(ref["maze_poses"], ref["maze_obs"]):
        got = (np.asarray(mo.ego_view_map_maze(tuple(pose), ref["maze_robot_verts"], boxes, walls, norm)) * 255).astype(np.uint8)   # :142-202
        assert np.array_equal(got, want), int((got != want).sum())


def test_position_controller_waypoints(ref):
    """PositionController.get_waypoints_to_spatial_action (position_controller.py:56-181) — the reference class run with the
    reference's own compiled spfa on the walled_env_with_columns configuration space — against the oracle's restatement:
    same waypoint count, positions and move sign on straight-line, clamped and spfa + Douglas-Peucker + pruning cases."""
    import random
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    acfg = {"env": "walled_env_with_columns", "num_obstacles": 4}
    ora = AreaClearingOracleEnv(cfg=acfg, layouts=[sample_layout(random.Random(3), acfg)])
    ora.reset()
    assert ora.configuration_space.sum() == ref["pc_cspace_sum"][0]
    off, multi = 0, 0
    for x, y, h, action, sign, n in ref["pc_rows"]:
        n = int(n)
        want = ref["pc_paths"][off:off + 2 * n].reshape(n, 2); off += 2 * n
        path, s = ora.get_waypoints_to_spatial_action([x, y], h, int(action))
        assert len(path) == n and float(s) == float(sign), (len(path), n, s, sign)
        np.testing.assert_allclose(np.asarray([p[:2] for p in path], dtype=np.float64), want, rtol=0, atol=1e-12)
        multi += n > 2
    assert multi >= 2, "the fixture should contain paths that go around the columns"


def test_box_delivery_layouts_follow_the_reference_sampler():
    """get_random_robot_start + generate_boundary (random columns / divider) + generate_boxes of the reference's BoxDeliveryEnv
    (box_delivery_env.py:313-320, 326-507, 509-559), run on seeded RandomStates by tools/make_ref_bd_layouts.py: the product's
    sampler and the oracle's give the same start, obstacles and boxes for every obstacle_config."""
    from benchpush_b200.envs.box_delivery import DEFAULT_CFG
    from benchpush_b200.layouts import sample_box_delivery_layout
    from benchpush_b200.utils import DotDict, merge_cfg
    from oracle.box_delivery_oracle import sample_layout
    g = np.load(os.path.join(GOLD, "ref_bd_layouts.npz"))
    n_obst = 0
    for oc in g["configs"]:
        cfg = {"env": {"obstacle_config": str(oc)}}
        full = merge_cfg(DotDict(DEFAULT_CFG), cfg)
        for seed in g["seeds"]:
            k = "%s_%d_" % (oc, seed)
            for lay in (sample_box_delivery_layout(np.random.RandomState(int(seed)), full), sample_layout(np.random.RandomState(int(seed)), cfg)):
                assert np.array_equal(np.array(lay["start"]), g[k + "start"]), k
                assert np.array_equal(np.array(lay["columns"], dtype=np.float64).reshape(-1, 2), g[k + "columns"]), k
                assert np.array_equal(np.array([lay["divider"]] if lay["divider"] is not None else [], dtype=np.float64).reshape(-1, 2), g[k + "divider"]), k
                assert np.array_equal(np.array(lay["boxes"]), g[k + "boxes"]), k
            n_obst += len(g[k + "columns"]) + len(g[k + "divider"])
    assert n_obst > 40                # the fixtures do contain columns and dividers
    # boundary order of a divider room: receptacle, walls, divider, 3 room corners, 2 divider corners (-90 deg below, 180 deg above)
    assert list(g["large_divider_0_types"]) == ["receptacle"] + ["wall"] * 4 + ["divider"] + ["corner"] * 5
    c = g["large_divider_0_corners"]; d = g["large_divider_0_divider"][0]
    np.testing.assert_allclose(c[3], [5.0, d[1] - 0.25, np.radians(-90)], rtol=0, atol=1e-15)
    np.testing.assert_allclose(c[4], [5.0, d[1] + 0.25, np.radians(180)], rtol=0, atol=1e-15)
