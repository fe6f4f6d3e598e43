"""area-clearing-v0 kernel logic on a machine without a GPU: the kernel sources compiled for the host with one lane
(tests/emu/emu_bd.cpp) against the oracle, through the Python mirror the product uses (AreaClearingVecEnv)."""
import random

import numpy as np

from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout, convex_intersects
from benchpush_b200.envs.area_clearing import AreaClearingVecEnv, DEFAULT_CFG, ENVS
from benchpush_b200.layouts import sample_area_clearing_layout
from benchpush_b200.utils import DotDict
from tests.areaclr_common import CLEARING_CFG, CLEARING_LAYOUT, compare_step
from tests.emu import emu_bd


def test_layout_sampler_and_intersection_predicate():
    a = sample_layout(random.Random(3))
    b = sample_area_clearing_layout(random.Random(3), DotDict.to_dot_dict(DEFAULT_CFG), ENVS["clear_env"])
    assert a["start"] == b["start"] and a["boxes"] == b["boxes"] and len(a["boxes"]) == 10
    sq = [[-5.0, -5.0], [-5.0, 5.0], [5.0, 5.0], [5.0, -5.0]]
    box = lambda cx, cy: [[cx + .5, cy + .5], [cx - .5, cy + .5], [cx - .5, cy - .5], [cx + .5, cy - .5]]
    assert convex_intersects(sq, box(0, 0)) and convex_intersects(sq, box(5.2, 0)) and convex_intersects(sq, box(5.5, 5.5))   # touching counts
    assert not convex_intersects(sq, box(5.51, 0)) and not convex_intersects(sq, box(-6, -6))


def test_maps_reset_and_heading_episode_match_the_oracle(portable_trig):
    lay = [sample_layout(random.Random(1))]
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, layouts=lay)
    ora = AreaClearingOracleEnv(layouts=lay)
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    r = env.room
    assert r.shape == (468, 468)
    assert np.array_equal(r.configuration_space, ora.configuration_space)
    assert np.array_equal(r.closest_cspace_indices, ora.closest_cspace_indices)
    assert np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
    assert np.array_equal(np.array(env.goal_points), np.array(ora.goal_points)) and len(env.goal_points) == 40
    assert np.array_equal(obs[0], oobs)
    assert np.array_equal(env.get_bodies(0)[:, :6], ora.space.state()[:11])
    rng = np.random.default_rng(1)
    for k in range(12):
        a = float(np.float32(rng.uniform(-1, 1)))
        compare_step(env, ora, 0, a, env.step(np.array([a], dtype=np.float32)), state_tol=1e-9)
    env.close()


def test_clearing_a_box_resets_the_clock_and_marks_it_completed(portable_trig):
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg=CLEARING_CFG, layouts=[CLEARING_LAYOUT])
    ora = AreaClearingOracleEnv(cfg=CLEARING_CFG, layouts=[CLEARING_LAYOUT])
    env.reset(); ora.reset()
    counts, ts = [], []
    for k in range(6):
        out = env.step(np.array([-1.0], dtype=np.float32))
        compare_step(env, ora, 0, -1.0, out, state_tol=1e-9)
        counts.append(int(out[4]["box_count"][0])); ts.append(int(out[4]["t"][0]))
    assert counts == [0, 0, 1, 1, 1, 1] and ts[2] == 0 and ts[3] == 1            # D13: the truncation clock restarts when a box is cleared
    assert float(env._buf["reward"][0]) == -0.25                                  # pressed against the outer wall: BOUNDARY_PENALTY
    env.close()


def test_configure_env_for_sam_matches_the_oracle(portable_trig):
    """``configure_env_for_SAM`` (area_clearing.py:221-223): the 96-pixel observation of the spatial-action-map baselines —
    room maps, goal-point map, rasters and position actions at 4 pixels per metre, heading and position actions."""
    lay = [sample_layout(random.Random(1))]
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, layouts=lay)
    ora = AreaClearingOracleEnv(layouts=lay)
    assert env.local_map_pixel_width == 224
    env.configure_env_for_SAM(); ora.configure_env_for_SAM()
    env.configure_env_for_SAM()                                                   # idempotent
    assert env.sam and env.local_map_pixel_width == 96 == ora.local_map_pixel_width and env.observation_shape == (96, 96, 4)
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    assert np.array_equal(env.room.configuration_space, ora.configuration_space)
    assert np.array_equal(env.room.closest_cspace_indices, ora.closest_cspace_indices)
    assert np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
    assert obs.shape[1:] == (96, 96, 4) and np.array_equal(obs[0], oobs)
    rng = np.random.default_rng(2)
    for k in range(6):
        a = float(np.float32(rng.uniform(-1, 1)))
        compare_step(env, ora, 0, a, env.step(np.array([a], dtype=np.float32)), state_tol=1e-9)
    env.update_path([(0.0, 0.0), (1.0, 1.0)])
    assert env.path == [(0.0, 0.0), (1.0, 1.0)]
    env.close()
    cfg = dict(CLEARING_CFG, agent={"action_type": "position"})
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg=cfg, layouts=[CLEARING_LAYOUT], sam=True)
    ora = AreaClearingOracleEnv(cfg=cfg, layouts=[CLEARING_LAYOUT]); ora.configure_env_for_SAM()
    env.reset(); ora.reset()
    for pix in (20 * 96 + 48, 48 * 96 + 12, 85 * 96 + 85):
        compare_step(env, ora, 0, float(pix), env.step(np.array([float(pix)], dtype=np.float32)), state_tol=1e-9)
    env.close()


def test_small_environment_matches_the_oracle(portable_trig):
    """clear_env_small: 12 m room, 8 m clearance square, 18.67 px / m, 542 x 542 grid."""
    cfg = {"env": "clear_env_small", "num_obstacles": 6, "min_obs_dist": 1.2}
    lay = [sample_layout(random.Random(4), cfg)]
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg=cfg, layouts=lay)
    ora = AreaClearingOracleEnv(cfg=cfg, layouts=lay)
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    assert env.room.shape == (542, 542) and np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
    assert np.array_equal(obs[0], oobs)
    rng = np.random.default_rng(4)
    for k in range(5):
        a = float(np.float32(rng.uniform(-1, 1)))
        compare_step(env, ora, 0, a, env.step(np.array([a], dtype=np.float32)), state_tol=1e-9)
    env.close()


import pytest


@pytest.mark.parametrize("name,shape", [("walled_env", (468, 468)), ("walled_env_with_columns", (542, 542))])
def test_walled_environments_match_the_oracle(name, shape, portable_trig):
    """walled_env / walled_env_with_columns (area_clearing/envs/*.yaml): inner walls (polygons around their centre lines,
    area_clearing.py:455-480) and static obstacles (:437-453) are static shapes of the handle and obstacle pixels of the
    configuration space; maps, reset and a heading episode against the oracle."""
    cfg = {"env": name, "num_obstacles": 6}
    lay = [sample_layout(random.Random(7), cfg)]
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg=cfg, layouts=lay)
    ora = AreaClearingOracleEnv(cfg=cfg, layouts=lay)
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    r = env.room
    assert r.shape == shape
    assert len(r.statics) == 4 + len(ENVS[name].get("walls", [])) + len(ENVS[name].get("static_obstacles", []))
    assert np.array_equal(r.configuration_space, ora.configuration_space)
    assert np.array_equal(r.closest_cspace_indices, ora.closest_cspace_indices)
    assert np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
    assert np.array_equal(obs[0], oobs)
    nb = len(lay[0]["boxes"])
    assert np.array_equal(env.get_bodies(0)[:1 + nb, :6], ora.space.state()[:1 + nb])
    rng = np.random.default_rng(2)
    for k in range(8):
        a = float(np.float32(rng.uniform(-1, 1)))
        compare_step(env, ora, 0, a, env.step(np.array([a], dtype=np.float32)), state_tol=1e-9)
    env.close()


def test_velocity_actions_match_the_oracle(portable_trig):
    """action_type velocity (area_clearing.py:656-663): yaw rate = max_yaw_rate_step * a[1] / 2, surge = target_speed * a[0]
    set once per step, then cfg.sim.steps substeps."""
    cfg = {"env": "walled_env", "num_obstacles": 5, "agent": {"action_type": "velocity"}}
    lay = [sample_layout(random.Random(11), cfg)]
    env = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg=cfg, layouts=lay)
    ora = AreaClearingOracleEnv(cfg=cfg, layouts=lay)
    env.reset(); ora.reset()
    assert env.single_action_space.shape == (2,)
    rng = np.random.default_rng(3)
    for k in range(12):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        a[0] = abs(a[0])                                   # drive forward into the boxes
        compare_step(env, ora, 0, a, env.step(a[None, :]), state_tol=1e-9)
    assert ora.space.state()[0, 1] > lay[0]["start"][1] + 0.2
    env.close()
