"""maze-NAMO on the CPU: (1) the restated third-party raster primitives are PINNED against the scipy in this
image and against committed golden vectors; (2) the maze oracle reproduces its committed trajectory pins;
(3) the kernel sources compiled as host C++ (tests/emu, test-only) are bit-identical to the oracle over whole
episodes — body states, observations, termination, reward within 1e-9 (numpy summation order in the oracle)."""
import os
import random

import numpy as np
import pytest

from tests.conftest import GOLD
from oracle import cpo
from oracle.maze_oracle import (MazeNamoMetricOracle, MazeNamoOracleEnv, MazeOccupancyGrid, cosdg, generate_obstacle_centres,
                                rotate_bilinear, sindg)


@pytest.fixture(scope="module")
def mg():
    return np.load(os.path.join(GOLD, "maze_golden.npz"))


def test_sindg_cosdg_match_scipy_and_golden(mg):
    x = mg["dg_x"]
    assert np.array_equal(np.array([sindg(float(v)) for v in x]), mg["dg_sin"])
    assert np.array_equal(np.array([cosdg(float(v)) for v in x]), mg["dg_cos"])
    from scipy import special
    xs = np.random.default_rng(3).uniform(-1000, 1000, 20000)
    assert np.array_equal(np.array([sindg(float(v)) for v in xs]), special.sindg(xs))
    assert np.array_equal(np.array([cosdg(float(v)) for v in xs]), special.cosdg(xs))


def test_rotate_restatement_matches_golden_and_live_scipy(mg):
    for k in range(3):
        ang, cval = mg["rot%d_angle" % k]
        assert np.array_equal(rotate_bilinear(mg["rot%d_in" % k], float(ang), float(cval)), mg["rot%d_out" % k])
    from scipy.ndimage import rotate
    rng = np.random.default_rng(11)
    for trial in range(24):
        n = [288, 40, 31, 64][trial % 4]
        img = (rng.random((n, n)) < 0.3).astype(np.float64) if trial % 3 else rng.random((n, n))
        ang = float(rng.uniform(-400, 400)) if trial % 5 else float(rng.integers(-4, 5) * 45)
        cval = float(trial % 2)
        assert np.array_equal(rotate_bilinear(img, ang, cval), rotate(img, ang, reshape=False, cval=cval, order=1)), (trial, ang)


def test_numpy_offset_product_is_the_fma_the_kernel_uses():
    """rotate() computes offset = centre - M @ centre with numpy; the kernel evaluates the 2x2 @ 2 product as
    fma(M[h,0], v0, M[h,1] * v1) (maze_raster.cuh).  Check that claim against numpy on this machine, with an exact
    rational model of the fused multiply-add."""
    from fractions import Fraction as F
    rng = np.random.default_rng(2)
    for _ in range(2000):
        ang = float(rng.uniform(-360, 360))
        c, s = cosdg(ang), sindg(ang)
        R = np.array([[c, s], [-s, c]])
        v = np.array([143.5, 143.5])
        got = R @ v
        for h, (a, b) in enumerate([(c, s), (-s, c)]):
            assert float(F(a) * F(v[0]) + F(b * v[1])) == got[h]


def test_point_query_and_layout_sampler_match_oracle():
    from benchpush_b200.layouts import _segment_query, generate_maze_layout, maze_walls
    walls = maze_walls(1, 15, 15)
    sp = cpo.Space(10, 0.0)
    for a, b in walls:
        bd = sp.add_body(cpo.STATIC, 0, 0, 0)
        sp.set_material(sp.add_segment(bd, a, b, 0.5), 0.5, 0.5, 3)
    rng = np.random.default_rng(0)
    for _ in range(4000):
        p = rng.uniform(-1, 16, 2)
        md = float(rng.choice([0.0, 0.62, 2.5, -0.1]))
        assert any(_segment_query(p, a, b, 0.5, md) for a, b in walls) == sp.point_query_any(p, md)
    for seed in range(5):
        assert generate_maze_layout(random.Random(seed), walls, 5, 15, 0.62) == generate_obstacle_centres(random.Random(seed), sp, 5, 15, 0.62)


def test_wavefront_literal_equals_level_bfs():
    """global_goal_point_dist_transform restated level by level equals the reference's literal FIFO loop
    (occupancy_map.py:459-478) on a small maze."""
    g = MazeOccupancyGrid(grid_width=0.25, grid_height=0.25, map_width=6, map_height=6, local_width=3, local_height=3, meter_to_pixel_scale=4)
    walls = [[(0, 0), (6, 0)], [(3, 0), (3, 4)]]
    norm, raw = g.global_goal_point_dist_transform((1.0, 1.0), walls)
    H, W = g.occ_map_height, g.occ_map_width
    wall = g.compute_occ_img_walls(walls, W, H)
    edt = np.zeros((H, W)); gy = gx = int(1.0 / (6 / W)); edt[gy, gx] = 1.0
    queue = [(gy, gx)]; visited = {(gy, gx)}
    while queue:
        y, x = queue.pop(0)
        for dy, dx in [(0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (1, -1), (-1, 1), (-1, -1)]:
            ny, nx = y + dy, x + dx
            if ny < 0 or ny >= H or nx < 0 or nx >= W or (ny, nx) in visited or wall[ny, nx] == 1.0:
                continue
            visited.add((ny, nx)); edt[ny, nx] = edt[y, x] + 1; queue.append((ny, nx))
    assert np.array_equal(raw, edt)
    n2 = edt / edt.max(); n2[wall == 1.0] = 1.0
    assert np.array_equal(norm, n2)


def _layouts(mg):
    return [[list(map(float, c)) for c in l] for l in mg["layouts"]]


def test_oracle_reproduces_its_pins(mg, portable_trig):
    layouts = _layouts(mg)
    env = MazeNamoOracleEnv(layouts=layouts, order_mode=1)
    for ep in (0, 1):
        obs, info = env.reset()
        assert np.array_equal(obs, mg["ep%d_reset_obs" % ep])
        assert np.array_equal(env.space.state(), mg["ep%d_reset_state" % ep])
        met = MazeNamoMetricOracle("pin", 1)
        met.reset(info)
        for k, a in enumerate(mg["ep%d_actions" % ep]):
            obs, r, term, _, info = env.step(float(a))
            met.update(info, r, term)
            row = mg["ep%d_rows" % ep][k]
            st = env.space.state()
            assert (st[11, 0], st[11, 1], st[11, 2], float(term), float(info["trial_success"])) == (row[0], row[1], row[2], row[5], row[6])
            np.testing.assert_allclose(r, row[3], rtol=1e-12, atol=1e-12)
        assert term
        assert np.array_equal(env.space.state(), mg["ep%d_final_state" % ep])
        assert np.array_equal(obs, mg["ep%d_final_obs" % ep])
        assert len(met.efficiency_scores) == 1 and (met.efficiency_scores[0] > 0) == (ep == 0)
        assert 0 < met.effort_scores[0] < 1


@pytest.mark.parametrize("ep", [0, 1, 2, 3])
def test_emulated_maze_kernels_match_oracle_bitwise(ep, mg, portable_trig):
    from tests.emu.emu import MazeEmu
    layouts = _layouts(mg)
    emu = MazeEmu(1, layouts)
    env = MazeNamoOracleEnv(layouts=layouts[ep:ep + 1], order_mode=1)
    obs, info = env.reset()
    wall, dn, dr = emu.maps()
    assert np.array_equal(wall.astype(np.float64), env._wall_map)
    assert np.array_equal(dn, env.global_distance_map) and np.array_equal(dr, env.unnormalized_dist_map)
    emu.reset(0, ep)
    assert np.array_equal(obs, emu.render_obs(0, True))
    assert np.array_equal(emu.bodies(0)[6:12], env.space.state()[6:12])
    acts = mg["ep%d_actions" % ep]
    n = len(acts) if ep == 1 else min(len(acts), 90)         # bounded CPU time; the GPU tests run the whole episodes
    for k in range(n):
        obs, r, term, _, info = env.step(float(acts[k]))
        emu.step(0, float(acts[k]))
        gb, ob = emu.bodies(0), env.space.state()
        assert np.array_equal(gb[6:12], ob[6:12]), k                   # obstacles + chassis: identical to the last bit
        assert all(np.array_equal(gb[11], gb[12 + w]) for w in range(4))   # the wheels are rigidly the chassis body
        o = emu.out(0)
        assert o["status"] == 0 and bool(o["terminated"]) == term and bool(o["trial_success"]) == info["trial_success"]
        np.testing.assert_allclose(o["reward"], r, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(o["total_work"], info["total_work"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(o["dist_reward"], info["dist increment reward"], rtol=0, atol=1e-12)
        assert (o["sx"], o["sy"], o["sa"]) == tuple(float(x) for x in info["state"])
        if k % 7 == 0 or term:
            assert np.array_equal(obs, emu.render_obs(0, False)), k
    if n == len(acts):
        assert term


def test_emulated_maze_batched_path_equals_per_env_path(mg, portable_trig):
    from tests.emu.emu import MazeEmu
    layouts = _layouts(mg)
    BKn = MazeEmu(1, layouts).batch_size()
    a_emu = MazeEmu(BKn + 1, layouts)
    b_emu = MazeEmu(BKn + 1, layouts)
    for e in range(BKn + 1):
        a_emu.reset(e, e % len(layouts)); b_emu.reset(e, e % len(layouts))
    rng = np.random.default_rng(4)
    for k in range(12):
        acts = rng.uniform(-1, 1, BKn + 1).astype(np.float32)
        if k < 6:
            acts[:] = 0.0                    # drive straight into the obstacles first
        for e in range(BKn + 1):
            a_emu.step(e, float(acts[e]))
        b_emu.step_batch(0, acts)
        b_emu.step_batch(BKn, acts)
        for e in range(BKn + 1):
            assert np.array_equal(a_emu.bodies(e), b_emu.bodies(e)), (k, e)
            assert a_emu.out(e) == b_emu.out(e)


def test_maze_metric_matches_reference_formulas():
    import torch
    from benchpush_b200.metrics import MazeNamoMetric
    goal_dt = np.arange(240 * 240, dtype=np.float64).reshape(240, 240)
    ora = MazeNamoMetricOracle("o", 1)
    met = MazeNamoMetric("m", 1)
    st0 = (11.25, 3.75, 1.57)
    ora.reset({"state": st0, "goal_dt": goal_dt, "m_to_pix_scale": 16})
    met.reset({"state": torch.tensor([st0], dtype=torch.float64), "goal_dt": goal_dt, "m_to_pix_scale": 16})
    states = [(11.25, 3.87, 1.6), (11.2, 4.0, 1.7), (11.1, 4.2, 1.8)]
    for k, st in enumerate(states):
        done = k == len(states) - 1
        info = {"state": st, "total_work": 0.5 * (k + 1), "trial_success": done}
        ora.update(info, 1.0 + k, done)
        tinfo = {"state": torch.tensor([st], dtype=torch.float64), "total_work": torch.tensor([0.5 * (k + 1)], dtype=torch.float64),
                 "trial_success": torch.tensor([done])}
        met.update(tinfo, torch.tensor([1.0 + k], dtype=torch.float64), torch.tensor([done]))
    np.testing.assert_allclose(met.efficiency_scores, ora.efficiency_scores, rtol=1e-12)
    np.testing.assert_allclose(met.effort_scores, ora.effort_scores, rtol=1e-12)
    np.testing.assert_allclose(met.rewards, ora.rewards, rtol=1e-12)


def test_emulated_maze_without_obstacles_and_with_ragged_counts(portable_trig):
    """Edge cases of the layout input: a maze with NO movable obstacle and one with two, in the same handle (ragged counts,
    padded to the largest): reset image, per-step state triple, status and images against the oracle."""
    from tests.emu.emu import MazeEmu
    layouts = [[], [[11.3, 5.2], [11.0, 7.5]]]
    emu = MazeEmu(2, layouts)
    for ep in (0, 1):
        env = MazeNamoOracleEnv(layouts=layouts[ep:ep + 1], order_mode=1)
        obs, info = env.reset()
        emu.reset(ep, ep)
        assert len(info["obs"]) == len(layouts[ep])
        assert np.array_equal(obs, emu.render_obs(ep, True))
        rng = np.random.default_rng(ep)
        for k in range(5):
            a = float(np.float32(rng.uniform(-0.3, 0.3)))
            obs, r, term, _, info = env.step(a)
            emu.step(ep, a)
            o = emu.out(ep)
            assert o["status"] == 0 and bool(o["terminated"]) == term
            assert (o["sx"], o["sy"], o["sa"]) == tuple(float(x) for x in info["state"])
            np.testing.assert_allclose(o["reward"], r, rtol=1e-9, atol=1e-9)
            assert np.array_equal(obs, emu.render_obs(ep, False)), (ep, k)
