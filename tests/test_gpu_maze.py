"""Parity of the maze-NAMO CUDA path (through the C ABI) against the CPU oracle and the committed pins.
Tolerances: with the oracle on the portable cos / sin the kernels use, body poses and velocities are
BIT-IDENTICAL after every env step over whole episodes; observation rasters (4 x 192 x 192 u8, incl. the
bilinear-rotated channels) and the rounded info['state'] are bit-exact; termination / success flags equal;
reward and work within 1e-9 (numpy summation order in the oracle)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLD
from oracle.maze_oracle import MazeNamoMetricOracle, MazeNamoOracleEnv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    return np.load(os.path.join(GOLD, "maze_golden.npz"))


def _layouts(mg):
    return [[list(map(float, c)) for c in l] for l in mg["layouts"]]


def _env(layouts, n, **kw):
    from benchpush_b200.envs import MazeNamoVecEnv
    return MazeNamoVecEnv(num_envs=n, layouts=layouts, device=0, **kw)


def test_constant_maps_match_oracle(mg):
    layouts = _layouts(mg)
    env = _env(layouts, 1)
    ora = MazeNamoOracleEnv(layouts=layouts, order_mode=1)
    ora.reset()
    assert np.array_equal(env.wall_map.astype(np.float64), ora._wall_map)
    assert np.array_equal(env.global_distance_map, ora.global_distance_map)
    assert np.array_equal(env.unnormalized_dist_map, ora.unnormalized_dist_map)
    env.close()


def test_whole_episodes_bitwise_vs_oracle_and_pins(mg, portable_trig):
    layouts = _layouts(mg)
    E = len(layouts)
    env = _env(layouts, E)                     # env e runs layout e
    obs, info = env.reset()
    oras = []
    for e in range(E):
        o = MazeNamoOracleEnv(layouts=layouts[e:e + 1], order_mode=1)
        oobs, oinfo = o.reset()
        assert np.array_equal(obs[e].cpu().numpy(), oobs), e
        assert np.array_equal(obs[e].cpu().numpy(), mg["ep%d_reset_obs" % e])
        assert np.array_equal(env.get_bodies(e)[6:12], o.space.state()[6:12])
        oras.append(o)
    acts = [mg["ep%d_actions" % e] for e in range(E)]
    T = max(len(a) for a in acts)
    alive = np.ones(E, dtype=bool)
    for k in range(T):
        a = np.array([acts[e][k] if k < len(acts[e]) else 0.0 for e in range(E)], dtype=np.float32)
        mask = torch.as_tensor(alive.astype(np.uint8))
        obs, rew, term, trunc, info = env.step(a, mask=mask)
        for e in range(E):
            if not alive[e]:
                continue
            oobs, orew, oterm, _, oinfo = oras[e].step(float(a[e]))
            gb, ob = env.get_bodies(e), oras[e].space.state()
            assert np.array_equal(gb[6:12], ob[6:12]), (e, k)
            assert all(np.array_equal(gb[11], gb[12 + w]) for w in range(4))
            assert bool(term[e]) == oterm and bool(info["trial_success"][e]) == oinfo["trial_success"]
            np.testing.assert_allclose(float(rew[e]), orew, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(float(info["total_work"][e]), oinfo["total_work"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(float(info["dist increment reward"][e]), oinfo["dist increment reward"], rtol=0, atol=1e-12)
            assert tuple(info["state"][e].tolist()) == tuple(float(x) for x in oinfo["state"])
            if k % 5 == 0 or oterm:
                assert np.array_equal(obs[e].cpu().numpy(), oobs), (e, k)
            row = mg["ep%d_rows" % e][k]
            assert (gb[11, 0], gb[11, 1], gb[11, 2]) == (row[0], row[1], row[2])
            if oterm:
                alive[e] = False
                assert k == len(acts[e]) - 1
                assert np.array_equal(gb[6:12], mg["ep%d_final_state" % e][6:12])
                assert np.array_equal(obs[e].cpu().numpy(), mg["ep%d_final_obs" % e])
    assert not alive.any()
    assert int(env._status.max()) == 0
    env.close()


def test_pins_without_oracle(mg):
    """The committed pins alone (no oracle in the loop): obstacle / robot states and observations after the
    recorded actions."""
    layouts = _layouts(mg)
    env = _env(layouts[:2], 2)
    env.reset()
    for k in range(61):
        a = np.array([mg["ep0_actions"][k], mg["ep1_actions"][min(k, len(mg["ep1_actions"]) - 1)]], dtype=np.float32)
        mask = torch.tensor([1, 1 if k < len(mg["ep1_actions"]) else 0], dtype=torch.uint8)
        obs, rew, term, trunc, info = env.step(a, mask=mask)
        if k in (10, 60):
            assert np.array_equal(obs[0].cpu().numpy(), mg["ep0_obs_k%d" % k])
        if k == 10:
            assert np.array_equal(obs[1].cpu().numpy(), mg["ep1_obs_k10"])
    env.close()


def test_auto_reset_and_metric_vs_oracle(mg, portable_trig):
    from benchpush_b200.metrics import MazeNamoMetric
    layouts = _layouts(mg)[2:4]                    # the two randomly sampled layouts, random actions
    env = _env(layouts, 1, auto_reset=True)
    obs, info = env.reset()
    ora = MazeNamoOracleEnv(layouts=layouts, order_mode=1)
    oobs, oinfo = ora.reset()
    met = MazeNamoMetric("gpu", robot_mass=1)
    omet = MazeNamoMetricOracle("cpu", 1)
    met.reset(info)
    omet.reset(oinfo)
    rng = np.random.default_rng(9)
    episodes = 0
    for k in range(400):
        a = float(np.float32(rng.uniform(-1, 1)))
        obs, rew, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora.step(a)
        assert bool(term[0]) == oterm
        met.update(info, rew, term | trunc)
        omet.update(oinfo, orew, oterm)
        if oterm:
            episodes += 1
            # next step call resets (next layout) and returns the reset observation with reward 0
            obs, rew, term, trunc, info = env.step(np.array([0.0], dtype=np.float32))
            oobs, oinfo = ora.reset()
            assert np.array_equal(obs[0].cpu().numpy(), oobs)
            assert float(rew[0]) == 0.0 and not bool(term[0])
            met.reset({**info, "goal_dt": env.unnormalized_dist_map, "m_to_pix_scale": env.cfg.occ.m_to_pix_scale})
            omet.reset(oinfo)
            if episodes == 3:
                break
    assert episodes == 3
    np.testing.assert_allclose(met.efficiency_scores, omet.efficiency_scores, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(met.effort_scores, omet.effort_scores, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(met.rewards, omet.rewards, rtol=1e-9, atol=1e-9)
    env.close()


def test_host_buffer_step_equals_device_step(mg):
    layouts = _layouts(mg)
    a_env = _env(layouts, 4)
    b_env = _env(layouts, 4)
    a_env.reset(); b_env.reset()
    rng = np.random.default_rng(1)
    for k in range(8):
        a = rng.uniform(-1, 1, 4).astype(np.float32) * (k >= 4)
        obs, rew, term, trunc, info = a_env.step(a)
        hb = b_env.step_host(a, want_obs=True)
        assert np.array_equal(hb["obs"], obs.cpu().numpy())
        assert np.array_equal(hb["reward"], rew.cpu().numpy())
        assert np.array_equal(hb["state"], info["state"].cpu().numpy())
        assert np.array_equal(hb["terminated"].astype(bool), term.cpu().numpy())
    a_env.close(); b_env.close()


def test_full_size_properties_16384_envs():
    """BASELINE config 4 size: 16 384 envs, randomised obstacles.  Size-independent properties: envs that share a
    layout and an action sequence are bit-identical wherever they sit in the batch; obstacles out of the robot's
    reach and clear of the walls do not move at all (an obstacle sampled inside a wall — the reference never tests
    the first one, maze_NAMO_env.py:287-296 — keeps relaxing out of it); the wheels stay rigidly on the chassis;
    every status word is 0."""
    from benchpush_b200.layouts import _segment_query, generate_maze_layouts
    E, Lc = 16384, 64
    layouts = generate_maze_layouts(Lc, seed=3)
    env = _env(layouts, E)
    obs, info = env.reset()
    assert obs.shape == (E, 4, 192, 192) and obs.dtype == torch.uint8
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    base = torch.rand((6, Lc), generator=g, device="cuda") * 2 - 1
    b0 = [env.get_bodies(e) for e in (0, 5, 63)]
    for k in range(6):
        a = base[k].repeat(E // Lc)                       # env e gets the action of its layout e % Lc
        obs, rew, term, trunc, info = env.step(a)
    torch.cuda.synchronize()
    assert int(env._status.max()) == 0
    o = obs.view(E // Lc, Lc, -1)
    assert bool((o == o[0:1]).all())
    r = rew.view(E // Lc, Lc)
    assert bool((r == r[0:1]).all())
    for i, e in enumerate((0, 5, 63)):
        b1 = env.get_bodies(e)
        assert np.array_equal(b1, env.get_bodies(e + 77 * Lc))
        assert np.array_equal(b1[:6], b0[i][:6])                               # static walls
        assert all(np.array_equal(b1[11], b1[12 + w]) for w in range(4))
        far = np.hypot(b0[i][6:11, 0] - 11.25, b0[i][6:11, 1] - 4.5) > 4.0     # the robot moved < 0.8 m from (11.25, 3.75)
        clear = np.array([not any(_segment_query(c, a, b, 0.5, 0.9) for a, b in env.maze_walls) for c in b0[i][6:11, :2]])
        quiet = far & clear & np.array([np.all(np.hypot(*(b0[i][6:11, :2] - c).T)[np.arange(5) != j] > 1.6) for j, c in enumerate(b0[i][6:11, :2])])
        assert np.array_equal(b1[6:11][quiet], b0[i][6:11][quiet])
    # observation sanity: the footprint channel holds the robot in the middle of the window, the wall channel is non-empty
    assert int(obs[0, 0, 96, 96]) == 255 and int(obs[:64, 2].max()) == 255
    env.close()


def test_maze_random_start(portable_trig):
    """cfg.random_start (maze_NAMO_env.py:229-238): start drawn in [1, start_range) x [1, start_range), rejected while a wall
    segment (radius 0.5) is closer than robot.min_obstacle_dist, heading 3 pi / 2; reset and steps from the drawn pose are
    bit-identical to the oracle started at the same pose, and the oracle's own sampler accepts exactly such poses."""
    import random
    from benchpush_b200.envs import MazeNamoVecEnv
    from oracle import cpo
    from oracle.maze_oracle import MazeNamoOracleEnv
    lay = [[[8.5, 11], [10, 9], [11.25, 11.5], [6, 10], [3.5, 8.5]]]
    cfg = {"random_start": True, "start_x_range": 14, "start_y_range": 14}
    env = MazeNamoVecEnv(num_envs=4, cfg=cfg, layouts=lay, device=0, start_seed=5)
    obs, info = env.reset()
    starts = env.start_poses.copy()
    assert ((starts[:, :2] >= 1) & (starts[:, :2] < 14)).all() and (starts[:, 2] == np.pi * 3 / 2).all() and len(set(starts[:, 0])) == 4
    for e in (0, 3):
        ora = MazeNamoOracleEnv(cfg, layouts=lay, order_mode=1)
        ora.start_override = starts[e]
        oobs, oinfo = ora.reset()
        walls_only = cpo.Space()                                  # the reference samples the start when only the walls exist
        for a_, b_ in ora.maze_walls:
            walls_only.add_segment(walls_only.add_body(cpo.STATIC, 0.0, 0.0, 0.0), a_, b_, 0.5)
        assert not walls_only.point_query_any((starts[e, 0], starts[e, 1]), ora.cfg["robot"]["min_obstacle_dist"])    # an admissible draw
        assert np.array_equal(env.get_bodies(e)[6:12], ora.space.state()[6:12])        # 5 obstacles + robot chassis
        assert np.array_equal(obs[e].cpu().numpy(), oobs)
        if e == 0:
            ora0 = ora
    env.redraw_starts = False
    rng = np.random.default_rng(2)
    for k in range(4):
        a = rng.uniform(-1, 1, 4)
        obs, rew, term, trunc, info = env.step(a.astype(np.float32))
        oobs, orew, oterm, _, oinfo = ora0.step(float(np.float32(a[0])))
        assert np.array_equal(env.get_bodies(0)[6:12], ora0.space.state()[6:12])
        assert np.array_equal(obs[0].cpu().numpy(), oobs)
        np.testing.assert_allclose(float(rew[0]), orew, rtol=1e-9, atol=1e-7)
    env.close()
    ora = MazeNamoOracleEnv(cfg, layouts=lay, order_mode=1)
    ora.rng = random.Random(1)
    ora.reset(settle=False)
    assert 1 <= ora.start[0] < 14 and 1 <= ora.start[1] < 14 and ora.start[2] == np.pi * 3 / 2
