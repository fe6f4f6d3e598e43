"""``cfg.low_dim_state`` on the CUDA path (``bp_shipice_low_dim`` / ``bp_maze_low_dim``, kernel ``k_low_dim``) against the
CPU oracle.  Tolerance: the body states are bit-identical (portable cos / sin); the observation entries agree to 1e-9 m
absolute (measured: 1.4e-10) — the kernel sums the shoelace terms sequentially, the reference's ``np.dot`` pairwise, and
the cancellation in those sums (coordinates up to 40 m, areas below 1 m^2) amplifies the last-bit difference; the same
tolerance as the work term — and the padding, the lengths and the maze's index quirks are exact."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLD, fixture_conc
from benchpush_b200.layouts import unpack_trials
from oracle.maze_oracle import MazeNamoOracleEnv
from oracle.ship_ice_oracle import ShipIceOracleEnv

pytestmark = pytest.mark.gpu


def test_ship_ice_low_dim_steps(golden_layouts, pins, portable_trig):
    from benchpush_b200.envs import ShipIceVecEnv
    name = "c10_t7"
    p = golden_layouts[name]
    conc = fixture_conc(name)
    cfg = {"concentration": conc, "low_dim_state": True, "fixed_trial_idx": 0}
    env = ShipIceVecEnv(num_envs=3, cfg=cfg, layouts=p, device=0)
    obs, info = env.reset()
    ora = ShipIceOracleEnv(cfg, experiment=unpack_trials(p), order_mode=1)
    oobs, oinfo = ora.reset()
    n2 = oobs.shape[0]
    assert obs.dtype == torch.float64 and obs.shape == (3, 2 * p.max_floes)
    assert env._low_dim_counts.tolist() == [n2] * 3
    np.testing.assert_allclose(obs[0, :n2].cpu().numpy(), oobs, rtol=0, atol=1e-9)
    assert float(obs[:, n2:].abs().sum()) == 0.0
    l0 = env.launch_count()
    for a in pins["%s_o1_actions" % name][:6]:
        act = np.array([a, a, -a], dtype=np.float32)
        obs, rew, term, trunc, info = env.step(act)
        oobs, orew, oterm, _, oinfo = ora.step(float(np.float32(a)))
        assert np.array_equal(env.get_bodies(0), ora.space.state())
        for e in (0, 1):
            np.testing.assert_allclose(obs[e, :n2].cpu().numpy(), oobs, rtol=0, atol=1e-9)
        assert float(obs[:, n2:].abs().sum()) == 0.0
        np.testing.assert_allclose(float(rew[0]), orew, rtol=1e-9, atol=1e-7)
    assert not torch.equal(obs[0], obs[2])                        # env 2 steered the other way
    assert env.launch_count() - l0 == 2 * 6                       # step + low-dim kernel: the image kernel is not launched
    env.close()
    # the same vector is available next to the image observation, and the single-env view cuts the padding
    from benchpush_b200.envs.gym_single import ShipIceEnvB200
    one = ShipIceEnvB200(cfg=cfg, layouts=p, device=0)
    o1, _ = one.reset()
    assert isinstance(o1, np.ndarray) and o1.shape == (n2,) and o1.dtype == np.float64
    one.close()
    img = ShipIceVecEnv(num_envs=1, cfg={"concentration": conc}, layouts=p, device=0)
    img.reset()
    assert img.low_dim_observation().shape == (1, 2 * p.max_floes)
    img.close()


def test_maze_low_dim_steps(portable_trig):
    from benchpush_b200.envs import MazeNamoVecEnv
    mg = np.load(os.path.join(GOLD, "maze_golden.npz"))
    layouts = [[list(map(float, c)) for c in l] for l in mg["layouts"]]
    E = len(layouts)
    env = MazeNamoVecEnv(num_envs=E, cfg={"low_dim_state": True}, layouts=layouts, device=0)
    obs, info = env.reset()
    oras = [MazeNamoOracleEnv({"low_dim_state": True}, layouts=layouts[e:e + 1], order_mode=1) for e in range(E)]
    for e, o in enumerate(oras):
        oobs, _ = o.reset()
        n2 = oobs.shape[0]
        assert int(env._low_dim_counts[e]) == n2
        np.testing.assert_allclose(obs[e, :n2].cpu().numpy(), oobs, rtol=0, atol=1e-9)
        assert obs[e, n2 - 2:].abs().sum() == 0
    for k in range(5):
        a = np.array([mg["ep%d_actions" % e][k] for e in range(E)], dtype=np.float32)
        obs, rew, term, trunc, info = env.step(a)
        for e, o in enumerate(oras):
            oobs, orew, oterm, _, oinfo = o.step(float(a[e]))
            n2 = oobs.shape[0]
            got = obs[e, :n2].cpu().numpy()
            assert got[0] == oobs[0] and got[1] == oobs[1]         # robot position: bit-identical
            np.testing.assert_allclose(got, oobs, rtol=0, atol=1e-9)
            assert obs[e, n2 - 2:].abs().sum() == 0
    env.close()


def test_ship_ice_global_observation(golden_layouts, pins, portable_trig):
    """cfg.egocentric_obs = false: u8[E, 2, 200, 60] (0.2 m cells: block-mean concentration + footprint) from k_shipice_obs_global, bit-exact against the oracle at reset (raw
    trial polygons) and after every step (hull polygons); the ship is steered to the map border so that part of the
    footprint's vertices leave the grid."""
    from benchpush_b200.envs import ShipIceVecEnv
    name = "c10_t0"
    p = golden_layouts[name]
    conc = fixture_conc(name)
    cfg = {"concentration": conc, "egocentric_obs": False}
    env = ShipIceVecEnv(num_envs=2, cfg=cfg, layouts=p, device=0)
    obs, info = env.reset()
    ora = ShipIceOracleEnv(cfg, experiment=unpack_trials(p), order_mode=1)
    oobs, oinfo = ora.reset()
    assert obs.shape == (2, 2, 200, 60) and obs.dtype == torch.uint8
    assert np.array_equal(obs[0].cpu().numpy(), oobs)
    assert int((obs[0, 0] > 0).sum()) > 300 and int((obs[0, 1] == 255).sum()) > 10
    for k in range(8):
        a = 1.0 if k < 4 else -1.0
        obs, rew, term, trunc, info = env.step(np.array([a, -a], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora.step(float(np.float32(a)))
        assert np.array_equal(env.get_bodies(0), ora.space.state())
        assert np.array_equal(obs[0].cpu().numpy(), oobs), k
        if oterm:
            break
    assert not torch.equal(obs[0, 1], obs[1, 1])
    env.close()


def test_ship_ice_random_start(golden_layouts, pins, portable_trig):
    """cfg.random_start (ship_ice_env.py:201-203): x ~ 1 + U * (start_x_range - 1), y = 1, theta = pi / 2 per episode.  The full
    reset (init + 1000 settle substeps with the ship at the drawn pose) and the following steps are bit-identical to the
    oracle started at the same pose; the auto reset inside step() consumes a fresh draw."""
    from benchpush_b200.envs import ShipIceVecEnv
    name = "c30_t7"
    p = golden_layouts[name]
    conc = fixture_conc(name)
    cfg = {"concentration": conc, "random_start": True}
    env = ShipIceVecEnv(num_envs=3, cfg=cfg, layouts=p, device=0, start_seed=11)
    obs, info = env.reset()
    starts = env.start_poses.copy()
    assert ((starts[:, 0] >= 1) & (starts[:, 0] < float(env.cfg.start_x_range))).all() and len(set(starts[:, 0])) == 3
    assert (starts[:, 1] == 1.0).all() and (starts[:, 2] == np.pi / 2).all()
    for e in (0, 2):
        ora = ShipIceOracleEnv(cfg, experiment=unpack_trials(p), order_mode=1)
        ora.start_override = starts[e]
        oobs, oinfo = ora.reset()
        assert np.array_equal(env.get_bodies(e), ora.space.state()), e
        assert np.array_equal(obs[e].cpu().numpy(), oobs)
        assert tuple(info["state"][e].tolist()) == tuple(float(x) for x in oinfo["state"])
        if e == 0:
            ora0 = ora
    env.redraw_starts = False
    for a in pins["%s_o1_actions" % name][:4]:
        obs, rew, term, trunc, info = env.step(np.array([a, 0.0, 0.0], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora0.step(float(np.float32(a)))
        assert np.array_equal(env.get_bodies(0), ora0.space.state())
        assert np.array_equal(obs[0].cpu().numpy(), oobs)
    env.close()
    # the oracle's own draw follows the reference's formula
    import random
    ora = ShipIceOracleEnv(cfg, experiment=unpack_trials(p), order_mode=1)
    ora.rng = random.Random(3)
    ora.episode_idx = 0
    ora.init_ship_ice_sim(); ora.init_ship_ice_env(settle=False)
    assert ora.start == (1 + random.Random(3).random() * (ora.cfg["start_x_range"] - 1), 1.0, np.pi / 2)
    # auto reset: a finished env restarts from the pose drawn after the previous step
    env = ShipIceVecEnv(num_envs=2, cfg=dict(cfg, goal_y=1.5), layouts=p, device=0, start_seed=4, auto_reset=True)
    env.reset()
    seen = []
    for k in range(6):
        pending = env.start_poses.copy()
        obs, rew, term, trunc, info = env.step(np.zeros(2, dtype=np.float32))
        pose = info["pose"].cpu().numpy()
        for e in range(2):
            if abs(pose[e, 1] - 1.0) < 1e-9 and k > 0:            # this step was the auto reset
                assert abs(pose[e, 0] - pending[e, 0]) < 1e-9
                seen.append(e)
    assert seen, "goal_y = 1.5 should end an episode within 6 steps"
    env.close()


def test_spaces_of_the_image_and_low_dim_modes(golden_layouts):
    """action_space / observation_space of ship-ice and maze (ship_ice_env.py:72, 84-99; maze_NAMO_env.py:102-124)."""
    from benchpush_b200.envs import MazeNamoVecEnv, ShipIceVecEnv
    from benchpush_b200.envs.gym_single import ShipIceEnvB200
    p = golden_layouts["c10_t0"]
    e = ShipIceVecEnv(num_envs=3, cfg={"concentration": 0.1}, layouts=p, device=0)
    assert e.single_action_space.shape == () and e.single_action_space.dtype == np.float32
    assert e.single_observation_space.shape == (4, 150, 150) and e.single_observation_space.dtype == np.uint8
    assert e.observation_space.shape == (3, 4, 150, 150) and e.action_space.shape == (3,)
    obs, _ = e.reset()
    assert tuple(obs.shape) == e.observation_space.shape
    e.close()
    g = ShipIceVecEnv(num_envs=1, cfg={"concentration": 0.1, "egocentric_obs": False}, layouts=p, device=0)
    assert g.single_observation_space.shape == (2, 200, 60)
    g.close()
    l = ShipIceVecEnv(num_envs=1, cfg={"concentration": 0.1, "low_dim_state": True, "fixed_trial_idx": 0}, layouts=p, device=0)
    assert l.single_observation_space.shape == (2 * p.max_floes,) and l.single_observation_space.dtype == np.float64
    assert float(l.single_observation_space.low[0]) == -10 and float(l.single_observation_space.high[0]) == 30
    l.close()
    m = MazeNamoVecEnv(num_envs=2, device=0)
    assert m.single_action_space.dtype == np.float64 and m.single_observation_space.shape == (4, 192, 192)
    m.close()
    s = ShipIceEnvB200(cfg={"concentration": 0.1}, layouts=p, device=0)
    assert s.action_space.shape == () and s.observation_space.shape == (4, 150, 150)
    s.close()
