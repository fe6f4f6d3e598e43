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
    """cfg.egocentric_obs = false: u8[E, 2, 1000, 300] from k_shipice_obs_global, bit-exact against the oracle at reset (raw
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
    assert obs.shape == (2, 2, 1000, 300) and obs.dtype == torch.uint8
    assert np.array_equal(obs[0].cpu().numpy(), oobs)
    assert int((obs[0, 0] == 255).sum()) > 1000 and int((obs[0, 1] == 255).sum()) > 100
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
