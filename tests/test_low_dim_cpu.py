"""Oracle side of ``generate_observation_low_dim`` (ship_ice_env.py:358-370, maze_NAMO_env.py:488-504), CPU only:
the restatement against a literal evaluation of the reference's formula on the oracle's own polygons, and the
reference's quirks the CUDA path has to reproduce (fixed trial in ship-ice; maze: robot first, obstacle 0 skipped,
last pair zero)."""
import os

import numpy as np

from tests.conftest import GOLD, fixture_conc
from benchpush_b200.layouts import unpack_trials
from oracle.maze_oracle import MazeNamoOracleEnv
from oracle.ship_ice_oracle import ShipIceOracleEnv


def _centroid_literal(v):
    # geometry/polygon.py:25-41, written out with np.roll exactly as the reference does
    x, y = np.asarray(v).T
    A = 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))
    u = x * np.roll(y, 1) - np.roll(x, 1) * y
    return np.abs((1 / (6 * A) * np.dot(x + np.roll(x, 1), u), 1 / (6 * A) * np.dot(y + np.roll(y, 1), u)))


def test_ship_ice_low_dim_oracle(golden_layouts):
    name = "c10_t7"
    trials = unpack_trials(golden_layouts[name])
    ora = ShipIceOracleEnv({"concentration": fixture_conc(name), "low_dim_state": True, "fixed_trial_idx": 0}, experiment=trials, order_mode=1)
    obs, info = ora.reset()
    n = len(info["obs"])
    assert obs.dtype == np.float64 and obs.shape == (2 * n,)
    for k in range(2):
        obs, rew, term, trunc, info = ora.step(0.5)
        assert obs.shape == (2 * n,)
        lit = np.concatenate([_centroid_literal(p) for p in info["obs"]])
        assert np.array_equal(obs, lit)
        assert (obs >= 0).all()
    # the fixed trial is used for every episode (ship_ice_env.py:189-190)
    obs2, info2 = ora.reset()
    assert ora.episode_idx == 1 and len(info2["obs"]) == n


def test_maze_low_dim_oracle_keeps_the_reference_indexing():
    mg = np.load(os.path.join(GOLD, "maze_golden.npz"))
    layouts = [[list(map(float, c)) for c in l] for l in mg["layouts"]][:1]
    ora = MazeNamoOracleEnv({"low_dim_state": True}, layouts=layouts, order_mode=1)
    obs, info = ora.reset()
    n = len(info["obs"])
    assert obs.shape == (2 * (n + 1),)
    px, py, _ = ora.robot_pose()
    assert obs[0] == px and obs[1] == py
    assert obs[2 * n] == 0.0 and obs[2 * n + 1] == 0.0                 # never written by the reference's loop
    for i in range(1, n):
        assert np.array_equal(obs[2 * i: 2 * i + 2], _centroid_literal(info["obs"][i]))
    c0 = _centroid_literal(info["obs"][0])
    assert not np.any(np.all(obs.reshape(-1, 2)[1:] == c0, axis=1))     # obstacle 0 is absent


def test_ship_ice_global_observation_oracle(golden_layouts):
    """cfg.egocentric_obs = false (ship_ice_env.py:96, 397-407): 2 x 200 x 60 on 0.2 m cells — the 5 x 5 block mean of the
    1000 x 300 occupancy image of ALL floes (also those beyond the 12 m local range of the egocentric channel) and the planner
    footprint; checked against the literal skimage loops and a literal block mean."""
    from oracle.raster_oracle import draw_polygon_literal
    name = "c10_t0"
    trials = unpack_trials(golden_layouts[name])
    ora = ShipIceOracleEnv({"concentration": fixture_conc(name), "egocentric_obs": False}, experiment=trials, order_mode=1)
    obs, info = ora.reset(settle=False)
    assert obs.dtype == np.uint8 and obs.shape == (2, 200, 60)
    assert (ora.occupancy.occ_map_height, ora.occupancy.occ_map_width) == (200, 60)
    lit = np.zeros((1000, 300))
    for ob in ora.obstacles:
        v = np.asarray(ob) * 25
        rr, cc = draw_polygon_literal(v[:, 1], v[:, 0], shape=lit.shape)
        lit[rr, cc] = 1.0
    conc = np.zeros((200, 60))
    for i in range(200):
        for j in range(60):
            conc[i, j] = np.mean(lit[5 * i: 5 * i + 5, 5 * j: 5 * j + 5])
    assert np.array_equal(obs[0], (conc * 255).astype(np.uint8))
    assert len(np.unique(obs[0])) > 5 and obs[0].max() == 255          # partial cells: k / 25 * 255 truncated
    far = (np.abs(np.arange(200) * 0.2 - ora.ship_pose()[1]) > 14)
    assert obs[0][far].any(), "floes outside the egocentric range must appear"
    # footprint: zeros outside, 255 inside the ship polygon scaled to cells (2 x 0.5 m -> ~ 25 cells)
    assert set(np.unique(obs[1])) == {0, 255}
    assert 12 <= int((obs[1] == 255).sum()) <= 45
    px, py, _ = ora.ship_pose()
    assert obs[1][int(py * 5), int(px * 5)] == 255


def test_emulated_global_observation_and_centroid_cache_match_oracle(golden_layouts, pins, portable_trig):
    """The kernel sources of k_shipice_obs_global (render_obs_global) and the |centroid| cache that k_low_dim compacts,
    compiled for the host (tests/emu), against the oracle: rasters bit-exact at reset (raw polygons) and after steps
    (hull polygons); low-dim entries within 1e-9 (sequential vs pairwise summation, see test_gpu_low_dim)."""
    from tests.emu.emu import Emu
    from tests.test_kernel_logic_cpu import SHIP_VERTS
    name = "c10_t7"
    p = golden_layouts[name]
    emu = Emu(1, p, SHIP_VERTS)
    cfg = {"concentration": fixture_conc(name), "egocentric_obs": False}
    ora = ShipIceOracleEnv(cfg, experiment=unpack_trials(p), order_mode=1)
    obs, info = ora.reset()
    emu.reset(0, 0)
    assert np.array_equal(obs, emu.render_obs_global(0, True))
    for a in pins["%s_o1_actions" % name][:3]:
        obs, r, term, _, info = ora.step(float(a))
        emu.step(0, float(a))
        assert np.array_equal(ora.space.state(), emu.bodies(0, 0))
        assert np.array_equal(obs, emu.render_obs_global(0, False))
        low = ora.generate_observation_low_dim(info["obs"])
        np.testing.assert_allclose(emu.centroid_cache(0, p.max_floes), low, rtol=0, atol=1e-9)
