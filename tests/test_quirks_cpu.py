"""SURVEY.md appendix D, the parity checklist of reference quirks, item by item — the ones that have no dedicated test elsewhere.
Each test names the quirk (D<n>), the reference lines, and checks the oracle and the kernel sources (CPU emulation, tests/emu)
for exactly that behaviour.  (D4, D7, D8, D10, D11, D13, D17, D20 are asserted in the env / metric / pin tests next to the
code they concern; D3, D5, D6, D19 by the pins of tests/test_ref_pins_cpu.py, which compare with the reference's own modules.
D15 needs nothing: the ``move_sign`` expression of box_delivery_env.py:932-935 only feeds ``robot_new_position`` /
``robot_new_heading`` (:919-938), which the reference never reads again — the robot is driven by the DP controller.)"""
import os
import random
import sys

import numpy as np
import pytest

from oracle import cpo
from tests.conftest import ROOT


# ------------------------------------------------------------------ D18: spfa corner cutting, unreachable = 0
def _spfa_impls():
    from tests.emu.emu import Emu
    from tests.conftest import SHIP_VERTS
    from benchpush_b200.layouts import generate_experiment, pack_trials
    emu = Emu(1, pack_trials(generate_experiment(1, 0.1, seed=1)), SHIP_VERTS)
    impls = [("oracle", cpo.spfa), ("k_spfa (emulated)", emu.spfa), ("k_spfa_serial (emulated)", emu.spfa_serial)]
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(ref_dir) and any(f.startswith("spfa") for f in os.listdir(ref_dir)):
        sys.path.insert(0, ref_dir)
        import spfa as ref_spfa
        impls.append(("reference main.cpp", ref_spfa.spfa))
    return impls


def test_d18_spfa_cuts_corners_and_reports_unreachable_cells_as_zero():
    """deps/spfa/src/main.cpp:50-61 has no check of the two orthogonal neighbours of a diagonal move, and :101 writes 0 for
    cells the queue never reached (box_delivery_env.py:1117-1123 then turns them into distance 0 with np.minimum)."""
    m = np.array([[1, 0, 1, 0, 1],
                  [0, 1, 0, 0, 0],
                  [1, 0, 1, 0, 1]], dtype=bool)           # (0,0) -> (1,1) -> (0,2) / (2,0) / (2,2) only through blocked corners
    for name, f in _spfa_impls():
        d, p = f(m, (0, 0))
        s2 = np.float32(np.sqrt(np.float32(2.0)))
        assert d[1, 1] == s2, name                                           # diagonal step between two blocked orthogonal neighbours
        assert d[0, 2] == np.float32(s2 + s2) and d[2, 2] == np.float32(s2 + s2), name
        assert d[0, 4] == 0 and d[2, 4] == 0, name                           # walled off: reported as 0, like the source itself
        assert d[0, 1] == 0 and d[1, 0] == 0, name                           # blocked cells as well
        assert p[0, 0] == -1 or p[0, 0] == 0, name
    # a blocked source reaches nothing
    for name, f in _spfa_impls():
        d, _ = f(m, (0, 1))
        assert not d.any(), name


# ------------------------------------------------------------------ D14: inactivity truncation sets both flags, terminal obs None
@pytest.mark.parametrize("action_type", ["heading", "position"])
def test_d14_inactivity_truncation_sets_terminated_and_truncated(action_type, portable_trig):
    """box_delivery_env.py:801-804: ``terminated = truncated = True`` once the inactivity counter reaches the cut-off;
    :1048-1049: with position actions the terminal observation is None (the vector env keeps its previous image)."""
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
    from tests.emu import emu_bd
    lay = [sample_layout(np.random.RandomState(3))]
    cfg = {"agent": {"action_type": action_type}, "misc": {"inactivity_cutoff": 2}}
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg=cfg, layouts=lay)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=lay)
    env.reset(); ora.reset()
    # two steps that deliver nothing: a small turn on the spot / the robot's own pixel
    a = 0.05 if action_type == "heading" else float(112 * 224 + 112)
    flags = []
    for k in range(2):
        obs, r, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        oobs, orr, oterm, otrunc, oinfo = ora.step(a if action_type == "heading" else int(a))
        assert (bool(term[0]), bool(trunc[0])) == (bool(oterm), bool(otrunc))
        assert int(info["inactivity"][0]) == oinfo["inactivity"] == k + 1
        flags.append((bool(term[0]), bool(trunc[0])))
    assert flags == [(False, False), (True, True)]
    assert (oobs is None) == (action_type == "position")
    env.close()


# ------------------------------------------------------------------ D16: the first maze obstacle is never tested against the walls
def test_d16_first_maze_obstacle_ignores_the_walls():
    """maze_NAMO_env.py:291-297: the wall test sits inside the loop over the obstacles placed so far, so obstacle 0 may
    overlap a wall while every later obstacle keeps its distance."""
    from benchpush_b200.layouts import generate_maze_layout, maze_walls, _segment_query
    walls = maze_walls(1, 15.0, 15.0)
    touching_first, later_ok = 0, True
    for seed in range(400):
        lay = generate_maze_layout(random.Random(seed), walls, 5, 15.0, 1.5)
        hit = [any(_segment_query((x, y), a, b, 0.5, 1.5) for a, b in walls) for (x, y) in lay]
        touching_first += int(hit[0])
        later_ok = later_ok and not any(hit[1:])
    assert touching_first > 0, "with 400 seeds some first obstacle lands within min_dist of a wall"
    assert later_ok


# ------------------------------------------------------------------ D12: maze low-dim observation skips obstacle 0
def test_d12_maze_low_dim_observation_skips_obstacle_zero():
    """maze_NAMO_env.py:488-504: slots 0, 1 hold the robot, the obstacle loop runs from 1 and writes obstacle i to slots 2i, 2i + 1 —
    obstacle 0 never appears and the last pair stays zero.  Oracle and the kernel (k_low_dim, CPU emulation in
    tests/test_low_dim_cpu.py) agree; here the rule itself."""
    from oracle.maze_oracle import MazeNamoOracleEnv
    from oracle.ship_ice_oracle import poly_centroid
    layouts = [[[11.3, 5.2], [11.0, 7.5], [9.4, 12.4], [5.0, 12.6], [3.6, 8.0]]]
    ora = MazeNamoOracleEnv(layouts=layouts, order_mode=1)
    _, info = ora.reset()
    low = np.asarray(ora.generate_observation_low_dim(info["obs"]))
    n = len(info["obs"])
    assert low.shape == (2 * (n + 1),)
    px, py, _ = ora.robot_pose()
    assert low[0] == px and low[1] == py
    cents = [np.abs(poly_centroid(v)) for v in info["obs"]]
    for i in range(1, n):
        assert low[2 * i] == cents[i][0] and low[2 * i + 1] == cents[i][1], i
    assert low[2 * n] == 0 and low[2 * n + 1] == 0                                  # the last pair is never written
    assert not any(low[2 * i] == cents[0][0] and low[2 * i + 1] == cents[0][1] for i in range(1, n + 1))   # obstacle 0 is nowhere


# ------------------------------------------------------------------ D21 / D22: no TimeLimit in evaluation, reset(seed) ignored
def test_d21_d22_registry_limits_are_data_and_reset_ignores_the_seed(portable_trig):
    """environments/__init__.py:3-51 registers max_episode_steps, but the evaluation loops call ``.unwrapped`` so it never
    applies (the vector envs do not truncate on it either; the table is exported for callers who want a TimeLimit);
    ``reset(seed=...)`` is accepted and ignored (ship_ice_env.py:223)."""
    from benchpush_b200.envs import MAX_EPISODE_STEPS
    assert MAX_EPISODE_STEPS == {"ship-ice-v0": 300, "box-delivery-v0": 30000, "maze-NAMO-v0": 400, "area-clearing-v0": 30000}
    from oracle.box_delivery_oracle import sample_layout
    from tests.emu import emu_bd
    lay = [sample_layout(np.random.RandomState(3))]
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg={"agent": {"action_type": "heading"}}, layouts=lay)
    o1, i1 = env.reset(seed=123)
    o1, s1 = np.array(o1), np.array(i1["state"])
    o2, i2 = env.reset(seed=7)
    assert np.array_equal(o1, np.array(o2)) and np.array_equal(s1, np.array(i2["state"]))
    env.close()


# ------------------------------------------------------------------ D1 / D2: velocity direction frozen per env step, yaw limit zeroes w only
def test_d1_d2_velocity_direction_is_frozen_and_the_yaw_limit_only_zeroes_omega(portable_trig):
    """ship_ice_env.py:270-274 sets the ship's velocity VECTOR once per env step from the heading at step start, while the heading
    keeps integrating through the 400 substeps: in open water the ship moves on a straight segment along its initial heading and
    ends the step turned by omega * dt.  :284-286: past the yaw limits (theta <= 0 or >= pi) the angular velocity is zeroed, theta
    itself is never clamped, and the next env step sets omega again."""
    from benchpush_b200.layouts import generate_experiment, pack_trials, unpack_trials
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    from tests.conftest import SHIP_VERTS
    from tests.emu.emu import Emu
    t = generate_experiment(1, 0.1, seed=3)[0]
    packed = pack_trials([{"obstacles": [], "ship_state": t["ship_state"]}])
    emu = Emu(1, packed, SHIP_VERTS)
    ora = ShipIceOracleEnv({"concentration": 0.1}, experiment=unpack_trials(packed), order_mode=1)
    ora.reset(); emu.reset(0, 0)
    ship = ora.space.state().shape[0] - 1
    row = lambda: emu.bodies(0, 0)[ship]
    assert np.array_equal(ora.space.state(), emu.bodies(0, 0))
    x0, y0, th0 = row()[:3]
    ora.step(1.0); emu.step(0, 1.0)
    assert np.array_equal(ora.space.state(), emu.bodies(0, 0))
    x1, y1, th1 = row()[:3]
    dt, speed = float(ora.cfg["dt"]), float(ora.cfg["target_speed"])
    # straight segment along the heading at step start, although the heading itself has turned by max_yaw_rate_step * dt
    assert abs(np.arctan2(y1 - y0, x1 - x0) - th0) < 1e-12
    assert abs(np.hypot(x1 - x0, y1 - y0) - speed * dt) < 1e-9
    assert abs((th1 - th0) - ora.max_yaw_rate_step * dt) < 1e-9 and th1 - th0 > 0.05
    # drive the heading across pi: omega is zeroed in that substep (theta stops a little PAST the limit, unclamped) ...
    for k in range(40):
        ora.step(1.0); emu.step(0, 1.0)
        assert np.array_equal(ora.space.state(), emu.bodies(0, 0)), k
        if row()[2] >= np.pi:
            break
    th_lim = row()[2]
    assert th_lim >= np.pi and th_lim - np.pi < ora.max_yaw_rate_step * dt / 400 * 1.0001 and row()[5] == 0.0
    # ... and the next env step sets omega again: one more substep of turning before it is zeroed again, theta keeps growing
    ora.step(1.0); emu.step(0, 1.0)
    assert np.array_equal(ora.space.state(), emu.bodies(0, 0))
    assert row()[2] > th_lim and row()[5] == 0.0
