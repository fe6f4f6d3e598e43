"""Box-delivery kernel logic on a machine without a GPU: the kernel sources compiled for the host with one lane
(tests/emu/emu_bd.cpp, same entry points as the C ABI) against the oracle, through the same Python mirror
(BoxDeliveryVecEnv) the product uses."""
import numpy as np
import pytest

from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv, DEFAULT_CFG
from benchpush_b200.layouts import sample_box_delivery_layout
from benchpush_b200.utils import DotDict
from tests.boxdel_common import DELIVERY_CFG, DELIVERY_LAYOUT, compare_step
from tests.emu import emu_bd


def _pair(cfg, layouts, n=1, force_box_spfa=False):
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=n, cfg=cfg, layouts=layouts, force_box_spfa=force_box_spfa)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=layouts)
    return env, ora


def test_layout_sampler_matches_the_oracle_restatement():
    a = sample_layout(np.random.RandomState(42))
    b = sample_box_delivery_layout(np.random.RandomState(42), DotDict.to_dot_dict(DEFAULT_CFG))
    assert a["start"] == b["start"] and a["boxes"] == b["boxes"] and len(a["boxes"]) == 10


def test_room_maps_and_reset_observation_match_the_oracle(portable_trig):
    lay = [sample_layout(np.random.RandomState(5))]
    env, ora = _pair({"agent": {"action_type": "heading"}}, lay)
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    r = env.room
    assert r.shape == (430, 542)
    assert np.array_equal(r.configuration_space, ora.configuration_space)
    assert np.array_equal(r.configuration_space_thin, ora.configuration_space_thin)
    assert np.array_equal(r.closest_cspace_indices, ora.closest_cspace_indices)
    assert np.array_equal(r.small_obstacle_map, ora.small_obstacle_map)
    assert np.array_equal(r.robot_state_channel, ora.robot_state_channel)
    assert np.array_equal(obs[0], oobs)
    gb = env.get_bodies(0)
    assert np.array_equal(gb[:, :6], ora.space.state()[:11])
    assert np.allclose(info["state"][0], np.array(oinfo["state"]))
    env.close()


def test_heading_episode_matches_the_oracle(portable_trig):
    lay = [sample_layout(np.random.RandomState(42))]
    env, ora = _pair({"agent": {"action_type": "heading"}}, lay)
    env.reset(); ora.reset()
    rng = np.random.default_rng(0)
    worst, spfa = 0.0, 0
    for k in range(16):
        a = float(np.float32(rng.uniform(-1, 1)))
        d = compare_step(env, ora, 0, a, a, state_tol=1e-9)
        worst = max(worst, d["state"]); spfa += d["spfa"]
    assert spfa >= 1, "the scenario should exercise the spfa path branch (robot next to a wall)"
    env.close()


def test_visibility_shortcut_equals_the_per_box_spfa_solves(portable_trig):
    """The room's receptacle cell sees every free cell, so shortest_path_distance(box, receptacle) is Euclidean
    (BdParams::recept_visible); with the shortcut disabled the kernels run spfa + path extraction per box, as the
    reference does, and must produce the same episode."""
    lay = [sample_layout(np.random.RandomState(42))]
    cfg = {"agent": {"action_type": "heading"}}
    fast, ora = _pair(cfg, lay)
    slow, _ = _pair(cfg, lay, force_box_spfa=True)
    assert fast._L.bp_boxdel_launch_count(fast._h) == 1 and slow._L.bp_boxdel_launch_count(slow._h) == 0
    fast.reset(); slow.reset(); ora.reset()
    rng = np.random.default_rng(4)
    for k in range(6):
        a = float(np.float32(rng.uniform(-1, 1)))
        _, r1, *_ = fast.step(np.array([a], dtype=np.float32))
        out = slow.step(np.array([a], dtype=np.float32))
        compare_step(slow, ora, 0, a, a, step_out=out, state_tol=1e-9)
        assert float(r1[0]) == float(out[1][0])
        assert np.array_equal(fast.get_step_debug(0)["init_dist"], slow.get_step_debug(0)["init_dist"])
        assert np.array_equal(fast.get_bodies(0), slow.get_bodies(0))
    fast.close(); slow.close()


def test_delivery_removal_and_position_actions_match_the_oracle(portable_trig):
    env, ora = _pair(DELIVERY_CFG, [DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for k in range(3):
        compare_step(env, ora, 0, -1.0, -1.0, state_tol=1e-9)
    assert int(env._buf["info_i"][0, 0]) == 1 and int(env._buf["info_i"][0, 2]) == 3
    env.close()
    cfg = dict(DELIVERY_CFG, agent={"action_type": "position"})
    env, ora = _pair(cfg, [DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for pix in (40 * 224 + 112, 112 * 224 + 30, 200 * 224 + 200):
        compare_step(env, ora, 0, float(pix), pix, state_tol=1e-9)
    env.close()


def test_sam_job_type_matches_the_oracle(portable_trig):
    """``train.job_type: sam`` (box_delivery_env.py:118-147, config.yaml): 96 x 96 observation (``local_map_pixel_width_sam``),
    the ``rewards_sam`` table and ``inactivity_cutoff_sam``; heading and position actions against the oracle."""
    lay = [sample_layout(np.random.RandomState(42))]
    cfg = {"agent": {"action_type": "heading"}, "train": {"job_type": "sam"}}
    env, ora = _pair(cfg, lay)
    assert env.local_map_pixel_width == 96 and env.observation_shape == (96, 96, 4) and ora.local_map_pixel_width == 96
    obs, _ = env.reset(); oobs, _ = ora.reset()
    assert obs.shape[1:] == (96, 96, 4) and np.array_equal(obs[0], oobs)
    rng = np.random.default_rng(1)
    for k in range(6):
        a = float(np.float32(rng.uniform(-1, 1)))
        compare_step(env, ora, 0, a, a, state_tol=1e-9)
    env.close()
    cfg = dict(DELIVERY_CFG, agent={"action_type": "position"}, train={"job_type": "sam"})
    env, ora = _pair(cfg, [DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for pix in (20 * 96 + 48, 48 * 96 + 12, 85 * 96 + 85):
        compare_step(env, ora, 0, float(pix), pix, state_tol=1e-9)
    env.close()


def test_envs_of_a_batch_are_independent(portable_trig):
    lays = [sample_layout(np.random.RandomState(s)) for s in (1, 2)]
    cfg = {"agent": {"action_type": "heading"}}
    both = emu_bd.EmuBoxDeliveryVecEnv(num_envs=2, cfg=cfg, layouts=lays)
    one = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg=cfg, layouts=lays[1:])
    both.reset(); one.reset()
    for a in (0.3, -0.6):
        both.step(np.array([0.9, a], dtype=np.float32))
        one.step(np.array([a], dtype=np.float32))
    assert np.array_equal(both.get_bodies(1), one.get_bodies(0))
    assert np.array_equal(both.obs[1], one.obs[0])
    # masked step leaves the other env untouched
    before = both.get_bodies(0).copy()
    both.step(np.array([0.1, 0.1], dtype=np.float32), mask=np.array([0, 1], dtype=np.uint8))
    assert np.array_equal(both.get_bodies(0), before)
    both.close(); one.close()


def test_velocity_actions_match_the_oracle(portable_trig):
    cfg = dict(DELIVERY_CFG, agent={"action_type": "velocity"})
    env, ora = _pair(cfg, [DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for act in ([0.25, 0.1], [0.9, -0.4], [-0.2, 0.0], [0.3, 0.2]):
        a = np.array([act], dtype=np.float32)
        out = env.step(a)
        oobs, orr, oterm, otrunc, oinfo = ora.step(a[0])
        gb = env.get_bodies(0); ob = ora.space.state()
        alive = gb[1:, 6] == 1
        assert np.abs(gb[0, :3] - ob[0, :3]).max() <= 1e-9 and np.abs(gb[1:][alive][:, :3] - ob[list(ora.box_bodies)][:, :3]).max() <= 1e-9
        assert abs(float(out[1][0]) - orr) <= 1e-9 and env.get_step_debug(0)["substeps"] == 1000 + ora.n_substeps
        assert abs(float(out[4]["cumulative_distance"][0]) - oinfo["cumulative_distance"]) <= 1e-9
        assert int((out[0][0] != oobs).sum()) <= 4
    env.close()


def test_masked_reset_advances_only_the_masked_envs_episode():
    """Every env keeps its own episode counter: a masked reset moves only the masked envs to their next layout, and the
    default pool holds several (env, episode) layouts per env (the reference re-randomises at every reset)."""
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from tests.emu import emu_bd
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=3, cfg={"agent": {"action_type": "heading"}})
    assert len(env.layouts) == 3 * 4
    env.reset()
    s0 = [tuple(env.get_bodies(e)[0, :3]) for e in range(3)]
    env.reset(mask=np.array([0, 1, 0], dtype=np.uint8))
    s1 = [tuple(env.get_bodies(e)[0, :3]) for e in range(3)]
    assert s1[0] == s0[0] and s1[2] == s0[2] and s1[1] != s0[1]
    assert env.episode_idx.tolist() == [0, 1, 0]
    assert np.allclose(s1[1], env.layouts[1 + 1 * 3]["start"])
    env.reset(mask=np.array([1, 0, 0], dtype=np.uint8))
    assert np.allclose(env.get_bodies(0)[0, :3], env.layouts[0 + 1 * 3]["start"])     # env 0's episode 1, whenever it is reached
    env.close()


def test_room_env_spaces_follow_the_reference():
    """action_space / observation_space (box_delivery_env.py:157-173, area_clearing.py:200-221): shapes, bounds, dtypes."""
    from benchpush_b200.envs.area_clearing import AreaClearingVecEnv
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from tests.emu import emu_bd
    bd = emu_bd.EmuBoxDeliveryVecEnv(num_envs=2, cfg={"agent": {"action_type": "heading"}})
    assert bd.single_action_space.shape == (1,) and bd.single_action_space.dtype == np.float32
    assert float(bd.single_action_space.low[0]) == -1 and float(bd.single_action_space.high[0]) == 1
    assert bd.single_observation_space.shape == (bd.local_map_pixel_width, bd.local_map_pixel_width, 4) and bd.single_observation_space.dtype == np.uint8
    assert bd.action_space.shape == (2, 1) and bd.observation_space.shape == (2,) + bd.single_observation_space.shape
    assert bd.single_action_space.sample() in bd.single_action_space
    bd.close()
    bp = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg={"agent": {"action_type": "position"}})
    assert bp.single_action_space.shape == () and float(bp.single_action_space.high) == bp.local_map_pixel_width ** 2
    bp.close()
    bv = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg={"agent": {"action_type": "velocity"}})
    assert bv.single_action_space.shape == (2,)
    bv.close()
    ac = emu_bd.EmuAreaClearingVecEnv(num_envs=1, cfg={"agent": {"action_type": "position"}})
    assert ac.single_action_space.shape == (1,) and ac.single_action_space.dtype == np.int32
    ac.close()


def test_low_dim_state_observation_of_the_room_envs():
    """cfg.low_dim_state (box_delivery_env.py:166-169, 1025-1037; area_clearing.py:212-215, 908-920): reset / step return
    <|cx|, |cy|> of every box left instead of the image, the space is Box(-10, 30, (2 * num_boxes,))."""
    import random
    from benchpush_b200.envs.area_clearing import AreaClearingVecEnv
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from oracle.area_clearing_oracle import sample_layout as ac_layout
    from oracle.raster_oracle import poly_centroid
    from tests.emu import emu_bd
    for mk in (lambda: emu_bd.EmuBoxDeliveryVecEnv(num_envs=2, cfg={"low_dim_state": True, "agent": {"action_type": "heading"}}),
               lambda: emu_bd.EmuAreaClearingVecEnv(num_envs=2, cfg={"low_dim_state": True}, layouts=[ac_layout(random.Random(s)) for s in (1, 2)])):
        env = mk()
        obs, info = env.reset()
        NB = env.num_boxes
        assert obs.shape == (2, 2 * NB) and env.single_observation_space.shape == (2 * NB,)
        obs, r, term, trunc, info = env.step(np.array([0.2, -0.4], dtype=np.float32))
        for e in range(2):
            want = np.concatenate([np.abs(poly_centroid(np.asarray(info["obs"][e][i]))) for i in range(int(info["n_boxes_left"][e]))])
            np.testing.assert_allclose(obs[e][:len(want)], want, rtol=1e-12)
            assert not obs[e][len(want):].any()
        env.close()


# ---- obstacle_config small_columns / large_columns / large_divider: per-layout rooms (static shapes + maps indexed by layout)
@pytest.mark.parametrize("oc,seed,steps,device_rooms", [("small_columns", 4, 10, True), ("large_columns", 7, 4, True), ("large_divider", 2, 9, True),
                                                        ("small_columns", 4, 3, False), ("large_divider", 2, 2, False)])
def test_column_and_divider_rooms_match_the_oracle(portable_trig, oc, seed, steps, device_rooms):
    """The reference redraws columns / the divider at every reset and rebuilds its configuration space
    (box_delivery_env.py:358-471, 1140-1175).  device_rooms: the env's maps are built by the kernel sources at reset
    (csrc/boxdel_room.cuh: fillPoly, disk dilation, scipy's distance_transform_edt indices incl. tie-breaks, receptacle distance
    field) and equal, byte for byte, the host build that calls cv2 / scipy and the oracle's maps; otherwise the host-built maps
    of the layout are uploaded.  With the reference solver's own parent order (exact_spfa_parents) reset and every step — poses,
    path, per-box path distances through the reward, observation bytes — match the oracle; the receptacle no longer sees every
    free cell, so the per-box solves run."""
    cfg = {"agent": {"action_type": "heading"}, "env": {"obstacle_config": oc}}
    lay = [sample_layout(np.random.RandomState(seed), cfg)]
    assert lay[0]["columns"] or lay[0]["divider"]
    from benchpush_b200.envs.box_delivery import RoomMaps
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg=cfg, layouts=lay, exact_spfa_parents=True, device_rooms=device_rooms)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=lay)
    assert env.device_rooms == device_rooms
    assert env._L.bp_boxdel_launch_count(env._h) == 0            # emulation: reports recept_visible
    obs, info = env.reset()
    oobs, oinfo = ora.reset()
    m = env.get_room_maps(0)
    host = RoomMaps(env.cfg, lay[0])                             # cv2.fillPoly / scipy binary_dilation / distance_transform_edt
    for key, o in (("configuration_space", ora.configuration_space), ("configuration_space_thin", ora.configuration_space_thin),
                   ("closest_cspace_indices", ora.closest_cspace_indices), ("small_obstacle_map", ora.small_obstacle_map)):
        assert np.array_equal(m[key], np.asarray(getattr(host, key)).astype(m[key].dtype)), key
        assert np.array_equal(m[key], np.asarray(o).astype(m[key].dtype)), key
    assert np.array_equal(m["receptacle_map"], ora.create_global_shortest_path_to_receptacle_map().astype(np.float32))
    assert np.array_equal(obs[0], oobs)
    assert np.array_equal(env.get_bodies(0)[:, :6], ora.space.state()[:1 + env.num_boxes])
    rng = np.random.default_rng(0)
    spfa = 0
    for k in range(steps):
        a = float(np.float32(rng.uniform(-1, 1)))
        spfa += compare_step(env, ora, 0, a, a, state_tol=1e-9)["spfa"]
    if oc != "large_columns" and steps >= 9:
        assert spfa >= 1
    env.close()


def test_rooms_are_indexed_by_layout_and_deterministic_parents_keep_the_physics(portable_trig):
    """Three envs on three different rooms of one handle (0, 1 and 2 columns): each matches its own oracle at reset (maps, raster,
    bodies), and with the deterministic parents of the fast mode (exact_spfa_parents=False) the poses still match while the robot's path needs no spfa solve
    (box path distances may follow another equal-cost staircase: DESIGN 5.4)."""
    cfg = {"agent": {"action_type": "heading"}, "env": {"obstacle_config": "small_columns"}}
    lays = [sample_layout(np.random.RandomState(s), cfg) for s in (2, 0, 4)]
    assert [len(l["columns"]) for l in lays] == [0, 1, 2]
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=3, cfg=cfg, layouts=lays, exact_spfa_parents=False)
    oras = [BoxDeliveryOracleEnv(cfg=cfg, layouts=[l]) for l in lays]
    obs, info = env.reset()
    for e, ora in enumerate(oras):
        oobs, _ = ora.reset()
        assert np.array_equal(env.get_room_maps(e)["configuration_space"], ora.configuration_space)
        assert np.array_equal(obs[e], oobs), e
        assert np.array_equal(env.get_bodies(e)[:, :6], ora.space.state()[:1 + env.num_boxes])
    acts = np.array([0.3, -0.6, 0.9], dtype=np.float32)
    env.step(acts)
    for e, ora in enumerate(oras):
        ora.step(float(acts[e]), want_obs=False) if "want_obs" in ora.step.__code__.co_varnames else ora.step(float(acts[e]))
        if env.get_step_debug(e)["need_spfa"] == 0:
            gb = env.get_bodies(e)
            alive = gb[1:, 6] == 1
            assert np.abs(gb[0, :3] - ora.space.state()[0, :3]).max() <= 1e-9
            assert np.abs(gb[1:][alive][:, :3] - ora.space.state()[list(ora.box_bodies)][:, :3]).max() <= 1e-9
    env.close()


def test_device_feature_transform_reproduces_scipy_including_ties():
    """bd_feature_transform (csrc/boxdel_room.cuh: the nearest-free-cell map of the on-device room build) against
    scipy.ndimage.distance_transform_edt(..., return_indices=True) itself: random maps of several densities, block-structured
    maps (many equidistant candidates), a single free cell, a fully free map."""
    import ctypes as C
    from scipy.ndimage import distance_transform_edt
    L = emu_bd.lib()
    rng = np.random.default_rng(0)
    cases = []
    for t in range(40):
        H, W = int(rng.integers(6, 70)), int(rng.integers(6, 70))
        blocked = (rng.random((H, W)) < rng.choice([0.3, 0.8, 0.97])).astype(np.uint8)
        if t % 3 == 0:
            blocked[:] = 1
            for _ in range(int(rng.integers(1, 6))):
                a, b = int(rng.integers(0, H - 2)), int(rng.integers(0, W - 2))
                blocked[a:a + int(rng.integers(1, 9)), b:b + int(rng.integers(1, 9))] = 0
        if (blocked == 0).sum() == 0:
            blocked[H // 2, W // 2] = 0
        cases.append(blocked)
    one = np.ones((31, 47), np.uint8); one[13, 20] = 0
    cases += [one, np.zeros((9, 12), np.uint8)]
    for blocked in cases:
        H, W = blocked.shape
        ref = distance_transform_edt(blocked, return_distances=False, return_indices=True)
        free = np.ascontiguousarray(1 - blocked)
        out = np.zeros((2, H, W), dtype=np.int32)
        L.emu_feature_transform(free.ctypes.data_as(C.c_void_p), C.c_int(H), C.c_int(W), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, ref.astype(np.int32)), (H, W)
