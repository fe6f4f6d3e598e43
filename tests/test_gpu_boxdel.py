"""box-delivery-v0 on the B200 through the C ABI (libbenchpush_b200.so) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from tests.boxdel_common import DELIVERY_CFG, DELIVERY_LAYOUT, compare_step

pytestmark = pytest.mark.gpu


def _mk(cfg, layouts, n):
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    return BoxDeliveryVecEnv(num_envs=n, cfg=cfg, layouts=layouts, device=0)


def test_reset_and_heading_episode_match_the_oracle(portable_trig):
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
    cfg = {"agent": {"action_type": "heading"}}
    lays = [sample_layout(np.random.RandomState(40 + s)) for s in range(48)]
    env = _mk(cfg, lays, 48)
    picks = [2, 37]                                  # layout 2 == RandomState(42): starts next to the wall (spfa paths)
    oras = {e: BoxDeliveryOracleEnv(cfg=cfg, layouts=[lays[e]]) for e in picks}
    obs, info = env.reset()
    for e, ora in oras.items():
        oobs, oinfo = ora.reset()
        assert np.array_equal(obs[e].cpu().numpy(), oobs), "reset observation differs from the oracle"
        assert np.array_equal(env.get_bodies(e)[:, :6], ora.space.state()[:11])
    rng = np.random.default_rng(0)
    spfa = 0
    for k in range(10):
        a = rng.uniform(-1, 1, size=48).astype(np.float32)
        out = env.step(a)
        for e, ora in oras.items():
            # tolerance: the controller's atan2 / sin / cos are CUDA's vs numpy's (<= 1 ulp apart); contacts amplify that slowly
            d = compare_step(env, ora, e, None, float(a[e]), step_out=out, state_tol=1e-7)
            spfa += d["spfa"]
    assert spfa >= 1
    assert int((env._buf["info_i"][:, 3] & ~16).max()) == 0, "an env reported a capacity status"
    env.close()


def test_delivery_and_position_actions_match_the_oracle(portable_trig):
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv
    env = _mk(DELIVERY_CFG, [DELIVERY_LAYOUT], 1)
    ora = BoxDeliveryOracleEnv(cfg=DELIVERY_CFG, layouts=[DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for k in range(3):
        compare_step(env, ora, 0, -1.0, -1.0, state_tol=1e-8)
    assert int(env._buf["info_i"][0, 0]) == 1 and int(env._buf["info_i"][0, 2]) == 3
    env.close()
    cfg = dict(DELIVERY_CFG, agent={"action_type": "position"})
    env = _mk(cfg, [DELIVERY_LAYOUT], 1)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=[DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for pix in (40 * 224 + 112, 112 * 224 + 30, 200 * 224 + 200):
        compare_step(env, ora, 0, float(pix), pix, state_tol=1e-8)
    env.close()


def test_batch_independence_masks_and_host_entry_point():
    import torch
    from oracle.box_delivery_oracle import sample_layout
    cfg = {"agent": {"action_type": "heading"}}
    lays = [sample_layout(np.random.RandomState(s)) for s in range(300, 332)]
    big = _mk(cfg, lays, 32)
    one = _mk(cfg, lays[17:18], 1)
    big.reset(); one.reset()
    rng = np.random.default_rng(3)
    for k in range(3):
        a = rng.uniform(-1, 1, size=32).astype(np.float32)
        big.step(a)
        one.step(a[17:18])
    assert np.array_equal(big.get_bodies(17), one.get_bodies(0)), "an env's trajectory depends on its batch"
    assert torch.equal(big.obs[17], one.obs[0])
    before = big.get_bodies(5).copy()
    m = np.ones(32, dtype=np.uint8); m[5] = 0
    big.step(rng.uniform(-1, 1, size=32).astype(np.float32), mask=m)
    assert np.array_equal(big.get_bodies(5), before)
    # host-buffer entry point == device entry point
    from benchpush_b200.envs.box_delivery import BoxDelOutputs
    two = _mk(cfg, lays[17:18], 1)
    two.reset()
    a = np.array([0.25], dtype=np.float32)
    ho = BoxDelOutputs()
    hobs = np.zeros((1, 224, 224, 4), dtype=np.uint8); hrew = np.zeros(1); hinfo = np.zeros((1, 8))
    ho.obs = hobs.ctypes.data; ho.reward = hrew.ctypes.data; ho.info = hinfo.ctypes.data
    rc = two._L.bp_boxdel_step_host(two._h, a.ctypes.data, C.byref(ho))
    assert rc == 0
    three = _mk(cfg, lays[17:18], 1)
    three.reset()
    obs, rew, term, trunc, info = three.step(a)
    assert np.array_equal(obs.cpu().numpy(), hobs) and float(rew[0]) == hrew[0]
    assert np.array_equal(info["total_work"].cpu().numpy(), hinfo[:, 5])
    for e in (big, one, two, three):
        e.close()


def test_full_size_batch_runs_clean():
    """BASELINE.json config 2: 4096 envs, 10 boxes, heading actions — properties that do not need the oracle."""
    import torch
    n = 4096
    env = _mk({"agent": {"action_type": "heading"}}, None, n)
    obs, info = env.reset()
    assert obs.shape == (n, 224, 224, 4)
    rng = np.random.default_rng(0)
    work0 = None
    for k in range(2):
        obs, rew, term, trunc, info = env.step(torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32)))
    torch.cuda.synchronize()
    assert int((info["status"] & ~16).max()) == 0       # 16 = an env ran into the reference's STEP_LIMIT (a robot spinning against boxes): informational
    assert bool(torch.isfinite(rew).all()) and bool((info["total_work"] >= 0).all())
    assert int(info["n_boxes_left"].max()) <= 10 and bool((info["cumulative_boxes"] + info["n_boxes_left"] == 10).all())
    o = obs.cpu().numpy()
    assert set(np.unique(o[..., 1])) <= {0, 255}
    assert set(np.unique(o[..., 0])) <= {0, 31, 95, 127, 191}
    # the robot footprint is always at the window centre
    assert (o[:, 112, 112, 0] == 191).all()
    env.close()


def test_phase_per_kernel_loop_equals_the_fused_loop(monkeypatch):
    """cfg.phase_kernels = 1 runs the substep loop as one kernel per phase over the whole batch (same device functions,
    global-memory loop state): the trajectories must be identical to the default fused kernel."""
    import torch
    from oracle.box_delivery_oracle import sample_layout
    cfg = {"agent": {"action_type": "heading"}}
    lays = [sample_layout(np.random.RandomState(s)) for s in range(500, 516)]
    monkeypatch.setenv("BENCHPUSH_BD_PHASED", "0")
    a = _mk(cfg, lays, 16)
    monkeypatch.setenv("BENCHPUSH_BD_PHASED", "1")
    b = _mk(cfg, lays, 16)
    a.reset(); b.reset()
    rng = np.random.default_rng(9)
    for k in range(3):
        act = rng.uniform(-1, 1, size=16).astype(np.float32)
        oa = a.step(act); ob = b.step(act)
        assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1])
    for e in (0, 7, 15):
        assert np.array_equal(a.get_bodies(e), b.get_bodies(e))
    a.close(); b.close()


def test_velocity_actions_match_the_oracle(portable_trig):
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv
    cfg = dict(DELIVERY_CFG, agent={"action_type": "velocity"})
    env = _mk(cfg, [DELIVERY_LAYOUT], 1)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=[DELIVERY_LAYOUT])
    env.reset(); ora.reset()
    for act in ([0.25, 0.1], [0.9, -0.4], [-0.2, 0.0]):
        a = np.array([act], dtype=np.float32)
        out = env.step(a)
        oobs, orr, oterm, otrunc, oinfo = ora.step(a[0])
        gb = env.get_bodies(0); ob = ora.space.state()
        alive = gb[1:, 6] == 1
        assert np.abs(gb[0, :3] - ob[0, :3]).max() <= 1e-8 and np.abs(gb[1:][alive][:, :3] - ob[list(ora.box_bodies)][:, :3]).max() <= 1e-8
        assert abs(float(out[1][0]) - orr) <= 1e-8 and env.get_step_debug(0)["substeps"] == 1000 + ora.n_substeps
        assert int((out[0][0].cpu().numpy() != oobs).sum()) <= 4
    env.close()


def test_literal_per_box_spfa_path_on_the_gpu(portable_trig):
    """force_box_spfa: k_bd_dist solves spfa per box, walks the parents and runs Douglas-Peucker + line-of-sight pruning (the
    reference's literal shortest_path_distance, box_delivery_env.py:1209-1268) instead of the exact visibility shortcut.
    On the GPU both must give the same episode as the oracle and as each other."""
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
    cfg = {"agent": {"action_type": "heading"}}
    lays = [sample_layout(np.random.RandomState(42 + s)) for s in range(4)]
    fast = BoxDeliveryVecEnv(num_envs=4, cfg=cfg, layouts=lays, device=0)
    slow = BoxDeliveryVecEnv(num_envs=4, cfg=cfg, layouts=lays, device=0, force_box_spfa=True)
    ora = BoxDeliveryOracleEnv(cfg=cfg, layouts=[lays[0]])
    fast.reset(); slow.reset(); ora.reset()
    rng = np.random.default_rng(4)
    for k in range(6):
        a = rng.uniform(-1, 1, size=4).astype(np.float32)
        out_f = fast.step(a)
        out_s = slow.step(a)
        compare_step(slow, ora, 0, None, float(a[0]), step_out=out_s, state_tol=1e-7)
        assert np.array_equal(out_f[1].cpu().numpy(), out_s[1].cpu().numpy())                  # rewards (partial rewards use the box distances)
        for e in range(4):
            assert np.array_equal(fast.get_step_debug(e)["init_dist"], slow.get_step_debug(e)["init_dist"])
            assert np.array_equal(fast.get_bodies(e), slow.get_bodies(e))
    fast.close(); slow.close()


@pytest.mark.parametrize("oc,seeds,steps", [("small_columns", (4, 0, 2, 7), 8), ("large_columns", (7, 1), 4), ("large_divider", (2, 5), 6)])
def test_column_and_divider_rooms_match_the_oracle_gpu(portable_trig, oc, seeds, steps):
    """obstacle_config small_columns / large_columns / large_divider (box_delivery_env.py:358-471): every layout carries its own
    room (bp_boxdel_load_rooms).  With the reference solver's parent order (exact_spfa_parents) reset and every step of every
    env match the oracle; the fast-mode handle (deterministic parents) reproduces the same reset and the same physics
    for the steps whose robot path is a straight line."""
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv, sample_layout
    cfg = {"agent": {"action_type": "heading"}, "env": {"obstacle_config": oc}}
    lays = [sample_layout(np.random.RandomState(s), cfg) for s in seeds]
    n = len(lays)
    from benchpush_b200.envs.box_delivery import RoomMaps
    env = BoxDeliveryVecEnv(num_envs=n, cfg=cfg, layouts=lays, device=0, exact_spfa_parents=True)            # rooms built on the device at reset
    fast = BoxDeliveryVecEnv(num_envs=n, cfg=cfg, layouts=lays, device=0, exact_spfa_parents=False, device_rooms=False)   # host-built maps uploaded
    assert env.device_rooms and not fast.device_rooms
    oras = [BoxDeliveryOracleEnv(cfg=cfg, layouts=[l]) for l in lays]
    obs, info = env.reset()
    fobs, _ = fast.reset()
    assert np.array_equal(obs.cpu().numpy(), fobs.cpu().numpy())
    for e, ora in enumerate(oras):
        oobs, _ = ora.reset()
        m, host = env.get_room_maps(e), RoomMaps(env.cfg, lays[e])
        for key in ("configuration_space", "configuration_space_thin", "closest_cspace_indices", "small_obstacle_map"):
            assert np.array_equal(m[key], np.asarray(getattr(host, key)).astype(m[key].dtype)), (e, key)      # device build == cv2 / scipy
            assert np.array_equal(m[key], np.asarray(getattr(ora, key)).astype(m[key].dtype)), (e, key)
            assert np.array_equal(fast.get_room_maps(e)[key], m[key]), (e, key)
        assert np.array_equal(m["receptacle_map"], ora.create_global_shortest_path_to_receptacle_map().astype(np.float32))
        assert np.array_equal(fast.get_room_maps(e)["receptacle_map"], m["receptacle_map"])
        assert np.array_equal(obs[e].cpu().numpy(), oobs), e
        assert np.array_equal(env.get_bodies(e)[:, :6], ora.space.state()[:1 + env.num_boxes])
    rng = np.random.default_rng(0)
    same = np.ones(n, dtype=bool)
    for k in range(steps):
        a = rng.uniform(-1, 1, size=n).astype(np.float32)
        out = env.step(a)
        fast.step(a)
        for e, ora in enumerate(oras):
            compare_step(env, ora, e, None, float(a[e]), step_out=out, state_tol=1e-7)
            same[e] &= fast.get_step_debug(e)["need_spfa"] == 0
            if same[e]:
                assert np.array_equal(fast.get_bodies(e), env.get_bodies(e)), (k, e)
    assert int((env._buf["info_i"][:, 3] & ~16).max()) == 0
    env.close(); fast.close()
