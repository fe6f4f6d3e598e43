"""area-clearing-v0 on the B200 through the C ABI (libbenchpush_b200.so, env_kind 1) against the CPU oracle."""
import random

import numpy as np
import pytest

from tests.areaclr_common import CLEARING_CFG, CLEARING_LAYOUT, compare_step

pytestmark = pytest.mark.gpu


def test_reset_and_episode_match_the_oracle(portable_trig):
    from benchpush_b200.envs import AreaClearingVecEnv
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    lays = [sample_layout(random.Random(s)) for s in range(1, 33)]
    env = AreaClearingVecEnv(num_envs=32, layouts=lays, device=0)
    picks = [0, 21]
    oras = {e: AreaClearingOracleEnv(layouts=[lays[e]]) for e in picks}
    obs, info = env.reset()
    for e, ora in oras.items():
        oobs, _ = ora.reset()
        assert np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
        assert np.array_equal(obs[e].cpu().numpy(), oobs), "reset observation differs from the oracle"
        assert np.array_equal(env.get_bodies(e)[:, :6], ora.space.state()[:11])
    rng = np.random.default_rng(1)
    for k in range(8):
        a = rng.uniform(-1, 1, size=32).astype(np.float32)
        out = env.step(a)
        for e, ora in oras.items():
            compare_step(env, ora, e, float(a[e]), out, state_tol=1e-7)
    env.close()


def test_clearing_scenario_matches_the_oracle(portable_trig):
    from benchpush_b200.envs import AreaClearingVecEnv
    from oracle.area_clearing_oracle import AreaClearingOracleEnv
    env = AreaClearingVecEnv(num_envs=1, cfg=CLEARING_CFG, layouts=[CLEARING_LAYOUT], device=0)
    ora = AreaClearingOracleEnv(cfg=CLEARING_CFG, layouts=[CLEARING_LAYOUT])
    env.reset(); ora.reset()
    counts = []
    for k in range(5):
        out = env.step(np.array([-1.0], dtype=np.float32))
        compare_step(env, ora, 0, -1.0, out, state_tol=1e-8)
        counts.append(int(out[4]["box_count"][0]))
    assert counts == [0, 0, 1, 1, 1]
    env.close()


def test_full_size_batch_properties():
    """BASELINE.json config 3: 16 384 envs, 10 boxes, heading actions."""
    import torch
    from benchpush_b200.envs import AreaClearingVecEnv
    n = 16384
    env = AreaClearingVecEnv(num_envs=n, device=0)
    obs, info = env.reset()
    assert obs.shape == (n, 224, 224, 4)
    rng = np.random.default_rng(0)
    for k in range(2):
        obs, rew, term, trunc, info = env.step(torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32)))
    torch.cuda.synchronize()
    assert int((info["status"] & ~16).max()) == 0
    assert bool(torch.isfinite(rew).all()) and bool((info["total_work"] >= 0).all())
    assert int(info["box_count"].max()) <= 10 and int(info["t"].max()) <= 2
    o = obs[:256].cpu().numpy()
    assert set(np.unique(o[..., 0])) <= {0, 31, 95, 127, 159, 223} and set(np.unique(o[..., 1])) <= {0, 255}
    assert (o[:, 112, 112, 0] == 159).all()                       # the robot footprint is at the window centre
    env.close()


@pytest.mark.parametrize("name", ["walled_env", "walled_env_with_columns"])
def test_walled_environments_match_the_oracle_gpu(name, portable_trig):
    """Inner walls / static obstacles (area_clearing/envs/walled_env*.yaml) on the GPU: maps, reset and an episode per env."""
    from benchpush_b200.envs import AreaClearingVecEnv
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    cfg = {"env": name, "num_obstacles": 6}
    lays = [sample_layout(random.Random(s), cfg) for s in range(7, 15)]
    env = AreaClearingVecEnv(num_envs=8, cfg=cfg, layouts=lays, device=0, exact_spfa_parents=False)      # fast mode: deterministic parents
    picks = [0, 5]
    oras = {e: AreaClearingOracleEnv(cfg=cfg, layouts=[lays[e]]) for e in picks}
    obs, info = env.reset()
    for e, ora in oras.items():
        oobs, _ = ora.reset()
        assert np.array_equal(env.room.configuration_space, ora.configuration_space)
        assert np.array_equal(env.goal_point_global_map, ora.goal_point_global_map)
        assert np.array_equal(obs[e].cpu().numpy(), oobs)
    rng = np.random.default_rng(2)
    compared = 0
    for k in range(8):
        a = rng.uniform(-1, 1, size=8).astype(np.float32)
        out = env.step(a)
        for e in list(oras):
            d = compare_step(env, oras[e], e, float(a[e]), out, state_tol=1e-7, allow_path_ties=True)
            if d.get("tie"):
                del oras[e]              # equal-cost spfa parent chains: the trajectories part ways here (statistical parity only)
            else:
                compared += 1
    assert compared >= 8
    env.close()


def test_walled_columns_with_the_reference_parent_order_gpu(portable_trig):
    """exact_spfa_parents: the robot's spfa paths follow the reference solver's own parent tie-breaks (one thread restates its
    queue, csrc/spfa.cuh spfa_serial_parents), so every step of every compared env matches the oracle — no tie allowance."""
    from benchpush_b200.envs import AreaClearingVecEnv
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    cfg = {"env": "walled_env_with_columns", "num_obstacles": 6}
    lays = [sample_layout(random.Random(s), cfg) for s in range(7, 11)]
    env = AreaClearingVecEnv(num_envs=4, cfg=cfg, layouts=lays, device=0, exact_spfa_parents=True)
    oras = {e: AreaClearingOracleEnv(cfg=cfg, layouts=[lays[e]]) for e in (0, 1, 3)}
    env.reset()
    for ora in oras.values():
        ora.reset()
    rng = np.random.default_rng(2)
    spfa = 0
    for k in range(8):
        a = rng.uniform(-1, 1, size=4).astype(np.float32)
        out = env.step(a)
        for e, ora in oras.items():
            compare_step(env, ora, e, float(a[e]), out, state_tol=1e-7)
            spfa += env.get_step_debug(e)["need_spfa"]
    assert spfa >= 1, "the scenario should route at least one path around a wall or column"
    env.close()


def test_velocity_actions_and_low_dim_state_gpu(portable_trig):
    """action_type velocity (area_clearing.py:656-663) against the oracle; cfg.low_dim_state returns the centroid vector."""
    from benchpush_b200.envs import AreaClearingVecEnv
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    cfg = {"env": "walled_env", "num_obstacles": 5, "agent": {"action_type": "velocity"}}
    lays = [sample_layout(random.Random(s), cfg) for s in (11, 12)]
    env = AreaClearingVecEnv(num_envs=2, cfg=cfg, layouts=lays, device=0)
    ora = AreaClearingOracleEnv(cfg=cfg, layouts=[lays[1]])
    env.reset(); ora.reset()
    rng = np.random.default_rng(3)
    for k in range(10):
        a = rng.uniform(-1, 1, (2, 2)).astype(np.float32)
        a[:, 0] = np.abs(a[:, 0])
        compare_step(env, ora, 1, a[1], env.step(a), state_tol=1e-7)
    env.close()
    low = AreaClearingVecEnv(num_envs=2, cfg={"low_dim_state": True, "num_obstacles": 5}, layouts=lays, device=0)
    obs, _ = low.reset()
    assert tuple(obs.shape) == (2, 10) and low.single_observation_space.shape == (10,)
    from oracle.raster_oracle import poly_centroid
    boxes = low._info()["obs"][0].cpu().numpy()
    want = np.concatenate([np.abs(poly_centroid(b)) for b in boxes])
    np.testing.assert_allclose(obs[0].cpu().numpy(), want, rtol=1e-12)
    low.close()
