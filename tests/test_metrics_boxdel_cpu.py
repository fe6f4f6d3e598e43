"""TaskDrivenMetric (task_driven_metric.py:8-157) batched: the MST cost is pinned against networkx (the library the
reference calls), the scores against a direct restatement of the reference formulas on the emulated env's info."""
import numpy as np

from benchpush_b200.metrics import TaskDrivenMetric, _mst_cost
from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
from tests.boxdel_common import DELIVERY_CFG, DELIVERY_LAYOUT
from tests.emu import emu_bd


def test_mst_cost_is_pinned_to_networkx():
    import networkx as nx
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(2, 12))
        p = rng.uniform(-5, 5, (n, 2))
        w = np.full((n, n), np.inf)
        G = nx.Graph()
        G.add_nodes_from(range(n))
        for i in range(n):
            for j in range(i + 1, n):
                if rng.random() < 0.7:
                    w[i, j] = w[j, i] = np.linalg.norm(p[i] - p[j])
                    G.add_edge(i, j, weight=w[i, j])
        mst = nx.minimum_spanning_tree(G)
        ref = sum(mst.edges[e]["weight"] for e in mst.edges)
        assert abs(_mst_cost(w) - ref) < 1e-9


def test_task_driven_metric_on_a_delivery_episode():
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg=DELIVERY_CFG, layouts=[DELIVERY_LAYOUT])
    m = TaskDrivenMetric("test", robot_mass=env.cfg.agent.mass)
    obs, info = env.reset()
    m.reset(info)
    start = info["state"][0].copy()
    box0 = info["obs"][0, 0].copy()
    path = 0.0
    prev = start
    for k in range(3):
        obs, rew, term, trunc, info = env.step(np.array([-1.0], dtype=np.float32))
        path += np.linalg.norm(info["state"][0, :2] - prev[:2])
        prev = info["state"][0].copy()
        m.update(info, rew, eps_complete=(k == 2))
    assert m.success_rates == [0.25]
    # one completed box: MST = |start - c0| + |c0 - receptacle|
    c0 = box0.mean(axis=0)
    goal = np.array(env.goal_points[0])
    mst = np.linalg.norm(start[:2] - c0) + np.linalg.norm(c0 - goal)
    assert abs(m.efficiency_scores[0] - mst / path) < 1e-9
    area = 0.44 * 0.44
    effort = (1.0 * path + np.linalg.norm(c0 - goal) * area) / (1.0 * path + float(info["total_work"][0]))
    assert abs(m.effort_scores[0] - effort) < 1e-6
    assert abs(m.rewards[0] - float(info["cumulative_reward"][0])) < 1e-9
    r = m.reduce()
    assert r["episodes"] == 1 and r["success_rate"] == 0.25
    env.close()


def test_single_env_views_return_the_reference_types():
    from benchpush_b200.envs.gym_single import AreaClearingEnvB200, BoxDeliveryEnvB200
    from tests.areaclr_common import CLEARING_CFG, CLEARING_LAYOUT
    env = emu_bd.EmuBoxDeliveryEnvB200(cfg=DELIVERY_CFG, layouts=[DELIVERY_LAYOUT])
    obs, info = env.reset()
    assert obs.shape == (224, 224, 4) and obs.dtype == np.uint8 and isinstance(info["state"], tuple) and len(info["obs"]) == 4
    assert set(info) == {"state", "cumulative_distance", "cumulative_boxes", "cumulative_reward", "total_work", "obs",
                         "box_completed_statuses", "goal_positions", "ministeps", "inactivity"}
    for _ in range(2):
        obs, r, term, trunc, info = env.step(-1.0)
    assert isinstance(r, float) and isinstance(term, bool) and info["cumulative_boxes"] == 1 and len(info["obs"]) == 3
    assert info["goal_positions"][0].x == 4.25 and info["goal_positions"][0].y == 1.75
    env.close()
    ac = emu_bd.EmuAreaClearingEnvB200(cfg=CLEARING_CFG, layouts=[CLEARING_LAYOUT])
    obs, info = ac.reset()
    for _ in range(3):
        obs, r, term, trunc, info = ac.step(-1.0)
    assert info["box_count"] == 1 and info["box_completed_statuses"] == [True, False, False] and r > 9
    assert set(info) >= {"state", "total_work", "collision reward", "diff_reward", "box_completed_reward", "obs", "box_completed_statuses", "box_count", "ministeps"}
    # the batched TaskDrivenMetric also consumes the area-clearing info (40 goal points on the clearance boundary)
    m = TaskDrivenMetric("t", robot_mass=1)
    v = ac._v
    o, i0 = v.reset(); m.reset(i0)
    for k in range(3):
        o, rr, te, tr, i1 = v.step(np.array([-1.0], dtype=np.float32))
        m.update(i1, rr, eps_complete=(k == 2))
    assert m.success_rates == [1 / 3] and 0 < m.efficiency_scores[0] and 0 < m.effort_scores[0] <= 1.0 + 1e-9
    ac.close()


def test_low_dim_observation_and_batched_evaluate():
    from oracle.raster_oracle import poly_centroid
    from benchpush_b200.evaluate import evaluate
    from tests.areaclr_common import CLEARING_CFG, CLEARING_LAYOUT
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=1, cfg=DELIVERY_CFG, layouts=[DELIVERY_LAYOUT])
    obs, info = env.reset()
    for _ in range(2):
        obs, r, te, tr, info = env.step(np.array([-1.0], dtype=np.float32))
    ld = env.low_dim_observation()
    assert ld.shape == (1, 8) and np.all(ld[0, 6:] == 0)              # one box delivered: 3 left, zero padded
    ref = np.concatenate([np.abs(poly_centroid(info["obs"][0, k])) for k in range(3)])
    assert np.allclose(ld[0, :6], ref, atol=1e-12)
    env.close()
    # batched evaluate(): two area-clearing envs, each episode ends by truncation after t_max = 2 steps
    from benchpush_b200.envs.area_clearing import AreaClearingVecEnv
    cfg = dict(CLEARING_CFG, sim={"t_max": 2})
    ac = emu_bd.EmuAreaClearingVecEnv(num_envs=2, cfg=cfg, layouts=[CLEARING_LAYOUT, CLEARING_LAYOUT])
    m = TaskDrivenMetric("t", robot_mass=1)
    res = evaluate(ac, lambda o: np.array([-1.0, -1.0], dtype=np.float32), m, num_episodes=4)
    assert len(res["rewards"]) == 4 and len(res["efficiency_scores"]) == 4 and res["success_rates"] == [0.0] * 4
    assert res["rewards"][0] == res["rewards"][1] == res["rewards"][2]   # identical layouts and actions: identical episodes, also after the masked reset
    ac.close()
