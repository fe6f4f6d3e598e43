"""BASELINE config 1 / north-star acceptance: ship-ice-v0, default config, 200 seeded episodes with random actions.
Golden scores come from the CPU oracle (tests/golden/shipice_200ep.npz, made by tools/make_golden_200ep.py):
``libm_*`` with the oracle's default cos / sin (what Chipmunk calls), ``portable_*`` with the routine the kernels use.

* CPU: the oracle still reproduces its own golden episodes (no drift), and the two trig modes give statistically
  indistinguishable score distributions.
* GPU: the 200 episodes run as one batch of 200 envs through the C ABI.  Against ``portable_*``: episode lengths and success
  flags exact, efficiency / effort / reward within 1e-9 relative (summation order of the work term).  Against ``libm_*``
  (the reference's arithmetic): success rate within the binomial 95 % interval, two-sample KS p > 0.05 on efficiency and
  effort."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import GOLD, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "shipice_200ep.npz"))


def _run_batch(env, met, actions, max_steps):
    """All episodes as one batch: step the envs still alive with their recorded actions, collect per-env scores."""
    import torch
    E = env.num_envs
    obs, info = env.reset()
    met.reset(info)
    acts = torch.as_tensor(actions, device=env.device)
    alive = torch.ones(E, dtype=torch.bool, device=env.device)
    steps = np.zeros(E, dtype=np.int32)
    succ = np.zeros(E, dtype=bool); eff = np.zeros(E); effort = np.zeros(E); rew = np.zeros(E)
    for t in range(max_steps):
        obs, r, term, trunc, info = env.step(acts[:, t], mask=alive)
        done = (term | (t == max_steps - 1)) & alive
        met.update(info, torch.where(alive, r, torch.zeros_like(r)), done)
        steps += alive.cpu().numpy()
        if bool(done.any()):
            idx = torch.nonzero(done).flatten().cpu().numpy()
            succ[idx] = info["trial_success"][idx].cpu().numpy()
            eff[idx] = met.compute_efficiency_score()[idx].cpu().numpy()
            effort[idx] = met.compute_effort_score()[idx].cpu().numpy()
            rew[idx] = met.eps_reward[idx].cpu().numpy()
        alive = alive & ~done
        if not bool(alive.any()):
            break
    assert len(met.rewards) == E
    return steps, succ, eff, effort, rew


def _stat_agree(a_succ, b_succ, a_eff, b_eff, a_effort, b_effort):
    from scipy.stats import ks_2samp
    n = len(a_succ)
    p = float(np.mean(b_succ))
    half = 1.96 * np.sqrt(p * (1 - p) / n)
    assert abs(float(np.mean(a_succ)) - p) <= half + 1e-12, (np.mean(a_succ), p, half)
    assert ks_2samp(a_eff, b_eff).pvalue > 0.05
    assert ks_2samp(a_effort, b_effort).pvalue > 0.05


def test_oracle_reproduces_golden_episodes_and_trig_modes_agree(gold):
    import make_golden_200ep as g
    assert np.array_equal(gold["actions"], g.actions_table())
    for k, portable in ((3, 0), (117, 1)):
        tag = "portable" if portable else "libm"
        _, n, succ, eff, effort, rew, work = g.run_episode((k, portable))
        assert n == int(gold["%s_steps" % tag][k]) and succ == bool(gold["%s_success" % tag][k])
        np.testing.assert_allclose([eff, effort, rew, work], [gold["%s_efficiency" % tag][k], gold["%s_effort" % tag][k],
                                                               gold["%s_reward" % tag][k], gold["%s_total_work" % tag][k]], rtol=1e-12, atol=1e-12)
    from oracle import cpo
    cpo.set_portable_trig(False)
    assert 0.5 < gold["libm_success"].mean() < 1.0 and gold["libm_steps"].max() <= 300
    _stat_agree(gold["portable_success"], gold["libm_success"], gold["portable_efficiency"], gold["libm_efficiency"],
                gold["portable_effort"], gold["libm_effort"])


@pytest.mark.gpu
def test_200_seeded_episodes_on_the_gpu_match_the_oracle_distributions(gold):
    import torch
    import make_golden_200ep as g
    from benchpush_b200.envs import ShipIceVecEnv
    from benchpush_b200.layouts import generate_experiment, pack_trials
    from benchpush_b200.metrics import ShipIceMetric
    E = g.N_EP
    packed = pack_trials(generate_experiment(100, g.CONC, seed=g.SEED_LAYOUT))
    env = ShipIceVecEnv(num_envs=E, cfg={"concentration": g.CONC}, layouts=packed, device=0, first_trial=np.arange(E) % 100)
    met = ShipIceMetric("gpu", ship_mass=env.cfg.ship.mass, goal=env.goal)
    steps, succ, eff, effort, rew = _run_batch(env, met, gold["actions"], g.MAX_STEPS)
    env.close()
    # the kernels' own arithmetic: episode by episode
    assert np.array_equal(steps, gold["portable_steps"])
    assert np.array_equal(succ, gold["portable_success"].astype(bool))
    np.testing.assert_allclose(eff, gold["portable_efficiency"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(effort, gold["portable_effort"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rew, gold["portable_reward"], rtol=1e-9, atol=1e-7)
    # the reference's arithmetic (libm cos / sin): statistically indistinguishable
    _stat_agree(succ, gold["libm_success"], eff, gold["libm_efficiency"], effort, gold["libm_effort"])


@pytest.fixture(scope="module")
def gold_ref():
    return np.load(os.path.join(GOLD, "shipice_200ep_refdata.npz"))


def test_oracle_reproduces_golden_episodes_on_the_reference_ice_fields(gold_ref):
    """BASELINE config 1 on the reference's own data: the 100 trials of experiments_10_100_r06_d40x12.pk (packed in
    tests/golden/ice_c10_all100.npz), episode k on trial k % 100 (ship_ice_env.py:188).  The canonical-order run (order_mode 0)
    and the greedy-colour-order run are statistically indistinguishable: the solver order does not move the episode scores."""
    import make_golden_200ep as g
    for k, portable, order, tag in ((5, 0, 1, "libm"), (142, 1, 1, "portable"), (77, 0, 0, "canonical")):
        _, n, succ, eff, effort, rew, work = g.run_episode((k, portable, True, order))
        assert n == int(gold_ref["%s_steps" % tag][k]) and succ == bool(gold_ref["%s_success" % tag][k])
        np.testing.assert_allclose([eff, effort, rew, work], [gold_ref["%s_efficiency" % tag][k], gold_ref["%s_effort" % tag][k],
                                                               gold_ref["%s_reward" % tag][k], gold_ref["%s_total_work" % tag][k]], rtol=1e-12, atol=1e-12)
    from oracle import cpo
    cpo.set_portable_trig(False)
    for tag in ("portable", "libm"):
        _stat_agree(gold_ref["%s_success" % tag], gold_ref["canonical_success"], gold_ref["%s_efficiency" % tag], gold_ref["canonical_efficiency"],
                    gold_ref["%s_effort" % tag], gold_ref["canonical_effort"])


@pytest.mark.gpu
def test_200_seeded_episodes_on_the_reference_ice_fields_gpu(gold_ref):
    """The same 200 episodes as one GPU batch: episode by episode against the oracle in the kernels' arithmetic, statistically
    against the reference's arithmetic (libm) and against the canonical solver order."""
    import make_golden_200ep as g
    from benchpush_b200.envs import ShipIceVecEnv
    from benchpush_b200.layouts import load_packed
    from benchpush_b200.metrics import ShipIceMetric
    E = g.N_EP
    packed = load_packed(g.REF_PACK)
    env = ShipIceVecEnv(num_envs=E, cfg={"concentration": g.CONC}, layouts=packed, device=0, first_trial=np.arange(E) % 100)
    met = ShipIceMetric("gpu", ship_mass=env.cfg.ship.mass, goal=env.goal)
    steps, succ, eff, effort, rew = _run_batch(env, met, gold_ref["actions"], g.MAX_STEPS)
    env.close()
    assert np.array_equal(steps, gold_ref["portable_steps"])
    assert np.array_equal(succ, gold_ref["portable_success"].astype(bool))
    np.testing.assert_allclose(eff, gold_ref["portable_efficiency"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(effort, gold_ref["portable_effort"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rew, gold_ref["portable_reward"], rtol=1e-9, atol=1e-7)
    for tag in ("libm", "canonical"):
        _stat_agree(succ, gold_ref["%s_success" % tag], eff, gold_ref["%s_efficiency" % tag], effort, gold_ref["%s_effort" % tag])


@pytest.fixture(scope="module")
def gold_maze():
    return np.load(os.path.join(GOLD, "maze_200ep.npz"))


def test_maze_oracle_reproduces_golden_episodes(gold_maze):
    import make_golden_200ep_maze as g
    assert np.array_equal(gold_maze["actions"], g.actions_table())
    for k, portable in ((11, 0), (150, 1)):
        tag = "portable" if portable else "libm"
        _, n, succ, eff, effort, rew, work = g.run_episode((k, portable))
        assert n == int(gold_maze["%s_steps" % tag][k]) and succ == bool(gold_maze["%s_success" % tag][k])
        np.testing.assert_allclose([eff, effort, rew, work], [gold_maze["%s_efficiency" % tag][k], gold_maze["%s_effort" % tag][k],
                                                               gold_maze["%s_reward" % tag][k], gold_maze["%s_total_work" % tag][k]], rtol=1e-12, atol=1e-12)
    from oracle import cpo
    cpo.set_portable_trig(False)
    # random steering never threads the maze: every episode ends on a wall (success 0, efficiency 0), lengths 46..132
    assert gold_maze["libm_success"].sum() == 0 and gold_maze["libm_steps"].max() < g.MAX_STEPS
    _stat_agree(gold_maze["portable_success"], gold_maze["libm_success"], gold_maze["portable_efficiency"], gold_maze["libm_efficiency"],
                gold_maze["portable_effort"], gold_maze["libm_effort"])


@pytest.mark.gpu
def test_200_seeded_maze_episodes_on_the_gpu_match_the_oracle(gold_maze):
    """maze-NAMO-v0, 200 seeded layouts with 5 movable obstacles, random actions: same criteria as ship-ice."""
    import make_golden_200ep_maze as g
    from benchpush_b200.envs import MazeNamoVecEnv
    from benchpush_b200.metrics import MazeNamoMetric
    E = g.N_EP
    env = MazeNamoVecEnv(num_envs=E, layouts=g.layouts(), device=0, first_layout=np.arange(E))
    met = MazeNamoMetric("gpu", robot_mass=1)
    steps, succ, eff, effort, rew = _run_batch(env, met, gold_maze["actions"], g.MAX_STEPS)
    env.close()
    assert np.array_equal(steps, gold_maze["portable_steps"])
    assert np.array_equal(succ, gold_maze["portable_success"].astype(bool))
    np.testing.assert_allclose(eff, gold_maze["portable_efficiency"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(effort, gold_maze["portable_effort"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rew, gold_maze["portable_reward"], rtol=1e-9, atol=1e-7)
    _stat_agree(succ, gold_maze["libm_success"], eff, gold_maze["libm_efficiency"], effort, gold_maze["libm_effort"])


@pytest.mark.gpu
def test_200_seeded_area_clearing_episodes_match_the_oracle_distributions():
    """area-clearing-v0 (BASELINE config 3): 200 seeded layouts, random heading actions, TaskDrivenMetric scores.  The room
    envs are not bit-comparable over whole episodes (the controller's atan2 / sin / cos are libm's on the device and numpy's
    in the oracle), so the criterion is the statistical one: two-sample KS p > 0.05 on episode length, success
    fraction, efficiency and effort against the oracle's golden run (tools/make_golden_200ep_areaclr.py)."""
    from scipy.stats import ks_2samp
    import make_golden_200ep_areaclr as g
    from benchpush_b200.envs import AreaClearingVecEnv
    from benchpush_b200.metrics import TaskDrivenMetric
    gold = np.load(os.path.join(GOLD, "areaclr_200ep.npz"))
    E = g.N_EP
    assert np.array_equal(gold["actions"], g.actions_table())
    env = AreaClearingVecEnv(num_envs=E, layouts=[g.layout(k) for k in range(E)], device=0)
    obs, info = env.reset()
    met = TaskDrivenMetric("gpu", robot_mass=1)
    met.reset(info)
    acts = gold["actions"]
    alive = np.ones(E, dtype=bool)
    steps = np.zeros(E, dtype=np.int32)
    succ = np.zeros(E); eff = np.zeros(E); effort = np.zeros(E)
    for t in range(g.MAX_STEPS):
        obs, r, term, trunc, info = env.step(acts[:, t], mask=alive.astype(np.uint8))
        done = (term.cpu().numpy() | trunc.cpu().numpy() | (t == g.MAX_STEPS - 1)) & alive
        rr = r.cpu().numpy().copy(); rr[~alive] = 0.0
        n0 = len(met.rewards)
        met.update(info, rr, done)
        steps += alive
        idx = np.nonzero(done)[0]
        succ[idx] = met.success_rates[n0:]; eff[idx] = met.efficiency_scores[n0:]; effort[idx] = met.effort_scores[n0:]
        alive &= ~done
        if not alive.any():
            break
    env.close()
    assert len(met.rewards) == E
    same = float(np.mean(steps == gold["steps"]))
    print("area-clearing: %.0f %% of the episodes have the oracle's length; success %.3f vs %.3f, effort %.4f vs %.4f"
          % (100 * same, succ.mean(), gold["success"].mean(), effort.mean(), gold["effort"].mean()))
    assert ks_2samp(steps, gold["steps"]).pvalue > 0.05
    assert ks_2samp(succ, gold["success"]).pvalue > 0.05
    assert ks_2samp(eff, gold["efficiency"]).pvalue > 0.05
    assert ks_2samp(effort, gold["effort"]).pvalue > 0.05
    assert abs(effort.mean() - gold["effort"].mean()) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("oc,exact", [("small_empty", False), ("small_columns", False), ("small_columns", True)])
def test_200_seeded_box_delivery_episodes_match_the_oracle_distributions(oc, exact):
    """box-delivery-v0 (BASELINE config 2's env): 200 seeded layouts (the reference's sampler, RandomState(42 + k)), random heading
    actions for up to 100 steps, TaskDrivenMetric scores; statistical criterion as for area-clearing.  small_columns: rooms with
    1 - 2 random columns per layout, with the default deterministic spfa parents (paths may take another equal-cost staircase than
    the reference's: the score distributions must not notice) and with the reference-order parents."""
    from scipy.stats import ks_2samp
    import make_golden_200ep_boxdel as g
    from benchpush_b200.envs import BoxDeliveryVecEnv
    from benchpush_b200.metrics import TaskDrivenMetric
    path = os.path.join(GOLD, "boxdel_200ep.npz" if oc == "small_empty" else "boxdel_200ep_%s.npz" % oc)
    if not os.path.exists(path):
        pytest.skip("%s not generated (BOXDEL_OC=%s python tools/make_golden_200ep_boxdel.py)" % (path, oc))
    gold = np.load(path)
    E = g.N_EP
    assert np.array_equal(gold["actions"], g.actions_table())
    g.OC = oc
    g.CFG = {"agent": {"action_type": "heading"}} if oc == "small_empty" else {"agent": {"action_type": "heading"}, "env": {"obstacle_config": oc}}
    env = BoxDeliveryVecEnv(num_envs=E, cfg=g.CFG, layouts=[g.layout(k) for k in range(E)], device=0, exact_spfa_parents=exact)
    obs, info = env.reset()
    met = TaskDrivenMetric("gpu", robot_mass=1)
    met.reset(info)
    acts = gold["actions"]
    alive = np.ones(E, dtype=bool)
    steps = np.zeros(E, dtype=np.int32)
    succ = np.zeros(E); eff = np.zeros(E); effort = np.zeros(E)
    for t in range(g.MAX_STEPS):
        obs, r, term, trunc, info = env.step(acts[:, t], mask=alive.astype(np.uint8))
        done = (term.cpu().numpy() | trunc.cpu().numpy() | (t == g.MAX_STEPS - 1)) & alive
        rr = r.cpu().numpy().copy(); rr[~alive] = 0.0
        n0 = len(met.rewards)
        met.update(info, rr, done)
        steps += alive
        idx = np.nonzero(done)[0]
        succ[idx] = met.success_rates[n0:]; eff[idx] = met.efficiency_scores[n0:]; effort[idx] = met.effort_scores[n0:]
        alive &= ~done
        if not alive.any():
            break
    env.close()
    assert len(met.rewards) == E
    same = float(np.mean(steps == gold["steps"]))
    print("box-delivery: %.0f %% of the episodes have the oracle's length; success %.3f vs %.3f, efficiency %.4f vs %.4f, effort %.4f vs %.4f"
          % (100 * same, succ.mean(), gold["success"].mean(), eff.mean(), gold["efficiency"].mean(), effort.mean(), gold["effort"].mean()))
    assert ks_2samp(steps, gold["steps"]).pvalue > 0.05
    assert ks_2samp(succ, gold["success"]).pvalue > 0.05
    assert ks_2samp(eff, gold["efficiency"]).pvalue > 0.05
    assert ks_2samp(effort, gold["effort"]).pvalue > 0.05
    assert abs(effort.mean() - gold["effort"].mean()) < 0.02
