"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Tolerances:
  * with the oracle switched to the portable cos / sin the kernels use (conftest.portable_trig):
    body poses and velocities are BIT-IDENTICAL after every env step, over whole episodes;
  * with the oracle in its default mode (libm cos / sin, what Chipmunk calls): 1e-9 absolute over
    the first ~10 env steps (the last-bit difference of cos / sin, amplified by the contact dynamics)
    — far inside the 1e-5 m tolerance SURVEY.md §8d states;
  * observation rasters and the rounded info['state']: bit-exact;
  * work / reward: 1e-9 absolute on work, 1e-7 on reward (numpy summation order in the oracle).
"""
import numpy as np
import pytest
import torch

from tests.conftest import FIXTURES, fixture_conc
from benchpush_b200.layouts import PackedLayouts, generate_experiment, pack_trials, unpack_trials
from oracle.ship_ice_oracle import ShipIceMetricOracle, ShipIceOracleEnv

pytestmark = pytest.mark.gpu


def _make_env(packed, conc, n=1, **kw):
    from benchpush_b200.envs import ShipIceVecEnv
    return ShipIceVecEnv(num_envs=n, cfg={"concentration": conc}, layouts=packed, device=0, **kw)


def _concat(ps):
    F = max(p.max_floes for p in ps)
    def pad(a, axis=1):
        w = [(0, 0)] * a.ndim; w[axis] = (0, F - a.shape[axis]); return np.pad(a, w)
    return PackedLayouts(np.concatenate([p.n_floes for p in ps]), np.concatenate([pad(p.n_verts) for p in ps]),
                         np.concatenate([pad(p.verts) for p in ps]), np.concatenate([pad(p.centres) for p in ps]),
                         np.concatenate([p.ship_state for p in ps]))


def test_step_parity_vs_libm_oracle(golden_layouts, pins):
    name = "c30_t7"
    p = golden_layouts[name]
    conc = fixture_conc(name)
    env = _make_env(p, conc)
    env.reset()
    ora = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=1)
    ora.reset()
    for a in pins["%s_o1_actions" % name]:
        obs, rew, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora.step(float(np.float32(a)))
        np.testing.assert_allclose(env.get_bodies(0), ora.space.state(), rtol=0, atol=1e-9)
        assert np.array_equal(obs[0].cpu().numpy(), oobs)
        np.testing.assert_allclose(float(rew[0]), orew, rtol=1e-9, atol=1e-7)
    # pinned final state of the (libm) oracle — guards against oracle drift on the GPU box
    np.testing.assert_allclose(env.get_bodies(0), pins["%s_final_state" % name], rtol=0, atol=1e-9)
    env.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_step_parity_on_reference_layouts(name, golden_layouts, pins, portable_trig):
    p = golden_layouts[name]
    conc = fixture_conc(name)
    env = _make_env(p, conc)
    obs, info = env.reset()
    ora = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=1)
    oobs, oinfo = ora.reset()
    assert np.array_equal(obs[0].cpu().numpy(), oobs)
    assert tuple(info["state"][0].tolist()) == tuple(float(x) for x in oinfo["state"])
    assert np.array_equal(env.get_bodies(0), ora.space.state())
    for a in pins["%s_o1_actions" % name]:
        obs, rew, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora.step(float(np.float32(a)))
        assert np.array_equal(env.get_bodies(0), ora.space.state())          # bit-identical transitions
        assert np.array_equal(obs[0].cpu().numpy(), oobs)
        assert tuple(info["state"][0].tolist()) == tuple(float(x) for x in oinfo["state"])
        np.testing.assert_allclose(float(info["total_work"][0]), oinfo["total_work"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(float(rew[0]), orew, rtol=1e-9, atol=1e-7)
        assert bool(term[0]) == oterm and not bool(trunc[0])
        assert bool(info["trial_success"][0]) == oinfo["trial_success"]
    # pinned final state of the (libm) oracle (golden fixture) — guards against oracle drift on the GPU box
    np.testing.assert_allclose(env.get_bodies(0), pins["%s_final_state" % name], rtol=0, atol=1e-9)
    assert np.array_equal(obs[0].cpu().numpy(), pins["%s_final_obs" % name])
    ob_list = info["obs"][0]
    oo = oinfo["obs"]
    assert len(ob_list) == len(oo)
    for g, o in zip(ob_list[:50], oo[:50]):
        np.testing.assert_allclose(g, o, rtol=0, atol=1e-9)
    env.close()


def test_batch_of_heterogeneous_layouts_and_episode_metrics(golden_layouts, portable_trig):
    """Many envs with different layouts in one launch; full episodes with termination; metric parity."""
    from benchpush_b200.metrics import ShipIceMetric
    packed = _concat([golden_layouts[n] for n in ["c10_t0", "c10_t7", "c30_t7"]])
    E = 6
    first = np.array([0, 1, 2, 2, 1, 0], dtype=np.int32)
    env = _make_env(packed, 0.1, n=E, first_trial=first)
    obs, info = env.reset()
    trials = unpack_trials(packed)
    oras, omets = [], []
    for e in range(E):
        o = ShipIceOracleEnv({"concentration": 0.1}, experiment=[trials[first[e]]], order_mode=1)
        oo, oi = o.reset()
        assert np.array_equal(obs[e].cpu().numpy(), oo)
        m = ShipIceMetricOracle("o", ship_mass=1, goal=(0, 9)); m.reset(oi)
        oras.append(o); omets.append(m)
    met = ShipIceMetric("g", ship_mass=env.cfg.ship.mass, goal=env.goal)
    met.reset(info)
    rng = np.random.default_rng(3)
    alive = np.ones(E, dtype=bool)
    for k in range(60):
        a = (0.3 * rng.uniform(-1, 1, E)).astype(np.float32)    # gentle steering: every episode reaches the goal line
        a[0] = 0.0                                        # env 0 sails straight to the goal
        obs, rew, term, trunc, info = env.step(a)
        done = met.update(info, rew, term | trunc)
        for e in range(E):
            if not alive[e]:
                continue
            oo, orr, ot, _, oi = oras[e].step(float(a[e]))
            omets[e].update(oi, orr, ot)
            assert bool(term[e]) == ot
            # whole episodes, same cos / sin on both sides: identical to the last bit
            assert np.array_equal(env.get_bodies(e), oras[e].space.state())
            assert np.array_equal(obs[e].cpu().numpy(), oo)
            if ot:
                alive[e] = False
        if not alive.any():
            break
        # finished envs keep stepping on the GPU (no auto reset here); ignore them afterwards
    assert not alive.any(), "every env should terminate within 60 steps (goal_y = 9)"
    env.close()
    # first completed episode of each env: scores agree with the reference formulas
    got_eff = np.array(met.efficiency_scores[:]); got_effort = np.array(met.effort_scores[:])
    exp_eff = np.array([m.efficiency_scores[0] for m in omets]); exp_effort = np.array([m.effort_scores[0] for m in omets])
    assert len(got_eff) >= E
    np.testing.assert_allclose(np.sort(got_eff[:E]), np.sort(exp_eff), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.sort(got_effort[:E]), np.sort(exp_effort), rtol=1e-9, atol=1e-9)


def test_auto_reset_moves_to_next_trial(golden_layouts):
    packed = _concat([golden_layouts[n] for n in ["c10_t0", "c10_t7"]])
    env = _make_env(packed, 0.1, n=2, auto_reset=True, first_trial=np.array([0, 1], dtype=np.int32))
    obs, info = env.reset()
    start = info["state"].clone()
    steps_to_done = None
    for k in range(50):
        obs, rew, term, trunc, info = env.step(np.zeros(2, dtype=np.float32))
        if bool(term.all()):
            steps_to_done = k + 1
            break
    assert steps_to_done is not None and steps_to_done == 34        # (9 - 1) / 0.24 -> 34 steps straight ahead
    obs, rew, term, trunc, info = env.step(np.zeros(2, dtype=np.float32))
    assert not bool(term.any()) and float(rew.abs().max()) == 0.0      # the reset step returns reward 0
    st = info["state"]
    # env 0 moved on to trial 1, env 1 wrapped to trial 0
    assert torch.equal(st[0], start[1]) and torch.equal(st[1], start[0])
    assert float(info["total_work"].max()) == 0.0
    env.close()


def test_full_size_properties_synthetic_high_concentration():
    """BASELINE config-5 sized layouts (concentration 0.5, ~640 floes) at a batch the oracle cannot
    follow: size-independent properties."""
    trials = generate_experiment(8, 0.5, seed=21)
    packed = pack_trials(trials)
    E = 64
    env = _make_env(packed, 0.5, n=E)
    obs, info = env.reset()
    b0 = [env.get_bodies(e) for e in (0, 9)]
    work_prev = info["total_work"].clone()
    y_prev = info["pose"][:, 1].clone()
    for k in range(4):
        obs, rew, term, trunc, info = env.step(np.zeros(E, dtype=np.float32))
        pose = info["pose"]
        # kinematic ship: exact closed-form advance (400 substeps of v*dt straight ahead)
        np.testing.assert_allclose((pose[:, 1] - y_prev).cpu().numpy(), 0.24, atol=1e-9)
        y_prev = pose[:, 1].clone()
        assert bool((info["total_work"] >= work_prev).all())          # work is a sum of distances
        work_prev = info["total_work"].clone()
        assert int(env._status.max()) & 7 == 0
    # envs that share a layout and saw the same actions are identical (determinism across warps)
    assert np.array_equal(env.get_bodies(1), env.get_bodies(9))
    assert torch.equal(obs[1], obs[9])
    # bodies far from the ship never moved at all (exact zero motion outside the awake set)
    b1 = env.get_bodies(0)
    far = np.abs(b0[0][:-1, 1] - b1[-1, 1]) > 8.0
    assert far.sum() > 100 and np.array_equal(b0[0][:-1][far], b1[:-1][far])
    # rasters: channel value sets
    o = obs.cpu().numpy()
    assert set(np.unique(o[:, 0])) <= {0, 127, 255} and set(np.unique(o[:, 2])) <= {0, 127, 255} and set(np.unique(o[:, 3])) <= {0, 255}
    c = env.counters(0)
    assert c["n_awake"] < 300 and c["status"] & 7 == 0
    env.close()


def test_no_cpu_fallback_errors():
    from benchpush_b200 import _lib
    from benchpush_b200.envs import ShipIceVecEnv
    with pytest.raises(NotImplementedError):                    # options that are not built fail loudly instead of being ignored
        ShipIceVecEnv(num_envs=1, cfg={"concentration": 0.1, "sim": {"gravity": (0, -1)}}, layouts=pack_trials(generate_experiment(1, 0.1, seed=1)), device=0)
    with pytest.raises(_lib.BenchPushError):
        trials = generate_experiment(1, 0.1, seed=1)
        ShipIceVecEnv(num_envs=1, cfg={"concentration": 0.1}, layouts=pack_trials(trials), device=99)


def test_batched_evaluate_loop(golden_layouts):
    """benchpush_b200.evaluate.evaluate on ship-ice: explicit masked resets, one score per finished episode."""
    import torch
    from benchpush_b200.envs import ShipIceVecEnv
    from benchpush_b200.evaluate import evaluate
    from benchpush_b200.metrics import ShipIceMetric
    env = ShipIceVecEnv(num_envs=3, cfg={"concentration": 0.1}, layouts=golden_layouts["c10_t0"], device=0)
    m = ShipIceMetric("straight", ship_mass=env.cfg.ship.mass, goal=env.goal)
    res = evaluate(env, lambda obs: torch.zeros(3, dtype=torch.float32, device=obs.device), m, num_episodes=3, max_steps=80)
    assert len(res["rewards"]) == 3 and len(res["efficiency_scores"]) == 3
    assert all(0.0 <= e <= 1.0 + 1e-9 for e in res["effort_scores"])
    assert res["rewards"][0] == res["rewards"][1] == res["rewards"][2]       # same layout, same actions
    env.close()


def test_resident_step_equals_batched_step(monkeypatch):
    """BP_STEP_IMPL=res (sim_res.cuh: the awake world of a batch staged in shared memory with cp.async.bulk, all 400 substeps
    on the pool) against the default batched kernel: bit-identical body states, rewards and observations over whole episodes
    with auto reset, at a batch that fills several pools and forces parking."""
    packed = pack_trials(generate_experiment(6, 0.5, seed=5))
    E = 96
    envs = {}
    for impl in ("batch", "res"):
        monkeypatch.setenv("BP_STEP_IMPL", impl)
        envs[impl] = _make_env(packed, 0.5, n=E, auto_reset=True)
        envs[impl].reset()
    monkeypatch.delenv("BP_STEP_IMPL")
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for k in range(45):
        a = torch.rand(E, generator=g, device="cuda") * 2 - 1
        ra = envs["batch"].step(a)
        rb = envs["res"].step(a)
        assert torch.equal(ra[0], rb[0]), k
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]), k
        for e in (0, 7, 8, 50, E - 1):
            assert np.array_equal(envs["batch"].get_bodies(e), envs["res"].get_bodies(e)), (k, e)
    for v in envs.values():
        v.close()


def test_evaluate_scores_every_episode_under_auto_reset(golden_layouts):
    """evaluate() with next-step auto reset must score the second and later episodes of an env like the explicit
    masked-reset path does (accumulators restarted from the reset state the auto-resetting step returns)."""
    from benchpush_b200.envs import ShipIceVecEnv
    from benchpush_b200.evaluate import evaluate
    from benchpush_b200.metrics import ShipIceMetric
    packed = _concat([golden_layouts["c10_t0"], golden_layouts["c10_t7"]])
    res = {}
    for auto in (False, True):
        env = ShipIceVecEnv(num_envs=2, cfg={"concentration": 0.1}, layouts=packed, device=0, auto_reset=auto, first_trial=np.array([0, 1], dtype=np.int32))
        m = ShipIceMetric("straight", ship_mass=env.cfg.ship.mass, goal=env.goal)
        res[auto] = evaluate(env, lambda obs: torch.full((2,), 0.3, dtype=torch.float32, device=obs.device), m, num_episodes=4, max_steps=200)
        env.close()
    for key in ("rewards", "efficiency_scores", "effort_scores"):
        assert len(res[True][key]) == 4
        np.testing.assert_allclose(sorted(res[True][key]), sorted(res[False][key]), rtol=1e-12, atol=1e-12)
    assert len(set(np.round(res[True]["rewards"], 9))) > 1          # the two trials differ, so a carried-over accumulator would show


def test_checkpoint_resume_is_bit_identical(golden_layouts):
    """bp_shipice_get_state / set_state: an env restored from its state block continues exactly as the original did."""
    name = "c30_t7"
    p = golden_layouts[name]
    env = _make_env(p, fixture_conc(name), n=2)
    env.reset()
    rng = np.random.default_rng(4)
    for _ in range(6):
        env.step(rng.uniform(-1, 1, 2).astype(np.float32))
    ck = env.get_state(0)
    acts = rng.uniform(-1, 1, (4, 2)).astype(np.float32)
    want = []
    for a in acts:
        obs, rew, *_ = env.step(a)
        want.append((env.get_bodies(0).copy(), float(rew[0]), obs[0].cpu().numpy().copy()))
    env.set_state(0, ck)
    for a, (bod, rew0, obs0) in zip(acts, want):
        obs, rew, *_ = env.step(a)
        assert np.array_equal(env.get_bodies(0), bod) and float(rew[0]) == rew0 and np.array_equal(obs[0].cpu().numpy(), obs0)
    with pytest.raises(ValueError):
        env.set_state(0, ck[:-8])
    env.close()


@pytest.mark.parametrize("name", ["c30_t7", "c50_t0"])
def test_set_mid_episode_oracle_state_then_step(name, golden_layouts, portable_trig):
    """bp_shipice_set_bodies: a body state taken from the oracle in the MIDDLE of an episode is injected into a freshly reset
    env (arbiter cache empty on both sides) and both advance: bit-identical transitions from an externally given state."""
    p = golden_layouts[name]
    conc = fixture_conc(name)
    ora = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=1)
    ora.reset()
    rng = np.random.default_rng(9)
    for _ in range(14):
        ora.step(float(np.float32(rng.uniform(-1, 1))), want_obs=False)
    st = ora.space.state()
    ora.space.clear_cache()
    ora.space.set_state(st)            # poses and velocities as exported; the bias velocities (not part of the state) are dropped
    ora.prev_obs = ora.get_obs_from_poly()
    env = _make_env(p, conc)
    env.reset()
    env.set_bodies(0, st)
    assert np.array_equal(env.get_bodies(0)[:, :3], st[:, :3])
    for _ in range(3):
        a = np.float32(rng.uniform(-1, 1))
        obs, rew, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        oobs, orew, oterm, _, oinfo = ora.step(float(a))
        assert np.array_equal(env.get_bodies(0), ora.space.state())
        assert np.array_equal(obs[0].cpu().numpy(), oobs)
        np.testing.assert_allclose(float(rew[0]), orew, rtol=1e-9, atol=1e-7)
    env.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_step_vs_canonical_order_oracle(name, golden_layouts, portable_trig):
    """Ordering error measured on the GPU path: the kernels solve the arbiters in greedy-colour order, the oracle's canonical
    order is ascending shape pair (Chipmunk's own order is unknowable offline).  At every env step the canonical-order oracle
    starts from the state the GPU is in (the order-1 oracle tracks the GPU bit for bit, asserted) and both take the step:
    per-step pose difference inside the envelope recorded in tests/test_oracle_cpu.py (concentration 0.1: exactly 0;
    0.3 / 0.5: mean <= 1.5e-3 m, max <= 1e-2 m — the max exceeds SURVEY §8d's 1e-3 m wish, see the finding there)."""
    from tests.test_oracle_cpu import ORDER_ENVELOPE
    p = golden_layouts[name]
    conc = fixture_conc(name)
    env = _make_env(p, conc)
    env.reset()
    A = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=1)
    A.reset()
    rng = np.random.default_rng(1)
    rows = []
    for k in range(20):
        a = np.float32(rng.uniform(-1, 1))
        B = A.clone(order_mode=0)
        env.step(np.array([a], dtype=np.float32))
        _, _, term, _, _ = A.step(float(a), want_obs=False)
        B.step(float(a), want_obs=False)
        g = env.get_bodies(0)
        assert np.array_equal(g, A.space.state())
        sb = B.space.state()
        rows.append((np.abs(g[:, :2] - sb[:, :2]).max(), np.abs(g[:, 2] - sb[:, 2]).max()))
        if term:
            break
    r = np.array(rows)
    if name.startswith("c10"):
        assert r.max() == 0.0
    else:
        assert r[:, 0].max() <= ORDER_ENVELOPE["pos_max"] and r[:, 0].mean() <= ORDER_ENVELOPE["pos_mean"]
        assert r[:, 1].max() <= ORDER_ENVELOPE["ang_max"] and r[:, 1].mean() <= ORDER_ENVELOPE["ang_mean"]
    env.close()
