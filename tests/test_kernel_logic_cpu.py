"""Kernel LOGIC checked on the CPU: the .cuh sources compiled as host C++ with a one-lane warp
(tests/emu, test-only) against the oracle.  The real CUDA build is checked by the -m gpu tests."""
import numpy as np
import pytest

from tests.conftest import FIXTURES, SHIP_VERTS, fixture_conc
from benchpush_b200.layouts import generate_experiment, pack_trials, unpack_trials
from oracle import cpo
from oracle.ship_ice_oracle import ShipIceOracleEnv, total_work_done
from tests.emu.emu import Emu


@pytest.mark.parametrize("name", ["c10_t7", "c30_t7", "c50_t0"])
def test_emulated_kernels_match_oracle_bitwise(name, golden_layouts, pins, portable_trig):
    p = golden_layouts[name]
    emu = Emu(1, p, SHIP_VERTS)
    env = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(p), order_mode=1)
    obs, info = env.reset()
    emu.reset(0, 0)
    assert np.array_equal(env.space.state(), emu.bodies(0, 0))
    assert np.array_equal(obs, emu.render_obs(0, True))
    for a in pins["%s_o1_actions" % name][:4]:
        obs, r, term, _, info = env.step(float(a))
        emu.step(0, float(a))
        # same solver order, same fp64 rounding sequence: identical to the last bit
        assert np.array_equal(env.space.state(), emu.bodies(0, 0))
        out = emu.out(0)
        assert out["status"] == 0
        # work / reward go through numpy (BLAS dot, pow) in the oracle: summation order differs -> 1e-9 tolerance
        np.testing.assert_allclose(out["reward"], r, rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(out["total_work"], info["total_work"], rtol=1e-9, atol=1e-9)
        assert (out["sx"], out["sy"], out["sa"]) == tuple(float(x) for x in info["state"])
        assert np.array_equal(obs, emu.render_obs(0, False))


def test_emulated_kernels_vs_libm_oracle(golden_layouts, pins):
    """Against the oracle in its default mode (libm cos/sin, what Chipmunk calls) the only difference
    is the last bit of cos/sin: 1e-9 over a few env steps."""
    name = "c30_t7"
    p = golden_layouts[name]
    emu = Emu(1, p, SHIP_VERTS)
    env = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(p), order_mode=1)
    env.reset(); emu.reset(0, 0)
    for a in pins["%s_o1_actions" % name][:4]:
        env.step(float(a)); emu.step(0, float(a))
        np.testing.assert_allclose(env.space.state(), emu.bodies(0, 0), rtol=0, atol=1e-9)


def test_portable_trig_accuracy():
    """The kernels' cos/sin (restated in the oracle) against libm: <= 1 ulp on the angles a ship or floe can reach."""
    import math
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-40, 40, 20000), rng.uniform(-1e-3, 1e-3, 1000), np.array([0.0, math.pi / 2, math.pi, -math.pi / 2, 1e-300, 0.78125, 0.3])])
    cpo.set_portable_trig(True)
    try:
        got = np.array([cpo.sincos(x) for x in xs])
    finally:
        cpo.set_portable_trig(False)
    ref = np.stack([np.sin(xs), np.cos(xs)], axis=1)
    ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
    assert ulp.max() <= 1.0, ulp.max()
    assert cpo.sincos(0.5) == (math.sin(0.5), math.cos(0.5))      # default mode is libm


@pytest.mark.parametrize("flags", ["", "-DBP_BWL=128 -DBP_BQ=64 -DBP_BQE=64", "-DBP_BWL=24 -DBP_BQ=64 -DBP_BQE=64 -DBP_RES_PA=24 -DBP_RES_PV=300 -DBP_RES_PS=20"])
def test_batched_path_equals_one_env_per_warp_path(flags, monkeypatch):
    """sim_batch.cuh (a warp advances 8 envs together) against sim_step.cuh (one env per warp): bit-identical, whatever the
    capacities of the per-warp queues (BP_BQ / BP_BQE: narrowphase queues flushed in rounds; BP_BWL: merged solver list — 24
    entries do not hold the arbiters of 8 envs, so those substeps are solved env by env)."""
    if flags:
        monkeypatch.setenv("BP_EMU_FLAGS", flags)
    nl = 3
    packed = pack_trials(generate_experiment(nl, 0.5, seed=11))
    E = 10
    kw = dict(pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
    A, B = Emu(E, packed, SHIP_VERTS, **kw), Emu(E, packed, SHIP_VERTS, **kw)
    for e in range(E):
        A.reset(e, e % nl); B.reset(e, e % nl)
    rng = np.random.default_rng(0)
    for k in range(8):
        acts = rng.uniform(-1, 1, E).astype(np.float32)
        for e in range(E):
            A.step(e, float(acts[e]))
        for e0 in range(0, E, B.batch_size()):
            B.step_batch(e0, acts)
        for e in range(E):
            assert np.array_equal(A.bodies(e, e % nl), B.bodies(e, e % nl))
            assert A.out(e) == B.out(e)
            # the schedule's cost counters (summed in shared memory by the batched path, flushed once per env step) agree with
            # the one-env path, which bumps them in the env block every substep
            sa, sb = A.scal(e), B.scal(e)
            assert (sa["substeps"], sa["live_sum"], sa["awake_sum"]) == (sb["substeps"], sb["live_sum"], sb["awake_sum"]), (k, e)


def test_awake_set_is_small_and_exact(golden_layouts):
    p = golden_layouts["c10_t0"]
    emu = Emu(1, p, SHIP_VERTS)
    emu.reset(0, 0)
    sc = emu.scal(0)
    assert sc["n_awake"] <= 3          # only the ship (and any deep-overlap pairs) stay awake in open water
    assert sc["status"] == 0


def test_emulated_reference_order_parents_are_bit_identical(spfa_golden, golden_layouts):
    """spfa_serial_parents (csrc/spfa.cuh, behind bp_spfa_batch_exact and the envs' exact_spfa_parents mode): dists and PARENTS
    equal the reference's own outputs (golden fixtures), the compiled reference run live (oracle/_ref, when built) and the
    oracle port on random maps incl. a blocked source."""
    import os, sys
    emu = Emu(1, golden_layouts["c10_t0"], SHIP_VERTS)
    for c in spfa_golden:
        d, p = emu.spfa_serial(c["map"], c["source"])
        assert np.array_equal(d, c["dists"]) and np.array_equal(p, c["parents"])
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    ref_spfa = None
    if os.path.isdir(ref_dir) and any(f.startswith("spfa") for f in os.listdir(ref_dir)):
        sys.path.insert(0, ref_dir)
        import spfa as ref_spfa
    rng = np.random.default_rng(3)
    for H, W in [(1, 1), (3, 4), (60, 45), (120, 150)]:
        m = rng.random((H, W)) > 0.25
        src = (int(rng.integers(0, H)), int(rng.integers(0, W)))
        for blocked in (False, True):
            m[src] = not blocked
            d, p = emu.spfa_serial(m, src)
            od, op = (ref_spfa.spfa(m, src) if ref_spfa is not None else cpo.spfa(m, src))
            assert np.array_equal(d, od) and np.array_equal(p, op), (H, W, blocked)


def test_emulated_spfa_matches_reference_golden(spfa_golden, golden_layouts):
    emu = Emu(1, golden_layouts["c10_t0"], SHIP_VERTS)
    for c in spfa_golden:
        d, p = emu.spfa(c["map"], c["source"])
        assert np.array_equal(d, c["dists"])                       # bit-identical distances
        assert np.array_equal(p >= 0, c["parents"] >= 0)           # same reachable set
        H, W = c["map"].shape
        idx = np.argwhere(p >= 0)
        pi, pj = np.divmod(p[p >= 0], W)
        assert np.all(np.maximum(np.abs(idx[:, 0] - pi), np.abs(idx[:, 1] - pj)) == 1)
        w = np.where((idx[:, 0] != pi) & (idx[:, 1] != pj), np.sqrt(np.float32(2.0)), np.float32(1.0)).astype(np.float32)
        assert np.array_equal((d[pi, pj] + w).astype(np.float32), d[p >= 0])    # a valid shortest-path tree


def test_product_hull_matches_oracle_hull_on_synthetic():
    trials = generate_experiment(1, 0.4, seed=3)
    p = pack_trials(trials)
    emu = Emu(1, p, SHIP_VERTS)
    env = ShipIceOracleEnv({"concentration": 0.4}, experiment=unpack_trials(p))
    env.reset(settle=False)
    for b, s in enumerate(env.poly_shapes):
        ev, mi = emu.layout_verts(0, b)
        assert np.array_equal(env.space.get_vertices(s), ev)
        m = env.space.shape_mass(s)
        np.testing.assert_allclose(1.0 / mi[0], m[0], rtol=1e-14)
        np.testing.assert_allclose(1.0 / mi[1], m[3], rtol=1e-12)


@pytest.mark.parametrize("flags", ["", "-DBP_RES_PA=24 -DBP_RES_PV=300 -DBP_RES_PS=20 -DBP_RES_DEBUG", "-DBP_RES_PA=12 -DBP_RES_PV=150 -DBP_RES_PS=8"])
def test_resident_path_equals_batched_path(flags, monkeypatch):
    """sim_res.cuh (awake world of a batch staged in a shared-memory pool) against sim_batch.cuh: bit-identical over whole
    episodes incl. auto reset.  The small pools force the rare paths: envs parked at a substep boundary, envs evicted in the
    middle of a substep (completed from the env block), envs that do not fit an empty pool (one-env path)."""
    if flags:
        monkeypatch.setenv("BP_EMU_FLAGS", flags)
    nl = 5
    packed = pack_trials(generate_experiment(nl, 0.5, seed=11))
    E = 10
    kw = dict(pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
    A, B = Emu(E, packed, SHIP_VERTS, **kw), Emu(E, packed, SHIP_VERTS, **kw)
    for e in range(E):
        A.reset(e, e % nl); B.reset(e, e % nl)
    rng = np.random.default_rng(0)
    big = int(np.argmax(packed.n_floes))
    for k in range(45):
        acts = rng.uniform(-1, 1, E).astype(np.float32)
        for e0 in range(0, E, A.batch_size()):
            A.step_batch(e0, acts, True)
            B.step_resident(e0, acts, True)
        for e in range(E):
            assert np.array_equal(A.bodies(e, big), B.bodies(e, big)), (k, e)
            assert A.out(e) == B.out(e), (k, e)


@pytest.mark.parametrize("mode", ["step", "batch", "res"])
def test_injected_mid_episode_state_matches_oracle(mode, golden_layouts, portable_trig):
    """bp_*_set_bodies semantics on the CPU build of the kernel sources: a body state exported by the oracle in the middle of
    an episode (arbiter cache dropped on both sides) continues bit-identically on all three step paths.  The first step runs
    with every body awake: a resting pair that overlaps deeper than the slop has no arbiter yet and must not be slept."""
    name = "c50_t0"
    p = golden_layouts[name]
    ora = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(p), order_mode=1)
    ora.reset()
    rng = np.random.default_rng(9)
    for _ in range(14):
        ora.step(float(np.float32(rng.uniform(-1, 1))), want_obs=False)
    st = ora.space.state()
    ora.space.clear_cache(); ora.space.set_state(st); ora.prev_obs = ora.get_obs_from_poly()
    emu = Emu(1, p, SHIP_VERTS, pmax=8 * (p.max_floes + 1), lmax=2 * (p.max_floes + 1))
    emu.reset(0, 0)
    emu.set_bodies(0, st)
    for k in range(2):
        a = np.float32(rng.uniform(-1, 1))
        if mode == "step":
            emu.step(0, float(a))
        elif mode == "batch":
            emu.step_batch(0, np.array([a], dtype=np.float32))
        else:
            emu.step_resident(0, np.array([a], dtype=np.float32))
        _, r, _, _, info = ora.step(float(a), want_obs=False)
        assert np.array_equal(emu.bodies(0, 0), ora.space.state()), k
        np.testing.assert_allclose(emu.out(0)["reward"], r, rtol=1e-9, atol=1e-7)
    assert emu.scal(0)["n_awake"] < 200          # the sleep pass is back after the first step


def test_fp32_build_of_the_kernel_sources_stays_within_tolerance(golden_layouts, monkeypatch):
    """-DBP_F32 (real = float): the same sources with fp32 body state / contacts / solver.  One env step from reset within
    1e-3 m / 1e-2 rad of the fp64 oracle, five steps within 2e-3 m; the three step paths stay bit-identical
    to each other in fp32 as well."""
    monkeypatch.setenv("BP_EMU_FLAGS", "-DBP_F32")
    name = "c30_t7"
    p = golden_layouts[name]
    ora = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(p), order_mode=1)
    ora.reset()
    kw = dict(pmax=8 * (p.max_floes + 1), lmax=2 * (p.max_floes + 1))
    A, B, C_ = Emu(1, p, SHIP_VERTS, **kw), Emu(1, p, SHIP_VERTS, **kw), Emu(1, p, SHIP_VERTS, **kw)
    for e in (A, B, C_):
        e.reset(0, 0)
    rng = np.random.default_rng(1)
    nfl = len(ora.space.state()) - 1
    prev = [A.world_verts(0, b) for b in range(nfl)]
    for k in range(5):
        a = np.float32(rng.uniform(-1, 1))
        A.step(0, float(a)); B.step_batch(0, np.array([a], dtype=np.float32)); C_.step_resident(0, np.array([a], dtype=np.float32))
        ora.step(float(a), want_obs=False)
        g, o = A.bodies(0, 0), ora.space.state()
        assert np.array_equal(g, B.bodies(0, 0)) and np.array_equal(g, C_.bodies(0, 0))
        lim = 1e-3 if k == 0 else 2e-3
        assert np.abs(g[:, :2] - o[:, :2]).max() <= lim and np.abs(g[:, 2] - o[:, 2]).max() <= 1e-2, k
        # the work term is accumulated in double from the fp32 vertices (shipice_env.cuh area_centroid_abs): it equals the
        # fp64 restatement of metrics.py:96-126 on the kernel's own before / after vertices
        cur = [A.world_verts(0, b) for b in range(nfl)]
        w = total_work_done(prev, cur)
        assert w > 0
        for e in (A, B, C_):
            np.testing.assert_allclose(-e.out(0)["collision_reward"], w, rtol=1e-7, atol=1e-12)
        prev = cur


def test_open_water_and_ragged_batches_match_the_oracle(portable_trig):
    """Edge cases of the layout input: a trial with NO floes (open water: the ship alone, no pairs, no arbiters) next to a dense
    field in the same packed batch (ragged floe counts, padded to the largest), stepped together in one warp batch; and a whole
    open-water episode up to ``terminated`` at the goal line."""
    dense = generate_experiment(1, 0.5, seed=11)[0]
    water = {"obstacles": [], "ship_state": dense["ship_state"]}
    packed = pack_trials([water, dense, water])
    assert packed.n_floes.tolist()[0] == 0 and packed.n_floes.tolist()[2] == 0 and packed.n_floes[1] > 100
    E = 3
    emu = Emu(E, packed, SHIP_VERTS, pmax=8 * (packed.max_floes + 1), lmax=2 * (packed.max_floes + 1))
    oras = [ShipIceOracleEnv({"concentration": 0.5}, experiment=unpack_trials(packed)[l:l + 1], order_mode=1) for l in range(E)]
    for e in range(E):
        emu.reset(e, e)
        obs, _ = oras[e].reset()
        assert np.array_equal(oras[e].space.state(), emu.bodies(e, e))
        assert np.array_equal(obs, emu.render_obs(e, True))
    assert emu.scal(0)["n_pairs"] == 0 and emu.scal(0)["n_order"] == 0
    rng = np.random.default_rng(5)
    for k in range(3):
        acts = rng.uniform(-1, 1, E).astype(np.float32)
        emu.step_batch(0, acts)
        for e in range(E):
            obs, r, term, _, info = oras[e].step(float(acts[e]))
            assert np.array_equal(oras[e].space.state(), emu.bodies(e, e)), (k, e)
            out = emu.out(e)
            assert out["status"] == 0 and bool(out["terminated"]) == bool(term)
            np.testing.assert_allclose(out["reward"], r, rtol=1e-9, atol=1e-7)
            assert np.array_equal(obs, emu.render_obs(e, False)), (k, e)
    # open water to the goal: straight ahead, terminates when the ship crosses goal_y, work stays exactly 0
    steps = 0
    while True:
        obs, r, term, _, info = oras[0].step(0.0)
        emu.step(0, 0.0)
        steps += 1
        out = emu.out(0)
        assert np.array_equal(oras[0].space.state(), emu.bodies(0, 0))
        assert bool(out["terminated"]) == bool(term) and out["total_work"] == 0.0 == info["total_work"]
        if term or steps > 400:
            break
    assert term and steps < 400


def test_emulated_spfa_on_degenerate_and_random_shapes(golden_layouts):
    """Edge cases of the map input: single cell, single row / column, nothing free, everything free, a walled-off source — and 40
    random maps of random shapes: distances of both kernels bit-identical to the port of main.cpp, parents of the
    reference-order kernel too, parents of the frontier kernel valid (a neighbour that attains the minimum)."""
    emu = Emu(1, golden_layouts["c10_t0"], SHIP_VERTS)
    rng = np.random.default_rng(123)
    cases = [(np.ones((1, 1), bool), (0, 0)), (np.ones((1, 17), bool), (0, 5)), (np.ones((23, 1), bool), (22, 0)),
             (np.zeros((4, 6), bool), (1, 1)), (np.ones((9, 13), bool), (8, 12))]
    walled = np.ones((7, 7), bool); walled[2, 2:5] = walled[4, 2:5] = walled[2:5, 2] = walled[2:5, 4] = False
    cases.append((walled, (3, 3)))
    for _ in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        m = rng.random((H, W)) > rng.uniform(0.1, 0.6)
        cases.append((m, (int(rng.integers(0, H)), int(rng.integers(0, W)))))
    s2 = np.float32(np.sqrt(np.float32(2.0)))
    for m, src in cases:
        rd, rp = cpo.spfa(m, src)
        d1, p1 = emu.spfa(m, src)
        d2, p2 = emu.spfa_serial(m, src)
        assert np.array_equal(rd, d1) and np.array_equal(rd, d2), (m.shape, src)
        assert np.array_equal(rp, p2), (m.shape, src)
        H, W = m.shape
        for i, j in zip(*np.nonzero(d1 > 0)):                      # frontier kernel: any minimising neighbour is a valid parent
            pi, pj = divmod(int(p1[i, j]), W)
            assert m[pi, pj] and max(abs(pi - i), abs(pj - j)) == 1
            w = s2 if (pi != i and pj != j) else np.float32(1.0)
            assert np.float32(d1[pi, pj] + w) == d1[i, j]
    assert cpo.spfa(walled, (3, 3))[0].sum() == 0                   # the walled-off source reaches nothing: all zeros (D18)
