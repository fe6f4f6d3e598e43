"""CPU tests of the oracle itself: pins against the reference where the reference pins anything
(spfa), drift pins + physical invariants where it does not (physics: PARITY UNPINNED)."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import FIXTURES, ROOT, fixture_conc
from oracle import cpo
from oracle import raster_oracle as ro
from oracle.ship_ice_oracle import ShipIceMetricOracle, ShipIceOracleEnv
from benchpush_b200.layouts import unpack_trials


# ------------------------------------------------------------------ spfa: pinned by the reference's own code
def test_spfa_port_matches_reference_golden(spfa_golden):
    for c in spfa_golden:
        d, p = cpo.spfa(c["map"], c["source"])
        assert np.array_equal(d, c["dists"])
        assert np.array_equal(p, c["parents"])


def test_spfa_port_matches_compiled_reference_live():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not any(f.startswith("spfa") for f in (os.listdir(ref_dir) if os.path.isdir(ref_dir) else [])):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    sys.path.insert(0, ref_dir)
    import spfa as ref_spfa
    rng = np.random.default_rng(7)
    for H, W in [(3, 3), (40, 55), (90, 70)]:
        m = rng.random((H, W)) > 0.3
        src = (int(rng.integers(0, H)), int(rng.integers(0, W)))
        rd, rp = ref_spfa.spfa(m, src)
        d, p = cpo.spfa(m, src)
        assert np.array_equal(rd, d) and np.array_equal(rp, p)


# ------------------------------------------------------------------ rasters
def test_polygon_vectorised_equals_literal():
    rng = np.random.default_rng(0)
    for _ in range(40):
        n = int(rng.integers(3, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(2, 9)
        r = 12 + rad * np.sin(ang) + rng.uniform(-3, 3)
        c = 12 + rad * np.cos(ang) + rng.uniform(-3, 3)
        if rng.random() < 0.3:                      # integer vertices: pixels exactly on vertices / edges
            r, c = np.round(r), np.round(c)
        shape = (24, 24) if rng.random() < 0.7 else None
        rr0, cc0 = ro.draw_polygon_literal(r, c, shape)
        rr1, cc1 = ro.draw_polygon(r, c, shape)
        assert np.array_equal(rr0, rr1) and np.array_equal(cc0, cc1)


def test_polygon_inclusive_edges():
    # axis-aligned square with integer corners: boundary pixels count as inside (vertex / edge rule)
    rr, cc = ro.draw_polygon([2, 2, 5, 5], [2, 5, 5, 2])
    assert len(rr) == 16 and rr.min() == 2 and rr.max() == 5 and cc.min() == 2 and cc.max() == 5


def test_crop_vectorised_equals_literal():
    og = ro.OccupancyGrid(0.04, 0.04, 12, 40, 6, 6, 25)
    rng = np.random.default_rng(1)
    g = rng.random((og.occ_map_height, og.occ_map_width))
    for pose in [(5.0, 1.0, 1.5), (0.3, 0.2, 0.1), (11.9, 39.5, 2.0), (-0.4, 3.0, 1.0), (12.3, 20.0, 1.0)]:
        a = og._crop(g, pose, 2, 1.0)
        b = og._crop_literal(g, pose, 2, 1.0)
        assert np.array_equal(a, b)


def test_goal_dist_transform_matches_loop():
    og = ro.OccupancyGrid(0.04, 0.04, 12, 40, 6, 6, 25)
    g = og.global_goal_dist_transform(9)
    g2m = og.map_height / og.occ_map_height
    for i in [0, 1, 224, 225, 226, 999]:
        d = 9 - i * g2m
        d = 0 if d < 0 else d
        assert g[i, 0] == d / 9 and g[i, 299] == d / 9


# ------------------------------------------------------------------ physics oracle: drift pins + invariants
@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_pins(name, golden_layouts, pins):
    if name == "c50_t0":
        order_modes = (1,)
    else:
        order_modes = (0, 1)
    for order in order_modes:
        env = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(golden_layouts[name]), order_mode=order)
        obs, info = env.reset()
        assert np.array_equal(obs.astype(np.int64).reshape(4, -1).sum(1), pins["%s_o%d_obs_sums" % (name, order)][0])
        for k, a in enumerate(pins["%s_o%d_actions" % (name, order)]):
            obs, r, term, _, info = env.step(float(a))
            st = env.space.state()
            row = pins["%s_o%d_rows" % (name, order)][k]
            got = np.array([st[-1, 0], st[-1, 1], st[-1, 2], r, info["total_work"], st[:-1, 0].sum(), st[:-1, 1].sum(), st[:-1, 2].sum(), float(term)])
            np.testing.assert_allclose(got, row, rtol=1e-9, atol=1e-9)
            assert np.array_equal(obs.astype(np.int64).reshape(4, -1).sum(1), pins["%s_o%d_obs_sums" % (name, order)][k + 1])


def _square(s):
    return np.array([[-s, -s], [s, -s], [s, s], [-s, s]], dtype=np.float64)


def test_hull_and_mass_properties():
    sp = cpo.Space()
    b = sp.add_body(cpo.DYNAMIC, 0, 0, 0)
    # duplicate + collinear vertices are dropped by the hull; order is CCW from the min-x/min-y vertex
    v = np.array([[1, 1], [1, 1], [-1, 1], [-1, 0], [-1, -1], [1, -1], [0, -1]], dtype=np.float64)
    s = sp.add_poly(b, v, 0.0)
    hv = sp.get_vertices(s)
    assert hv.shape == (4, 2)
    assert np.array_equal(hv, np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64))
    sp.set_density(s, 2.0)
    m, i_unit, area, body_i = sp.shape_mass(s)
    assert area == 4.0 and m == 8.0
    np.testing.assert_allclose(i_unit, (2 * 2 + 2 * 2) / 12.0)      # unit moment of a 2x2 box
    np.testing.assert_allclose(body_i, 8.0 * 8.0 / 12.0)


def test_momentum_conserved_without_damping():
    sp = cpo.Space(iterations=10, damping=1.0)
    ids = []
    for x, vx in [(-0.6, 1.0), (0.6, -0.5)]:
        b = sp.add_body(cpo.DYNAMIC, x, 0.05 * x, 0.0)
        s = sp.add_poly(b, _square(0.5), 0.02)
        sp.set_density(s, 1.0)
        sp.set_material(s, 0.5, 0.3)
        sp.set_velocity(b, vx, 0.0)
        ids.append(b)
    m = [sp.body(b)[6] for b in ids]
    p0 = sum(m[k] * sp.body(ids[k])[3] for k in range(2))
    for _ in range(200):
        sp.step(0.002)
    p1 = sum(m[k] * sp.body(ids[k])[3] for k in range(2))
    assert abs(p1 - p0) < 1e-12 * max(1.0, abs(p0))
    assert sp.stats()["arb_solved"] > 0                       # they did collide


def test_kinematic_pusher_and_quasi_static_world():
    # damping = 0: a box moves only while the kinematic pusher is in contact; it never penetrates
    # beyond the slop; a far-away box never moves at all.
    sp = cpo.Space(iterations=10, damping=0.0)
    k = sp.add_body(cpo.KINEMATIC, 0.0, 0.0, 0.0)
    ks = sp.add_poly(k, _square(0.5), 0.02); sp.set_material(ks, 0.01, 1.0)
    near = sp.add_body(cpo.DYNAMIC, 1.2, 0.0, 0.0)
    ns = sp.add_poly(near, _square(0.5), 0.02); sp.set_density(ns, 0.001); sp.set_material(ns, 0.01, 1.0)
    far = sp.add_body(cpo.DYNAMIC, 10.0, 0.0, 0.0)
    fs = sp.add_poly(far, _square(0.5), 0.02); sp.set_density(fs, 0.001); sp.set_material(fs, 0.01, 1.0)
    sp.set_velocity(k, 0.3, 0.0)
    x_contact = None
    for step in range(1500):
        sp.step(0.002)
        kx, nx = sp.body(k)[0], sp.body(near)[0]
        np.testing.assert_allclose(kx, 0.3 * 0.002 * (step + 1), rtol=1e-12)     # kinematic exactness
        assert nx - kx > 1.0 + 0.04 - 0.1 - 1e-9                                  # non-penetration beyond slop
        if x_contact is None and nx != 1.2:
            x_contact = kx
    assert x_contact is not None and abs(x_contact - (1.2 - 1.04)) < 2e-3          # contact starts at the rounded surfaces
    assert sp.body(near)[0] > 1.2 + 0.3 * 3.0 - 0.25                               # pushed along
    assert sp.body(far)[0] == 10.0 and sp.body(far)[3] == 0.0                      # untouched body: exactly still
    sp.set_velocity(k, 0.0, 0.0)
    for _ in range(3):
        sp.step(0.002)
    assert abs(sp.body(near)[3]) < 1e-12                                            # velocity dies with the contact (damping 0)


def test_metric_oracle_formulas():
    m = ShipIceMetricOracle("x", ship_mass=1, goal=(0, 9))
    m.reset({"state": (5.0, 1.0, 1.57)})
    m.update({"state": (5.0, 4.0, 1.57), "total_work": 1.0, "trial_success": False}, -1.0, False)
    m.update({"state": (9.0, 9.0, 1.57), "total_work": 2.5, "trial_success": True}, 200.0, True)
    L, l0 = 8.0, 3.0 + np.hypot(4.0, 5.0)
    assert m.rewards == [199.0]
    np.testing.assert_allclose(m.efficiency_scores, [L / l0])
    np.testing.assert_allclose(m.effort_scores, [l0 / (l0 + 2.5)])


# ---------------------------------------------------------------- ordering error (finding)
ORDER_ENVELOPE = {"pos_max": 1e-2, "pos_mean": 1.5e-3, "ang_max": 2.5e-2, "ang_mean": 5e-3}


def ordering_error_single_steps(name, n_steps, golden_layouts, seed=1):
    """One env step (400 substeps) from IDENTICAL mid-episode states (bodies + arbiter cache, Space.clone) with the kernels'
    greedy-colour solver order (order_mode 1) and with the canonical order (ascending shape pair, order_mode 0): max over
    bodies of the pose difference, per step."""
    from tests.conftest import fixture_conc
    from benchpush_b200.layouts import unpack_trials
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    A = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(golden_layouts[name]), order_mode=1)
    A.reset()
    rng = np.random.default_rng(seed)
    rows = []
    for k in range(n_steps):
        a = float(rng.uniform(-1, 1))
        B = A.clone(order_mode=0)
        _, _, term, _, _ = A.step(a, want_obs=False)
        B.step(a, want_obs=False)
        sa, sb = A.space.state(), B.space.state()
        rows.append((np.abs(sa[:, :2] - sb[:, :2]).max(), np.abs(sa[:, 2] - sb[:, 2]).max()))
        if term:
            break
    return np.array(rows)


@pytest.mark.parametrize("name", ["c10_t7", "c30_t7", "c50_t0"])
def test_solver_ordering_error_of_one_env_step(name, golden_layouts):
    """FINDING (not a pass/fail of the 1e-3 m wish): Chipmunk's arbiter order comes from its BBTree and is unknowable offline, so
    the kernels' order (greedy colours over ascending shape pairs) and the oracle's canonical order are two members of the same
    family.  Measured over whole episodes, per env step from identical states: concentration 0.1: exactly 0 (no coupled
    contacts); 0.3: <= 2.9e-3 m (mean 5.8e-4), <= 7.7e-3 rad; 0.5: <= 5.4e-3 m (mean 8.5e-4), <= 1.1e-2 rad.  The MEAN is
    inside SURVEY §8d's 1e-3 m / 1e-2 rad, the MAX is not: a sequential-impulse transition is only defined up to the solver
    order at that level, and episode scores are insensitive to it (tests/test_episodes_200.py).  The envelope below guards
    against regressions of either ordering."""
    r = ordering_error_single_steps(name, 34 if name != "c10_t7" else 12, golden_layouts)
    if name.startswith("c10"):
        assert r.max() == 0.0
        return
    assert r[:, 0].max() <= ORDER_ENVELOPE["pos_max"] and r[:, 0].mean() <= ORDER_ENVELOPE["pos_mean"], r[:, 0]
    assert r[:, 1].max() <= ORDER_ENVELOPE["ang_max"] and r[:, 1].mean() <= ORDER_ENVELOPE["ang_mean"], r[:, 1]
    assert r[:, 0].max() > 1e-6            # the two orders do differ once contacts are coupled


def test_space_clone_and_set_state_roundtrip(golden_layouts):
    from tests.conftest import fixture_conc
    from benchpush_b200.layouts import unpack_trials
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    name = "c30_t7"
    A = ShipIceOracleEnv({"concentration": fixture_conc(name)}, experiment=unpack_trials(golden_layouts[name]), order_mode=1)
    A.reset()
    for a in (0.3, -0.5, 0.8, 0.1):
        A.step(a, want_obs=False)
    B = A.clone()
    A.step(0.2, want_obs=False); B.step(0.2, want_obs=False)
    assert np.array_equal(A.space.state(), B.space.state())               # clone carries the arbiter cache
    st = A.space.state()
    C_ = A.clone(); C_.space.clear_cache(); C_.space.set_state(st)
    assert np.array_equal(C_.space.state()[:, :3], st[:, :3])
    np.testing.assert_array_equal(C_.space.state()[:, 3:], st[:, 3:])
