"""Box-delivery oracle: the third-party primitives it restates are pinned against the libraries in this image
(cv2.fillPoly, scipy.ndimage.rotate order 0), the absent ones (skimage line / approximate_polygon) are checked
against their defining properties, and the Chipmunk handler set of the env behaves as the reference's callbacks."""
import numpy as np
import pytest

from oracle import cpo
from oracle import box_delivery_oracle as bdo
from tests.boxdel_common import DELIVERY_CFG, DELIVERY_LAYOUT


def test_fill_poly_restatement_is_pinned_to_cv2():
    import cv2
    rng = np.random.default_rng(0)
    for t in range(1500):
        n = int(rng.integers(3, 7))
        c = rng.uniform(30, 210, 2)
        r = rng.uniform(2, 25)
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        if rng.random() < 0.5:
            ang = ang[::-1]
        pts = (c + r * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(np.int32)
        a = np.zeros((244, 244), np.float32)
        b = a.copy()
        cv2.fillPoly(a, [pts], color=0.5)
        bdo.fill_poly(b, pts, 0.5)
        assert np.array_equal(a, b), pts


def test_rotate_nearest_restatement_is_pinned_to_scipy():
    from scipy.ndimage import rotate
    rng = np.random.default_rng(1)
    for t in range(120):
        H, W = rng.integers(20, 60, 2)
        img = rng.random((H, W)).astype(np.float32)
        ang = [0, 90, 180, 270, -90, 45, 30, 360, -180, 135][t] if t < 10 else rng.uniform(-400, 400)
        a = rotate(img, ang, order=0)
        b = bdo.rotate_nearest(img, ang)
        assert a.shape == b.shape and np.array_equal(a, b), ang
    # the environment's case: a 318 x 318 crop
    img = rng.random((318, 318)).astype(np.float32)
    for ang in (12.5, -77.0, 90 - np.degrees(4.6)):
        assert np.array_equal(rotate(img, ang, order=0), bdo.rotate_nearest(img, ang))


def test_draw_line_is_bresenham_with_both_ends():
    rng = np.random.default_rng(2)
    for _ in range(300):
        r0, c0, r1, c1 = (int(v) for v in rng.integers(0, 60, 4))
        rr, cc = bdo.draw_line(r0, c0, r1, c1)
        assert (rr[0], cc[0]) == (r0, c0) and (rr[-1], cc[-1]) == (r1, c1)
        assert len(rr) == max(abs(r1 - r0), abs(c1 - c0)) + 1
        assert np.all(np.maximum(np.abs(np.diff(rr)), np.abs(np.diff(cc))) == 1) or len(rr) == 1
        # never further than half a pixel from the ideal segment along the minor axis
        if abs(c1 - c0) >= abs(r1 - r0) and c1 != c0:
            ideal = r0 + (cc - c0) * (r1 - r0) / (c1 - c0)
            assert np.all(np.abs(rr - ideal) <= 0.5 + 1e-9)


def test_approximate_polygon_keeps_corners_and_drops_collinear_points():
    line = np.array([[i, 2 * i] for i in range(20)])
    assert np.array_equal(bdo.approximate_polygon(line, 1), line[[0, -1]])
    ell = np.array([[0, j] for j in range(10)] + [[i, 9] for i in range(1, 10)])
    out = bdo.approximate_polygon(ell, 1)
    assert np.array_equal(out, np.array([[0, 0], [0, 9], [9, 9]]))
    # a one-pixel jog is within tolerance 1 -> dropped
    jog = np.array([[0, j] for j in range(6)] + [[1, j] for j in range(6, 12)])
    assert len(bdo.approximate_polygon(jog, 1)) == 2


def test_boundary_handler_pushes_a_box_out_of_the_wall_and_flags_the_robot():
    sp = cpo.Space(iterations=10, damping=0.0)
    sp.set_handler_mode(1)
    rb = sp.add_body(cpo.KINEMATIC, 0.0, 0.0, 0.0)
    s = sp.add_poly(rb, np.array([[-.3, -.3], [.3, -.3], [.3, .3], [-.3, .3]]), 0.0)
    sp.set_material(s, 0.01, 1.0, 1)
    bx = sp.add_body(cpo.DYNAMIC, 2.0, 0.9, 0.0)
    s = sp.add_poly(bx, np.array([[-.22, -.22], [.22, -.22], [.22, .22], [-.22, .22]]), 0.02)
    sp.set_density(s, 0.001)
    sp.set_material(s, 0.01, 1.0, 2)
    wall = sp.add_body(cpo.STATIC, 0, 0, 0)
    s = sp.add_poly(wall, np.array([[-5, 1.0], [5, 1.0], [5, 3.0], [-5, 3.0]]), 0.0)
    sp.set_material(s, 0.01, 1.0, 3)
    assert sp.robot_hit() == -1
    sp.step(0.002)
    # the box overlapped the wall by 0.12 + 0.02: the (2,3) pre_solve moved it out by the penetration depth
    assert abs(sp.body(bx)[1] - (1.0 - 0.24)) < 1e-9 and sp.robot_hit() == -1
    sp.set_velocity(rb, 0.0, 100.0)          # drive the robot into the wall: 0.2 per substep
    for _ in range(4):
        sp.step(0.002)
    assert sp.robot_hit() == 1
    assert sp.body(rb)[1] <= 0.7 + 0.2 + 1e-9      # corrected back every step
    # removal: the body stops integrating and leaves the broadphase
    sp.set_velocity(bx, 1.0, 0.0)
    sp.remove_body(bx)
    x = sp.body(bx)[0]
    sp.step(0.002)
    assert sp.body(bx)[0] == x
    assert sp.point_query_types((0.0, 2.0), 0) == (1 << 3)


def test_env_delivery_scenario(portable_trig):
    ora = bdo.BoxDeliveryOracleEnv(cfg=DELIVERY_CFG, layouts=[DELIVERY_LAYOUT])
    obs, info = ora.reset()
    assert obs.shape == (224, 224, 4) and obs.dtype == np.uint8
    assert set(info) == {"state", "cumulative_distance", "cumulative_boxes", "cumulative_reward", "total_work", "obs",
                         "box_completed_statuses", "goal_positions", "ministeps", "inactivity"}
    assert info["state"] == (1.5, 1.75, 0.0) and len(info["obs"]) == 4
    assert sorted(np.unique(obs[:, :, 0]).tolist()) == [0, 31, 95, 127, 191] and set(np.unique(obs[:, :, 1])) == {0, 255}
    o, r, term, trunc, info = ora.step(-1.0)                 # straight ahead: pushes box 0 towards the receptacle
    assert 0 < r < 1 and info["cumulative_boxes"] == 0
    o, r, term, trunc, info = ora.step(-1.0)
    assert r > 10 and info["cumulative_boxes"] == 1 and info["box_completed_statuses"] == [True, False, False, False]
    assert len(info["obs"]) == 3 and info["inactivity"] == 0 and not term
    o, r, term, trunc, info = ora.step(-1.0)
    assert info["inactivity"] == 1
