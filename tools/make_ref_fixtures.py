"""Golden vectors from the REFERENCE ITSELF for the env layer of the hot path (tests/golden/ref_envlayer.npz).

The four env modules of the reference cannot be imported here (pymunk / gymnasium absent), but the plain-Python layer they
call can: this script imports, from /root/reference,
    benchpush/common/geometry/polygon.py            poly_area, poly_centroid
    benchpush/common/evaluation/metrics.py          total_work_done, obs_to_goal_difference, euclid_dist
    benchpush/common/cost_map.py                    CostMap.get_obs_from_poly
    benchpush/common/utils/utils.py                 DotDict
    benchpush/common/metrics/*.py                   ShipIceMetric, MazeNamoMetric, TaskDrivenMetric (networkx is installed)
    benchpush/common/controller/dp.py               DP.ideal_control, TargetCourse
    benchpush/common/occupancy_grid/occupancy_map.py  every OccupancyGrid method the 2D envs call
runs them on seeded inputs and stores inputs + outputs.  tests/test_ref_pins_cpu.py holds the oracle (and through it the
kernels) to these outputs.

What the stubs stand in for (absent third-party packages; everything else is the reference's own code, unmodified):
  * matplotlib: imported at module top by the reference, never called on these paths -> empty modules;
  * skimage.draw.polygon / skimage.draw.line / skimage.measure.block_reduce / approximate_polygon: the restatements of
    oracle/ (published algorithms; these primitives themselves stay UNPINNED, the reference loops around them are pinned);
  * shapely Polygon / Point: area, centroid, x, y, intersects for convex polygons (closed-set SAT), restated.
Run in the build container only (it needs /root/reference):  python tools/make_ref_fixtures.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_envlayer.npz")


def install_stubs():
    from oracle import raster_oracle as ro
    from oracle import box_delivery_oracle as bo
    from oracle.area_clearing_oracle import convex_intersects

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    plt = mod("matplotlib.pyplot")
    plt.__getattr__ = lambda k: _Any()
    mpl = mod("matplotlib", pyplot=plt, patches=mod("matplotlib.patches"), ticker=mod("matplotlib.ticker"), colors=mod("matplotlib.colors"),
              cm=mod("matplotlib.cm"))
    for sub in ("patches", "ticker", "colors", "cm"):
        sys.modules["matplotlib." + sub].__getattr__ = lambda k: _Any()
    mpl.__getattr__ = lambda k: _Any()

    def polygon(r, c, shape=None):
        return ro.draw_polygon(np.asarray(r, dtype=np.float64), np.asarray(c, dtype=np.float64), shape)

    def line(r0, c0, r1, c1):
        return bo.draw_line(int(r0), int(c0), int(r1), int(c1))

    def block_reduce(image, block_size, func=np.sum, cval=0, func_kwargs=None):
        image = np.asarray(image)
        bs = tuple(block_size) if hasattr(block_size, "__len__") else (block_size,) * image.ndim
        pad = [(0, (-image.shape[i]) % bs[i]) for i in range(image.ndim)]
        image = np.pad(image, pad, mode="constant", constant_values=cval)
        H, W = image.shape
        v = image.reshape(H // bs[0], bs[0], W // bs[1], bs[1])
        return func(v, axis=(1, 3))

    drawmod = mod("skimage.draw.draw", polygon=polygon, line=line)
    drawpkg = mod("skimage.draw", draw=drawmod, polygon=polygon, line=line)
    meas = mod("skimage.measure", block_reduce=block_reduce, approximate_polygon=lambda coords, tolerance: bo.approximate_polygon(np.asarray(coords), tolerance))
    morph = mod("skimage.morphology", disk=bo.disk)
    morph.__getattr__ = lambda k: _Any()
    tr = mod("skimage.transform")
    tr.__getattr__ = lambda k: _Any()
    mod("skimage", draw=drawpkg, measure=meas, morphology=morph, transform=tr)

    class Point:
        def __init__(self, x, y=None):
            if y is None:
                x, y = x
            self.x, self.y = float(x), float(y)

    class Polygon:
        def __init__(self, coords):
            self.v = np.asarray(coords, dtype=np.float64).reshape(-1, 2)

        @property
        def area(self):
            x, y = self.v[:, 0], self.v[:, 1]
            return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))

        @property
        def centroid(self):
            x, y = self.v[:, 0], self.v[:, 1]
            cr = x * np.roll(y, -1) - np.roll(x, -1) * y
            A = 0.5 * cr.sum()
            return Point(((x + np.roll(x, -1)) * cr).sum() / (6 * A), ((y + np.roll(y, -1)) * cr).sum() / (6 * A))

        def intersects(self, other):
            return bool(convex_intersects(self.v, other.v))

    class _Geo:
        def __init__(self, *a, **k):
            pass

    geom = mod("shapely.geometry", Polygon=Polygon, Point=Point, LinearRing=_Geo, LineString=_Geo, MultiPolygon=_Geo)
    mod("shapely", geometry=geom)
    mod("dubins")
    return Polygon, Point


class _Vec(tuple):      # pymunk.Vec2d is a NamedTuple (x, y)
    def __new__(cls, x, y):
        return tuple.__new__(cls, (float(x), float(y)))

    x = property(lambda self: self[0])
    y = property(lambda self: self[1])


class _Body:           # the attributes of pymunk.Body that get_obs_from_poly reads
    def __init__(self, x, y, a):
        self.position, self.angle = _Vec(x, y), float(a)


class _Poly:           # ... and of pymunk.Poly
    def __init__(self, verts, body):
        self._v, self.body = [_Vec(*v) for v in verts], body

    def get_vertices(self):
        return self._v


def main():
    Polygon, Point = install_stubs()
    sys.path.insert(0, REF)
    from benchpush.common.geometry.polygon import poly_area, poly_centroid
    from benchpush.common.evaluation.metrics import total_work_done, obs_to_goal_difference, euclid_dist
    from benchpush.common.cost_map import CostMap
    from benchpush.common.utils.utils import DotDict
    from benchpush.common.metrics.ship_ice_metric import ShipIceMetric
    from benchpush.common.metrics.maze_namo_metric import MazeNamoMetric
    from benchpush.common.metrics.task_driven_metric import TaskDrivenMetric
    from benchpush.common.controller.dp import DP
    from benchpush.common.occupancy_grid.occupancy_map import OccupancyGrid

    rng = np.random.default_rng(20261018)
    out = {}

    # ---- polygons, work, push reward (E9, A1's obs_to_goal_difference) --------------------------------------------
    def rand_poly(cx, cy, r, n):
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        return np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], axis=1)

    polys_a = [rand_poly(rng.uniform(-6, 6), rng.uniform(-6, 12), rng.uniform(0.2, 0.6), int(rng.integers(4, 12))) for _ in range(40)]
    polys_b = []
    for p in polys_a:
        th = rng.uniform(-0.3, 0.3)
        Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        c = p.mean(axis=0)
        polys_b.append((p - c) @ Rm.T + c + rng.uniform(-0.4, 0.4, 2) * (rng.random() < 0.6))
    out["poly_nv"] = np.array([len(p) for p in polys_a])
    out["poly_a"] = np.concatenate(polys_a)
    out["poly_b"] = np.concatenate(polys_b)
    out["poly_area"] = np.array([poly_area(p) for p in polys_a])
    out["poly_centroid"] = np.array([poly_centroid(p) for p in polys_a])
    out["total_work"] = np.array([total_work_done(polys_a, polys_b), total_work_done(polys_a[:7], polys_b[:7])])
    out["euclid"] = np.array([euclid_dist(tuple(polys_a[i][0]), tuple(polys_b[i][1])) for i in range(5)])
    goal_pts = [(float(x), float(y)) for x, y in rng.uniform(-8, 8, (12, 2))]
    square = np.array([[-5.0, -5.0], [5.0, -5.0], [5.0, 5.0], [-5.0, 5.0]])
    out["goal_pts"] = np.array(goal_pts)
    out["square"] = square
    gp = [Point(x, y) for x, y in goal_pts]          # the envs pass shapely Points (area_clearing.py:233-261)
    out["push_diff"] = np.array([obs_to_goal_difference(polys_a, polys_b, gp, Polygon(square)),
                                 obs_to_goal_difference(polys_a[5:20], polys_b[5:20], gp[:3], Polygon(square))])

    # ---- CostMap.get_obs_from_poly (E8) ----------------------------------------------------------------------------
    bodies = [(_Body(rng.uniform(0, 12), rng.uniform(0, 40), rng.uniform(-3, 3))) for _ in range(10)]
    locs = [rand_poly(0, 0, rng.uniform(0.2, 0.6), int(rng.integers(3, 10))) for _ in range(10)]
    got = CostMap.get_obs_from_poly([_Poly(l, b) for l, b in zip(locs, bodies)])
    out["gop_pose"] = np.array([[b.position.x, b.position.y, b.angle] for b in bodies])
    out["gop_nv"] = np.array([len(l) for l in locs])
    out["gop_local"] = np.concatenate(locs)
    out["gop_world"] = np.concatenate([np.asarray(g) for g in got])

    # ---- DotDict / cfg merge (E1) ----------------------------------------------------------------------------------
    base = DotDict.to_dot_dict({"a": 1, "sim": {"steps": 400, "gravity": (0, 0)}, "ship": {"mass": 1}})
    user = {"sim": {"steps": 10}, "b": 2, "new": {"k": 7}}
    for cfg_type in user:                            # the merge block of ship_ice_env.py:47-56, verbatim
        if type(user[cfg_type]) is DotDict or type(user[cfg_type]) is dict:
            if cfg_type not in base:
                base[cfg_type] = DotDict()
            for param in user[cfg_type]:
                base[cfg_type][param] = user[cfg_type][param]
        else:
            base[cfg_type] = user[cfg_type]
    out["dotdict"] = np.array([base.a, base.sim.steps, base.sim.gravity[0], base.ship.mass, base.b, base.new.k], dtype=np.float64)

    # ---- metrics (E11, T1, T2) -------------------------------------------------------------------------------------
    T = 40
    st = np.zeros((T + 1, 3)); st[0] = (5.0, 1.0, np.pi / 2)
    for t in range(T):
        st[t + 1] = st[t] + (rng.normal(0, 0.02), 0.24, rng.normal(0, 0.05))
    st = np.round(st, 2)
    work = np.cumsum(np.abs(rng.normal(0.02, 0.02, T)))
    rew = rng.normal(0.5, 1.0, T)
    m = ShipIceMetric("pin", ship_mass=1.0, goal=(0, 9.0))
    rows = []
    for ep, succ in enumerate((True, False)):
        m.reset({"state": tuple(st[0])})
        for t in range(T):
            m.update({"state": tuple(st[t + 1]), "total_work": float(work[t]), "trial_success": succ and t == T - 1}, float(rew[t]), eps_complete=(t == T - 1))
    out["sm_state"], out["sm_work"], out["sm_rew"] = st, work, rew
    out["sm_scores"] = np.array([m.efficiency_scores, m.effort_scores, m.rewards])
    goal_dt = rng.integers(0, 200, (240, 240)).astype(np.float64)
    mst = st.copy(); mst[:, 0] += 3.0; mst[:, 1] += 2.0
    mm = MazeNamoMetric("pin", robot_mass=2.5)
    for ep, succ in enumerate((True, False)):
        mm.reset({"state": tuple(mst[0]), "goal_dt": goal_dt, "m_to_pix_scale": 16})
        for t in range(T):
            mm.update({"state": tuple(mst[t + 1]), "total_work": float(work[t]), "trial_success": succ and t == T - 1, "goal_dt": goal_dt}, float(rew[t]), eps_complete=(t == T - 1))
    out["mm_state"], out["mm_goal_dt"] = mst, goal_dt
    out["mm_scores"] = np.array([mm.efficiency_scores, mm.effort_scores, mm.rewards])
    boxes = [np.array([[cx - .22, cy - .22], [cx + .22, cy - .22], [cx + .22, cy + .22], [cx - .22, cy + .22]]) for cx, cy in rng.uniform(-4, 4, (6, 2))]
    goals = [Point(4.5, 2.0), Point(-4.5, -2.0), Point(0.0, 4.9)]
    tm = TaskDrivenMetric("pin", robot_mass=1.2)
    tm.box_mass = None                                # 2D policies never pass it (box_delivery/ppo/policy.py:88); SURVEY D11
    tstate = np.round(np.cumsum(rng.normal(0, 0.3, (T + 1, 3)), axis=0), 3)
    status_seq = np.zeros((T, 6), dtype=bool)
    for k, t0 in enumerate((5, 12, 20, 33)):
        status_seq[t0:, k] = True
    tm.reset({"state": tuple(tstate[0]), "obs": boxes, "goal_positions": goals})
    for t in range(T):
        tm.update({"state": tuple(tstate[t + 1]), "total_work": float(work[t]), "box_completed_statuses": list(status_seq[t])}, float(rew[t]), eps_complete=(t == T - 1))
    out["tm_boxes"] = np.array(boxes); out["tm_goals"] = np.array([[g.x, g.y] for g in goals]); out["tm_state"] = tstate; out["tm_status"] = status_seq
    out["tm_scores"] = np.array([tm.success_rates[0], tm.efficiency_scores[0], tm.effort_scores[0], tm.rewards[0]])

    # ---- DP.ideal_control + TargetCourse through BoxDeliveryEnv.controller (C1; box_delivery_env.py:867-885) ------------
    dp_rows = []
    for case, (ctrl, path) in enumerate([({"dt": 0.2 / 100, "target_speed": 0.2}, np.array([[0.0, 0.0, 0.3], [1.2, 0.7, 0.5]])),
                                          ({"dt": 0.2 / 100, "target_speed": 0.2, "Lfc": 0.5}, np.array([[0.0, 0.0, 0.3], [0.3, 0.2, 0.5]])),
                                          ({"dt": 0.2 / 100, "target_speed": 0.2, "Lfc": 0.5}, np.array([[1.0, 1.0, -1.0], [-0.8, 2.5, 2.0]]))]):
        dp = None
        x, y, h = float(path[0, 0]) + 0.01, float(path[0, 1]) - 0.02, float(path[0, 2])
        for k in range(120):
            if dp is None:                                  # the body of BoxDeliveryEnv.controller, verbatim calls into dp.py
                cx, cy, ch = path.T[0][0:2], path.T[1][0:2], path.T[2][0:2]
                dp = DP(x=x, y=y, yaw=h, cx=cx, cy=cy, ch=ch, **ctrl)
            omega, v = dp.ideal_control(x, y, h)
            x_s, y_s, h_s = dp.get_setpoint()
            dp.setpoint = np.asarray([x_s, y_s, h_s])
            dp_rows.append([case, x, y, h, omega, v[0], v[1], x_s, y_s])
            x += 2 * v[0] * ctrl["dt"]; y += 2 * v[1] * ctrl["dt"]; h += float(np.clip(3 * omega, -20, 20)) * ctrl["dt"]
    out["dp_rows"] = np.array(dp_rows)

    # ---- OccupancyGrid (E7, global mode, M1) -----------------------------------------------------------------------
    from benchpush_b200.layouts import load_packed, unpack_trials
    trial = unpack_trials(load_packed(os.path.join(ROOT, "tests", "golden", "ice_c30_t7.npz")))[0]
    obstacles = [np.asarray(o["vertices"], dtype=np.float64) for o in trial["obstacles"]]
    ship_verts = [[1., -0.], [0.9, 0.10], [0.5, 0.25], [0.25, 0.25], [0, 0.25], [-0.25, 0.25], [-0.5, 0.25], [-0.75, 0.25], [-1., 0.25], [-1., -0.25],
                  [-0.75, -0.25], [-0.5, -0.25], [-0.25, -0.25], [0, -0.25], [0.25, -0.25], [0.5, -0.25], [0.9, -0.10]]
    occ = OccupancyGrid(grid_width=1 / 25, grid_height=1 / 25, map_width=12, map_height=40, local_width=6, local_height=6, ship_body=None, meter_to_pixel_scale=25)
    poses = [(5.3, 1.0, np.pi / 2), (0.3, 4.77, 2.1), (11.9, 8.2, 0.4), (6.01, 2.5, np.pi - 0.05), (-0.2, 3.3, 1.3)]
    obs_rows = []
    for pose in poses:
        raw = occ.compute_occ_img(obstacles=obstacles, ice_binary_w=300, ice_binary_h=1000, local_range=[12, 12], ship_state=pose)
        ch = [occ.ego_view_footprint(ship_state=pose, ship_vertices=ship_verts, vertical_shift=2),
              occ.ego_view_goal_dist_transform(goal_y=9.0, ship_state=pose, vertical_shift=2),
              occ.ego_view_orientation_map(ship_state=pose, head=[1., 0.], tail=[-1., 0.], vertical_shift=2),
              occ.ego_view_obstacle_map(raw_ice_binary=raw, ship_state=pose, vertical_shift=2)]
        obs_rows.append((np.stack([np.array(c) for c in ch]) * 255).astype(np.uint8))
    out["occ_poses"] = np.array(poses)
    out["occ_obs"] = np.stack(obs_rows)
    occ_g = OccupancyGrid(grid_width=0.2, grid_height=0.2, map_width=12, map_height=40, local_width=6, local_height=6, ship_body=None, meter_to_pixel_scale=25)
    glob = []
    for pose in poses[:3]:
        raw = occ_g.compute_occ_img(obstacles=obstacles, ice_binary_w=300, ice_binary_h=1000)
        occ_g.compute_con_gridmap(raw_ice_binary=raw, save_fig_dir=None)
        o = np.copy(occ_g.occ_map)
        occ_g.compute_ship_footprint_planner(ship_state=pose, ship_vertices=np.array(ship_verts))
        f = np.copy(occ_g.footprint)
        glob.append((np.stack([o, f]) * 255).astype(np.uint8))
    out["occ_global"] = np.stack(glob)
    # maze: wall image, goal-point wavefront map, the 4-channel rotated egocentric view
    walls = [[(0, 0), (15, 0)], [(0, 0), (0, 15)], [(15, 0), (15, 15)], [(0, 15), (15, 15)], [(15, 15), (15, 5)], [(7.5, 0), (7.5, 10)]]
    mo = OccupancyGrid(grid_width=1 / 16, grid_height=1 / 16, map_width=15, map_height=15, local_width=12, local_height=12, ship_body=None, meter_to_pixel_scale=16)
    wall_img = mo.compute_occ_img_walls(walls, 240, 240, wall_radius=0.5)
    dist_map, dist_raw = mo.global_goal_point_dist_transform((3.75, 3.75), walls, wall_radius=0.5)
    rv = np.array([[0.7, -0.5], [0.55, -0.6], [-0.55, -0.6], [-0.7, -0.5], [-0.7, 0.5], [-0.55, 0.6], [0.55, 0.6], [0.7, 0.5]])
    mboxes = [np.array([[cx - .5, cy - .5], [cx + .5, cy - .5], [cx + .5, cy + .5], [cx - .5, cy + .5]]) for cx, cy in ((10.0, 6.0), (12.5, 9.0), (3.0, 11.0), (5.0, 5.5), (11.0, 12.5))]
    mposes = [(11.25, 3.75, np.pi / 2), (10.2, 7.1, 2.0), (4.4, 12.0, -0.7)]
    out["maze_wall_img"] = np.asarray(wall_img)
    out["maze_dist_map"] = np.asarray(dist_map)
    out["maze_dist_raw"] = np.asarray(dist_raw)
    out["maze_boxes"] = np.array(mboxes)
    out["maze_poses"] = np.array(mposes)
    out["maze_obs"] = np.stack([(np.asarray(mo.ego_view_map_maze(p, rv, mboxes, walls, dist_map)) * 255).astype(np.uint8) for p in mposes])
    out["maze_robot_verts"] = rv
    # ---- PositionController (C1): the reference class with the reference's own compiled spfa (oracle/_ref), on the
    # configuration space of walled_env_with_columns so that both the straight-line and the spfa + Douglas-Peucker branch run
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from benchpush.common.controller.position_controller import PositionController
    from oracle.area_clearing_oracle import AreaClearingOracleEnv, sample_layout
    import random as _random
    acfg = {"env": "walled_env_with_columns", "num_obstacles": 4}
    ora = AreaClearingOracleEnv(cfg=acfg, layouts=[sample_layout(_random.Random(3), acfg)])
    ora.reset()
    pc = PositionController(None, ora.robot_radius, ora.map_width, ora.map_height, ora.configuration_space, ora.configuration_space_thin,
                            ora.closest_cspace_indices, ora.local_map_pixel_width, ora.local_map_width, ora.local_map_pixels_per_meter,
                            np.radians(15), 0.05, 0.03, np.radians(1))
    LW = ora.local_map_pixel_width
    pc_rows, pc_paths = [], []
    for (x, y, h, i, j) in [(0.0, -7.0, np.pi / 2, 20, 112), (-3.0, -5.5, 0.3, 60, 200), (3.5, 0.0, -2.0, 112, 10), (0.0, -7.0, np.pi / 2, 5, 112),
                            (-6.0, 6.5, -1.2, 30, 180), (2.0, 2.0, 2.5, 200, 60), (0.5, -3.0, 1.6, 40, 112), (6.5, -7.0, 3.0, 100, 20)]:
        path, sign = pc.get_waypoints_to_spatial_action([x, y], h, int(i * LW + j))
        pc_rows.append([x, y, h, i * LW + j, sign, len(path)])
        pc_paths.append(np.asarray(path, dtype=np.float64)[:, :2].reshape(-1))
    out["pc_rows"] = np.array(pc_rows)
    out["pc_paths"] = np.concatenate(pc_paths)
    out["pc_cspace_sum"] = np.array([ora.configuration_space.sum()])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
