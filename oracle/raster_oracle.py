"""ORACLE — test infrastructure only.

Restatement of the egocentric observation rasters of the reference
(benchpush/common/occupancy_grid/occupancy_map.py) and of the third-party primitives that are
absent from this image (skimage.draw.polygon).  cv2.line is called directly (importable,
bit-exact).  The per-pixel Python double loops of the reference (e.g. occupancy_map.py:127-137,
:420-431) are vectorised with numpy slicing; ``*_literal`` variants keep the loops for the
equality tests on small cases.
"""
import math
import numpy as np
import cv2


# ---------------------------------------------------------------- skimage.draw.polygon (restated)
def point_in_polygon(xp, yp, x, y):
    """skimage/_shared/geometry.pyx::point_in_polygon (O'Rourke crossing test; a point on a vertex
    or an edge counts as inside).  Returns 0 outside, 1 inside, 2 vertex, 3 edge.  Literal scalar form."""
    n = len(xp)
    eps = 1e-12
    r_cross = 0
    l_cross = 0
    x1 = xp[n - 1] - x
    y1 = yp[n - 1] - y
    for i in range(n):
        x0 = xp[i] - x
        y0 = yp[i] - y
        if (-eps < x0 < eps) and (-eps < y0 < eps):
            return 2
        if (y0 > 0) != (y1 > 0):
            if ((x0 * y1 - x1 * y0) / (y1 - y0)) > 0:
                r_cross += 1
        if (y0 < 0) != (y1 < 0):
            if ((x0 * y1 - x1 * y0) / (y1 - y0)) < 0:
                l_cross += 1
        x1 = x0
        y1 = y0
    if (r_cross & 1) != (l_cross & 1):
        return 3
    if r_cross & 1:
        return 1
    return 0


def draw_polygon_literal(r, c, shape=None):
    """skimage.draw.polygon(r, c, shape) — literal loops (skimage/draw/_draw.pyx::_polygon)."""
    r = np.asarray(r, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    minr = int(max(0, r.min()))
    maxr = int(math.ceil(r.max()))
    minc = int(max(0, c.min()))
    maxc = int(math.ceil(c.max()))
    if shape is not None:
        maxr = min(shape[0] - 1, maxr)
        maxc = min(shape[1] - 1, maxc)
    rr, cc = [], []
    for r_ in range(minr, maxr + 1):
        for c_ in range(minc, maxc + 1):
            if point_in_polygon(c, r, float(c_), float(r_)):
                rr.append(r_)
                cc.append(c_)
    return np.array(rr, dtype=np.intp), np.array(cc, dtype=np.intp)


def polygon_mask(r, c, shape=None):
    """Vectorised draw.polygon: returns (minr, minc, mask) with mask[r_-minr, c_-minc] == inside."""
    r = np.asarray(r, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    minr = int(max(0, r.min()))
    maxr = int(math.ceil(r.max()))
    minc = int(max(0, c.min()))
    maxc = int(math.ceil(c.max()))
    if shape is not None:
        maxr = min(shape[0] - 1, maxr)
        maxc = min(shape[1] - 1, maxc)
    if maxr < minr or maxc < minc:
        return minr, minc, np.zeros((0, 0), dtype=bool)
    Y, X = np.meshgrid(np.arange(minr, maxr + 1, dtype=np.float64),
                       np.arange(minc, maxc + 1, dtype=np.float64), indexing="ij")
    n = len(r)
    eps = 1e-12
    vertex = np.zeros(Y.shape, dtype=bool)
    rc = np.zeros(Y.shape, dtype=np.int64)
    lc = np.zeros(Y.shape, dtype=np.int64)
    x1 = c[n - 1] - X
    y1 = r[n - 1] - Y
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            x0 = c[i] - X
            y0 = r[i] - Y
            vertex |= (-eps < x0) & (x0 < eps) & (-eps < y0) & (y0 < eps)
            q = (x0 * y1 - x1 * y0) / (y1 - y0)
            rc += ((y0 > 0) != (y1 > 0)) & (q > 0)
            lc += ((y0 < 0) != (y1 < 0)) & (q < 0)
            x1, y1 = x0, y0
    inside = vertex | ((rc & 1) != (lc & 1)) | ((rc & 1) == 1)
    return minr, minc, inside


def draw_polygon(r, c, shape=None):
    minr, minc, m = polygon_mask(r, c, shape)
    rr, cc = np.nonzero(m)
    return rr + minr, cc + minc


# ---------------------------------------------------------------- polygon.py:25-41
def poly_area(vertices):
    x, y = np.asarray(vertices).T
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def poly_centroid(vertices):
    x, y = np.asarray(vertices).T
    A = poly_area(vertices)
    u = x * np.roll(y, 1) - np.roll(x, 1) * y
    return np.abs((1 / (6 * A) * np.dot(x + np.roll(x, 1), u),
                   1 / (6 * A) * np.dot(y + np.roll(y, 1), u)))


class OccupancyGrid:
    """occupancy_map.py:9-35 (constructor) and the ship-ice egocentric channel methods."""

    def __init__(self, grid_width, grid_height, map_width, map_height, local_width=6, local_height=6,
                 meter_to_pixel_scale=50):
        self.grid_width = grid_width
        self.grid_height = grid_height
        self.map_width = map_width
        self.map_height = map_height
        self.occ_map_width = int(self.map_width / self.grid_width)
        self.occ_map_height = int(self.map_height / self.grid_height)
        self.meter_to_pixel_scale = meter_to_pixel_scale
        self.local_width = local_width
        self.local_height = local_height
        self.local_window_height = int(self.local_height * meter_to_pixel_scale)
        self.local_window_width = int(self.local_width * meter_to_pixel_scale)

    # occupancy_map.py:37-65
    def compute_occ_img(self, obstacles, ice_binary_w, ice_binary_h, local_range=None, ship_state=None):
        s = self.meter_to_pixel_scale
        raw = np.zeros((ice_binary_h, ice_binary_w))
        for ob in obstacles:
            if local_range is not None:
                robot_x, robot_y = ship_state[:2]
                range_x, range_y = local_range
                center_x, center_y = np.abs(poly_centroid(ob))
                if abs(robot_x - center_x) > range_x or abs(robot_y - center_y) > range_y:
                    continue
            ob = np.asarray(ob) * s
            minr, minc, m = polygon_mask(ob[:, 1], ob[:, 0], shape=raw.shape)
            if m.size == 0 or not m.any():
                continue
            raw[minr:minr + m.shape[0], minc:minc + m.shape[1]][m] = 1.0
        return raw

    # occupancy_map.py:97-109 — block_reduce(mean) with block (int(gh*s), int(gw*s)); for the
    # egocentric grid (grid = 1/s) the block is (1,1) and the reduction is the identity.
    def compute_con_gridmap(self, raw):
        by = int(self.grid_height * self.meter_to_pixel_scale)
        bx = int(self.grid_width * self.meter_to_pixel_scale)
        if by == 1 and bx == 1:
            return raw
        H, W = raw.shape
        Hp, Wp = -(-H // by) * by, -(-W // bx) * bx
        pad = np.zeros((Hp, Wp))
        pad[:H, :W] = raw
        return pad.reshape(Hp // by, by, Wp // bx, bx).mean(axis=(1, 3))

    def _window(self, ship_state, vertical_shift):
        sx = self.occ_map_width / self.map_width
        sy = self.occ_map_height / self.map_height
        robot_x, robot_y = ship_state[:2]
        window_x = int(robot_x * sx)
        window_y = int((robot_y + vertical_shift) * sy)
        return window_x, window_y

    def _crop(self, global_map, ship_state, vertical_shift, fill):
        """Vectorised form of the crop loops (occupancy_map.py:127-137, :397-407, :509-518, :575-584)."""
        wx, wy = self._window(ship_state, vertical_shift)
        Hh, Ww = self.local_window_height, self.local_window_width
        out = np.full((Hh, Ww), fill, dtype=np.float64)
        gi = (np.arange(Hh) + wy - (Hh / 2)).astype(np.int64)
        gj = (np.arange(Ww) + wx - (Ww / 2)).astype(np.int64)
        vi = (gi >= 0) & (gi < global_map.shape[0])
        vj = (gj >= 0) & (gj < global_map.shape[1])
        out[np.ix_(vi, vj)] = global_map[np.ix_(gi[vi], gj[vj])]
        return out

    def _crop_literal(self, global_map, ship_state, vertical_shift, fill):
        wx, wy = self._window(ship_state, vertical_shift)
        out = np.full((self.local_window_height, self.local_window_width), fill, dtype=np.float64)
        for li in range(self.local_window_height):
            for lj in range(self.local_window_width):
                gi = int(li + wy - (self.local_window_height / 2))
                gj = int(lj + wx - (self.local_window_width / 2))
                if gi < 0 or gi >= global_map.shape[0] or gj < 0 or gj >= global_map.shape[1]:
                    continue
                out[li, lj] = global_map[gi, gj]
        return out

    # occupancy_map.py:112-140
    def ego_view_obstacle_map(self, raw_ice_binary, ship_state, vertical_shift):
        return self._crop(self.compute_con_gridmap(raw_ice_binary), ship_state, vertical_shift, 0.0)

    # occupancy_map.py:300-337
    def _compute_global_footprint(self, ship_state, ship_vertices):
        g = np.zeros((self.occ_map_height, self.occ_map_width)) + 0.5
        sx = self.occ_map_width / self.map_width
        sy = self.occ_map_height / self.map_height
        position = ship_state[:2]
        heading = ship_state[2]
        R = np.asarray([[math.cos(heading), -math.sin(heading)], [math.sin(heading), math.cos(heading)]])
        vertices = np.asarray(ship_vertices) @ R.T + np.asarray(position)
        r, c = [], []
        for x, y in vertices:
            gx = x * sx
            gy = y * sy
            if gy < 0 or gy >= self.occ_map_height or gx < 0 or gx >= self.occ_map_width:
                continue
            r.append(gy)
            c.append(gx)
        if len(r) == 0:
            return None
        rr, cc = draw_polygon(r, c)
        g[rr, cc] = 1.0
        return g

    # occupancy_map.py:253-296 (the planner / global-observation footprint: zeros, 1.0 inside the in-grid vertices' polygon)
    def compute_ship_footprint_planner(self, ship_state, ship_vertices):
        self.footprint = np.zeros((self.occ_map_height, self.occ_map_width))
        sx = self.occ_map_width / self.map_width
        sy = self.occ_map_height / self.map_height
        position = ship_state[:2]
        heading = ship_state[2]
        R = np.asarray([[math.cos(heading), -math.sin(heading)], [math.sin(heading), math.cos(heading)]])
        vertices = np.asarray(ship_vertices) @ R.T + np.asarray(position)
        r, c = [], []
        for x, y in vertices:
            gx = x * sx
            gy = y * sy
            if gy < 0 or gy >= self.occ_map_height or gx < 0 or gx >= self.occ_map_width:
                continue
            r.append(gy)
            c.append(gx)
        if len(r) == 0 or len(c) == 0:
            return
        rr, cc = draw_polygon(r, c)
        self.footprint[rr, cc] = 1.0

    # occupancy_map.py:379-410
    def ego_view_footprint(self, ship_state, ship_vertices, vertical_shift=2):
        g = self._compute_global_footprint(ship_state, ship_vertices)
        if g is None:
            # the reference dereferences None here (AttributeError); unreachable for in-bounds ships
            raise AttributeError("'NoneType' object has no attribute 'shape'")
        return self._crop(g, ship_state, vertical_shift, 0.0)

    # occupancy_map.py:413-433 (double loop -> one column broadcast)
    def global_goal_dist_transform(self, goal_y):
        g2m = self.map_height / self.occ_map_height
        i = np.arange(self.occ_map_height, dtype=np.float64)
        d = goal_y - i * g2m
        d = np.where(d < 0, 0.0, d)
        col = d / goal_y
        return np.repeat(col[:, None], self.occ_map_width, axis=1)

    # occupancy_map.py:492-521
    def ego_view_goal_dist_transform(self, goal_y, ship_state, vertical_shift=2):
        return self._crop(self.global_goal_dist_transform(goal_y), ship_state, vertical_shift, 1.0)

    # occupancy_map.py:524-554
    def global_orientation_map(self, ship_state, head, tail):
        g = np.zeros((self.occ_map_height, self.occ_map_width))
        sx = self.occ_map_width / self.map_width
        sy = self.occ_map_height / self.map_height
        position = ship_state[:2]
        heading = ship_state[2]
        R = np.asarray([[math.cos(heading), -math.sin(heading)], [math.sin(heading), math.cos(heading)]])
        head_pos = np.array(head) @ R.T + np.array(position)
        tail_pos = np.array(tail) @ R.T + np.array(position)
        with np.errstate(invalid="ignore"):
            head_pix = np.array([head_pos[0] * sx, head_pos[1] * sy]).astype(np.uint16)
            tail_pix = np.array([tail_pos[0] * sx, tail_pos[1] * sy]).astype(np.uint16)
        cv2.line(g, head_pix, tail_pix, color=0.5, thickness=1)
        head_pix = np.clip(head_pix, a_min=[0, 0], a_max=[g.shape[1] - 1, g.shape[0] - 1])
        g[head_pix[1], head_pix[0]] = 1.0
        return g

    # occupancy_map.py:557-587
    def ego_view_orientation_map(self, ship_state, head, tail, vertical_shift):
        return self._crop(self.global_orientation_map(ship_state, head, tail), ship_state, vertical_shift, 0.0)
