"""Ice-field layouts for ship-ice-v0: packing of the reference's trial format and a seeded synthetic
generator.

The reference keeps 100 pre-generated trials per concentration in pickles
(``ship_ice_env.py:75-80``: ``ddict['exp'][concentration][trial] = {goal, ship_state, obstacles}``)
and picks ``trial = experiment[episode_idx % 100]`` at reset (``ship_ice_env.py:188-192``).  Layouts
are INPUTS here too: :func:`pack_trials` turns that structure into flat arrays for the CUDA
library; :func:`load_reference_pickle` reads a reference pickle when the user has one; and
:func:`generate_ice_field` builds statistically similar fields (non-overlapping random circles,
a random convex polygon in each circle's bounding square — the recipe of ``cost_map.py:219-260`` and
``geometry/polygon.py:53-146``) from a seed, for machines without the reference data.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

VMAX = 20


@dataclass
class PackedLayouts:
    n_floes: np.ndarray      # int32 [L]
    n_verts: np.ndarray      # int32 [L, F]
    verts: np.ndarray        # float64 [L, F, VMAX, 2] world coordinates (raw trial vertices)
    centres: np.ndarray      # float64 [L, F, 2]
    ship_state: np.ndarray   # float64 [L, 3]

    @property
    def n_layouts(self) -> int:
        return int(self.n_floes.shape[0])

    @property
    def max_floes(self) -> int:
        return int(self.n_verts.shape[1])


def _poly_area(v: np.ndarray) -> float:
    # geometry/polygon.py:25-29
    x, y = v.T
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def pack_trials(trials: Sequence[dict], max_floes: int | None = None) -> PackedLayouts:
    """Pack a list of reference-format trials.  Zero-area obstacles are dropped (ship_ice_env.py:206)."""
    kept = []
    for t in trials:
        obs = [ob for ob in t["obstacles"] if _poly_area(np.asarray(ob["vertices"], dtype=np.float64)) != 0]
        kept.append(obs)
    F = max(len(o) for o in kept)
    if max_floes is not None:
        if F > max_floes:
            raise ValueError(f"layout has {F} floes > max_floes={max_floes}")
        F = max_floes
    L = len(trials)
    n_floes = np.zeros(L, dtype=np.int32)
    n_verts = np.zeros((L, F), dtype=np.int32)
    verts = np.zeros((L, F, VMAX, 2), dtype=np.float64)
    centres = np.zeros((L, F, 2), dtype=np.float64)
    ship_state = np.zeros((L, 3), dtype=np.float64)
    for l, (t, obs) in enumerate(zip(trials, kept)):
        n_floes[l] = len(obs)
        ship_state[l] = np.asarray(t["ship_state"], dtype=np.float64)
        for f, ob in enumerate(obs):
            v = np.asarray(ob["vertices"], dtype=np.float64)
            if v.shape[0] > VMAX:
                raise ValueError(f"obstacle with {v.shape[0]} vertices > {VMAX}")
            n_verts[l, f] = v.shape[0]
            verts[l, f, : v.shape[0]] = v
            centres[l, f] = np.asarray(ob["centre"], dtype=np.float64)
    return PackedLayouts(n_floes, n_verts, verts, centres, ship_state)


def unpack_trials(p: PackedLayouts) -> List[dict]:
    """Inverse of :func:`pack_trials` (reference trial format), e.g. to feed a CPU checker."""
    out = []
    for l in range(p.n_layouts):
        obs = []
        for f in range(int(p.n_floes[l])):
            n = int(p.n_verts[l, f])
            obs.append({"vertices": p.verts[l, f, :n].copy(), "centre": (p.centres[l, f, 0], p.centres[l, f, 1])})
        out.append({"goal": (2, 40), "ship_state": tuple(p.ship_state[l]), "obstacles": obs})
    return out


def load_reference_pickle(path: str, concentration: float) -> List[dict]:
    """Read one of the reference's ``experiments_XX_100_r06_d40x12.pk`` files."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    exp = d["exp"][concentration]
    return [exp[i] for i in range(len(exp))]


def save_packed(path: str, p: PackedLayouts) -> None:
    np.savez_compressed(path, n_floes=p.n_floes, n_verts=p.n_verts, verts=p.verts, centres=p.centres, ship_state=p.ship_state)


def load_packed(path: str) -> PackedLayouts:
    z = np.load(path)
    return PackedLayouts(z["n_floes"], z["n_verts"], z["verts"], z["centres"], z["ship_state"])


# ---------------------------------------------------------------------------------- synthetic generator
def _random_convex_polygon(rng: np.random.Generator, diameter: float, n: int) -> np.ndarray:
    """Valtr's random convex polygon in a ``diameter`` x ``diameter`` box (geometry/polygon.py:53-146)."""
    xs = np.sort(rng.uniform(0, diameter, n))
    ys = np.sort(rng.uniform(0, diameter, n))

    def chains(vals):
        lo, hi = vals[0], vals[-1]
        top = bot = lo
        vec = []
        for v in vals[1:-1]:
            if rng.integers(2):
                vec.append(v - top)
                top = v
            else:
                vec.append(bot - v)
                bot = v
        vec.append(hi - top)
        vec.append(bot - hi)
        return np.array(vec)

    xv = chains(xs)
    yv = chains(ys)
    rng.shuffle(yv)
    ang = np.arctan2(xv, yv)
    order = np.argsort(ang, kind="stable")
    xv, yv = xv[order], yv[order]
    px = np.concatenate(([0.0], np.cumsum(xv)[:-1]))
    py = np.concatenate(([0.0], np.cumsum(yv)[:-1]))
    px += xs[0] - px.min()
    py += ys[0] - py.min()
    return np.stack([px, py], axis=1)


def _centroid(v: np.ndarray) -> np.ndarray:
    x, y = v.T
    a = 0.5 * (np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    u = x * np.roll(y, -1) - np.roll(x, -1) * y
    return np.array([np.dot(x + np.roll(x, -1), u), np.dot(y + np.roll(y, -1), u)]) / (6 * a)


def generate_ice_field(seed: int, concentration: float, map_width: float = 12.0, min_y: float = 2.0, max_y: float = 40.0,
                       min_r: float = 0.2, max_r: float = 0.6, max_floes: int = 653) -> dict:
    """One synthetic trial in the reference's format.  Circles of radius U(min_r, max_r) are dropped
    without overlap until their area reaches ``concentration`` of the field; each carries a random
    convex polygon with 10..20 vertices whose centroid sits at the circle centre."""
    rng = np.random.default_rng(seed)
    field_area = map_width * (max_y - min_y)
    cell = 2 * max_r
    gx, gy = int(np.ceil(map_width / cell)), int(np.ceil((max_y - min_y) / cell))
    grid: List[List[List[int]]] = [[[] for _ in range(gx)] for _ in range(gy)]
    cx, cy, cr = [], [], []
    covered = 0.0
    tries = 0
    while covered < concentration * field_area and len(cx) < max_floes and tries < 200000:
        tries += 1
        r = rng.uniform(min_r, max_r)
        x = rng.uniform(r, map_width - r)
        y = rng.uniform(min_y + r, max_y - r)
        ix, iy = int(x / cell), int((y - min_y) / cell)
        ok = True
        for jy in range(max(0, iy - 1), min(gy, iy + 2)):
            for jx in range(max(0, ix - 1), min(gx, ix + 2)):
                for k in grid[jy][jx]:
                    if (cx[k] - x) ** 2 + (cy[k] - y) ** 2 < (cr[k] + r) ** 2:
                        ok = False
                        break
                if not ok:
                    break
            if not ok:
                break
        if not ok:
            continue
        grid[iy][ix].append(len(cx))
        cx.append(x); cy.append(y); cr.append(r)
        covered += np.pi * r * r
    obstacles = []
    for x, y, r in zip(cx, cy, cr):
        n = int(rng.integers(10, 21))
        poly = _random_convex_polygon(rng, 2 * r, n)
        poly = poly - _centroid(poly) + np.array([x, y])
        poly = poly[::-1].copy()          # the reference's polygons are clockwise
        obstacles.append({"vertices": poly, "centre": (x, y), "radius": r})
    ship_x = rng.uniform(1.0, 10.0)
    return {"goal": (2, 40), "ship_state": (float(ship_x), 1.0, float(np.pi / 2)), "obstacles": obstacles}


def generate_experiment(n_trials: int, concentration: float, seed: int = 99, **kw) -> List[dict]:
    return [generate_ice_field(seed * 100003 + i, concentration, **kw) for i in range(n_trials)]


# ------------------------------------------------------------------------------------------------ maze-NAMO
def maze_walls(maze_version: int, width: float, length: float):
    """construct_maze_walls (maze_NAMO_env.py:361-379): six wall centre lines [(x0, y0), (x1, y1)]."""
    W, L = width, length
    if maze_version == 1:
        return [[(0, 0), (W, 0)], [(0, 0), (0, L)], [(W, 0), (W, L)], [(0, L), (W, L)],
                [(2 * W / 2, L), (2 * W / 2, 5)], [(W / 2, 0), (W / 2, L - L / 3)]]
    if maze_version == 2:
        return [[(0, 0), (W, 0)], [(0, 0), (0, L)], [(W, 0), (W, L)], [(0, L), (W, L)],
                [(W / 3, 0), (W / 3, 2 * L / 3)], [(2 * W / 3, L), (2 * W / 3, L / 3)]]
    raise Exception("Invalid Maze Version!")


def _segment_query(p, a, b, radius, max_dist) -> bool:
    """pymunk ``space.point_query(p, max_dist, ShapeFilter())`` restricted to one Segment shape: True when the
    signed distance from p to the rounded segment is strictly below max_dist (Chipmunk NearestPointQuery), with
    the spatial-index prefilter (AABB of the segment against the box p +- max(max_dist, 0))."""
    md = max(max_dist, 0.0)
    l, r = min(a[0], b[0]) - radius, max(a[0], b[0]) + radius
    bt, tp = min(a[1], b[1]) - radius, max(a[1], b[1]) + radius
    if not (p[0] - md <= r and l <= p[0] + md and p[1] - md <= tp and bt <= p[1] + md):
        return False
    dx, dy = a[0] - b[0], a[1] - b[1]
    t = (dx * (p[0] - b[0]) + dy * (p[1] - b[1])) / (dx * dx + dy * dy)
    t = max(0.0, min(t, 1.0))
    cx, cy = b[0] + dx * t, b[1] + dy * t
    ex, ey = p[0] - cx, p[1] - cy
    return (ex * ex + ey * ey) ** 0.5 - radius < max_dist


def generate_maze_layout(rng, walls, num_obstacles: int, length: float, min_dist: float, wall_radius: float = 0.5):
    """MazeNAMO.generate_obstacles, randomised branch (maze_NAMO_env.py:275-301), with ``rng`` (a
    ``random.Random``) in place of the reference's unseeded global ``random`` module.  The space holds only the
    wall segments at that point (:224-226 vs :253).  Quirk kept: the wall test sits inside the loop over the
    previously placed obstacles, so the first obstacle is never tested against the walls."""
    obstacles = []
    while len(obstacles) < num_obstacles:
        cx = rng.random() * (length - 0) + 0
        cy = rng.random() * (length - 0) + 0
        overlapped = False
        for px, py in obstacles:
            if ((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 <= min_dist:
                overlapped = True
                break
            if any(_segment_query((cx, cy), a, b, wall_radius, min_dist) for a, b in walls):
                overlapped = True
                break
        if not overlapped:
            obstacles.append([cx, cy])
    return obstacles


def generate_maze_layouts(n: int, seed: int = 0, maze_version: int = 1, num_obstacles: int = 5, width: float = 15.0,
                          length: float = 15.0, min_dist: float = 0.62):
    """``n`` seeded obstacle layouts (lists of [x, y] centres); layout k depends on (seed, k) only."""
    import random
    walls = maze_walls(maze_version, width, length)
    return [generate_maze_layout(random.Random(seed * 1000003 + k), walls, num_obstacles, length, min_dist) for k in range(n)]


def pack_maze_layouts(layouts, max_boxes: int | None = None):
    """-> (n_boxes int32 [L], centres float64 [L, max_boxes, 2])"""
    B = max(len(l) for l in layouts) if max_boxes is None else max_boxes
    nb = np.array([len(l) for l in layouts], dtype=np.int32)
    if int(nb.max()) > B:
        raise ValueError("layout has more obstacles than max_boxes")
    cen = np.zeros((len(layouts), max(B, 1), 2), dtype=np.float64)
    for i, l in enumerate(layouts):
        if len(l):
            cen[i, :len(l)] = np.asarray(l, dtype=np.float64).reshape(-1, 2)
    return nb, cen


def _bd_robot_start(random_state, ag, room_length, room_width):
    """get_random_robot_start (box_delivery_env.py:313-320)"""
    size = max(ag["length"], ag["width"])
    x = random_state.uniform(-room_length / 2 + size, room_length / 2 - size)
    y = random_state.uniform(-room_width / 2 + size, room_width / 2 - size)
    heading = random_state.uniform(0, 2 * np.pi)
    return (float(x), float(y), float(heading))


def _bd_columns(random_state, max_num_columns, room_length, room_width, rpos, rsize, start, robot_radius):
    """add_random_columns (box_delivery_env.py:358-415): 1 .. max - 1 unit columns, 100 tries each.  A try that lands on the
    receptacle or on the robot gives the column up (the reference ``break``s out of the retry loop there); a try that lands on
    an earlier column is retried."""
    num_columns = random_state.randint(1, max_num_columns)
    column_length = column_width = 1
    buffer_width = 0.8
    col_min_dist = 2
    cols = []
    for _ in range(num_columns):
        for _ in range(100):
            x = random_state.uniform(-room_length / 2 + 2 * buffer_width + column_length / 2, room_length / 2 - 2 * buffer_width - column_length / 2)
            y = random_state.uniform(-room_width / 2 + 2 * buffer_width + column_width / 2, room_width / 2 - 2 * buffer_width - column_width / 2)
            if ((x - rpos[0]) ** 2 + (y - rpos[1]) ** 2) ** 0.5 <= col_min_dist / 2 + rsize / 2:
                break
            if ((x - start[0]) ** 2 + (y - start[1]) ** 2) ** 0.5 <= col_min_dist / 2 + robot_radius:
                break
            if any(((x - px) ** 2 + (y - py) ** 2) ** 0.5 <= col_min_dist for px, py in cols):
                continue
            cols.append([float(x), float(y)])
            break
    return cols


def _bd_divider(random_state, ag, room_length, room_width, start, robot_radius):
    """add_random_horiz_divider (box_delivery_env.py:417-462): an 8 x 0.5 divider attached to the right wall at a random y;
    after 100 tries that touch the robot the ROBOT START is redrawn (robot_info['start_pos'], which is what the sim uses)."""
    divider_length, divider_width, buffer_width = 8, 0.5, 3.5
    for _ in range(100):
        for _ in range(100):
            x = room_length / 2 - divider_length / 2
            y = random_state.uniform(-room_width / 2 + buffer_width + divider_width / 2, room_width / 2 - buffer_width - divider_width / 2)
            if not (((x - start[0]) ** 2 + (y - start[1]) ** 2) ** 0.5 <= 3 * robot_radius):
                return [float(x), float(y)], start
        start = _bd_robot_start(random_state, ag, room_length, room_width)
    return None, start


def sample_box_delivery_layout(random_state, cfg) -> dict:
    """One box-delivery episode layout drawn the way the reference draws it at every reset (box_delivery_env.py:246-252:
    get_random_robot_start :313-320, generate_boundary :326-507 — the random columns / the divider of the obstacle_config —
    then generate_boxes :509-546 by rejection sampling against the receptacle, the columns / divider and the boxes placed so
    far).  ``random_state`` is a numpy RandomState (the reference shares RandomState(cfg.misc.random_seed) across episodes; a
    vector env gives each (env, episode) its own stream).  Keys: start, boxes, columns [[x, y], ...], divider [x, y] | None."""
    env, ag, bx = cfg["env"], cfg["agent"], cfg["boxes"]
    room_length = env["room_length"]
    oc = env["obstacle_config"]
    small = oc.split("_")[0] == "small"
    room_width = env["room_width_small"] if small else env["room_width_large"]
    num_boxes = bx["num_boxes_small"] if small else bx["num_boxes_large"]
    start = _bd_robot_start(random_state, ag, room_length, room_width) if ag["random_start"] else (5.0, 1.5, float(np.pi * 3 / 2))
    rsize = env["receptacle_width"]
    rx, ry = room_length / 2 - rsize / 2, room_width / 2 - rsize / 2
    robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
    columns, divider = [], None
    if oc in ("small_columns", "large_columns"):
        columns = _bd_columns(random_state, 3 if small else 8, room_length, room_width, (rx, ry), rsize, start, robot_radius)
    elif oc == "large_divider":
        divider, start = _bd_divider(random_state, ag, room_length, room_width, start, robot_radius)
    elif not oc.endswith("_empty"):
        raise ValueError("Invalid obstacle config: %r" % (oc,))
    half = bx["box_size"] / 2
    dmin = bx["min_box_dist"]
    min_x, max_x = -room_length / 2 + half, room_length / 2 - half
    min_y, max_y = -room_width / 2 + half, room_width / 2 - half
    # obstacles generate_boxes tests: (position, 'width'); the divider by its y distance only
    round_obstacles = [((rx, ry), rsize)] + [((cx, cy), 1) for cx, cy in columns]
    boxes = []
    while len(boxes) < num_boxes:
        cx = random_state.uniform(min_x, max_x)
        cy = random_state.uniform(min_y, max_y)
        hd = random_state.uniform(0, 2 * np.pi)
        ok = True
        for (ox, oy), width in round_obstacles:
            if ((cx - ox) ** 2 + (cy - oy) ** 2) ** 0.5 <= (dmin / 2 + width / 2) * 1.2:
                ok = False
                break
        if ok and divider is not None and abs(cy - divider[1]) <= (dmin / 2 + 0.5 / 2) * 1.2:
            ok = False
        if ok:
            for px, py, _ in boxes:
                if ((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 <= dmin:
                    ok = False
                    break
        if ok:
            boxes.append([float(cx), float(cy), float(hd)])
    return {"start": start, "boxes": boxes, "columns": columns, "divider": divider}


def sample_area_clearing_layout(rng, cfg, env_cfg) -> dict:
    """One area-clearing episode layout drawn the way the reference draws it (area_clearing.py:363-368 robot start,
    :515-541 generate_obstacles) from a ``random.Random`` (the reference uses the unseeded global ``random`` module)."""
    b = env_cfg["boundary"]
    min_x, max_x = min(p[0] for p in b), max(p[0] for p in b)
    min_y, max_y = min(p[1] for p in b), max(p[1] for p in b)
    if cfg["random_start"]:
        start = ((min_x + 1) + rng.random() * ((max_x - min_x) - 2), min_y + 1.0, float(np.pi / 2))
    else:
        start = ((min_x + max_x) / 2, min_y + 1.0, float(np.pi / 2))
    lo_x, hi_x, lo_y, hi_y = min_x + 1, max_x - 1, min_y + 1, max_y - 1
    obs = []
    while len(obs) < cfg["num_obstacles"]:
        cx = rng.random() * (hi_x - lo_x) + lo_x
        cy = rng.random() * (hi_y - lo_y) + lo_y
        if all(((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 > cfg["min_obs_dist"] for px, py in obs):
            obs.append([float(cx), float(cy)])
    return {"start": tuple(float(v) for v in start), "boxes": obs}
