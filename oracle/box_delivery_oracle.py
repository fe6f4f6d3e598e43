"""ORACLE — test infrastructure only.

CPU restatement of ``BoxDeliveryEnv`` (reference benchpush/environments/box_delivery/box_delivery_env.py)
and of the position-control chain it calls (benchpush/common/controller/position_controller.py,
dp.py:217-248) without gymnasium / pymunk / matplotlib / skimage: same reset/step flow, same constants,
same info keys.

* Physics: the Chipmunk restatement (oracle/chipmunk_oracle.cpp, handler_mode 1) — PARITY UNPINNED there.
* ``cv2.fillPoly``, ``scipy.ndimage.rotate(order=0)``, ``scipy.ndimage.binary_dilation``,
  ``scipy.ndimage.distance_transform_edt`` are the libraries the reference itself calls and are
  called directly.  ``rotate_nearest`` below restates scipy's rotate(order=0, reshape=True) and is PINNED
  against the scipy in this image (tests/test_oracle_boxdel_cpu.py); ``fill_poly`` restates
  ``cv2.fillPoly`` for convex int32 polygons and is PINNED against cv2 the same way.
* ``skimage.draw.line``, ``skimage.measure.approximate_polygon``, ``skimage.morphology.disk`` are
  absent from this image: restated from their published algorithms (Bresenham with the both-end-
  inclusive update order of skimage/draw/_draw.pyx::_line; iterative Douglas-Peucker of
  skimage/measure/_polygon.py) — unpinned.
* ``spfa.spfa``: oracle/spfa_oracle.cpp, pinned against the reference's own source (oracle/_ref).

Layouts are INPUTS: the reference draws the robot start and the boxes from ``RandomState(42)``
shared across episodes (box_delivery_env.py:152, 313-320, 509-546); ``sample_layout`` restates that
sampler for one episode from a caller-supplied RandomState.
"""
import math
import numpy as np
import cv2
from scipy.ndimage import distance_transform_edt, rotate as rotate_image, binary_dilation

from . import cpo
from .ship_ice_oracle import merge_cfg, total_work_done

OBSTACLE_SEG_INDEX, FLOOR_SEG_INDEX, RECEPTACLE_SEG_INDEX, BOX_SEG_INDEX, ROBOT_SEG_INDEX, MAX_SEG_INDEX = 0, 1, 3, 4, 6, 8
MOVE_STEP_SIZE = 0.05                          # box_delivery_env.py:53-62
TURN_STEP_SIZE = np.radians(15)
WAYPOINT_MOVING_THRESHOLD = 0.6
WAYPOINT_TURNING_THRESHOLD = np.radians(10)
NOT_MOVING_THRESHOLD = 0.005
NONMOVEMENT_DIST_THRESHOLD = 0.05
NONMOVEMENT_TURN_THRESHOLD = np.radians(0.05)
STEP_LIMIT = 10000

# box_delivery/config.yaml restated (reference defaults)
DEFAULT_CFG = {
    "teleop_mode": False, "low_dim_state": False, "fixed_trial_idx": 1, "output_dir": "logs/",
    "render": {"log_obs": False, "show": False, "show_obs": False, "frequency": 20},
    "sim": {"t_max": 2000, "steps": 100, "gravity": (0, 0), "iterations": 10, "damping": 0},
    "controller": {"dt": 0.2, "target_speed": 0.3},
    "boxes": {"num_boxes_small": 10, "num_boxes_large": 20, "box_size": 0.44, "min_box_dist": 0.62, "box_density": 0.001},
    "agent": {
        "action_type": "position", "step_size": 1.75, "random_start": True, "mass": 1, "padding": 0.25,
        "length": 0.6, "width": 0.92,
        "vertices": [[0.26, -0.35], [-0.31, -0.35], [-0.31, 0.35], [0.26, 0.35]],
        "front_bumper_vertices": [[0.31, -0.53], [0.20, -0.39], [0.20, 0.39], [0.31, 0.53]],
        "wheel_vertices": [
            [[0.09, 0.35], [0.09, 0.39], [0.24, 0.39], [0.24, 0.35]],
            [[0.09, -0.39], [0.09, -0.35], [0.24, -0.35], [0.24, -0.39]],
            [[-0.09, 0.35], [-0.09, 0.39], [-0.24, 0.39], [-0.24, 0.35]],
            [[-0.09, -0.35], [-0.09, -0.39], [-0.24, -0.39], [-0.24, -0.35]],
        ],
    },
    "env": {"obstacle_config": "small_empty", "room_length": 10, "room_width_small": 5, "room_width_large": 10,
            "receptacle_width": 1.5, "shortest_path_channel_scale": 0.25, "local_map_pixel_width": 224,
            "local_map_pixel_width_sam": 96, "local_map_width": 10, "wall_thickness": 14, "invert_receptacle_map": False},
    "misc": {"ministep_size": 2.5, "inactivity_cutoff_sam": 100, "inactivity_cutoff": 200, "random_seed": 42},
    "rewards": {"partial_rewards_scale": 0.2, "goal_reward": 10.0, "collision_penalty": 0.25, "non_movement_penalty": 0,
                "correct_direction_reward_scale": 3},
    "rewards_sam": {"partial_rewards_scale": 0.2, "goal_reward": 1.0, "collision_penalty": 0.25, "non_movement_penalty": 0.25,
                    "correct_direction_reward_scale": 1},
    "train": {"job_type": "ppo", "use_correct_direction_reward": False},
}


def R(theta):
    return np.asarray([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])


def restrict_heading_range(heading):
    return np.mod(heading + np.pi, 2 * np.pi) - np.pi


def heading_difference(h1, h2):
    return restrict_heading_range(h1 - h2)


def distance(p1, p2):
    return np.linalg.norm(np.asarray(p1)[:2] - np.asarray(p2)[:2])


# ---------------------------------------------------------------- skimage restatements
def draw_line(r0, c0, r1, c1):
    """skimage.draw.line (skimage/draw/_draw.pyx::_line): Bresenham, both end points included."""
    r0, c0, r1, c1 = int(r0), int(c0), int(r1), int(c1)
    steep = 0
    r, c = r0, c0
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sc = 1 if (c1 - c) > 0 else -1
    sr = 1 if (r1 - r) > 0 else -1
    if dr > dc:
        steep = 1
        c, r = r, c
        dc, dr = dr, dc
        sc, sr = sr, sc
    d = (2 * dr) - dc
    rr = np.zeros(dc + 1, dtype=np.intp)
    cc = np.zeros(dc + 1, dtype=np.intp)
    for i in range(dc):
        if steep:
            rr[i], cc[i] = c, r
        else:
            rr[i], cc[i] = r, c
        while d >= 0:
            r = r + sr
            d = d - (2 * dc)
        c = c + sc
        d = d + (2 * dr)
    rr[dc], cc[dc] = r1, c1
    return rr, cc


def approximate_polygon(coords, tolerance):
    """skimage.measure.approximate_polygon (skimage/measure/_polygon.py): iterative Douglas-Peucker."""
    coords = np.asarray(coords)
    if tolerance <= 0:
        return coords
    chain = np.zeros(coords.shape[0], "bool")
    dists = np.zeros(coords.shape[0])
    chain[0] = True
    chain[-1] = True
    pos_stack = [(0, chain.shape[0] - 1)]
    end_of_chain = False
    while not end_of_chain:
        start, end = pos_stack.pop()
        r0, c0 = coords[start, :]
        r1, c1 = coords[end, :]
        dr = r1 - r0
        dc = c1 - c0
        segment_angle = -np.arctan2(dr, dc)
        segment_dist = c0 * np.sin(segment_angle) + r0 * np.cos(segment_angle)
        segment_coords = coords[start + 1:end, :]
        segment_dists = dists[start + 1:end]
        dr0 = segment_coords[:, 0] - r0
        dc0 = segment_coords[:, 1] - c0
        dr1 = segment_coords[:, 0] - r1
        dc1 = segment_coords[:, 1] - c1
        projected_lengths0 = dr0 * dr + dc0 * dc
        projected_lengths1 = -dr1 * dr - dc1 * dc
        perp = np.logical_and(projected_lengths0 > 0, projected_lengths1 > 0)
        eucl = np.logical_not(perp)
        segment_dists[perp] = np.abs(segment_coords[perp, 0] * np.cos(segment_angle)
                                     + segment_coords[perp, 1] * np.sin(segment_angle) - segment_dist)
        segment_dists[eucl] = np.minimum(np.sqrt(dc0[eucl] ** 2 + dr0[eucl] ** 2), np.sqrt(dc1[eucl] ** 2 + dr1[eucl] ** 2))
        if np.any(segment_dists > tolerance):
            new_end = start + np.argmax(segment_dists) + 1
            pos_stack.append((new_end, end))
            pos_stack.append((start, new_end))
            chain[new_end] = True
        if len(pos_stack) == 0:
            end_of_chain = True
    return coords[chain, :]


def disk(radius):
    """skimage.morphology.disk"""
    radius = int(radius)
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.uint8)


# ---------------------------------------------------------------- scipy / cv2 restatements (pinned by tests)
def rotate_nearest_geometry(in_shape, angle_deg):
    """Output shape, matrix and offset of scipy.ndimage.rotate(x, angle, reshape=True) (ndimage/_interpolation.py)."""
    from .maze_oracle import cosdg, sindg
    c, s = cosdg(float(angle_deg)), sindg(float(angle_deg))
    rot = np.array([[c, s], [-s, c]])
    iy, ix = in_shape
    out_bounds = rot @ [[0, 0, iy, iy], [0, ix, 0, ix]]
    out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
    out_center = rot @ ((out_shape - 1) / 2)
    in_center = (np.array(in_shape) - 1) / 2
    offset = in_center - out_center
    return tuple(int(v) for v in out_shape), rot, offset


def rotate_nearest(inp, angle_deg, cval=0.0):
    """scipy.ndimage.rotate(inp, angle, order=0) (reshape=True, mode='constant'), restated:
    coordinate h of output pixel (i, j) = (offset[h] + i * M[h,0]) + j * M[h,1]  (NI_GeometricTransform);
    outside [-0.5 .. len - 0.5) is not the rule — scipy's constant mode tests  cc < 0 - eps? no:
    for mode 'constant' a coordinate is outside when cc < 0 or cc > len - 1 (map_coordinate), otherwise the
    order-0 sample is the input at floor(cc + 0.5)."""
    inp = np.asarray(inp)
    H, W = inp.shape
    (OH, OW), M, off = rotate_nearest_geometry((H, W), angle_deg)
    I = np.arange(OH, dtype=np.float64)[:, None]
    J = np.arange(OW, dtype=np.float64)[None, :]
    c0 = (off[0] + I * M[0, 0]) + J * M[0, 1]
    c1 = (off[1] + I * M[1, 0]) + J * M[1, 1]
    outside = (c0 < 0) | (c0 > H - 1) | (c1 < 0) | (c1 > W - 1)
    i0 = np.floor(np.where(outside, 0.0, c0) + 0.5).astype(np.int64)
    i1 = np.floor(np.where(outside, 0.0, c1) + 0.5).astype(np.int64)
    out = inp[np.clip(i0, 0, H - 1), np.clip(i1, 0, W - 1)].copy()
    out[outside] = cval
    return out


def fill_poly(img, pts, color):
    """cv2.fillPoly(img, [pts], color) for ONE convex polygon with int32 vertices (line_type 8, shift 0),
    restated from OpenCV's imgproc/src/drawing.cpp:
      CollectPolyEdges — every edge is drawn with the 8-connected ``Line`` and, unless horizontal, enters
        the edge table with x in 16.16 fixed point and dx = (x1 - x0) / (y1 - y0) (int64, truncating);
      FillEdgeCollection — for each row y in [y_min, min(y_max, H)) the two active edges (y0 <= y < y1)
        give the span [ceil(x_left / 2^16), floor(x_right / 2^16)] (as cv2 4.13 in this image does), clipped to the row.
    Pinned against cv2 by tests/test_oracle_boxdel_cpu.py (polygons inside the image; a polygon that leaves
    the image goes through cv::clipLine in OpenCV, which is not restated)."""
    H, W = img.shape
    pts = [(int(q[0]), int(q[1])) for q in pts]
    n = len(pts)
    XY_SHIFT = 16
    edges = []
    p0 = pts[-1]
    for i in range(n):
        p1 = pts[i]
        cv_line(img, p0, p1, color)
        if p0[1] != p1[1]:
            num = (p1[0] - p0[0]) << XY_SHIFT
            den = p1[1] - p0[1]
            q = abs(num) // abs(den)
            dx = q if (num >= 0) == (den > 0) else -q
            if p0[1] < p1[1]:
                edges.append((p0[1], p1[1], p0[0] << XY_SHIFT, dx))
            else:
                edges.append((p1[1], p0[1], p1[0] << XY_SHIFT, dx))
        p0 = p1
    if len(edges) < 2:
        return
    ymin = min(e[0] for e in edges)
    ymax = min(max(e[1] for e in edges), H)
    for y in range(ymin, ymax):
        if y < 0:
            continue
        act = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(act) - 1, 2):
            x1, x2 = (act[k] + (1 << XY_SHIFT) - 1) >> XY_SHIFT, act[k + 1] >> XY_SHIFT
            if x1 < W and x2 >= 0:
                img[y, max(x1, 0):min(x2, W - 1) + 1] = color


def cv_line(img, p0, p1, color):
    """OpenCV Line(img, pt1, pt2, color, 8) = LineIterator(img, pt1, pt2, 8, leftToRight=true) for end points
    inside the image: start from the end with the smaller x (pt1 when equal); err = major - 2 minor; every
    step advances along the major axis and, when err < 0 before the update, also along the minor axis;
    err += -2 minor (+ 2 major when it was negative)."""
    H, W = img.shape
    x0, y0 = p0
    x1, y1 = p1
    dx, dy = x1 - x0, y1 - y0
    sx = sy = 1
    if dx < 0:
        dx, dy = -dx, -dy
        x0, y0 = x1, y1
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    major, minor = (dy, dx) if vert else (dx, dy)
    err = major - (minor + minor)
    x, y = x0, y0
    for _ in range(major + 1):
        if 0 <= x < W and 0 <= y < H:
            img[y, x] = color
        neg = err < 0
        err += -(minor + minor) + ((major + major) if neg else 0)
        if vert:
            y += sy
            if neg:
                x += sx
        else:
            x += sx
            if neg:
                y += sy


# ---------------------------------------------------------------- layouts
def sample_layout(random_state, cfg=None):
    """One episode's robot start, random obstacles and boxes in the order init_box_delivery_env draws them
    (box_delivery_env.py:241-252): get_random_robot_start :313-320, generate_boundary :326-507 (add_random_columns /
    add_random_horiz_divider), generate_boxes :509-546.  Pinned to the reference's own methods by
    tests/golden/ref_bd_layouts.npz (tools/make_ref_bd_layouts.py)."""
    cfg = merge_cfg(DEFAULT_CFG, cfg or {})
    room_length = cfg["env"]["room_length"]
    oc = cfg["env"]["obstacle_config"]
    small = oc.split("_")[0] == "small"
    room_width = cfg["env"]["room_width_small"] if small else cfg["env"]["room_width_large"]
    num_boxes = cfg["boxes"]["num_boxes_small"] if small else cfg["boxes"]["num_boxes_large"]
    ag = cfg["agent"]

    def random_start():
        size = max(ag["length"], ag["width"])
        x = random_state.uniform(-room_length / 2 + size, room_length / 2 - size)
        y = random_state.uniform(-room_width / 2 + size, room_width / 2 - size)
        h = random_state.uniform(0, 2 * np.pi)
        return (x, y, h)

    start = random_start() if ag["random_start"] else (5, 1.5, np.pi * 3 / 2)
    rsize = cfg["env"]["receptacle_width"]
    rpos = (room_length / 2 - rsize / 2, room_width / 2 - rsize / 2)
    robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
    columns, divider = [], None
    if oc in ("small_columns", "large_columns"):
        # add_random_columns: randint(1, max) columns, up to 100 tries each; hitting the receptacle / robot abandons the column
        num_columns = random_state.randint(1, 3 if small else 8)
        buffer_width, col_min_dist = 0.8, 2
        for _ in range(num_columns):
            for _ in range(100):
                x = random_state.uniform(-room_length / 2 + 2 * buffer_width + 1 / 2, room_length / 2 - 2 * buffer_width - 1 / 2)
                y = random_state.uniform(-room_width / 2 + 2 * buffer_width + 1 / 2, room_width / 2 - 2 * buffer_width - 1 / 2)
                overlapped = False
                if ((x - rpos[0]) ** 2 + (y - rpos[1]) ** 2) ** 0.5 <= col_min_dist / 2 + rsize / 2:
                    break
                if ((x - start[0]) ** 2 + (y - start[1]) ** 2) ** 0.5 <= col_min_dist / 2 + robot_radius:
                    break
                for pc in columns:
                    if ((x - pc[0]) ** 2 + (y - pc[1]) ** 2) ** 0.5 <= col_min_dist:
                        overlapped = True
                        break
                if not overlapped:
                    columns.append([x, y])
                    break
    elif oc == "large_divider":
        # add_random_horiz_divider: 100 x 100 tries; the robot start is redrawn after every 100 failures
        for _ in range(100):
            for _ in range(100):
                x = room_length / 2 - 8 / 2
                y = random_state.uniform(-room_width / 2 + 3.5 + 0.5 / 2, room_width / 2 - 3.5 - 0.5 / 2)
                if not ((x - start[0]) ** 2 + (y - start[1]) ** 2) ** 0.5 <= 3 * robot_radius:
                    divider = [x, y]
                    break
            if divider is not None:
                break
            start = random_start()
    box_size = cfg["boxes"]["box_size"] / 2
    dmin = cfg["boxes"]["min_box_dist"]
    min_x, max_x = -room_length / 2 + box_size, room_length / 2 - box_size
    min_y, max_y = -room_width / 2 + box_size, room_width / 2 - box_size
    boxes = []
    while len(boxes) < num_boxes:
        cx = random_state.uniform(min_x, max_x)
        cy = random_state.uniform(min_y, max_y)
        hd = random_state.uniform(0, 2 * np.pi)
        overlapped = False
        # boundary_dicts that are neither wall nor corner: the receptacle ('width' = size), columns ('width' = 1), the divider (y only)
        for (ox, oy), width in [(rpos, rsize)] + [((c[0], c[1]), 1) for c in columns]:
            if ((cx - ox) ** 2 + (cy - oy) ** 2) ** 0.5 <= (dmin / 2 + width / 2) * 1.2:
                overlapped = True
                break
        if not overlapped and divider is not None and abs(cy - divider[1]) <= (dmin / 2 + 0.5 / 2) * 1.2:
            overlapped = True
        if not overlapped:
            for px, py, _ in boxes:
                if ((cx - px) ** 2 + (cy - py) ** 2) ** 0.5 <= dmin:
                    overlapped = True
                    break
        if not overlapped:
            boxes.append([cx, cy, hd])
    return {"start": tuple(float(v) for v in start), "boxes": [[float(a) for a in b] for b in boxes],
            "columns": [[float(a) for a in c] for c in columns], "divider": None if divider is None else [float(a) for a in divider]}


# ---------------------------------------------------------------- the environment
class BoxDeliveryOracleEnv:
    def __init__(self, cfg=None, layouts=None, order_mode=1):
        self.cfg = merge_cfg(DEFAULT_CFG, cfg or {})
        c = self.cfg
        assert c["env"]["obstacle_config"] in ("small_empty", "large_empty", "small_columns", "large_columns", "large_divider")
        self.order_mode = order_mode
        self.layouts = layouts
        self.local_map_pixel_width = c["env"]["local_map_pixel_width"] if c["train"]["job_type"] != "sam" else c["env"]["local_map_pixel_width_sam"]
        self.local_map_width = c["env"]["local_map_width"]
        self.local_map_pixels_per_meter = self.local_map_pixel_width / self.local_map_width
        self.room_length = c["env"]["room_length"]
        self.wall_thickness = c["env"]["wall_thickness"]
        small = c["env"]["obstacle_config"].split("_")[0] == "small"
        self.num_boxes = c["boxes"]["num_boxes_small"] if small else c["boxes"]["num_boxes_large"]
        self.room_width = c["env"]["room_width_small"] if small else c["env"]["room_width_large"]
        self.num_channels = 4
        ag = c["agent"]
        self.robot_radius = ((ag["length"] ** 2 + ag["width"] ** 2) ** 0.5 / 2) * 1.2
        self.robot_half_width = max(ag["length"], ag["width"]) / 2
        W = self.local_map_pixel_width
        robot_pixel_width = int(2 * self.robot_radius * self.local_map_pixels_per_meter)
        self.robot_state_channel = np.zeros((W, W), dtype=np.float32)
        start = int(np.floor(W / 2 - robot_pixel_width / 2))
        for i in range(start, start + robot_pixel_width):
            for j in range(start, start + robot_pixel_width):
                if (((i + 0.5) - W / 2) ** 2 + ((j + 0.5) - W / 2) ** 2) ** 0.5 < robot_pixel_width / 2:
                    self.robot_state_channel[i, j] = 1
        rw = c["rewards_sam"] if c["train"]["job_type"] == "sam" else c["rewards"]
        self.partial_rewards_scale = rw["partial_rewards_scale"]
        self.goal_reward = rw["goal_reward"]
        self.collision_penalty = rw["collision_penalty"]
        self.non_movement_penalty = rw["non_movement_penalty"]
        self.correct_direction_reward_scale = rw["correct_direction_reward_scale"]
        self.ministep_size = c["misc"]["ministep_size"]
        self.inactivity_cutoff = c["misc"]["inactivity_cutoff"] if c["train"]["job_type"] != "sam" else c["misc"]["inactivity_cutoff_sam"]
        self.episode_idx = None
        self.path = None
        self.dp = None
        # knobs the area-clearing restatement overrides (same control chain, different constants)
        self.bound_x_half = self.room_length / 2      # PositionController clamps x with map_height, y with map_width (D17)
        self.bound_y_half = self.room_width / 2
        self.lfc = 0.0                                # cfg.controller.Lfc (absent in box-delivery => 0)
        self.omega_gain, self.v_gain = 3, 2           # apply_controller (box_delivery_env.py:887-889)
        self.step_limit = STEP_LIMIT

    # ---- sim construction (box_delivery_env.py:185-292; sim_utils.py:20-165)
    def _build_space(self, layout):
        c = self.cfg
        self.steps = c["sim"]["steps"]
        self.dt = c["controller"]["dt"]
        self.target_speed = c["controller"]["target_speed"]
        sp = cpo.Space(iterations=c["sim"]["iterations"], damping=c["sim"]["damping"], gravity=c["sim"]["gravity"])
        sp.set_order(self.order_mode)
        sp.set_handler_mode(1)
        self.space = sp
        self.total_work = [0, []]
        size = c["env"]["receptacle_width"]
        self.receptacle_position = (self.room_length / 2 - size / 2, self.room_width / 2 - size / 2)
        self.receptacle_size = size
        self.goal_points = [self.receptacle_position]
        self.start = tuple(layout["start"])
        # robot: kinematic body, chassis recentred on its hull centroid (D8), wheels r=.02 friction 0, bumper r=.02
        ag = c["agent"]
        x, y, h = self.start
        rb = sp.add_body(cpo.KINEMATIC, x, y, h)
        self.robot_body = rb
        verts = np.asarray(ag["vertices"], dtype=np.float64)
        cg = cpo.poly_centroid_hull(verts)
        vs = np.array([[vx - cg[0], vy - cg[1]] for vx, vy in verts])
        self.robot_shape = sp.add_poly(rb, vs, 0.0)
        sp.set_mass(self.robot_shape, 10.0)
        sp.set_material(self.robot_shape, 0.01, 1.0, 1)
        for wv in ag["wheel_vertices"]:
            s = sp.add_poly(rb, np.asarray(wv, dtype=np.float64), 0.02)
            sp.set_mass(s, 1)
            sp.set_material(s, 0.01, 0.0, 0)
        s = sp.add_poly(rb, np.asarray(ag["front_bumper_vertices"], dtype=np.float64), 0.02)
        sp.set_mass(s, 1)
        sp.set_material(s, 0.01, 0.0, 0)
        # boxes (generate_boxes vertices :548-556; create_polygon ignores radius => 0.02, D7)
        bs = c["boxes"]["box_size"] / 2
        self.box_bodies, self.box_shapes, self.box_idx = [], [], []
        for i, (bx, by, bh) in enumerate(layout["boxes"]):
            vertices = np.array([[bx + bs, by + bs], [bx - bs, by + bs], [bx - bs, by - bs], [bx + bs, by - bs]])
            local = (vertices - np.array([bx, by])).tolist()
            b = sp.add_body(cpo.DYNAMIC, bx, by, bh)
            cg = cpo.poly_centroid_hull(np.asarray(local))
            vloc = np.array([[vx - cg[0], vy - cg[1]] for vx, vy in local])
            s = sp.add_poly(b, vloc, 0.02)
            sp.set_density(s, c["boxes"]["box_density"])
            sp.set_material(s, 0.01, 1.0, 2)
            self.box_bodies.append(b)
            self.box_shapes.append(s)
            self.box_idx.append(i)
        # boundaries: receptacle (type 4), 4 walls (type 3), 3 corners x 3 triangles r=.02 (type 3)
        self.boundary_shapes = []
        (rx, ry) = self.receptacle_position
        rect = lambda x, y, l, w: np.array([[x - l / 2, y - w / 2], [x + l / 2, y - w / 2], [x + l / 2, y + w / 2], [x - l / 2, y + w / 2]])
        b = sp.add_body(cpo.STATIC, 0, 0, 0)
        s = sp.add_poly(b, rect(rx, ry, size, size), 0.0)
        sp.set_material(s, 0.01, 1.0, 4)
        self.receptacle_shape, self.receptacle_body = s, b
        rl, rwid, wt = self.room_length, self.room_width, self.wall_thickness
        for (x, y, length, width) in [(-rl / 2 - wt / 2, 0, wt, rwid), (rl / 2 + wt / 2, 0, wt, rwid),
                                      (0, -rwid / 2 - wt / 2, rl + 2 * wt, wt), (0, rwid / 2 + wt / 2, rl + 2 * wt, wt)]:
            b = sp.add_body(cpo.STATIC, 0, 0, 0)
            s = sp.add_poly(b, rect(x, y, length, width), 0.0)
            sp.set_material(s, 0.01, 1.0, 3)
            self.boundary_shapes.append((b, s, "wall"))
        # random obstacles of the layout (generate_boundary :358-471; create_static: no radius): unit columns, or the 8 x 0.5 divider
        for (x, y) in layout.get("columns") or []:
            b = sp.add_body(cpo.STATIC, 0, 0, 0)
            s = sp.add_poly(b, rect(x, y, 1, 1), 0.0)
            sp.set_material(s, 0.01, 1.0, 3)
            self.boundary_shapes.append((b, s, "column"))
        divider = layout.get("divider")
        if divider is not None:
            b = sp.add_body(cpo.STATIC, 0, 0, 0)
            s = sp.add_poly(b, rect(divider[0], divider[1], 8, 0.5), 0.0)
            sp.set_material(s, 0.01, 1.0, 3)
            self.boundary_shapes.append((b, s, "divider"))
        s225, c225 = np.sin(22.5 * np.pi / 180), np.cos(22.5 * np.pi / 180)
        vs1 = [(0, 0), (0, -1), (0.5 * s225, -1 + 0.5 * c225)]
        vs2 = [(0, 0), (0.5 * s225, -1 + 0.5 * c225), (1 - 0.5 * c225, -0.5 * s225)]
        vs3 = [(0, 0), (1, 0), (1 - 0.5 * c225, -0.5 * s225)]
        corners = []
        for i, (x, y) in enumerate([(-rl / 2, rwid / 2), (rl / 2, rwid / 2), (rl / 2, -rwid / 2), (-rl / 2, -rwid / 2)]):
            if i == 1:
                continue
            corners.append((x, y, -np.radians(i * 90)))
        if divider is not None:             # corners where the divider meets the right wall (:491-505)
            corners.append((rl / 2, divider[1] - 0.5 / 2, np.radians(-90)))
            corners.append((rl / 2, divider[1] + 0.5 / 2, np.radians(180)))
        for (x, y, heading) in corners:
            for vs in (vs1, vs2, vs3):
                b = sp.add_body(cpo.STATIC, x, y, heading)
                s = sp.add_poly(b, np.asarray(vs, dtype=np.float64), 0.02)
                sp.set_material(s, 0.01, 1.0, 3)
                self.boundary_shapes.append((b, s, "corner"))
        self.box_local = {s: sp.get_vertices(s) for s in self.box_shapes}
        self.robot_local = sp.get_vertices(self.robot_shape)

    # ---- maps (box_delivery_env.py:1098-1207)
    def create_padded_room(self, value):
        ppm, W = self.local_map_pixels_per_meter, self.local_map_pixel_width
        shape = (int(2 * np.ceil((self.room_width * ppm + W * np.sqrt(2)) / 2)), int(2 * np.ceil((self.room_length * ppm + W * np.sqrt(2)) / 2)))
        return np.full(shape, value, dtype=np.float32)

    def _world_verts(self, body, local):
        return np.array([self.space.local_to_world(body, v[0], v[1]) for v in local])

    def _verts_px(self, vertices_np, H):
        ppm = self.local_map_pixels_per_meter
        px = (vertices_np * ppm).astype(np.int32)
        off = int(self.local_map_width * ppm / 2) + 10
        px[:, 0] += off
        px[:, 1] += off
        px[:, 1] = H - px[:, 1]
        return px

    def update_configuration_space(self):
        W = self.local_map_pixel_width
        ppm = self.local_map_pixels_per_meter
        obstacle_map = self.create_padded_room(1.0)
        small = np.zeros((W + 20, W + 20), dtype=np.float32)
        for (b, s, label) in self.boundary_shapes:
            vertices = self._world_verts(b, self.space.get_vertices(s))
            cv2.fillPoly(small, [self._verts_px(vertices, small.shape[0])], color=1)
        si, sj = int(obstacle_map.shape[0] / 2 - small.shape[0] / 2), int(obstacle_map.shape[1] / 2 - small.shape[1] / 2)
        obstacle_map[si:si + small.shape[0], sj:sj + small.shape[1]] = small
        selem = disk(np.floor(self.robot_radius * ppm))
        self.configuration_space = 1 - binary_dilation(obstacle_map, selem).astype(np.float32)
        selem_thin = disk(np.floor(self.robot_half_width * ppm))
        self.configuration_space_thin = 1 - binary_dilation(obstacle_map, selem_thin).astype(np.float32)
        self.closest_cspace_indices = distance_transform_edt(1 - self.configuration_space, return_distances=False, return_indices=True)
        self.small_obstacle_map = 1 - small
        self.small_start = (si, sj)

    def update_global_overhead_map(self):
        small = self.small_obstacle_map.copy()
        H = small.shape[0]
        polys = [(self.receptacle_body, self.space.get_vertices(self.receptacle_shape), RECEPTACLE_SEG_INDEX)]
        polys += [(b, self.box_local[s], BOX_SEG_INDEX) for b, s in zip(self.box_bodies, self.box_shapes)]
        polys += [(self.robot_body, self.robot_local, ROBOT_SEG_INDEX)]
        for body, local, seg in polys:
            px = self._verts_px(self._world_verts(body, local), H)
            small[small == 1] = FLOOR_SEG_INDEX / MAX_SEG_INDEX
            cv2.fillPoly(small, [px], color=seg / MAX_SEG_INDEX)
        si, sj = self.small_start
        self.global_overhead_map[si:si + H, sj:sj + small.shape[1]] = small

    def position_to_pixel_indices(self, x, y, shape):
        ppm = self.local_map_pixels_per_meter
        i = np.floor(shape[0] / 2 - y * ppm).astype(np.int32)
        j = np.floor(shape[1] / 2 + x * ppm).astype(np.int32)
        return np.clip(i, 0, shape[0] - 1), np.clip(j, 0, shape[1] - 1)

    def pixel_indices_to_position(self, i, j, shape):
        ppm = self.local_map_pixels_per_meter
        return (j - shape[1] / 2) / ppm, (shape[0] / 2 - i) / ppm

    def closest_valid_cspace_indices(self, i, j):
        return self.closest_cspace_indices[:, i, j]

    def get_local_map(self, global_map, robot_position, robot_heading):
        W = self.local_map_pixel_width
        ppm = self.local_map_pixels_per_meter
        crop_width = int(np.ceil(W * np.sqrt(2) / 2) * 2)
        rotation_angle = 90 - np.degrees(robot_heading)
        pi = int(np.floor(-robot_position[1] * ppm + global_map.shape[0] / 2))
        pj = int(np.floor(robot_position[0] * ppm + global_map.shape[1] / 2))
        crop = global_map[pi - crop_width // 2:pi + crop_width // 2, pj - crop_width // 2:pj + crop_width // 2]
        rot = rotate_image(crop, rotation_angle, order=0)
        return rot[rot.shape[0] // 2 - W // 2:rot.shape[0] // 2 + W // 2, rot.shape[1] // 2 - W // 2:rot.shape[1] // 2 + W // 2]

    def get_local_distance_map(self, global_map, robot_position, robot_heading):
        local_map = self.get_local_map(global_map, robot_position, robot_heading)
        local_map -= local_map.min()
        return local_map

    def _scale_sp(self, m):
        ppm, W = self.local_map_pixels_per_meter, self.local_map_pixel_width
        m /= ppm
        m /= (np.sqrt(2) * W) / ppm
        m *= self.cfg["env"]["shortest_path_channel_scale"]
        return m

    def create_global_shortest_path_to_receptacle_map(self):
        global_map = self.create_padded_room(0.0) + np.inf
        rx, ry = self.receptacle_position
        i, j = self.position_to_pixel_indices(rx, ry, self.configuration_space.shape)
        i, j = self.closest_valid_cspace_indices(i, j)
        img, _ = cpo.spfa(self.configuration_space, (i, j))
        img /= self.local_map_pixels_per_meter
        global_map = np.minimum(global_map, img)
        ppm, W = self.local_map_pixels_per_meter, self.local_map_pixel_width
        global_map /= (np.sqrt(2) * W) / ppm
        global_map *= self.cfg["env"]["shortest_path_channel_scale"]
        return global_map

    def create_global_shortest_path_map(self, robot_position):
        i, j = self.position_to_pixel_indices(robot_position[0], robot_position[1], self.configuration_space.shape)
        i, j = self.closest_valid_cspace_indices(i, j)
        m, _ = cpo.spfa(self.configuration_space, (i, j))
        return self._scale_sp(m)

    def robot_pose(self):
        b = self.space.body(self.robot_body)
        return [float(b[0]), float(b[1])], float(b[2])

    def generate_observation(self, done=False):
        self.update_global_overhead_map()
        if done and self.cfg["agent"]["action_type"] == "position":
            return None
        pos, ang = self.robot_pose()
        ch = [self.get_local_map(self.global_overhead_map, pos, ang), self.robot_state_channel,
              self.get_local_distance_map(self.create_global_shortest_path_map(pos), pos, ang),
              self.get_local_distance_map(self.create_global_shortest_path_to_receptacle_map(), pos, ang)]
        observation = np.stack(ch, axis=2)
        return (observation * 255).astype(np.uint8)

    # ---- paths (box_delivery_env.py:1209-1268 == position_controller.py:126-181)
    def shortest_path(self, source_position, target_position, check_straight=False):
        cs = self.configuration_space
        si, sj = self.position_to_pixel_indices(source_position[0], source_position[1], cs.shape)
        ti, tj = self.position_to_pixel_indices(target_position[0], target_position[1], cs.shape)
        if check_straight:
            rr, cc = draw_line(si, sj, ti, tj)
            if (1 - self.configuration_space_thin[rr, cc]).sum() == 0:
                return [source_position, target_position]
        si, sj = self.closest_valid_cspace_indices(si, sj)
        ti, tj = self.closest_valid_cspace_indices(ti, tj)
        _, parents = cpo.spfa(cs, (si, sj))
        Wd = parents.shape[1]
        i, j = int(ti), int(tj)
        coords = [[i, j]]
        while not (i == si and j == sj):
            p = parents[i, j]
            if p < 0:
                break            # parents_ij = [-1, -1] => i + j < 0
            i, j = int(p // Wd), int(p % Wd)
            coords.append([i, j])
        coords = approximate_polygon(np.asarray(coords), tolerance=1)
        new_coords = [coords[0]]
        for k in range(1, len(coords) - 1):
            rr, cc = draw_line(*new_coords[-1], *coords[k + 1])
            if (1 - cs[rr, cc]).sum() > 0:
                new_coords.append(coords[k])
        if len(coords) > 1:
            new_coords.append(coords[-1])
        coords = new_coords
        path = []
        for coord in coords[::-1]:
            px, py = self.pixel_indices_to_position(coord[0], coord[1], cs.shape)
            path.append([px, py])
        if len(path) < 2:
            path = [source_position, target_position]
        else:
            path[0] = source_position
            path[-1] = target_position
        return path

    def shortest_path_distance(self, source_position, target_position):
        path = self.shortest_path(source_position, target_position)
        return sum(distance(path[i - 1], path[i]) for i in range(1, len(path)))

    def get_waypoints_to_spatial_action(self, p0, h0, spatial_action):
        """position_controller.py:56-124 (map_width = room_width, map_height = room_length: D17)"""
        W, ppm = self.local_map_pixel_width, self.local_map_pixels_per_meter
        robot_action = np.unravel_index(spatial_action, (W, W))
        x_movement = -self.local_map_width / 2 + float(robot_action[1]) / ppm
        y_movement = self.local_map_width / 2 - float(robot_action[0]) / ppm
        straight_line_dist = np.sqrt(x_movement ** 2 + y_movement ** 2)
        turn_angle = np.arctan2(-x_movement, y_movement)
        straight_line_heading = restrict_heading_range(h0 + turn_angle)
        tgt = [p0[0] + straight_line_dist * np.cos(straight_line_heading), p0[1] + straight_line_dist * np.sin(straight_line_heading)]
        diff = np.asarray(tgt) - np.asarray(p0)
        ratio_x, ratio_y = (1, 1)
        bound_x = np.sign(tgt[0]) * (2 * self.bound_x_half) / 2
        bound_y = np.sign(tgt[1]) * (2 * self.bound_y_half) / 2
        if abs(tgt[0]) > abs(bound_x):
            ratio_x = (bound_x - p0[0]) / (tgt[0] - p0[0])
        if abs(tgt[1]) > abs(bound_y):
            ratio_y = (bound_y - p0[1]) / (tgt[1] - p0[1])
        ratio = min(ratio_x, ratio_y)
        tgt = (np.asarray(p0) + ratio * diff).tolist()
        wp = self.shortest_path(p0, tgt, check_straight=True)
        wp = [list(p) for p in wp]
        hd = [None]
        for i in range(1, len(wp)):
            hd.append(restrict_heading_range(np.arctan2(wp[i][1] - wp[i - 1][1], wp[i][0] - wp[i - 1][0])))
        d_end = distance(wp[-2], wp[-1])
        signed_dist = d_end - self.robot_radius
        move_sign = np.sign(signed_dist)
        if len(wp) > 2 and signed_dist < 0:
            wp[-2] = wp[-1]
            hd[-2] = restrict_heading_range(np.arctan2(wp[-2][1] - wp[-3][1], wp[-2][0] - wp[-3][0]))
            move_sign = 1
        path = np.array([[p[0], p[1], h] for p, h in zip(wp, hd)], dtype=object)
        return path, move_sign

    # ---- controller (box_delivery_env.py:867-889; dp.py:100-128,217-248)
    def controller(self, pos, heading):
        x, y, h = pos[0], pos[1], heading
        if self.dp is None:
            # DP(x, y, yaw, cx, cy, ch = first two waypoints): TargetCourse with Lfc = 0; setpoint = nearest course point
            cx = [self.path[0][0], self.path[1][0]]
            cy = [self.path[0][1], self.path[1][1]]
            chh = [self.path[0][2], self.path[1][2]]
            d = np.hypot([x - cx[0], x - cx[1]], [y - cy[0], y - cy[1]])
            ind = int(np.argmin(d))
            # TargetCourse.init_setpoint (dp.py:74-84): look-ahead while Lfc exceeds the distance to the course point
            while self.lfc > ((x - cx[ind]) ** 2 + (y - cy[ind]) ** 2) ** 0.5:
                if (ind + 1) >= len(cx):
                    break
                ind += 1
            # metrics.py:176-181 path_length(cumsum=True) of TWO points is the one-element array [d] (no leading 0), so
            # setpoint_al starts at d whatever `ind` is, and the first advance() already selects course point 1
            path_length = (np.sum(np.diff(np.asarray([cx, cy]).T, axis=0) ** 2, axis=1) ** 0.5).cumsum()
            self.dp = {"cx": cx, "cy": cy, "ch": chh, "pl": path_length,
                       "al": path_length[min(len(path_length) - 1, ind)], "setpoint": [cx[ind], cy[ind], chh[ind]]}
        dp = self.dp
        x_d, y_d = dp["setpoint"][0], dp["setpoint"][1]
        theta_d = np.arctan2(y_d - y, x_d - x)
        theta_e = theta_d - h
        theta_e = np.arctan2(np.sin(theta_e), np.cos(theta_e))
        omega = 1.0 * theta_e / self.dt
        v = R(h) @ [self.target_speed, 0]
        # get_setpoint -> TargetCourse.advance(target_speed, dt)
        dp["al"] = dp["al"] + self.target_speed * self.dt
        ind = int(np.sum(dp["pl"] < dp["al"]))
        dp["setpoint"] = [dp["cx"][ind], dp["cy"][ind], dp["ch"][ind]]
        return omega, v

    def apply_controller(self, omega, v):
        self.space.set_angular_velocity(self.robot_body, omega * 3 if self.omega_gain == 3 else omega / 2)
        vv = (v * self.v_gain).tolist()
        self.space.set_velocity(self.robot_body, vv[0], vv[1])

    def execute_robot_path(self, p_init, h_init, move_sign):
        pos = list(p_init)
        heading = h_init
        moving = True
        robot_distance = 0
        wi = 1
        wps = [(w[0], w[1]) for w in self.path]
        whs = [w[2] for w in self.path]
        prev_wp = wps[wi - 1]
        wp = wps[wi]
        wh = whs[wi]
        sim_steps = 0
        done_turning = False
        prev_heading_diff = 0
        while True:
            if not moving:
                break
            prev_pos = list(pos)
            prev_heading = heading
            heading_diff = heading_difference(heading, wh)
            if np.abs(heading_diff) > TURN_STEP_SIZE and np.abs(heading_diff - prev_heading_diff) > 0.001:
                pass
            else:
                done_turning = True
            omega, v = self.controller(prev_pos, prev_heading)
            if not done_turning:
                self.apply_controller(omega, v * 0)
            else:
                self.apply_controller(omega, v)
            self.space.step(self.dt / self.steps)
            self.n_substeps += 1
            pos, ang = self.robot_pose()
            heading = restrict_heading_range(ang)
            prev_heading_diff = heading_diff
            if distance(prev_wp, pos) > MOVE_STEP_SIZE:
                if self.space.robot_hit() == 1:
                    moving = False
                    break
            if (distance(pos, wps[wi]) < WAYPOINT_MOVING_THRESHOLD and np.abs(heading - whs[wi]) < WAYPOINT_TURNING_THRESHOLD):
                robot_distance += distance(prev_wp, pos)
                if wi == len(wps) - 1:
                    moving = False
                else:
                    wi += 1
                    prev_wp = wps[wi - 1]
                    wp = wps[wi]
                    wh = whs[wi]
                    done_turning = False
                    self.dp = None
                    self.path = self.path[1:]
            sim_steps += 1
            if sim_steps > self.step_limit:
                break
        _, ang = self.robot_pose()
        heading = restrict_heading_range(ang)
        return robot_distance, heading_difference(h_init, heading)

    def box_world_vertices(self, k):
        return self._world_verts(self.box_bodies[k], self.box_local[self.box_shapes[k]])

    def step_simulation_until_still(self):
        prev_positions = []
        sim_steps = 0
        done = False
        cs = self.configuration_space
        while not done:
            for k in range(len(self.box_bodies)):
                verts = self.box_world_vertices(k)
                if any(self.space.point_query_types(v, 0) & (1 << 3) for v in verts):
                    b = self.space.body(self.box_bodies[k])
                    i, j = self.position_to_pixel_indices(b[0], b[1], cs.shape)
                    ni, nj = self.closest_valid_cspace_indices(i, j)
                    nx, ny = self.pixel_indices_to_position(ni, nj, cs.shape)
                    self.space.set_position(self.box_bodies[k], nx, ny)
                    self.space.set_velocity(self.box_bodies[k], 0, 0)
            positions = [self.space.body(b)[:2].copy() for b in self.box_bodies + [self.robot_body]]
            if len(prev_positions) > 0:
                done = True
                for i, position in enumerate(positions):
                    if np.linalg.norm(np.asarray(prev_positions[i]) - np.asarray(position)) > NOT_MOVING_THRESHOLD:
                        done = False
                        break
            prev_positions = positions
            self.space.step(self.dt / self.steps)
            self.n_substeps += 1
            sim_steps += 1
            if sim_steps > STEP_LIMIT:
                break

    def get_obs_from_poly(self):
        out = []
        for b, s in zip(self.box_bodies, self.box_shapes):
            st = self.space.body(b)
            out.append(self.box_local[s] @ R(st[2]).T + st[:2])
        return out

    # ---- gym API
    def reset(self, seed=None, options=None):
        self.episode_idx = 0 if self.episode_idx is None else self.episode_idx + 1
        layout = self.layouts[self.episode_idx % len(self.layouts)]
        self._build_space(layout)
        self.update_configuration_space()
        self.box_clearance_statuses = [False for _ in self.box_bodies]
        self.n_substeps = 0
        for _ in range(1000):
            self.space.step(self.dt / self.steps)
        self.prev_boxes = self.get_obs_from_poly()
        self.global_overhead_map = self.create_padded_room(0.0)
        self.update_global_overhead_map()
        self.t = 0
        self.inactivity_counter = 0
        self.robot_cumulative_distance = 0
        self.robot_cumulative_boxes = 0
        self.robot_cumulative_reward = 0
        updated_boxes = self.get_obs_from_poly()
        self.observation = self.generate_observation()
        return self.observation, self._info(updated_boxes, 0)

    def _info(self, updated_boxes, ministeps):
        b = self.space.body(self.robot_body)
        return {"state": (round(float(b[0]), 2), round(float(b[1]), 2), round(float(b[2]), 2)),
                "cumulative_distance": self.robot_cumulative_distance, "cumulative_boxes": self.robot_cumulative_boxes,
                "cumulative_reward": self.robot_cumulative_reward, "total_work": self.total_work[0], "obs": updated_boxes,
                "box_completed_statuses": self.box_clearance_statuses, "goal_positions": self.goal_points,
                "ministeps": ministeps, "inactivity": self.inactivity_counter}

    def heading_to_pixel_action(self, action):
        W, ppm = self.local_map_pixel_width, self.local_map_pixels_per_meter
        angle = (action + 1) * np.pi + np.pi / 2
        step_size = self.cfg["agent"]["step_size"]
        x_movement = step_size * np.cos(angle)
        y_movement = step_size * np.sin(angle)
        x_pixel = int(W / 2 + x_movement * ppm)
        y_pixel = int(W / 2 - y_movement * ppm)
        return y_pixel * W + x_pixel

    def step(self, action, want_obs=True):
        self.t += 1
        self.dp = None
        self.space.set_robot_hit(0)
        robot_boxes = 0
        robot_reward = 0
        p_init, ang = self.robot_pose()
        h_init = restrict_heading_range(ang)
        initial = {}
        for k, b in enumerate(self.box_bodies):
            initial[self.box_idx[k]] = self.shortest_path_distance(list(self.space.body(b)[:2]), self.receptacle_position)
        at = self.cfg["agent"]["action_type"]
        assert at in ("position", "heading", "velocity")
        if at == "velocity":
            # box_delivery_env.py:667-700; the policy's action is a float32 array (action_space dtype)
            action = np.asarray(action, dtype=np.float32)
            linear_speed, angular_speed = action[0], action[1]
            if abs(linear_speed) >= self.target_speed:
                linear_speed = self.target_speed * np.sign(linear_speed)
            self.space.set_angular_velocity(self.robot_body, float(angular_speed))
            for _ in range(self.steps):
                _, ang_now = self.robot_pose()
                gv = R(ang_now) @ [linear_speed, 0]
                self.space.set_velocity(self.robot_body, float(gv[0]), float(gv[1]))
                self.space.step(self.dt / self.steps)
                self.n_substeps += 1
                if self.space.robot_hit() == 1:
                    break
            pos, _ = self.robot_pose()
            robot_distance = distance(p_init, pos)
            self.last_path = []
        else:
            if at == "heading":
                action = self.heading_to_pixel_action(float(np.asarray(action).reshape(-1)[0]))
            self.path, move_sign = self.get_waypoints_to_spatial_action(p_init, h_init, int(action))
            self.last_path = [list(p) for p in self.path]
            robot_distance, robot_turn_angle = self.execute_robot_path(p_init, h_init, move_sign)
        self.step_simulation_until_still()
        final = {}
        for k, b in enumerate(self.box_bodies):
            final[self.box_idx[k]] = self.shortest_path_distance(list(self.space.body(b)[:2]), self.receptacle_position)
        to_remove = []
        for k in range(len(self.box_bodies)):
            idx = self.box_idx[k]
            dist_moved = initial[idx] - final[idx]
            if self.cfg["train"]["use_correct_direction_reward"] and dist_moved > 0:
                dist_moved *= self.correct_direction_reward_scale
            robot_reward += self.partial_rewards_scale * dist_moved
            verts = self.box_world_vertices(k)
            if all(self.space.point_query_types(v, 0) & (1 << 4) for v in verts):
                to_remove.append(k)
                self.box_clearance_statuses[idx] = True
                self.inactivity_counter = 0
                robot_boxes += 1
                robot_reward += self.goal_reward
        for k in sorted(to_remove, reverse=True):
            self.space.remove_body(self.box_bodies[k])
            del self.box_bodies[k], self.box_shapes[k], self.box_idx[k]
        if self.space.robot_hit() == 1:
            robot_reward -= self.collision_penalty
        _, ang = self.robot_pose()
        robot_heading = restrict_heading_range(ang)
        robot_turn_angle = heading_difference(h_init, robot_heading)
        if robot_distance < NONMOVEMENT_DIST_THRESHOLD and abs(robot_turn_angle) < NONMOVEMENT_TURN_THRESHOLD:
            robot_reward -= self.non_movement_penalty
        self.robot_cumulative_distance += robot_distance
        self.robot_cumulative_boxes += robot_boxes
        self.robot_cumulative_reward += robot_reward
        updated_boxes = self.get_obs_from_poly()
        work = total_work_done(self.prev_boxes, updated_boxes)      # zip truncates / misaligns after removals (D10)
        self.total_work[0] += work
        self.total_work[1].append(work)
        self.prev_boxes = updated_boxes
        if robot_boxes == 0:
            self.inactivity_counter += 1
        terminated = self.robot_cumulative_boxes == self.num_boxes
        truncated = False
        if self.inactivity_counter >= self.inactivity_cutoff:
            terminated = True
            truncated = True
        self.observation = self.generate_observation(done=terminated) if want_obs else None
        ministeps = robot_distance / self.ministep_size
        return self.observation, robot_reward, terminated, truncated, self._info(updated_boxes, ministeps)
