"""ctypes binding of oracle/libbp_oracle.so (ORACLE — test infrastructure only).

Mirrors the slice of the pymunk API the reference environments call
(reference call sites: ship_ice_env.py:117-120,218-219,280-281; sim_utils.py:136-153; ship.py:77-98).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DYNAMIC, KINEMATIC, STATIC = 0, 1, 2


def set_portable_trig(on):
    """Process-wide switch: libm cos/sin (default, what Chipmunk calls) or the portable routine
    (bit-reproducible everywhere; used for whole-episode comparisons).  See chipmunk_oracle.cpp."""
    lib().cpo_set_portable_trig(1 if on else 0)


def portable_trig():
    return bool(lib().cpo_get_portable_trig())


def sincos(a):
    """(sin a, cos a) under the current trig mode."""
    s, c = C.c_double(), C.c_double()
    lib().cpo_sincos(C.c_double(float(a)), C.byref(s), C.byref(c))
    return s.value, c.value


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("chipmunk_oracle.cpp", "spfa_oracle.cpp")]
    if os.environ.get("BP_ORACLE_F32") == "1":
        # precision experiment only (tools/fp32_experiment.py): the same restatement with cpFloat = float
        so = os.path.join(_HERE, "libbp_oracle_f32.so")
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-DCPO_FLOAT32", "-shared", "-o", so] + srcs)
        return so
    so = os.path.join(_HERE, "libbp_oracle.so")
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libbp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.cpo_space_new.restype = C.c_void_p
        L.cpo_space_free.argtypes = [C.c_void_p]
        L.cpo_space_clone.restype = C.c_void_p
        L.cpo_space_clone.argtypes = [C.c_void_p]
        L.cpo_space_set.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.cpo_space_set_order.argtypes = [C.c_void_p, C.c_int]
        L.cpo_body_new.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.cpo_poly_new.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_double]
        L.cpo_shape_set_mass.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpo_shape_set_density.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpo_shape_set_material.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.cpo_poly_centroid.argtypes = [C.c_int, dp, dp]
        L.cpo_shape_count.argtypes = [C.c_void_p, C.c_int]
        L.cpo_shape_get_verts.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpo_shape_get_mass.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpo_body_set_velocity.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.cpo_body_set_angular_velocity.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpo_body_get.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpo_num_bodies.argtypes = [C.c_void_p]
        L.cpo_get_state.argtypes = [C.c_void_p, dp]
        L.cpo_step.argtypes = [C.c_void_p, C.c_double]
        L.cpo_set_state.argtypes = [C.c_void_p, dp]
        L.cpo_clear_cache.argtypes = [C.c_void_p]
        L.cpo_shipice_substeps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.cpo_segment_new.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        L.cpo_maze_substeps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.cpo_hits_13.argtypes = [C.c_void_p]
        L.cpo_hits_13.restype = C.c_longlong
        L.cpo_point_query_any.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.cpo_point_query_types.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.cpo_space_set_handler_mode.argtypes = [C.c_void_p, C.c_int]
        L.cpo_robot_hit.argtypes = [C.c_void_p]
        L.cpo_set_robot_hit.argtypes = [C.c_void_p, C.c_int]
        L.cpo_body_set_position.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.cpo_body_remove.argtypes = [C.c_void_p, C.c_int]
        L.cpo_body_local_to_world.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, dp]
        L.cpo_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.cpo_reset_stats.argtypes = [C.c_void_p]
        L.cpo_get_arbiters.argtypes = [C.c_void_p, dp, C.c_int]
        L.spfa_oracle.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Space:
    """pymunk.Space restricted to what the 2D envs use."""

    def __init__(self, iterations=10, damping=1.0, gravity=(0.0, 0.0)):
        self.L = lib()
        self.h = C.c_void_p(self.L.cpo_space_new())
        self.L.cpo_space_set(self.h, int(iterations), float(damping), float(gravity[0]), float(gravity[1]))

    def __del__(self):
        try:
            self.L.cpo_space_free(self.h)
        except Exception:
            pass

    def clone(self):
        """Deep copy (bodies, shapes, arbiter cache, counters)."""
        c = Space.__new__(Space)
        c.L = self.L
        c.h = C.c_void_p(self.L.cpo_space_clone(self.h))
        return c

    def set_order(self, mode):
        self.L.cpo_space_set_order(self.h, int(mode))

    def add_body(self, body_type, x, y, angle):
        return self.L.cpo_body_new(self.h, body_type, float(x), float(y), float(angle))

    def add_poly(self, body, verts, radius):
        v = np.ascontiguousarray(verts, dtype=np.float64)
        return self.L.cpo_poly_new(self.h, body, v.shape[0], _dp(v), float(radius))

    def add_segment(self, body, a, b, radius):
        return self.L.cpo_segment_new(self.h, body, float(a[0]), float(a[1]), float(b[0]), float(b[1]), float(radius))

    def maze_substeps(self, robot_body, steps, dt, width):
        return self.L.cpo_maze_substeps(self.h, robot_body, int(steps), float(dt), float(width))

    def hits_13(self):
        """pre_solve calls of the (1, 3) handler so far (robot chassis vs wall)."""
        return int(self.L.cpo_hits_13(self.h))

    def point_query_any(self, p, max_dist):
        """bool(space.point_query(p, max_dist, ShapeFilter()))"""
        return bool(self.L.cpo_point_query_any(self.h, float(p[0]), float(p[1]), float(max_dist)))

    def point_query_types(self, p, max_dist):
        """bit c set = a shape of collision_type c is in space.point_query(p, max_dist, ShapeFilter())"""
        return int(self.L.cpo_point_query_types(self.h, float(p[0]), float(p[1]), float(max_dist)))

    def set_handler_mode(self, mode):
        """1 = the box-delivery / area-clearing handler set (see chipmunk_oracle.cpp Space::handler_mode)"""
        self.L.cpo_space_set_handler_mode(self.h, int(mode))

    def robot_hit(self):
        return int(self.L.cpo_robot_hit(self.h))

    def set_robot_hit(self, v):
        self.L.cpo_set_robot_hit(self.h, int(v))

    def set_position(self, body, x, y):
        self.L.cpo_body_set_position(self.h, body, float(x), float(y))

    def remove_body(self, body):
        self.L.cpo_body_remove(self.h, body)

    def local_to_world(self, body, x, y):
        out = np.empty(2, dtype=np.float64)
        self.L.cpo_body_local_to_world(self.h, body, float(x), float(y), _dp(out))
        return out

    def set_density(self, shape, rho):
        self.L.cpo_shape_set_density(self.h, shape, float(rho))

    def set_mass(self, shape, m):
        self.L.cpo_shape_set_mass(self.h, shape, float(m))

    def set_material(self, shape, e, u, ctype=0):
        self.L.cpo_shape_set_material(self.h, shape, float(e), float(u), int(ctype))

    def get_vertices(self, shape):
        n = self.L.cpo_shape_count(self.h, shape)
        out = np.empty((n, 2), dtype=np.float64)
        self.L.cpo_shape_get_verts(self.h, shape, _dp(out))
        return out

    def shape_mass(self, shape):
        out = np.empty(4, dtype=np.float64)
        self.L.cpo_shape_get_mass(self.h, shape, _dp(out))
        return out

    def set_velocity(self, body, vx, vy):
        self.L.cpo_body_set_velocity(self.h, body, float(vx), float(vy))

    def set_angular_velocity(self, body, w):
        self.L.cpo_body_set_angular_velocity(self.h, body, float(w))

    def body(self, b):
        out = np.empty(8, dtype=np.float64)
        self.L.cpo_body_get(self.h, b, _dp(out))
        return out

    def num_bodies(self):
        return self.L.cpo_num_bodies(self.h)

    def state(self):
        out = np.empty((self.num_bodies(), 6), dtype=np.float64)
        self.L.cpo_get_state(self.h, _dp(out))
        return out

    def set_state(self, st):
        """Counterpart of :meth:`state`: poses and velocities of every body from ``st`` [n, 6]."""
        st = np.ascontiguousarray(st, dtype=np.float64)
        assert st.shape == (self.num_bodies(), 6)
        self.L.cpo_set_state(self.h, _dp(st))

    def clear_cache(self):
        self.L.cpo_clear_cache(self.h)

    def step(self, dt):
        self.L.cpo_step(self.h, float(dt))

    def shipice_substeps(self, ship_body, steps, dt, yaw_lo, yaw_hi, map_width):
        return self.L.cpo_shipice_substeps(self.h, ship_body, int(steps), float(dt), float(yaw_lo), float(yaw_hi), float(map_width))

    def stats(self):
        out = (C.c_longlong * 11)()
        self.L.cpo_get_stats(self.h, out)
        keys = ["steps", "pairs_bb", "pairs_touch", "arb_solved", "arb_nonzero", "gjk_iters", "epa_calls",
                "max_arb", "max_nonzero", "moved_bodies", "max_moved"]
        return dict(zip(keys, list(out)))

    def reset_stats(self):
        self.L.cpo_reset_stats(self.h)

    def arbiters(self, max_rows=8192):
        out = np.zeros((max_rows, 7), dtype=np.float64)
        n = self.L.cpo_get_arbiters(self.h, _dp(out), max_rows)
        return out[:min(n, max_rows)]


def poly_centroid_hull(verts):
    """pymunk.Poly(None, verts).center_of_gravity (sim_utils.py:140-141)."""
    v = np.ascontiguousarray(verts, dtype=np.float64)
    out = np.empty(2, dtype=np.float64)
    lib().cpo_poly_centroid(v.shape[0], _dp(v), _dp(out))
    return out


def spfa(input_map, source):
    """Port of spfa.spfa (deps/spfa/src/main.cpp:10): bool[H,W], (i,j) -> (f32[H,W], i32[H,W])."""
    m = np.ascontiguousarray(np.asarray(input_map).astype(bool).astype(np.uint8))
    H, W = m.shape
    d = np.empty((H, W), dtype=np.float32)
    p = np.empty((H, W), dtype=np.int32)
    lib().spfa_oracle(m.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, int(source[0]), int(source[1]),
                      d.ctypes.data_as(C.POINTER(C.c_float)), p.ctypes.data_as(C.POINTER(C.c_int32)))
    return d, p
