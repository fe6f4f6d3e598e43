"""ctypes loader for the TEST-ONLY host emulation of the env kernels (tests/emu/emu_harness.cpp)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "benchpush_b200", "csrc")


def build():
    # BP_EMU_FLAGS: extra compiler flags (e.g. "-DBP_RES_PA=20" to shrink the resident pool and exercise parking); each set of
    # flags builds its own library
    extra = os.environ.get("BP_EMU_FLAGS", "").split()
    tag = ("_" + "".join(c if c.isalnum() else "_" for c in "".join(extra))) if extra else ""
    so = os.path.join(_HERE, "libbp_emu%s.so" % tag)
    srcs = [os.path.join(_HERE, "emu_harness.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared"] + extra +
                              ["-o", so, os.path.join(_HERE, "emu_harness.cpp")])
    return so


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Emu:
    def __init__(self, n_envs, packed, ship_verts, pmax=8192, lmax=1024, goal_y=9.0, settle_steps=1000,
                 margin_body=0.3, margin_robot=0.6):
        L = C.CDLL(build())
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]
        L.emu_env_bytes.restype = C.c_size_t
        for f in ("emu_destroy", "emu_load_layouts", "emu_reset", "emu_step", "emu_set_robot_velocity", "emu_substeps",
                  "emu_init_only", "emu_sleep_pass", "emu_get_bodies", "emu_get_out", "emu_get_scal", "emu_layout_verts", "emu_env_bytes"):
            getattr(L, f).argtypes = None
        self.L = L
        self.p = packed
        self.h = C.c_void_p(L.emu_create(n_envs, packed.max_floes, pmax, lmax, float(goal_y), settle_steps, float(margin_body), float(margin_robot)))
        sv = np.ascontiguousarray(ship_verts, dtype=np.float64)
        rc = L.emu_load_layouts(self.h, C.c_int(packed.n_layouts), _ip(packed.n_floes), _ip(np.ascontiguousarray(packed.n_verts)),
                                _dp(np.ascontiguousarray(packed.verts)), _dp(np.ascontiguousarray(packed.centres)),
                                _dp(np.ascontiguousarray(packed.ship_state)), _dp(sv), C.c_int(sv.shape[0]))
        assert rc == 0

    def reset(self, env, layout):
        self.L.emu_reset(self.h, C.c_int(env), C.c_int(layout))

    def init_only(self, env, layout):
        self.L.emu_init_only(self.h, C.c_int(env), C.c_int(layout))

    def step(self, env, action):
        self.L.emu_step(self.h, C.c_int(env), C.c_double(action))

    def step_batch(self, env0, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self.L.emu_step_batch(self.h, C.c_int(env0), a.ctypes.data_as(C.c_void_p), C.c_int(1 if auto_reset else 0))

    def step_resident(self, env0, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self.L.emu_step_resident(self.h, C.c_int(env0), a.ctypes.data_as(C.c_void_p), C.c_int(1 if auto_reset else 0))

    def set_bodies(self, env, bodies):
        b = np.ascontiguousarray(bodies, dtype=np.float64)
        pad = np.zeros((self.p.max_floes + 1, 6)); pad[:len(b)] = b
        self.L.emu_set_bodies(self.h, C.c_int(env), _dp(pad))

    def batch_size(self):
        return int(self.L.emu_batch_size())

    def set_robot_velocity(self, env, vx, vy, w):
        self.L.emu_set_robot_velocity(self.h, C.c_int(env), C.c_double(vx), C.c_double(vy), C.c_double(w))

    def substeps(self, env, n, first=False):
        self.L.emu_substeps(self.h, C.c_int(env), C.c_int(n), C.c_int(1 if first else 0))

    def sleep_pass(self, env):
        self.L.emu_sleep_pass(self.h, C.c_int(env))

    def bodies(self, env, layout):
        nb = int(self.p.n_floes[layout]) + 1
        out = np.zeros((nb, 6), dtype=np.float64)
        self.L.emu_get_bodies(self.h, C.c_int(env), _dp(out))
        return out

    def out(self, env):
        o = np.zeros(13, dtype=np.float64)
        self.L.emu_get_out(self.h, C.c_int(env), _dp(o))
        keys = ["reward", "terminated", "truncated", "sx", "sy", "sa", "total_work", "collision_reward",
                "scaled_collision_reward", "dist_reward", "trial_success", "status", "angle"]
        return dict(zip(keys, o))

    def scal(self, env):
        o = (C.c_longlong * 18)()
        self.L.emu_get_scal(self.h, C.c_int(env), o)
        keys = ["stamp", "n_awake", "n_pairs", "n_free", "n_order", "n_colours", "status", "n_apairs", "n_cached", "gjk",
                "pairs_eval", "live_sum", "awake_sum", "recolour", "rebuild", "substeps", "epa", "skipped"]
        return dict(zip(keys, list(o)))

    def layout_verts(self, l, b):
        out = np.zeros((20, 2), dtype=np.float64)
        mi = np.zeros(4, dtype=np.float64)
        n = self.L.emu_layout_verts(self.h, C.c_int(l), C.c_int(b), _dp(out), _dp(mi))
        return out[:n], mi

    def render_obs(self, env, raw):
        out = np.zeros((4, 150, 150), dtype=np.uint8)
        self.L.emu_render_obs(self.h, C.c_int(env), C.c_int(1 if raw else 0), out.ctypes.data_as(C.c_void_p))
        return out

    def render_obs_global(self, env, raw, shape=(2, 200, 60)):
        out = np.full(shape, 7, dtype=np.uint8)              # the kernel must overwrite every byte
        self.L.emu_render_obs_global(self.h, C.c_int(env), C.c_int(1 if raw else 0), out.ctypes.data_as(C.c_void_p))
        return out

    def centroid_cache(self, env, max_floes):
        out = np.zeros(2 * max_floes, dtype=np.float64)
        n = self.L.emu_centroid_cache(self.h, C.c_int(env), _dp(out))
        return out[:2 * n]

    def spfa_serial(self, m, source):
        m = np.ascontiguousarray(np.asarray(m).astype(bool).astype(np.uint8))
        H, W = m.shape
        d = np.zeros((H, W), dtype=np.float32); p = np.zeros((H, W), dtype=np.int32)
        self.L.emu_spfa_serial(m.ctypes.data_as(C.c_void_p), C.c_int(H), C.c_int(W), C.c_int(int(source[0])), C.c_int(int(source[1])),
                               d.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
        return d, p

    def world_verts(self, env, b):
        out = np.zeros(2 * 64, dtype=np.float64)
        n = self.L.emu_world_verts(self.h, C.c_int(env), C.c_int(b), _dp(out))
        return out[:2 * n].reshape(n, 2).copy()

    def spfa(self, m, source):
        m = np.ascontiguousarray(np.asarray(m).astype(bool).astype(np.uint8))
        H, W = m.shape
        d = np.empty((H, W), dtype=np.float32)
        p = np.empty((H, W), dtype=np.int32)
        self.L.emu_spfa(m.ctypes.data_as(C.c_void_p), C.c_int(H), C.c_int(W), C.c_int(int(source[0])), C.c_int(int(source[1])),
                        d.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
        return d, p

    def env_bytes(self):
        return self.L.emu_env_bytes(self.h)


class MazeEmu(Emu):
    """Host emulation of the maze-NAMO kernels (default cfg of maze_version 1 or 2); layouts = list of centre lists."""

    def __init__(self, n_envs, layouts, max_boxes=5, maze_version=1, settle_steps=1000):
        L = C.CDLL(build())
        L.emu_maze_create.restype = C.c_void_p
        L.emu_maze_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_env_bytes.restype = C.c_size_t
        self.L = L
        self.max_boxes = max_boxes
        self.layouts = layouts
        self.h = C.c_void_p(L.emu_maze_create(n_envs, max_boxes, maze_version, settle_steps))
        nb = np.array([len(l) for l in layouts], dtype=np.int32)
        cen = np.zeros((len(layouts), max_boxes, 2), dtype=np.float64)
        for i, l in enumerate(layouts):
            cen[i, :len(l)] = np.asarray(l, dtype=np.float64).reshape(-1, 2)
        rc = L.emu_maze_load_layouts(self.h, C.c_int(len(layouts)), C.c_int(max_boxes), _ip(nb), _dp(cen))
        assert rc == 0
        self.grid = int((20.0 if maze_version == 2 else 15.0) * 16)

    def bodies(self, env, layout=None):
        sc = (C.c_int * 8)()
        self.L.emu_get_scal2(self.h, C.c_int(env), sc)
        out = np.zeros((int(sc[1]), 6), dtype=np.float64)
        self.L.emu_get_bodies(self.h, C.c_int(env), _dp(out))
        return out

    def scal2(self, env):
        sc = (C.c_int * 8)()
        self.L.emu_get_scal2(self.h, C.c_int(env), sc)
        return dict(robot=int(sc[0]), nbody=int(sc[1]), n_fixed=int(sc[2]), wall_hit=int(sc[3]), has_prev_dist=int(sc[4]))

    def maps(self):
        n = self.grid
        wall = np.zeros((n, n), dtype=np.uint8); dn = np.zeros((n, n)); dr = np.zeros((n, n))
        self.L.emu_maze_get_maps(self.h, wall.ctypes.data_as(C.c_void_p), _dp(dn), _dp(dr))
        return wall, dn, dr

    def render_obs(self, env, raw):
        out = np.zeros((4, 192, 192), dtype=np.uint8)
        self.L.emu_maze_render_obs(self.h, C.c_int(env), C.c_int(1 if raw else 0), out.ctypes.data_as(C.c_void_p))
        return out
