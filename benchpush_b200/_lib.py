"""ctypes binding of ``benchpush_b200/csrc/libbenchpush_b200.so`` (C ABI: include/benchpush_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is usable, loading / creating
a handle raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# BENCHPUSH_B200_LIB: development aid to load a differently-tuned build of the same CUDA library
LIB_PATH = os.environ.get("BENCHPUSH_B200_LIB") or os.path.join(CSRC, "libbenchpush_b200.so")
# The same sources built with fp32 body state / contacts / solver (csrc/Makefile, -DBP_F32): ship-ice and maze only.
LIB_PATH_F32 = os.environ.get("BENCHPUSH_B200_LIB_F32") or os.path.join(CSRC, "libbenchpush_b200_f32.so")

BP_VMAX = 20
OBS_SHAPE = (4, 150, 150)


class BenchPushError(RuntimeError):
    pass


class ShipIceConfig(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("max_floes", C.c_int32), ("device", C.c_int32), ("pair_capacity", C.c_int32),
        ("live_capacity", C.c_int32), ("auto_reset", C.c_int32),
        ("dt", C.c_double), ("steps", C.c_int32), ("iterations", C.c_int32), ("damping", C.c_double),
        ("obstacle_density", C.c_double), ("target_speed", C.c_double), ("goal_y", C.c_double),
        ("map_width", C.c_double), ("map_height", C.c_double), ("m_to_pix_scale", C.c_int32), ("settle_steps", C.c_int32),
        ("n_ship_verts", C.c_int32), ("ship_verts", C.c_double * 64), ("ship_head", C.c_double * 2), ("ship_tail", C.c_double * 2),
    ]


class ShipIceOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "obs", "reward", "terminated", "truncated", "state", "pose", "total_work", "collision_reward",
        "scaled_collision_reward", "dist_reward", "trial_success", "status")]


class MazeConfig(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("max_boxes", C.c_int32), ("device", C.c_int32), ("auto_reset", C.c_int32),
        ("dt", C.c_double), ("steps", C.c_int32), ("iterations", C.c_int32), ("damping", C.c_double),
        ("obstacle_density", C.c_double), ("obstacle_size", C.c_double), ("target_speed", C.c_double),
        ("width", C.c_double), ("length", C.c_double), ("goal_x", C.c_double), ("goal_y", C.c_double),
        ("goal_radius", C.c_double), ("min_r", C.c_double), ("m_to_pix_scale", C.c_int32), ("settle_steps", C.c_int32),
        ("start", C.c_double * 3), ("n_walls", C.c_int32), ("walls", C.c_double * 64),
        ("n_robot_verts", C.c_int32), ("robot_verts", C.c_double * 64),
        ("n_wheels", C.c_int32), ("wheel_nv", C.c_int32 * 8), ("wheel_verts", C.c_double * 128),
    ]


MAZE_OBS_SHAPE = (4, 192, 192)


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if res.returncode != 0:
        raise BenchPushError("nvcc build failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


_LIB = None
_LIBS = {}


def lib(precision: str = "fp64"):
    """``precision``: "fp64" = libbenchpush_b200.so (bit-comparable with the fp64 oracle; every env family), "fp32" =
    libbenchpush_b200_f32.so (ship-ice / maze with fp32 body state, contacts and solver; same C ABI, outputs stay double)."""
    global _LIB
    if precision not in ("fp64", "fp32"):
        raise ValueError("precision must be 'fp64' or 'fp32'")
    if precision in _LIBS:
        return _LIBS[precision]
    if True:
        path = LIB_PATH if precision == "fp64" else LIB_PATH_F32
        if not os.path.exists(path):
            raise BenchPushError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(path)
        L.bp_version.restype = C.c_int
        L.bp_last_error.restype = C.c_char_p
        L.bp_device_count.restype = C.c_int
        L.bp_shipice_default_config.argtypes = [C.POINTER(ShipIceConfig)]
        L.bp_shipice_default_config.restype = None
        L.bp_shipice_create.argtypes = [C.POINTER(ShipIceConfig), C.POINTER(C.c_void_p)]
        L.bp_shipice_destroy.argtypes = [C.c_void_p]
        L.bp_shipice_load_layouts.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
        L.bp_shipice_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_shipice_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_shipice_step_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_shipice_step_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs)]
        L.bp_shipice_step_host_obs_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_maze_step_host_obs_device.argtypes = L.bp_shipice_step_host_obs_device.argtypes
        L.bp_shipice_get_obstacles.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.bp_shipice_low_dim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bp_maze_low_dim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bp_shipice_set_start_poses.argtypes = [C.c_void_p, C.c_void_p]
        L.bp_maze_set_start_poses.argtypes = [C.c_void_p, C.c_void_p]
        L.bp_shipice_obs_global_shape.argtypes = [C.c_void_p, C.c_void_p]
        L.bp_shipice_obs_global.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bp_shipice_get_bodies.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        for fam in ("shipice", "maze"):
            getattr(L, "bp_%s_state_bytes" % fam).argtypes = [C.c_void_p]
            getattr(L, "bp_%s_state_bytes" % fam).restype = C.c_size_t
            getattr(L, "bp_%s_get_state" % fam).argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            getattr(L, "bp_%s_set_state" % fam).argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            getattr(L, "bp_%s_set_bodies" % fam).argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.bp_shipice_get_counters.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.bp_shipice_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        L.bp_shipice_get_kernel_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.bp_shipice_env_bytes.argtypes = [C.c_void_p]
        L.bp_shipice_env_bytes.restype = C.c_size_t
        L.bp_shipice_launch_count.argtypes = [C.c_void_p]
        L.bp_shipice_launch_count.restype = C.c_int64
        L.bp_maze_default_config.argtypes = [C.POINTER(MazeConfig), C.c_int32]
        L.bp_maze_default_config.restype = None
        L.bp_maze_create.argtypes = [C.POINTER(MazeConfig), C.POINTER(C.c_void_p)]
        L.bp_maze_destroy.argtypes = [C.c_void_p]
        L.bp_maze_load_layouts.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.bp_maze_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_maze_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_maze_step_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs), C.c_void_p]
        L.bp_maze_step_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(ShipIceOutputs)]
        L.bp_maze_get_obstacles.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.bp_maze_get_bodies.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.bp_maze_get_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bp_spfa_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bp_spfa_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.bp_spfa_batch_exact.argtypes = L.bp_spfa_batch.argtypes
        L.bp_spfa_batch_exact_host.argtypes = L.bp_spfa_batch_host.argtypes
        _LIBS[precision] = L
        if precision == "fp64":
            _LIB = L
    return _LIBS[precision]


EXPORTS = [
    "bp_version", "bp_last_error", "bp_device_count", "bp_shipice_default_config", "bp_shipice_create",
    "bp_shipice_destroy", "bp_shipice_load_layouts", "bp_shipice_reset", "bp_shipice_step", "bp_shipice_step_masked", "bp_shipice_step_host", "bp_shipice_step_host_obs_device",
    "bp_shipice_state_bytes", "bp_shipice_get_state", "bp_shipice_set_state", "bp_shipice_set_bodies",
    "bp_maze_state_bytes", "bp_maze_get_state", "bp_maze_set_state", "bp_maze_set_bodies",
    "bp_shipice_get_obstacles", "bp_shipice_get_bodies", "bp_shipice_get_counters", "bp_shipice_env_bytes", "bp_shipice_low_dim", "bp_maze_low_dim", "bp_shipice_obs_global", "bp_shipice_obs_global_shape", "bp_shipice_set_start_poses", "bp_maze_set_start_poses",
    "bp_shipice_set_profiling", "bp_shipice_get_kernel_ms",
    "bp_shipice_launch_count", "bp_spfa_batch", "bp_spfa_batch_host", "bp_spfa_batch_exact", "bp_spfa_batch_exact_host",
    "bp_maze_default_config", "bp_maze_create", "bp_maze_destroy", "bp_maze_load_layouts", "bp_maze_reset", "bp_maze_step",
    "bp_maze_step_masked", "bp_maze_step_host_obs_device", "bp_maze_step_host", "bp_maze_get_obstacles", "bp_maze_get_bodies", "bp_maze_get_maps",
    "bp_boxdel_default_config", "bp_boxdel_create", "bp_boxdel_destroy", "bp_boxdel_load", "bp_boxdel_load_rooms", "bp_boxdel_load_rooms_device", "bp_boxdel_get_room_maps", "bp_boxdel_reset", "bp_boxdel_step", "bp_boxdel_step_host_obs_device",
    "bp_boxdel_step_host", "bp_boxdel_get_bodies", "bp_boxdel_get_step_debug", "bp_boxdel_launch_count", "bp_boxdel_env_bytes", "bp_boxdel_set_profiling", "bp_boxdel_get_kernel_ms",
]


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().bp_last_error().decode("utf-8", "replace")
        raise BenchPushError(f"benchpush_b200 error {rc}: {msg}")
    return rc
