"""ctypes loader for the TEST-ONLY host emulation of the box-delivery kernels (tests/emu/emu_bd.cpp)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "benchpush_b200", "csrc")
_LIB = None


def build():
    so = os.path.join(_HERE, "libbp_emu_bd.so")
    srcs = [os.path.join(_HERE, "emu_bd.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        # BD_MOVE_K=3: the emulation steps three envs per "warp" so that the phase-major multi-env path (and its ragged tail) is covered
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-DBD_MOVE_K=3", "-shared", "-o", so, os.path.join(_HERE, "emu_bd.cpp")])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.bp_last_error.restype = C.c_char_p
    return _LIB
