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


# ---- the product's host mirrors on the emulation library: numpy buffers instead of CUDA tensors (tests only)
class _HostBuffers:
    def _open(self, device):
        from benchpush_b200.envs.box_delivery import bind
        self._L = bind(lib())
        self.device, self.device_index = None, 0

    def _alloc(self, name, shape, dtype):
        import numpy as np
        self._buf[name] = np.zeros(shape, dtype=dtype)

    def _ptr(self, name):
        return self._buf[name].ctypes.data

    def _put(self, name, values):
        import numpy as np
        self._buf[name][...] = np.asarray(values, dtype=self._buf[name].dtype).reshape(self._buf[name].shape)

    def _stream(self):
        return None

    @staticmethod
    def _as_bool(x):
        return x.astype(bool)

    def low_dim_observation(self):
        """numpy twin of BoxDeliveryVecEnv.low_dim_observation (generate_observation_low_dim)"""
        import numpy as np
        v = self._buf["box_verts"]
        n = self._buf["info_i"][:, 2]
        x, y = v[..., 0], v[..., 1]
        xp, yp = np.roll(x, 1, axis=2), np.roll(y, 1, axis=2)
        A = 0.5 * np.abs((x * yp).sum(2) - (y * xp).sum(2))
        u = x * yp - xp * y
        with np.errstate(divide="ignore", invalid="ignore"):
            cx = np.abs(((x + xp) * u).sum(2) / (6 * A)); cy = np.abs(((y + yp) * u).sum(2) / (6 * A))
        valid = np.arange(self.num_boxes)[None, :] < n[:, None]
        out = np.stack([np.where(valid, cx, 0.0), np.where(valid, cy, 0.0)], axis=2)
        return out.reshape(self.num_envs, 2 * self.num_boxes)


def _classes():
    from benchpush_b200.envs.area_clearing import AreaClearingVecEnv
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv

    class EmuBoxDeliveryVecEnv(_HostBuffers, BoxDeliveryVecEnv):
        pass

    class EmuAreaClearingVecEnv(_HostBuffers, AreaClearingVecEnv):
        pass

    from benchpush_b200.envs.gym_single import AreaClearingEnvB200, BoxDeliveryEnvB200

    class EmuBoxDeliveryEnvB200(BoxDeliveryEnvB200):
        _vec_class = staticmethod(lambda: EmuBoxDeliveryVecEnv)

    class EmuAreaClearingEnvB200(AreaClearingEnvB200):
        _vec_class = staticmethod(lambda: EmuAreaClearingVecEnv)

    return EmuBoxDeliveryVecEnv, EmuAreaClearingVecEnv, EmuBoxDeliveryEnvB200, EmuAreaClearingEnvB200


EmuBoxDeliveryVecEnv, EmuAreaClearingVecEnv, EmuBoxDeliveryEnvB200, EmuAreaClearingEnvB200 = _classes()
