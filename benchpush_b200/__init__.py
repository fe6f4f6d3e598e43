"""benchpush_b200 — B200-native batched simulator for BenchPush's 2D pushing environments.

Hot path only: reset/step/observation/reward of the pymunk environments and the
spfa distance maps, as hand-written sm_100a CUDA kernels behind a C ABI
(``include/benchpush_b200.h``).  No CPU fallback.
"""
from ._lib import BenchPushError, build, lib  # noqa: F401
from .utils import DotDict  # noqa: F401

__version__ = "0.1.0"


def make(env_id, num_envs, cfg=None, **kw):
    from .envs import make as _make
    return _make(env_id, num_envs, cfg=cfg, **kw)
