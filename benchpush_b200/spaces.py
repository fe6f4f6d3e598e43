"""`action_space` / `observation_space` of the envs (part of the gym.Env surface the reference's callers use —
SB3 reads `.shape`, `.dtype`, `.low`, `.high`, `sample()`).  gymnasium is not a dependency: when it is importable the real
`gymnasium.spaces.Box` is returned, otherwise this duck-typed stand-in with the same attributes."""
from __future__ import annotations

import numpy as np


class Box:
    """Minimal `gymnasium.spaces.Box`: elementwise bounds, shape, dtype, `sample`, `contains`."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self._shape = tuple(int(s) for s in shape)
        self.low = np.full(self._shape, low, dtype=self.dtype) if np.isscalar(low) else np.broadcast_to(np.asarray(low, dtype=self.dtype), self._shape).copy()
        self.high = np.full(self._shape, high, dtype=self.dtype) if np.isscalar(high) else np.broadcast_to(np.asarray(high, dtype=self.dtype), self._shape).copy()
        self._rng = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        if np.issubdtype(self.dtype, np.floating):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)
        return self._rng.integers(self.low.astype(np.int64), self.high.astype(np.int64) + 1).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self._shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min() if self.low.size else self.low, self.high.max() if self.high.size else self.high, self._shape, self.dtype)

    def __eq__(self, other):
        return (hasattr(other, "low") and hasattr(other, "high") and tuple(other.shape) == self._shape and np.dtype(other.dtype) == self.dtype
                and np.array_equal(np.asarray(other.low), self.low) and np.array_equal(np.asarray(other.high), self.high))


def box(low, high, shape=None, dtype=np.float32):
    try:                                              # the real class when gymnasium is installed (SB3 type-checks it)
        from gymnasium.spaces import Box as GymBox
        return GymBox(low=low, high=high, shape=shape, dtype=dtype)
    except Exception:
        return Box(low, high, shape, dtype)


def batched(space, n: int):
    """The vector envs' batch space: the single-env space with a leading axis of `n` envs."""
    return box(np.broadcast_to(np.asarray(space.low), (n,) + tuple(space.shape)), np.broadcast_to(np.asarray(space.high), (n,) + tuple(space.shape)),
               (n,) + tuple(space.shape), space.dtype)
