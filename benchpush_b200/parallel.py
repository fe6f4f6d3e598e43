"""Multi-GPU layout of the path: environments are independent (no cross-env coupling anywhere in
the reference), so the batch is split into contiguous blocks of global env ids, one process per
GPU, with NO data-path collective; the only exchange is the all-reduce of episode-metric sums
(``ShipIceMetric.reduce``).  Layout / trial selection is keyed by the GLOBAL env id so results do
not depend on the number of ranks."""
from __future__ import annotations


def shard_range(n_envs: int, rank: int, world: int):
    """Contiguous block [lo, hi) of global env ids owned by ``rank``."""
    base, rem = divmod(n_envs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi
