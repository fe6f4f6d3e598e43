"""Batched environments.  Ids follow the reference registry (benchpush/environments/__init__.py:3-51)."""
from .area_clearing import AreaClearingVecEnv
from .box_delivery import BoxDeliveryVecEnv
from .maze_namo import MazeNamoVecEnv
from .ship_ice import ShipIceVecEnv

REGISTRY = {"ship-ice-v0": ShipIceVecEnv, "maze-NAMO-v0": MazeNamoVecEnv, "box-delivery-v0": BoxDeliveryVecEnv, "area-clearing-v0": AreaClearingVecEnv}
# max_episode_steps of the reference registry (benchpush/environments/__init__.py:3-51)
MAX_EPISODE_STEPS = {"ship-ice-v0": 300, "box-delivery-v0": 30000, "maze-NAMO-v0": 400, "area-clearing-v0": 30000}


def make(env_id: str, num_envs: int, cfg=None, **kw):
    """Vector analogue of ``gym.make(id, cfg=cfg)``."""
    if env_id not in REGISTRY:
        raise KeyError(f"{env_id!r} is not available in benchpush_b200 (have: {sorted(REGISTRY)})")
    return REGISTRY[env_id](num_envs=num_envs, cfg=cfg, **kw)
